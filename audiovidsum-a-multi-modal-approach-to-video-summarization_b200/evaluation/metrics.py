"""Drop-in for the reference's ``evaluation/metrics.py`` (overlap F1 between shot lists).

``compute_temporal_f1`` keeps the reference signature and result
(/root/reference/evaluation/metrics.py:1-9: pairwise clipped overlap, precision over the
predicted length, recall over the ground-truth length, F1 with a 1e-8 guard) but evaluates
the O(P*G) pair loop in the batched CUDA kernel behind ``avs_temporal_f1``;
``compute_temporal_f1_batch`` scores many videos in one launch.  The unused ``total_frames``
argument is kept for signature compatibility, as in the reference.
"""
from __future__ import annotations

import math

from .. import runtime


def compute_temporal_f1_batch(pred_shots_list, gt_shots_list):
    """float64 F1 per video; NaN where the reference would divide by zero."""
    return runtime.temporal_f1_batch(pred_shots_list, gt_shots_list)


def compute_temporal_f1(pred_shots, gt_shots, total_frames):
    f1 = float(compute_temporal_f1_batch([list(pred_shots)], [list(gt_shots)])[0])
    if math.isnan(f1) or math.isinf(f1):
        raise ZeroDivisionError("division by zero")  # what the reference raises for empty shot lists
    return f1
