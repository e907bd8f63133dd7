"""Drop-in for the reference's ``utils/alignments.py``.

``align_shots_to_annotations`` (/root/reference/utils/alignments.py:4-22) is host-side
bookkeeping executed once per training step on a handful of shots: each shot (start, end) in
frames maps to the 2-second annotation bins [int((start/fps)//2), int((end/fps)//2) + 1) whose
mean is the shot's target.  It is not on the inference hot path; it is kept so that callers of
the reference find the same function.
"""
from __future__ import annotations

import numpy as np
import torch


def align_shots_to_annotations(shot_boundaries, annotations, fps):
    annotations = np.asarray(annotations)
    shot_scores = []
    for start, end in shot_boundaries:
        first = int((start / fps) // 2)
        last = int((end / fps) // 2) + 1
        shot_scores.append(annotations[first:last].mean())
    return torch.tensor(shot_scores)
