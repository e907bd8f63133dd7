"""Drop-in for the reference's ``utils/shot_metrics.py`` (a two-function split of
``evaluation/metrics.py``; /root/reference/utils/shot_metrics.py:4-16)."""
from __future__ import annotations

from ..evaluation.metrics import compute_temporal_f1


def calculate_overlap(pred_segments, gt_segments):
    """Total pairwise clipped overlap (shot_metrics.py:4-9); integer bookkeeping on the host."""
    overlap = 0
    for p_start, p_end in pred_segments:
        for g_start, g_end in gt_segments:
            overlap += max(0, min(p_end, g_end) - max(p_start, g_start))
    return overlap


def compute_f1(pred_segments, gt_segments, video_length):
    """shot_metrics.py:12-16 -- identical arithmetic to compute_temporal_f1, run on the GPU."""
    return compute_temporal_f1(pred_segments, gt_segments, video_length)
