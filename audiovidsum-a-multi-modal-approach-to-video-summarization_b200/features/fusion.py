"""Drop-in for the reference's ``features/fusion.py`` (DTW fusion helpers), B200-native.

Same function names, argument meaning and return types as /root/reference/features/fusion.py:7-32:

* ``compute_dtw(visual, audio)`` -- the pairwise Euclidean distance matrix (scipy ``cdist``), returned as a
  float64 NumPy array; computed by the ``cdist_kernel`` behind ``avs_cdist`` in scipy's own summation
  order, so the result is bit-identical.  Like ``cdist`` it raises ``ValueError`` when the two feature
  dimensions differ (e.g. 1024-d visual vs 128-d audio).
* ``compute_optimal_path(dtw_matrix)`` -- in the reference this calls ``fastdtw(dtw_matrix, radius=10)``,
  which raises ``TypeError`` (fastdtw's signature is ``fastdtw(x, y, radius=1, dist=None)``); the default
  here reproduces that behaviour.  ``exact=True`` (an extension) returns what the function evidently
  intends -- the optimal warping path through the cost matrix -- from the wavefront DTW kernel behind
  ``avs_dtw_path`` (fastdtw's published recurrence and tie order; exact, so no radius).
* ``interpolate_features(features, path, target_length)`` -- rows ``unique(path[:, 0])`` of ``features``
  scaled by ``count / counts.sum()``, stacked and truncated to ``target_length`` (fusion.py:21-32); the
  gather-scale runs in ``gather_scale_kernel`` behind ``avs_interpolate``.

There is no CPU fallback: the helpers need a CUDA sm_100 device.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import runtime


def compute_dtw(visual, audio):
    """Compute DTW cost matrix (fusion.py:7-12): [Tv, D], [Ta, D] torch tensors -> float64 numpy [Tv, Ta]."""
    if visual.dim() != 2 or audio.dim() != 2:
        raise ValueError("XA must be a 2-dimensional array.") if visual.dim() != 2 else \
            ValueError("XB must be a 2-dimensional array.")
    if visual.shape[1] != audio.shape[1]:
        raise ValueError("XA and XB must have the same number of columns "
                         "(i.e. feature dimension.)")   # scipy.spatial.distance.cdist's message
    return runtime.cdist_euclidean(visual, audio)


def compute_optimal_path(dtw_matrix, exact: bool = False):
    """Get optimal warping path (fusion.py:15-18).

    Default: the reference's behaviour -- ``fastdtw(dtw_matrix, radius=10)`` lacks the required ``y``
    argument and raises ``TypeError``.  ``exact=True``: exact DTW path through ``dtw_matrix`` as an
    ``np.array`` of (i, j) pairs.
    """
    if not exact:
        raise TypeError("fastdtw() missing 1 required positional argument: 'y'")
    _, path = runtime.dtw_path(dtw_matrix)
    return np.array(path)


def interpolate_features(features, path, target_length):
    """Interpolate features using alignment path (fusion.py:21-32)."""
    path = np.asarray(path)
    aligned_indices = path[:, 0]
    unique_indices, counts = np.unique(aligned_indices, return_counts=True)
    weights = counts / counts.sum()
    if unique_indices.size == 0:
        raise RuntimeError("stack expects a non-empty TensorList")   # torch.stack([]) in the reference
    # negative indices address from the end, as features[idx] does in the reference
    n = int(features.shape[0])
    idx = np.where(unique_indices < 0, unique_indices + n, unique_indices)
    if idx.min() < 0 or idx.max() >= n:
        raise IndexError(f"index out of range for features with {n} rows")
    keep = min(max(int(target_length), 0) if target_length >= 0 else max(idx.size + int(target_length), 0), idx.size) \
        if target_length is not None else idx.size
    out = runtime.gather_scale(features, idx[:keep], weights[:keep].astype(np.float32))
    return out
