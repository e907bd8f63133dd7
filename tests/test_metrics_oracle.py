"""CPU tests pinning the oracle restatements behind the metric / fusion kernels
to numpy / scipy (the libraries the reference calls) and to the goldens produced by the reference itself."""
import os

import numpy as np
import pytest
import torch

from oracle import av_oracle as o


def test_pairwise_mean_restatement_equals_np_mean():
    rng = np.random.default_rng(0)
    for n in [1, 5, 8, 9, 127, 128, 129, 320, 700, 1000, 4097, 8192]:
        for dt in (np.float32, np.float64):
            a = rng.random(n).astype(dt)
            assert o.np_mean_pairwise(a) == np.mean(a), (n, dt)


def test_kendall_integer_restatement_is_bit_exact_vs_scipy():
    from scipy.stats import kendalltau
    rng = np.random.default_rng(1)
    for n in [2, 3, 10, 57, 300]:
        for levels in (4, 1000):
            x = rng.integers(0, levels, n).astype(np.float64)
            y = rng.integers(0, levels, n).astype(np.float64)
            dis, xt, yt, nt, tot = o.kendall_counts(x, y)
            if xt == tot or yt == tot:
                continue
            tau = (tot - xt - yt + nt - 2 * dis) / np.sqrt(tot - xt) / np.sqrt(tot - yt)
            assert min(1.0, max(-1.0, tau)) == kendalltau(x, y).correlation
            # float32 inputs: scipy rounds the float64 result once to float32
            assert np.float32(tau) == kendalltau(x.astype(np.float32), y.astype(np.float32)).correlation


def test_cdist_and_interpolate_restatements_match_reference_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "helpers.npz"))
    gen = torch.Generator().manual_seed(11)
    fv = torch.randn(23, 64, generator=gen).numpy()
    fa = torch.randn(31, 64, generator=gen).numpy()
    assert np.array_equal(o.cdist_euclidean(fv, fa), g["dtw"])            # bit-exact vs the reference's output
    assert np.array_equal(o.interpolate_features(fv, g["path"], 20), g["interp"])


def test_dtw_path_is_optimal_and_monotone():
    rng = np.random.default_rng(2)
    for n, m in [(1, 1), (1, 6), (5, 1), (7, 9), (12, 12)]:
        cost = rng.integers(0, 4, (n, m)).astype(np.float64)             # many ties
        total, path = o.dtw_path(cost)
        acc = np.full((n + 1, m + 1), np.inf)
        acc[0, 0] = 0
        for i in range(1, n + 1):
            for j in range(1, m + 1):
                acc[i, j] = cost[i - 1, j - 1] + min(acc[i - 1, j], acc[i, j - 1], acc[i - 1, j - 1])
        assert total == acc[n, m]
        assert tuple(path[0]) == (0, 0) and tuple(path[-1]) == (n - 1, m - 1)
        steps = np.diff(path, axis=0)
        assert np.all((steps >= 0) & (steps <= 1)) and np.all(steps.sum(axis=1) >= 1)
        assert sum(cost[i, j] for i, j in path) == total


def test_eval_metrics_oracle_uses_reference_formulas():
    rng = np.random.default_rng(3)
    pred = rng.random(200).astype(np.float32)
    target = rng.random(200).astype(np.float32)
    f1, rho, tau = o.eval_metrics(pred, target)
    assert f1 == o.threshold_f1(pred, target) and -1 <= rho <= 1 and -1 <= tau <= 1
