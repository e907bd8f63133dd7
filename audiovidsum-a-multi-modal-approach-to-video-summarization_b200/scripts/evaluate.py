"""Drop-in for the reference's ``scripts/evaluate.py`` -- ``evaluate(model, dataset)``.

The reference (/root/reference/scripts/evaluate.py:6-42) scores one video at a time (B = 1, a synchronous
``.cuda()`` / ``.cpu()`` pair per video) and then computes, per video on the host, the mean-threshold F1,
Spearman's rho and Kendall's tau, returning their means.  Here the whole dataset is ONE packed
variable-length batch: one ``avs_forward`` (each video treated as its own B = 1 call, i.e. the reference's
semantics) and one ``avs_eval_metrics`` launch that evaluates the metric block for every video on the GPU;
only 3 doubles per video come back.  Return value: the same ``{"f1", "spearman", "kendall"}`` dict.
"""
from __future__ import annotations

import numpy as np
import torch


def evaluate(model, dataset, return_per_video: bool = False):
    model.eval()
    from ..data.dataset import DeviceDataset
    if isinstance(dataset, DeviceDataset):      # packed once, already on the GPU: no per-call H2D at all
        if len(dataset) == 0:
            nan = float("nan")
            return {"f1": nan, "spearman": nan, "kendall": nan}
        with torch.no_grad():
            axis = "literal_b1" if model.attn_axis == "literal" else model.attn_axis
            pred = model.native().forward_rows(dataset.visual, dataset.audio, dataset.row_start, dataset.lengths, axis,
                                               model.precision)
        return _metrics(pred, dataset.scores, dataset.row_start, dataset.lengths,
                        torch.float64 if dataset.scores.dtype == torch.float64 else torch.float32, return_per_video)
    visuals, audios, targets = [], [], []
    for features, scores in dataset:
        visuals.append(torch.as_tensor(features["visual"]))
        audios.append(torch.as_tensor(features["audio"]))
        targets.append(torch.as_tensor(scores).reshape(-1))
    if not visuals:
        nan = float("nan")   # np.mean([]) in the reference
        return {"f1": nan, "spearman": nan, "kendall": nan}
    lens = [int(v.shape[0]) for v in visuals]
    for n, t in zip(lens, targets):
        if int(t.numel()) != n:
            raise ValueError("every video needs one target score per frame")
    # rows are laid out longest video first (each recurrence group then owns one block of rows and the native call
    # runs its tail behind its own recurrence); the descriptors -- and so the per-video metrics -- stay in dataset order
    layout = sorted(range(len(lens)), key=lambda i: -lens[i])
    starts = np.zeros(len(lens), dtype=np.int32)
    at = 0
    for i in layout:
        starts[i] = at
        at += lens[i]
    dev = next(model.parameters()).device
    with torch.no_grad():
        axis = "literal_b1" if model.attn_axis == "literal" else model.attn_axis
        pred = model.native().forward_rows(torch.cat([visuals[i] for i in layout]).to(dev),
                                           torch.cat([audios[i] for i in layout]).to(dev), starts, lens, axis,
                                           model.precision)
    tgt_dtype = torch.float64 if any(t.dtype == torch.float64 for t in targets) else torch.float32
    target = torch.cat([targets[i].to(tgt_dtype) for i in layout]).to(dev)
    return _metrics(pred, target, starts, lens, tgt_dtype, return_per_video)


def _metrics(pred, target, starts, lens, tgt_dtype, return_per_video):
    from .. import runtime
    metrics, counts = runtime.eval_metrics_rows(pred, target, starts, lens)
    # scipy.stats.kendalltau returns the correlation in the dtype of its inputs when both are float32
    # (float64 arithmetic, rounded once), spearmanr always float64; np.mean then follows those dtypes
    kendalls = metrics[:, 2].astype(np.float32) if tgt_dtype == torch.float32 else metrics[:, 2]
    out = {
        "f1": np.mean(metrics[:, 0]),
        "spearman": np.mean(metrics[:, 1]),
        "kendall": np.mean(kendalls),
    }
    if return_per_video:
        return out, metrics, counts
    return out
