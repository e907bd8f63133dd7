"""Summary generation: shot pooling over change points + 0/1 knapsack at the length budget.

The reference contains no such code; BASELINE.json's north_star asks
for it and names it "the evaluation summary generator".  The arithmetic is specified, in
integers, by ``oracle/av_oracle.py`` and runs in the K7/K8 kernels of libavsum_b200.so
(``avs_summarize``); results are bit-exact against that oracle.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch


@dataclass
class VideoSummary:
    scores: torch.Tensor      # fp32 [T] frame scores
    picks: np.ndarray         # uint8 [S]: 1 where the shot is a keyshot
    seg_mean: np.ndarray      # int64 [S]: pooled shot score in 2^-24 fixed point
    summary: np.ndarray       # uint8 [n_frames] keyshot bitmap

    @property
    def keyshots(self):
        return np.flatnonzero(self.picks)


def generate_summary(model, ypred, cps, n_frames, positions, proportion=0.15):
    """Single-video convenience wrapper -> uint8 [n_frames] bitmap (de-facto TVSum/SumMe call shape).

    ``model`` is an ``AVBiLSTMModel`` (its native handle owns the device workspace).
    """
    dev = next(model.parameters()).device
    s = torch.as_tensor(np.asarray(ypred, dtype=np.float32)).to(dev)
    pos = torch.as_tensor(np.asarray(positions, dtype=np.int32)).to(dev)
    picks, _, summary, _, _ = model.native().summarize_rows(s, pos, [0], [s.numel()], [int(n_frames)], [cps], proportion)
    return summary.cpu().numpy()


@torch.no_grad()
def summarize_videos(model, videos, proportion=0.15, attn_axis: Optional[str] = None) -> List[VideoSummary]:
    """Score + pool + select for a batch of ``synth.Video``-like objects
    (attributes visual, audio, positions, n_frames, cps).

    Feature tensors may be host (pinned) tensors -- then every H2D / D2H copy happens inside
    the two native calls -- or tensors already on the model's GPU.
    """
    nat = model.native()
    # packed longest video first (the order data.dataset.packed_batches produces): the native call then pipelines
    # host-space batches by video group and runs each recurrence group's tail behind its own recurrence; the results
    # are handed back in the caller's order
    caller_order = sorted(range(len(videos)), key=lambda i: -int(videos[i].visual.shape[0]))
    videos = [videos[i] for i in caller_order]
    lens = [int(v.visual.shape[0]) for v in videos]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    visual = torch.cat([v.visual for v in videos], dim=0)
    audio = torch.cat([v.audio for v in videos], dim=0)
    axis = attn_axis or ("literal_b1" if model.attn_axis == "literal" else model.attn_axis)
    positions = torch.as_tensor(np.concatenate([np.asarray(v.positions, dtype=np.int32) for v in videos]))
    if visual.is_cuda:
        positions = positions.to(visual.device)
    if axis == "literal":     # cross-video mixing cannot be pipelined by video group: two calls
        scores = nat.forward_rows(visual, audio, starts, lens, axis, model.precision)
        picks, seg_mean, summary, cps_start, sum_start = nat.summarize_rows(
            scores, positions, starts, lens, [v.n_frames for v in videos], [v.cps for v in videos], proportion)
    else:
        scores, picks, seg_mean, summary, cps_start, sum_start = nat.score_and_summarize_rows(
            visual, audio, positions, starts, lens, [v.n_frames for v in videos], [v.cps for v in videos], proportion,
            axis, model.precision)
    picks_h, mean_h, sum_h, scores_h = picks.cpu().numpy(), seg_mean.cpu().numpy(), summary.cpu().numpy(), scores.cpu()
    out = [None] * len(videos)
    for i in range(len(videos)):
        out[caller_order[i]] = VideoSummary(scores_h[starts[i]:starts[i] + lens[i]], picks_h[cps_start[i]:cps_start[i + 1]],
                                            mean_h[cps_start[i]:cps_start[i + 1]], sum_h[sum_start[i]:sum_start[i + 1]])
    return out


@torch.no_grad()
def summarize_stream(model, batches, proportion=0.15, attn_axis: Optional[str] = None, depth: int = 2):
    """Score + pool + select a STREAM of packed host batches, two in flight (the ``scripts/evaluate.py:12-18`` loop
    over a whole dataset): while batch i is being computed the features of batch i+1 already cross PCIe
    (``avs_forward_summarize_async``, one device staging slot per batch in flight), so the steady-state cost of a
    batch is its H2D transfer.

    ``batches`` yields tuples ``(visual, audio, positions, row_start, lengths, shots)`` of pinned host tensors /
    descriptor arrays, ``shots`` being a ``runtime.ShotDesc``.  Yields, in order, what
    ``NativeModel.score_and_summarize_rows`` returns for each batch.
    """
    if depth not in (1, 2):
        raise ValueError("depth must be 1 or 2 (the library has two staging slots)")
    nat = model.native()
    axis = attn_axis or ("literal_b1" if model.attn_axis == "literal" else model.attn_axis)
    if axis == "literal":
        raise ValueError("literal attention mixes the videos of a batch and cannot be pipelined; use literal_b1 / temporal")
    pending = []
    slot = 0
    for visual, audio, positions, row_start, lengths, shots in batches:
        if depth == 1:     # nothing to overlap with: the synchronous call pipelines its own video groups more finely
            yield nat.score_and_summarize_rows(visual, audio, positions, row_start, lengths, None, shots,
                                               proportion, axis, model.precision)
            continue
        if len(pending) == depth:
            yield pending.pop(0).wait()
        pending.append(nat.score_and_summarize_rows(visual, audio, positions, row_start, lengths, None, shots,
                                                    proportion, axis, model.precision, slot=slot))
        slot = (slot + 1) % 2
    while pending:
        yield pending.pop(0).wait()
