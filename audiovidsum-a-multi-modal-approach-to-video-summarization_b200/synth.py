"""Seeded synthetic TVSum/SumMe-shaped workloads.

Shared by the tests, ``bench.py`` and ``tests/golden/make_golden.py`` so that the
oracle, the golden fixtures and the CUDA path all see identical bytes.  Uses
only torch/numpy CPU generators, which are platform independent.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

SAMPLE_STRIDE = 15  # 30 fps video sampled at 2 fps (reference configs/data_config.yaml:12)


@dataclass
class Video:
    visual: torch.Tensor       # [T, Dv] fp32 (host)
    audio: torch.Tensor        # [T, Da] fp32 (host)
    n_frames: int              # original frame count
    positions: np.ndarray      # int32[T] original index of every sampled frame
    cps: np.ndarray            # int32[S, 2] inclusive change-point segments covering [0, n_frames)

    @property
    def T(self) -> int:
        return int(self.visual.shape[0])


def make_features(T: int, visual_dim: int, audio_dim: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    visual = torch.randn(T, visual_dim, generator=g, dtype=torch.float32)
    audio = torch.randn(T, audio_dim, generator=g, dtype=torch.float32)
    return visual, audio


def make_change_points(n_frames: int, seed: int, lo: int = 30, hi: int = 301) -> np.ndarray:
    """Shot lengths drawn from integers[lo, hi) and cumulated to n_frames -> inclusive [start, end]."""
    rng = np.random.default_rng(seed)
    starts, ends = [], []
    s = 0
    while s < n_frames:
        ln = int(rng.integers(lo, hi))
        e = min(s + ln, n_frames)
        starts.append(s)
        ends.append(e - 1)
        s = e
    return np.stack([np.asarray(starts), np.asarray(ends)], axis=1).astype(np.int32)


def make_video(T: int, visual_dim: int, audio_dim: int, seed: int) -> Video:
    visual, audio = make_features(T, visual_dim, audio_dim, seed)
    n_frames = SAMPLE_STRIDE * T
    positions = (np.arange(T, dtype=np.int64) * SAMPLE_STRIDE).astype(np.int32)
    cps = make_change_points(n_frames, seed=100000 + seed)
    return Video(visual, audio, n_frames, positions, cps)


def config1(visual_dim=1024, audio_dim=128) -> Video:
    """B=1, T=320 (BASELINE.json configs[0])."""
    return make_video(320, visual_dim, audio_dim, seed=1234)


def video_batch(n_videos: int, t_lo: int, t_hi: int, visual_dim=1024, audio_dim=128,
                length_seed: int = 0, seed0: int = 1234) -> List[Video]:
    lengths = np.random.default_rng(length_seed).integers(t_lo, t_hi + 1, n_videos)
    return [make_video(int(t), visual_dim, audio_dim, seed0 + i) for i, t in enumerate(lengths)]


def config2(visual_dim=1024, audio_dim=128) -> List[Video]:
    """50 TVSum-length videos, T in [200, 700] (BASELINE.json configs[1]); sum T = 21,477."""
    return video_batch(50, 200, 700, visual_dim, audio_dim)


def config3(visual_dim=1024, audio_dim=128) -> List[Video]:
    """25 SumMe-shaped videos, T in [100, 1000] (BASELINE.json configs[2])."""
    return video_batch(25, 100, 1000, visual_dim, audio_dim, length_seed=3, seed0=5000)


def config4(n_videos: int = 8, T: int = 8192, visual_dim=1024, audio_dim=128) -> List[Video]:
    """Long-video stress test (BASELINE.json configs[3])."""
    return [make_video(T, visual_dim, audio_dim, 9000 + i) for i in range(n_videos)]


PEAK_FACTOR = 30.0   # "peaked" weight set: q and k projections x30 -> attention logits x900 (sigma ~ 4.7 nats on config 1)


def seeded_state_dict(visual_dim=1024, audio_dim=128, hidden_dim=512, seed=0, spread=False, peaked=False):
    """Reference-format state_dict with torch's default init under torch.manual_seed(seed).

    Builds the same torch.nn containers in the same order as the reference
    constructor (/root/reference/models/av_model.py:10-31), so the RNG stream --
    and therefore every weight -- equals ``torch.manual_seed(seed); AVBiLSTMModel(...)``.
    ``spread`` multiplies scorer.2.weight by 50 so scores leave the 0.52 +- 0.005 band.
    ``peaked`` multiplies the q and k rows of attention.in_proj_weight by 30: with torch's default init the
    attention logits have sigma ~ 0.005 (softmax == uniform weights, the context is mean(V) for every frame);
    x900 gives sigma ~ 4.7 nats, per-row spans of ~30 log2 units and weights up to 0.94, so the temporal
    branch depends on Q K^T, the running maximum and the rescaling of the flash-style kernel.
    """
    import torch.nn as nn
    torch.manual_seed(seed)
    mods = nn.ModuleDict()
    mods["visual_fc"] = nn.Sequential(nn.Linear(visual_dim, hidden_dim), nn.ReLU(), nn.Dropout(0.3))
    mods["audio_fc"] = nn.Sequential(nn.Linear(audio_dim, hidden_dim), nn.ReLU(), nn.Dropout(0.3))
    mods["visual_bilstm"] = nn.LSTM(hidden_dim, hidden_dim // 2, bidirectional=True, batch_first=True)
    mods["audio_bilstm"] = nn.LSTM(hidden_dim, hidden_dim // 2, bidirectional=True, batch_first=True)
    mods["attention"] = nn.MultiheadAttention(embed_dim=hidden_dim * 2, num_heads=4)
    mods["scorer"] = nn.Sequential(nn.Linear(hidden_dim * 2, 64), nn.ReLU(), nn.Linear(64, 1), nn.Sigmoid())
    sd = {k: v.detach().clone() for k, v in mods.state_dict().items()}
    if spread:
        sd["scorer.2.weight"] = sd["scorer.2.weight"] * 50.0
    if peaked:
        w = sd["attention.in_proj_weight"].clone()
        w[:4 * hidden_dim] *= PEAK_FACTOR
        sd["attention.in_proj_weight"] = w
    return sd


def state_dict_checksum(sd) -> float:
    """Order-independent float64 checksum used to prove two weight sets are identical."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].detach().double()
        tot += float(v.abs().sum()) + 3.0 * float(v.sum())
    return tot
