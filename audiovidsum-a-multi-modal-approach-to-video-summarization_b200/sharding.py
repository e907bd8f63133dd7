"""Batch sharding across the GPUs of one box.

Videos are independent through forward, pooling and knapsack, so a batch is split by video
with NO data-path collective: longest-processing-time greedy on
``cost_i = 16.0e6 * T_i + 4096 * T_i**2`` (FLOPs per video).  The only exchange is
the gather of the (tiny) per-video results to rank 0, done with ``torch.distributed``
(NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def video_cost(T: int) -> float:
    return 16.0e6 * T + 4096.0 * T * T


def shard_videos(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Indices of the videos each rank processes (deterministic LPT greedy, ties by index)."""
    order = sorted(range(len(lengths)), key=lambda i: (-video_cost(int(lengths[i])), i))
    loads = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += video_cost(int(lengths[i]))
    return [sorted(s) for s in shards]


def gather_picks(local_ids: Sequence[int], local_picks: Sequence[np.ndarray], n_videos: int, device="cpu"):
    """All ranks contribute {video id: picks}; every rank returns the full list ordered by id.

    One padded uint8 all_gather (KBs): header [n_local, then (id, S) pairs] + payload.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = [None] * n_videos
        for i, p in zip(local_ids, local_picks):
            out[i] = np.asarray(p, dtype=np.uint8)
        return out
    world = dist.get_world_size()
    meta = np.asarray([len(local_ids)] + [x for i, p in zip(local_ids, local_picks) for x in (i, len(p))], dtype=np.int64)
    payload = np.concatenate([np.asarray(p, dtype=np.uint8) for p in local_picks]) if len(local_picks) else np.zeros(0, np.uint8)
    sizes = torch.tensor([meta.size, payload.size], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    max_meta = int(max(int(s[0]) for s in all_sizes))
    max_pay = int(max(int(s[1]) for s in all_sizes))
    mbuf = torch.zeros(max_meta, dtype=torch.int64, device=device)
    mbuf[:meta.size] = torch.from_numpy(meta).to(device)
    pbuf = torch.zeros(max(max_pay, 1), dtype=torch.uint8, device=device)
    pbuf[:payload.size] = torch.from_numpy(payload).to(device)
    metas = [torch.zeros_like(mbuf) for _ in range(world)]
    pays = [torch.zeros_like(pbuf) for _ in range(world)]
    dist.all_gather(metas, mbuf)
    dist.all_gather(pays, pbuf)
    out = [None] * n_videos
    for r in range(world):
        m = metas[r].cpu().numpy()
        p = pays[r].cpu().numpy()
        off = 0
        for k in range(int(m[0])):
            vid, S = int(m[1 + 2 * k]), int(m[2 + 2 * k])
            out[vid] = p[off:off + S].copy()
            off += S
    return out
