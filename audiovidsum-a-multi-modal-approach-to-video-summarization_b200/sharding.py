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


def bind_process_to_gpu_numa(device_index: int) -> str:
    """One process per GPU: pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates
    pinned host buffers, so that the batches it feeds across PCIe are first-touched in memory local to that GPU's
    root complex (eight ranks streaming 50 GB/s each otherwise contend for one socket's memory and the inter-socket
    link).  Best effort: returns a one-line description; a no-op when sysfs exposes no NUMA topology (single node,
    containers without /sys/bus/pci), when the affinity call is not permitted, or when AVS_NO_NUMA_BIND is set.
    """
    import os
    if os.environ.get("AVS_NO_NUMA_BIND"):
        return "numa: binding disabled (AVS_NO_NUMA_BIND)"
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        with open(base + "/numa_node") as f:
            node = int(f.read().strip())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2:
            return f"numa: {bdf} reports node {node}, {len(nodes)} node(s) visible -- nothing to bind"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"numa: node {node} has no CPU this process may use"
        os.sched_setaffinity(0, cpus)
        return f"numa: {bdf} -> node {node}, bound to {len(cpus)} CPUs ({cpulist})"
    except Exception as e:   # sysfs layout, permissions, old torch: stay unbound
        return f"numa: not bound ({type(e).__name__}: {e})"
