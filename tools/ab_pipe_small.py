"""A/B of the pipelined tail on small / equal-length batches: python tools/ab_pipe_small.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsum_b200 import runtime, synth

native = runtime.NativeModel({k: v.cuda() for k, v in synth.seeded_state_dict().items()}, 1024, 128)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda").manual_seed(1)


def bench(lens, axis, steps=20):
    lens = sorted(lens, reverse=True)
    R = sum(lens)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    vd = torch.randn(R, 1024, generator=g, device="cuda")
    ad = torch.randn(R, 128, generator=g, device="cuda")
    res = {}
    outs = {}
    for setting in ("0", "1"):
        os.environ["AVS_PIPE_TAIL"] = setting
        outs[setting] = native.forward_rows(vd, ad, starts, lens, axis, "tf32").clone()
        for _ in range(3):
            native.forward_rows(vd, ad, starts, lens, axis, "tf32")
        ts = []
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            native.forward_rows(vd, ad, starts, lens, axis, "tf32")
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        res[setting] = ts[len(ts) // 2]
    print(f"{len(lens):3d} videos, T {min(lens)}-{max(lens)}, {axis:10s}: one-launch {res['0']:.4f} ms  pipelined {res['1']:.4f} ms  "
          f"({(res['1'] / res['0'] - 1) * 100:+.1f} %)  identical {bool(torch.equal(outs['0'], outs['1']))}", flush=True)


rng = np.random.default_rng(0)
for axis in ("literal_b1", "temporal"):
    bench([320] * 5, axis)
    bench([320] * 8, axis)
    bench([320] * 12, axis)
    bench([320] * 16, axis)
    bench(list(rng.integers(100, 400, 20)), axis)
    bench(list(rng.integers(200, 700, 32)), axis)
    bench(list(rng.integers(200, 700, 64)), axis)
    bench([2048] * 8, axis)
