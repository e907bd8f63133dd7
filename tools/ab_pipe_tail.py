"""A/B of the pipelined tail (AVS_PIPE_TAIL): config 2, device-resident, literal_b1 -- bit-identity of the scores and the
device time of one scored + summarised step (CUDA events, 256 MiB L2 flush between steps).
    python tools/ab_pipe_tail.py [steps]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avsum_b200 import runtime, synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
axis = sys.argv[2] if len(sys.argv) > 2 else "literal_b1"
vids = sorted(synth.config2(), key=lambda v: -v.T)
lens = [v.T for v in vids]
starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
vd = torch.cat([v.visual for v in vids]).cuda()
ad = torch.cat([v.audio for v in vids]).cuda()
native = runtime.NativeModel({k: v.cuda() for k, v in synth.seeded_state_dict(spread=True).items()}, 1024, 128)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(tag):
    out = native.forward_rows(vd, ad, starts, lens, axis, "tf32").clone()
    for _ in range(5):
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
    print(f"{tag}: forward median {ts[len(ts) // 2]:.4f} ms  min {ts[0]:.4f}  max {ts[-1]:.4f}", flush=True)
    print("   sorted:", " ".join(f"{t:.3f}" for t in ts), flush=True)
    return out


base = None
for setting in (sys.argv[3:] or ["0", "1"]):
    os.environ["AVS_PIPE_TAIL"] = setting
    o = run(f"AVS_PIPE_TAIL={setting}")
    if base is None:
        base = o
    else:
        print("  bit-identical to AVS_PIPE_TAIL=0:", bool(torch.equal(o, base)))
