"""Where does the end-to-end (host buffers) step go?  H2D alone, forward alone, summary alone."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa
from avsum_b200 import synth
from avsum_b200.models.av_model import AVBiLSTMModel

vids = sorted(synth.config2(), key=lambda v: -v.T)
lens = [v.T for v in vids]
starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
model = AVBiLSTMModel(1024, 128, 512, attn_axis="literal_b1").eval()
model.load_state_dict(synth.seeded_state_dict())
model = model.cuda()
nat = model.native()
vh = torch.cat([v.visual for v in vids]).pin_memory()
ah = torch.cat([v.audio for v in vids]).pin_memory()
ph = torch.from_numpy(np.concatenate([v.positions for v in vids]).astype(np.int32)).pin_memory()
vd, ad = torch.empty_like(vh, device="cuda"), torch.empty_like(ah, device="cuda")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def h2d():
    vd.copy_(vh, non_blocking=True)
    ad.copy_(ah, non_blocking=True)
    torch.cuda.synchronize()


print("H2D 99 MB alone          %.3f ms" % timeit(h2d))
print("forward device           %.3f ms" % timeit(lambda: nat.forward_rows(vd, ad, starts, lens, "literal_b1", "tf32")))
sc = nat.forward_rows(vh, ah, starts, lens, "literal_b1", "tf32")
print("forward host (pipelined) %.3f ms" % timeit(lambda: nat.forward_rows(vh, ah, starts, lens, "literal_b1", "tf32")))
os.environ["X"] = "1"
print("summarize host           %.3f ms" % timeit(lambda: nat.summarize_rows(sc, ph, starts, lens, [v.n_frames for v in vids], [v.cps for v in vids], 0.15)))
scd, pd = sc.cuda(), ph.cuda()
print("summarize device         %.3f ms" % timeit(lambda: nat.summarize_rows(scd, pd, starts, lens, [v.n_frames for v in vids], [v.cps for v in vids], 0.15)))
# per-group device compute (3 groups by rows, longest first)
R = sum(lens)
cuts = [0]
for gi in (1, 2):
    want = R * gi // 3
    b = cuts[-1] + 1
    while b < len(lens) - (3 - gi) and starts[b] < want:
        b += 1
    cuts.append(b)
cuts.append(len(lens))
for gi in range(3):
    a, b = cuts[gi], cuts[gi + 1]
    r0, r1 = int(starts[a]), int(starts[b - 1] + lens[b - 1])
    rs = (starts[a:b] - r0).astype(np.int32)
    f = lambda: nat.forward_rows(vd[r0:r1], ad[r0:r1], rs, lens[a:b], "literal_b1", "tf32")
    print("group %d: %d videos, %d rows, max len %d: %.3f ms" % (gi, b - a, r1 - r0, max(lens[a:b]), timeit(f)))
