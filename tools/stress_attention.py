"""Repeatability of the temporal attention core and of the whole device-space forward under background load."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa
from avsum_b200 import runtime, synth
from avsum_b200.models.av_model import AVBiLSTMModel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
vids = sorted(synth.config2(), key=lambda v: -v.T)
lens = [v.T for v in vids]
starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
R = int(sum(lens))
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(R, 3072, generator=g) * 0.5).cuda()
side = torch.cuda.Stream()
a = torch.randn(4096, 4096, device="cuda")

def noise(k):
    with torch.cuda.stream(side):
        for _ in range(k):
            torch.mm(a, a)

for label, nz in (("quiet", 0), ("noisy", 3)):
    ref = runtime.attention(qkv, 1024, 4, starts, np.ones_like(starts), lens).clone()
    bad = 0
    for i in range(n):
        if nz:
            noise(nz)
        out = runtime.attention(qkv, 1024, 4, starts, np.ones_like(starts), lens)
        if not torch.equal(out, ref):
            bad += 1
            d = (out != ref).any(dim=1).nonzero().flatten().cpu().numpy()
            v = np.searchsorted(starts, d, side="right") - 1
            print(f"attention {label} iter {i}: {d.size} bad rows, first {d[:4].tolist()} video {sorted(set(v.tolist()))[:4]} rel rows {(d - starts[v])[:4].tolist()}", flush=True)
    torch.cuda.synchronize()
    print(f"attention {label}: {bad} / {n} runs differ", flush=True)

model = AVBiLSTMModel(1024, 128, 512, attn_axis="temporal").eval()
model.load_state_dict(synth.seeded_state_dict(spread=True))
model = model.cuda()
nat = model.native()
vd = torch.cat([v.visual for v in vids]).cuda()
ad = torch.cat([v.audio for v in vids]).cuda()
for label, nz in (("quiet", 0), ("noisy", 2)):
    ref = nat.forward_rows(vd, ad, starts, lens, "temporal", "tf32").clone()
    bad = 0
    for i in range(n):
        if nz:
            noise(nz)
        out = nat.forward_rows(vd, ad, starts, lens, "temporal", "tf32")
        if not torch.equal(out, ref):
            bad += 1
            d = (out != ref).nonzero().flatten().cpu().numpy()
            v = np.searchsorted(starts, d, side="right") - 1
            print(f"forward {label} iter {i}: {d.size} bad rows, first {d[:4].tolist()} video {sorted(set(v.tolist()))[:4]} rel rows {(d - starts[v])[:4].tolist()}", flush=True)
    torch.cuda.synchronize()
    print(f"forward {label}: {bad} / {n} runs differ", flush=True)
