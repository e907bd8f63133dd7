"""Repeatability of the GEMM paths, the whole forward (all three precisions / both axes), the summary kernels and a
training step while another kernel shares the SMs: every repetition must equal the first bit for bit."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa
from avsum_b200 import runtime, synth
from avsum_b200.models.av_model import AVBiLSTMModel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
side = torch.cuda.Stream()
a = torch.randn(4096, 4096, device="cuda")
small = torch.randn(256, 256, device="cuda")


def noise(i):
    with torch.cuda.stream(side):
        if i % 3 == 0:
            for _ in range(3):
                torch.mm(a, a)
        elif i % 3 == 1:
            for _ in range(40):
                torch.mm(small, small)
        else:
            a.mul_(1.0)


def check(name, fn):
    ref = [t.clone() for t in fn()]
    bad = 0
    for i in range(n):
        noise(i)
        out = fn()
        if not all(torch.equal(x, y) for x, y in zip(out, ref)):
            bad += 1
    torch.cuda.synchronize()
    print(f"{name}: {bad} / {n} repetitions differ", flush=True)
    return bad


total = 0
g = torch.Generator().manual_seed(3)
for (M, N, K, prec) in ((38000, 512, 1024, "tf32"), (20001, 2048, 512, "bf16"), (5000, 1024, 1024, "bf16"), (3000, 64, 1024, "tf32")):
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    total += check(f"linear {M}x{N}x{K} {prec}", lambda: (runtime.linear(x, w, b, relu=True, precision=prec),))

vids = sorted(synth.config2(), key=lambda v: -v.T)
lens = [v.T for v in vids]
starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
vd = torch.cat([v.visual for v in vids]).cuda()
ad = torch.cat([v.audio for v in vids]).cuda()
pd = torch.from_numpy(np.concatenate([v.positions for v in vids]).astype(np.int32)).cuda()
shots = runtime.ShotDesc([v.n_frames for v in vids], [v.cps for v in vids])
for axis, prec in (("literal_b1", "tf32"), ("temporal", "bf16"), ("literal_b1", "bf16")):
    model = AVBiLSTMModel(1024, 128, 512, attn_axis=axis).eval()
    model.load_state_dict(synth.seeded_state_dict(spread=True))
    model = model.cuda()
    nat = model.native()

    def step():
        sc = nat.forward_rows(vd, ad, starts, lens, axis, prec)
        r = nat.summarize_rows(sc, pd, starts, lens, None, shots, 0.15)
        return sc, r[0], r[1], r[2]
    total += check(f"forward + summary {axis} {prec}", step)
    sub = 12
    total += check(f"forward {axis} {prec}, {sub} videos", lambda: (nat.forward_rows(vd[:starts[sub]], ad[:starts[sub]], starts[:sub], lens[:sub], axis, prec),))

# training step gradients
B, T = 8, 320
v = torch.randn(B, T, 1024, generator=g).cuda(); au = torch.randn(B, T, 128, generator=g).cuda(); tg = torch.rand(B, T, generator=g).cuda()
m = AVBiLSTMModel(1024, 128, 512, attn_axis="literal_b1")
m.load_state_dict(synth.seeded_state_dict())
m = m.cuda().train()
for seq in (m.visual_fc, m.audio_fc):
    seq[2].p = 0.0


def grads():
    m.zero_grad(set_to_none=True)
    loss = torch.nn.functional.mse_loss(m(v, au), tg)
    loss.backward()
    return [loss.detach()] + [p.grad for p in m.parameters()]


total += check("training forward + backward (28 gradients)", grads)
print("stress", "FAILED" if total else "ok")
sys.exit(1 if total else 0)
