"""BASELINE configs[3]: long-video stress test, T = 8192 frames, temporal attention.
Reports attention-only and full-forward numbers (CUDA events inside the library) and checks the
attention core against the numpy oracle on one (video, head)."""
import json
import os
os.environ.setdefault("AVS_PIPE_TAIL", "0")   # per-stage times / single launches: the one-launch schedule
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa
from avsum_b200 import _cabi, runtime, synth
from avsum_b200.models.av_model import AVBiLSTMModel


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    steps = 5
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    # ---- attention core alone on random q|k|v, checked against the oracle
    g = torch.Generator().manual_seed(0)
    qkv = (torch.randn(T, 3072, generator=g) * 0.5).cuda()
    ctx = runtime.attention(qkv, 1024, 4, [0], [1], [T])
    torch.cuda.synchronize()
    q = qkv[:, :256].cpu().numpy(); k = qkv[:, 1024:1280].cpu().numpy(); v = qkv[:, 2048:2304].cpu().numpy()
    rows = np.arange(0, T, max(1, T // 64))
    s = (q[rows] @ k.T) / 16.0
    s -= s.max(axis=1, keepdims=True)
    p = np.exp(s); p /= p.sum(axis=1, keepdims=True)
    want = p @ v
    err = float(np.max(np.abs(ctx[rows, :256].cpu().numpy() - want)) / np.max(np.abs(want)))
    print(f"attention core T={T}: max rel err vs oracle (head 0, {len(rows)} rows) = {err:.2e}")

    model = AVBiLSTMModel(1024, 128, 512, attn_axis="temporal").eval()
    model.load_state_dict(synth.seeded_state_dict())
    model = model.cuda()
    nat = model.native()
    vids = synth.config4(B, T)
    visual = torch.cat([v.visual for v in vids]).cuda()
    audio = torch.cat([v.audio for v in vids]).cuda()
    lens = [T] * B
    starts = [i * T for i in range(B)]
    for _ in range(2):
        nat.forward_rows(visual, audio, starts, lens, "temporal", "tf32")
    torch.cuda.synchronize()
    _cabi.profile(2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        nat.forward_rows(visual, audio, starts, lens, "temporal", "tf32")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = _cabi.profile_read()
    _cabi.profile(0)
    att_ms = st["attention_core"][0] / steps
    flops_att = 4.0 * T * T * 1024 * B
    out = {"config": f"config4: {B} videos x T={T}, temporal attention, tf32 mode (fp16 attention operands)",
           "forward_ms": ms, "frames_per_s": B * T / (ms * 1e-3),
           "attention_ms": att_ms, "attention_tflops": flops_att / (att_ms * 1e-3) / 1e12,
           "attention_frac_of_bf16_burst_peak": flops_att / (att_ms * 1e-3) / 1e12 / peaks["bf16_tflops"],
           "attention_frac_of_bf16_sustained_peak": flops_att / (att_ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
           "stages_ms": {k: v[0] / steps for k, v in st.items() if v[1]}, "attention_rel_err": err}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
