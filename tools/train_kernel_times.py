"""Per-kernel GPU times of the eager training step (config 5 shape) from the CUPTI activity records that
torch.profiler collects: kernels keep running concurrently on their streams (unlike an ncu launch list, which
serialises them), so the numbers add up to more than the step where side streams overlap.
    python tools/train_kernel_times.py [n_steps [chrome_trace.json]]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa: E402,F401
from avsum_b200 import synth  # noqa: E402
from avsum_b200.models.av_model import AVBiLSTMModel  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    g = torch.Generator().manual_seed(100)
    visual = torch.randn(8, 320, 1024, generator=g).cuda()
    audio = torch.randn(8, 320, 128, generator=g).cuda()
    target = torch.rand(8, 320, generator=g).cuda()
    model = AVBiLSTMModel(1024, 128, 512, attn_axis="literal_b1")
    model.load_state_dict(synth.seeded_state_dict())
    model = model.cuda().train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, capturable=True, fused=True)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(model(visual, audio), target)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(n):
            step()
        torch.cuda.synchronize()
    if len(sys.argv) > 2:   # chrome trace of the profiled steps (timeline of the streams)
        prof.export_chrome_trace(sys.argv[2])
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
        if t > 0 and e.device_type.name == "CUDA":
            rows.append((t / n, e.count / n, e.key[:110]))
    rows.sort(reverse=True)
    print(f"per-step GPU kernel time, {n} eager steps (us per step, launches per step, kernel)")
    for t, c, k in rows[:30]:
        print(f"{t:9.1f}  x{c:5.1f}  {k}")
    print(f"{sum(r[0] for r in rows):9.1f}  sum over all kernels")


if __name__ == "__main__":
    main()
