"""Phase trace of the tensor-core BPTT kernel (debugging aid): where does one backward time step go?

    python tools/bptt_trace.py            (sets AVS_BPTT_TRACE=1; config-5 shape: 8 videos x T = 320)
"""
import ctypes as C
import os
import sys

os.environ["AVS_BPTT_TRACE"] = "1"
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa: E402,F401
from avsum_b200 import _cabi, synth  # noqa: E402
from avsum_b200.models.av_model import AVBiLSTMModel  # noqa: E402

NAMES = ["partials landed -> sum, 4 multiplies, B operand staged, named barrier passed",
         "fence.proxy.async + 32 MMAs + commit issued", "commit issued -> epilogue awake (MMA latency)",
         "tcgen05.ld", "shuffles + st.async issue", None,
         "partials sent -> next partials landed (DSMEM exchange, slowest peer)"]


def main():
    n_videos = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 320
    g = torch.Generator().manual_seed(100)
    visual = torch.randn(n_videos, T, 1024, generator=g).cuda()
    audio = torch.randn(n_videos, T, 128, generator=g).cuda()
    target = torch.rand(n_videos, T, generator=g).cuda()
    model = AVBiLSTMModel(1024, 128, 512, attn_axis="literal_b1")
    model.load_state_dict(synth.seeded_state_dict())
    model = model.cuda().train()
    for _ in range(3):
        model.zero_grad(set_to_none=True)
        torch.nn.functional.mse_loss(model(visual, audio), target).backward()
    torch.cuda.synchronize()
    out = np.zeros(10, dtype=np.uint64)
    _cabi.check(_cabi.lib().avs_debug_bptt_trace(C.c_void_p(out.ctypes.data)))
    steps = max(int(out[8]) - 1, 1)
    tot = 0.0
    for n, v in zip(NAMES, out[:7]):
        if n is None:
            continue
        per = float(v) / steps
        tot += per
        print(f"{per:8.1f} clk/step  {n}")
    print(f"{tot:8.1f} clk/step  total  ({steps + 1} steps)")
    print(f"{float(out[7]) / steps:8.1f} clk/step  (partials sent -> next step's dh-independent math done; hidden if < exchange)")


if __name__ == "__main__":
    main()
