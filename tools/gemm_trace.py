"""Where does a tensor-core GEMM wait?  (AVS_GEMM_TRACE=1; block 0's clock64 totals)

    AVS_GEMM_TRACE=1 [AVS_GEMM_1CTA=1] python tools/gemm_trace.py

Runs single avs_linear calls of the config-2 GEMM shapes (fp32 outputs, like the LSTM input projection) and
prints, per call: kernel time (CUDA events) and the share of block 0's MMA thread spent waiting for operands /
for a drained accumulator, the producer's wait for free stages and the epilogue's wait for accumulators.
"""
import ctypes as C
import os
os.environ.setdefault("AVS_PIPE_TAIL", "0")   # per-stage times / single launches: the one-launch schedule
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("AVS_GEMM_TRACE", "1")
import avsum_b200  # noqa: E402,F401
from avsum_b200 import _cabi, runtime  # noqa: E402


def trace():
    out = np.zeros(8, dtype=np.uint64)
    _cabi.check(_cabi.lib().avs_debug_gemm_trace(C.c_void_p(out.ctypes.data)))
    return out.astype(np.float64)


def main():
    M = 21477
    for name, N, K, prec in (("fc_visual (tf32)", 512, 1024, "tf32"), ("lstm_input (bf16 operands, fp32 out)", 2048, 512, "bf16"),
                             ("out_proj shape (bf16 operands, fp32 out)", 1024, 1024, "bf16")):
        g = torch.Generator().manual_seed(1)
        x = torch.randn(M, K, generator=g).cuda()
        w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
        b = torch.randn(N, generator=g).cuda()
        for _ in range(3):
            runtime.linear(x, w, b, precision=prec)
        torch.cuda.synchronize()
        trace()
        n = 10
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        _cabi.profile(2)
        for a, e in ev:
            runtime.linear(x, w, b, precision=prec)
        st = _cabi.profile_read()
        _cabi.profile(0)
        t = trace() / n
        tiles = max(t[6], 1)
        print(f"{name}: M={M} N={N} K={K}")
        print(f"   block 0: {tiles:.1f} tiles, MMA thread span {t[0]:.0f} clk = {t[0] / tiles:.0f} clk/tile; waiting for operands "
              f"{100 * t[1] / max(t[0], 1):.1f} %, for a drained accumulator {100 * t[2] / max(t[0], 1):.1f} %")
        print(f"   producer waiting for a free stage {t[3]:.0f} clk ({100 * t[3] / max(t[0], 1):.1f} % of the span); epilogue warp span "
              f"{t[5]:.0f} clk, of which waiting for an accumulator {100 * t[4] / max(t[5], 1):.1f} %")


if __name__ == "__main__":
    main()
