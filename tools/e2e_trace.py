"""Timeline of the end-to-end (pinned host buffers) config-2 step: where do the milliseconds go?

    AVS_E2E_TRACE=1 python tools/e2e_trace.py [n_steps]

Prints, averaged over the steps: the host clock at the phases of avs_forward_summarize, the device timeline
(features of group g landed / forward of group g finished / knapsack / last D2H), the Python time around the C
call, and a bare H2D copy of the same pinned buffers for comparison.
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("AVS_E2E_TRACE", "1")
import avsum_b200  # noqa: E402,F401
from avsum_b200 import _cabi, synth  # noqa: E402
from avsum_b200.models.av_model import AVBiLSTMModel  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
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
    from avsum_b200.runtime import ShotDesc
    shots = ShotDesc([v.n_frames for v in vids], [v.cps for v in vids])
    noflush = len(sys.argv) > 2 and sys.argv[2] == "noflush"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    vd, ad = torch.empty_like(vh, device="cuda"), torch.empty_like(ah, device="cuda")

    def step():
        return nat.score_and_summarize_rows(vh, ah, ph, starts, lens, None, shots, 0.15, "literal_b1", "tf32")

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    acc = np.zeros(20)
    wall = 0.0
    out = np.zeros(20)
    for _ in range(n):
        if not noflush:
            flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        wall += time.perf_counter() - t0
        _cabi.check(_cabi.lib().avs_debug_e2e_trace(C.c_void_p(out.ctypes.data)))
        acc += out
    acc /= n
    G = int(round(acc[0]))
    print("python call (wall)            %.3f ms" % (wall / n * 1e3))
    print("C call: copies queued %.3f | groups queued %.3f | tail queued %.3f | synchronised %.3f ms after entry"
          % tuple(acc[2:6]))
    print("device: features landed  ", " ".join("%.3f" % x for x in acc[6:6 + G]))
    print("device: forward finished ", " ".join("%.3f" % x for x in acc[12:12 + G]))
    print("device: knapsack %.3f | last D2H %.3f ms after the first device timestamp" % (acc[18], acc[19]))

    def h2d():
        vd.copy_(vh, non_blocking=True)
        ad.copy_(ah, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(3):
        h2d()
    t0 = time.perf_counter()
    for _ in range(n):
        h2d()
    print("bare H2D of the features      %.3f ms" % ((time.perf_counter() - t0) / n * 1e3))


if __name__ == "__main__":
    main()
