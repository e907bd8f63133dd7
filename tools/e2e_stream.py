"""End-to-end (pinned host buffers) config-2 throughput: one synchronous call per batch vs. the streamed form
(two batches in flight, evaluation.summary.summarize_stream) vs. the bare H2D copy of the same features.

    python tools/e2e_stream.py [n_steps]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa: E402,F401
from avsum_b200 import synth  # noqa: E402
from avsum_b200.evaluation.summary import summarize_stream  # noqa: E402
from avsum_b200.models.av_model import AVBiLSTMModel  # noqa: E402
from avsum_b200.runtime import ShotDesc  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
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
    shots = ShotDesc([v.n_frames for v in vids], [v.cps for v in vids])
    batch = (vh, ah, ph, starts, lens, shots)
    R = sum(lens)

    def sync_call():
        return nat.score_and_summarize_rows(vh, ah, ph, starts, lens, None, shots, 0.15, "literal_b1", "tf32")

    ref = sync_call()
    for _ in range(3):
        sync_call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        sync_call()
    torch.cuda.synchronize()
    t_sync = (time.perf_counter() - t0) / n * 1e3
    print("synchronous call per batch   %.3f ms  -> %.2f M frames/s" % (t_sync, R / t_sync / 1e3))

    for depth in (1, 2):
        for res in summarize_stream(model, (batch for _ in range(4)), 0.15, depth=depth):
            pass
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for res in summarize_stream(model, (batch for _ in range(n)), 0.15, depth=depth):
            pass
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / n * 1e3
        same = all(torch.equal(a, b) for a, b in zip(res[:4], ref[:4]))
        print("streamed, %d in flight        %.3f ms  -> %.2f M frames/s   identical=%s" % (depth, t, R / t / 1e3, same))

    vd, ad = torch.empty_like(vh, device="cuda"), torch.empty_like(ah, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        vd.copy_(vh, non_blocking=True)
        ad.copy_(ah, non_blocking=True)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / n * 1e3
    print("bare H2D of the features     %.3f ms  (%.1f GB/s)" % (t, (vh.numel() + ah.numel()) * 4 / t / 1e6))


if __name__ == "__main__":
    main()
