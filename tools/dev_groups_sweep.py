"""Device-resident config-2 step under different video-group pipelining settings (AVS_DEV_GROUPS / AVS_DEV_SHARES).

    python tools/dev_groups_sweep.py [n_steps]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa: E402,F401
from avsum_b200 import synth  # noqa: E402
from avsum_b200.models.av_model import AVBiLSTMModel  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    axis = sys.argv[2] if len(sys.argv) > 2 else "literal_b1"
    vids = sorted(synth.config2(), key=lambda v: -v.T)
    lens = [v.T for v in vids]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    model = AVBiLSTMModel(1024, 128, 512, attn_axis=axis).eval()
    model.load_state_dict(synth.seeded_state_dict())
    model = model.cuda()
    nat = model.native()
    vd = torch.cat([v.visual for v in vids]).cuda()
    ad = torch.cat([v.audio for v in vids]).cuda()
    pd = torch.from_numpy(np.concatenate([v.positions for v in vids]).astype(np.int32)).cuda()
    nf = [v.n_frames for v in vids]
    cps = [v.cps for v in vids]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        sc = nat.forward_rows(vd, ad, starts, lens, axis, "tf32")
        return sc, nat.summarize_rows(sc, pd, starts, lens, nf, cps, 0.15)

    os.environ.pop("AVS_DEV_GROUPS", None)
    ref_sc, ref_sum = step()
    torch.cuda.synchronize()
    settings = [("1", None), ("2", None), ("3", None), ("4", None), ("6", None),
                ("2", "24,100"), ("2", "45,100"), ("3", "24,62,100"), ("3", "45,76,100"), ("4", "24,45,76,100"),
                ("4", "24,62,88,100"), ("6", "24,45,62,76,94,100")]
    for g, sh in settings:
        os.environ["AVS_DEV_GROUPS"] = g
        if sh:
            os.environ["AVS_DEV_SHARES"] = sh
        else:
            os.environ.pop("AVS_DEV_SHARES", None)
        for _ in range(3):
            sc, sm = step()
        torch.cuda.synchronize()
        same = bool(torch.equal(sc, ref_sc)) and bool(torch.equal(sm[0], ref_sum[0]))
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in ev:
            flush.fill_(1)
            a.record()
            step()
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        print("groups %s shares %-22s  mean %.4f  median %.4f  min %.4f ms  identical=%s"
              % (g, sh or "default", sum(ms) / n, ms[n // 2], ms[0], same), flush=True)


if __name__ == "__main__":
    main()
