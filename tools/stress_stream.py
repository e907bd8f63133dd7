"""Stress test of the streamed / synchronous / device-space call mix: random batch compositions, both attention
axes, slots reused hundreds of times; every streamed result must equal the synchronous call bit for bit.

    python tools/stress_stream.py [n_batches] [seed]
"""
import os
import sys

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
    n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    vids = synth.config2()
    pinned = [(v.visual.pin_memory(), v.audio.pin_memory()) for v in vids]
    bad = 0
    axes = sys.argv[3].split(",") if len(sys.argv) > 3 else ("literal_b1", "temporal")
    for axis in axes:
        model = AVBiLSTMModel(1024, 128, 512, attn_axis=axis).eval()
        model.load_state_dict(synth.seeded_state_dict(spread=True))
        model = model.cuda()
        nat = model.native()
        batches = []
        for _ in range(n_batches):
            k = int(rng.integers(1, 51))
            idx = rng.choice(50, size=k, replace=False)
            if rng.random() < 0.7:
                idx = sorted(idx, key=lambda i: -vids[i].T)
            sub = [vids[i] for i in idx]
            lens = [v.T for v in sub]
            starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
            batches.append((torch.cat([pinned[i][0] for i in idx]).pin_memory(), torch.cat([pinned[i][1] for i in idx]).pin_memory(),
                            torch.from_numpy(np.concatenate([v.positions for v in sub]).astype(np.int32)).pin_memory(),
                            starts, lens, ShotDesc([v.n_frames for v in sub], [v.cps for v in sub])))
        want = [nat.score_and_summarize_rows(b[0], b[1], b[2], b[3], b[4], None, b[5], 0.15, axis) for b in batches]
        # streamed, with device-space calls interleaved by the consumer of the stream
        k = 0
        for got in summarize_stream(model, iter(batches), 0.15, axis):
            w = want[k]
            if not all(torch.equal(a, b) for a, b in zip(got[:4], w[:4])):
                bad += 1
                lens_k, starts_k = batches[k][4], batches[k][3]
                diff = (got[0] != w[0]).nonzero().flatten().numpy()
                vid_of = np.searchsorted(np.asarray(starts_k), diff, side="right") - 1
                err = float((got[0] - w[0]).abs().max())
                print(f"MISMATCH axis={axis} batch={k} videos={len(lens_k)} rows={int(sum(lens_k))} bad_rows={diff.size} "
                      f"first={diff[:6].tolist()} videos_hit={sorted(set(vid_of.tolist()))[:12]} lens={[lens_k[i] for i in sorted(set(vid_of.tolist()))[:12]]} "
                      f"max_abs_diff={err:.3e} picks_equal={bool(torch.equal(got[1], w[1]))} sorted={lens_k == sorted(lens_k, reverse=True)}", flush=True)
                again = nat.score_and_summarize_rows(batches[k][0], batches[k][1], batches[k][2], starts_k, lens_k, None, batches[k][5], 0.15, axis)
                print("   sync call repeated: equal to first sync =", bool(torch.equal(again[0], w[0])), flush=True)
            if k % 7 == 3:      # a device-space call between two streamed steps (shares workspaces and streams)
                b = batches[k]
                sc = nat.forward_rows(b[0].cuda(), b[1].cuda(), b[3], b[4], axis, "tf32")
                if not torch.equal(sc.cpu(), w[0]):
                    bad += 1
                    print(f"MISMATCH (device call) axis={axis} batch={k}", flush=True)
            k += 1
        assert k == len(batches)
        print(f"axis {axis}: {k} streamed batches checked", flush=True)
    print("stress", "FAILED" if bad else "ok", f"({bad} mismatches)")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
