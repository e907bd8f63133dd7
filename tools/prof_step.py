"""A few config-2 steps (device-resident) and nothing else: the command ncu wraps.

    python tools/prof_step.py [n_steps] [axis]
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    axis = sys.argv[2] if len(sys.argv) > 2 else "literal_b1"
    prec = sys.argv[3] if len(sys.argv) > 3 else "tf32"
    vids = synth.config2()
    lens = [v.T for v in vids]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    model = AVBiLSTMModel(1024, 128, 512, attn_axis=axis).eval()
    model.load_state_dict(synth.seeded_state_dict())
    model = model.cuda()
    nat = model.native()
    visual = torch.cat([v.visual for v in vids]).cuda()
    audio = torch.cat([v.audio for v in vids]).cuda()
    pos = torch.from_numpy(np.concatenate([v.positions for v in vids]).astype(np.int32)).cuda()
    for _ in range(n):
        scores = nat.forward_rows(visual, audio, starts, lens, axis, prec)
        nat.summarize_rows(scores, pos, starts, lens, [v.n_frames for v in vids], [v.cps for v in vids], 0.15)
    torch.cuda.synchronize()
    print("ok", float(scores.mean()))


if __name__ == "__main__":
    main()
