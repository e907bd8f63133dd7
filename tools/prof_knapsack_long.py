"""Pooling + knapsack of n long videos (T = 8192: 122,880 frames, capacity 18,432, ~750 shots), timed with CUDA events;
the command ncu wraps for the long-video knapsack kernel.   python tools/prof_knapsack_long.py [n_videos] [T] [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa: E402,F401
from avsum_b200 import runtime, synth  # noqa: E402
from avsum_b200.runtime import ShotDesc  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    sd = synth.seeded_state_dict()
    nat = runtime.NativeModel({k: v.cuda() for k, v in sd.items()}, 1024, 128)
    vids = [synth.make_video(T, 4, 4, 9000 + i) for i in range(n)]
    lens = [v.T for v in vids]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    scores = torch.rand(sum(lens), device="cuda")
    pos = torch.from_numpy(np.concatenate([v.positions for v in vids]).astype(np.int32)).cuda()
    shots = ShotDesc([v.n_frames for v in vids], [v.cps for v in vids])
    print("shots per video", [len(v.cps) for v in vids][:4], "capacity", vids[0].n_frames * 15 // 100)
    for _ in range(2):
        nat.summarize_rows(scores, pos, starts, lens, None, shots, 0.15)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        nat.summarize_rows(scores, pos, starts, lens, None, shots, 0.15)
    e1.record()
    torch.cuda.synchronize()
    print(f"summarize_rows: {e0.elapsed_time(e1) / reps:.3f} ms per call ({n} videos x T={T})")


if __name__ == "__main__":
    main()
