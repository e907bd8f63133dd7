"""A configs[3]-shaped forward (n videos x T = 8192, temporal attention) and nothing else: the command ncu wraps for
the attention core's capture.

    python tools/prof_long.py [n_videos] [T] [n_forwards]
"""
import os
os.environ.setdefault("AVS_PIPE_TAIL", "0")   # per-stage times / single launches: the one-launch schedule
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa: E402,F401
from avsum_b200 import synth  # noqa: E402
from avsum_b200.models.av_model import AVBiLSTMModel  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    model = AVBiLSTMModel(1024, 128, 512, attn_axis="temporal").eval()
    model.load_state_dict(synth.seeded_state_dict())
    model = model.cuda()
    nat = model.native()
    g = torch.Generator(device="cuda").manual_seed(4)
    visual = torch.randn(B * T, 1024, generator=g, device="cuda")
    audio = torch.randn(B * T, 128, generator=g, device="cuda")
    for _ in range(n):
        scores = nat.forward_rows(visual, audio, [i * T for i in range(B)], [T] * B, "temporal", "tf32")
    torch.cuda.synchronize()
    print("ok", float(scores.mean()))


if __name__ == "__main__":
    main()
