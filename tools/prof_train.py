"""A few EAGER training steps (config 5 shape: 8 videos x T = 320) and nothing else: the command ncu wraps for the
launch list of the training step.   python tools/prof_train.py [n_steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa: E402,F401
from avsum_b200 import synth  # noqa: E402
from avsum_b200.models.av_model import AVBiLSTMModel  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    g = torch.Generator().manual_seed(100)
    visual = torch.randn(8, 320, 1024, generator=g).cuda()
    audio = torch.randn(8, 320, 128, generator=g).cuda()
    target = torch.rand(8, 320, generator=g).cuda()
    model = AVBiLSTMModel(1024, 128, 512, attn_axis="literal_b1")
    model.load_state_dict(synth.seeded_state_dict())
    model = model.cuda().train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, capturable=True, fused=True)   # as bench.py --config train
    for _ in range(n):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(model(visual, audio), target)
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    print("ok", float(loss.detach()))


if __name__ == "__main__":
    main()
