"""BASELINE configs[2]: SumMe-shaped long videos (T up to 1,000 frames) with variable-length masking, bf16, 1 x B200.

25 videos, T in [100, 1000], PADDED to Tmax with lengths[B] (the masked-batch call of AVBiLSTMModel.forward),
temporal attention, bf16 operands with fp32 accumulation; also the default (tf32 / fp16) mode for comparison.
Prints one JSON line per precision: device-resident ms per batch, frames/s over the VALID frames.
"""
import json
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
    vids = synth.config3()
    lens = [v.T for v in vids]
    T = max(lens)
    visual, audio = torch.zeros(len(vids), T, 1024), torch.zeros(len(vids), T, 128)
    for b, v in enumerate(vids):
        visual[b, :v.T], audio[b, :v.T] = v.visual, v.audio
    visual, audio = visual.cuda(), audio.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for prec in ("bf16", "tf32"):
        model = AVBiLSTMModel(1024, 128, 512, attn_axis="temporal", precision=prec).eval()
        model.load_state_dict(synth.seeded_state_dict())
        model = model.cuda()
        with torch.no_grad():
            for _ in range(3):
                model(visual, audio, lengths=lens)
            torch.cuda.synchronize()
            steps = 10
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a, b in ev:
                flush.fill_(1)
                a.record()
                model(visual, audio, lengths=lens)
                b.record()
            torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / steps
        print(json.dumps({"config": f"config3: {len(vids)} videos, T in [{min(lens)}, {T}] padded to {T} with lengths, "
                                    f"temporal attention, precision={prec}",
                          "ms_per_batch": ms, "valid_frames": int(sum(lens)), "padded_frames": len(vids) * T,
                          "frames_per_s": sum(lens) / (ms * 1e-3)}), flush=True)


if __name__ == "__main__":
    main()
