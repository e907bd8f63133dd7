"""BASELINE configs[0]: one TVSum-shaped video (T = 320, 1024-d visual + 128-d audio), single-call latency.
GPU: AVBiLSTMModel.forward on device tensors and on host tensors (H2D + D2H inside); CPU: the reference's module
tree on the host cores (torch CPU ops, restated inline for timing only)."""
import json
import os
import sys
import time

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import avsum_b200  # noqa: E402,F401
from avsum_b200 import synth  # noqa: E402
from avsum_b200.models.av_model import AVBiLSTMModel  # noqa: E402


class RefModel(nn.Module):   # models/av_model.py:7-46
    def __init__(self, vd, ad, hd):
        super().__init__()
        self.visual_fc = nn.Sequential(nn.Linear(vd, hd), nn.ReLU(), nn.Dropout(0.3))
        self.audio_fc = nn.Sequential(nn.Linear(ad, hd), nn.ReLU(), nn.Dropout(0.3))
        self.visual_bilstm = nn.LSTM(hd, hd // 2, bidirectional=True, batch_first=True)
        self.audio_bilstm = nn.LSTM(hd, hd // 2, bidirectional=True, batch_first=True)
        self.attention = nn.MultiheadAttention(embed_dim=hd * 2, num_heads=4)
        self.scorer = nn.Sequential(nn.Linear(hd * 2, 64), nn.ReLU(), nn.Linear(64, 1), nn.Sigmoid())

    def forward(self, visual, audio):
        v, _ = self.visual_bilstm(self.visual_fc(visual))
        a, _ = self.audio_bilstm(self.audio_fc(audio))
        fused = torch.cat([v, a], dim=-1)
        return self.scorer(self.attention(fused, fused, fused)[0]).squeeze()


def timeit(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    vid = synth.config1()
    sd = synth.seeded_state_dict()
    m = AVBiLSTMModel(1024, 128, 512).eval()
    m.load_state_dict(sd)
    m = m.cuda()
    vd, ad = vid.visual[None].cuda(), vid.audio[None].cuda()
    vh, ah = vid.visual[None].pin_memory(), vid.audio[None].pin_memory()
    out = {"config": "config1: B=1, T=320, 1024-d visual + 128-d audio"}
    with torch.no_grad():
        out["gpu_device_tensors_ms"] = timeit(lambda: m(vd, ad), 50)
        out["gpu_host_tensors_ms"] = timeit(lambda: m(vh, ah), 50)
        out["gpu_temporal_attention_ms"] = timeit(lambda: m(vd, ad, attn_axis="temporal"), 50)
        ref = RefModel(1024, 128, 512).eval()
        ref.load_state_dict(sd)
        torch.set_num_threads(os.cpu_count() or 1)
        out["cpu_reference_ms"] = timeit(lambda: ref(vid.visual[None], vid.audio[None]), 5)
        out["cpu_cores"] = os.cpu_count()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
