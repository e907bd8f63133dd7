"""CPU timing baseline: the reference's forward restated on the SAME torch CPU operators.  TEST/BENCH INFRASTRUCTURE ONLY.

The reference is pure Python over ``torch.nn`` (``/root/reference/models/av_model.py:7-46``),
so its CPU cost *is* the cost of ATen's ``nn.Linear`` / ``nn.LSTM`` /
``nn.MultiheadAttention`` kernels.  ``/root/reference`` does not exist on the
GPU box, therefore the bench's ``cpu_baseline`` / ``--impl reference`` legs time
this restatement ("kind": "port"): same modules, same call order, same B=1
per-video loop as ``scripts/evaluate.py:12-18`` (minus ``.cuda()``), followed by
the numpy summary oracle.  ``tests/golden/make_golden.py`` checks, in the container
that has the reference, that this module reproduces the imported reference
bit for bit (``torch.equal``) from the same state_dict.

Nothing under the product package imports this file.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class RefPortModel(nn.Module):
    """Module tree with the reference's parameter names (av_model.py:10-31)."""

    def __init__(self, visual_dim=4096, audio_dim=296, hidden_dim=512):
        super().__init__()
        self.visual_fc = nn.Sequential(nn.Linear(visual_dim, hidden_dim), nn.ReLU(), nn.Dropout(0.3))
        self.audio_fc = nn.Sequential(nn.Linear(audio_dim, hidden_dim), nn.ReLU(), nn.Dropout(0.3))
        self.visual_bilstm = nn.LSTM(hidden_dim, hidden_dim // 2, bidirectional=True, batch_first=True)
        self.audio_bilstm = nn.LSTM(hidden_dim, hidden_dim // 2, bidirectional=True, batch_first=True)
        self.attention = nn.MultiheadAttention(embed_dim=hidden_dim * 2, num_heads=4)
        self.scorer = nn.Sequential(nn.Linear(hidden_dim * 2, 64), nn.ReLU(), nn.Linear(64, 1), nn.Sigmoid())

    def forward(self, visual, audio, attn_axis="literal"):
        v_out, _ = self.visual_bilstm(self.visual_fc(visual))
        a_out, _ = self.audio_bilstm(self.audio_fc(audio))
        fused = torch.cat([v_out, a_out], dim=-1)
        if attn_axis == "temporal":  # frame self-attention: the same module on the transposed tensor
            f = fused.transpose(0, 1)
            attn_out = self.attention(f, f, f)[0].transpose(0, 1)
        else:  # literal av_model.py:44
            attn_out, _ = self.attention(fused, fused, fused)
        return self.scorer(attn_out).squeeze()


@torch.no_grad()
def run_videos(model, videos, attn_axis="literal"):
    """B=1 loop of scripts/evaluate.py:12-18 on host tensors; returns list of score vectors."""
    model.eval()
    out = []
    for visual, audio in videos:
        out.append(model(visual.unsqueeze(0), audio.unsqueeze(0), attn_axis).reshape(-1))
    return out
