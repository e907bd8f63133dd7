"""Drop-in for the reference's ``models/attention.py`` (``MultiHeadSelfAttention``).

Same constructor, sub-module names (``query``, ``key``, ``value``, ``out`` -- all
``nn.Linear(E, E)``), attributes (``num_heads``, ``dim_head``) and ``forward(x[B, T, E])``
as /root/reference/models/attention.py:5-25; the arithmetic (three projections, per-head
softmax(QK^T / sqrt(dh)) V, output projection) runs in libavsum_b200.so.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import runtime


class MultiHeadSelfAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, precision: str = "tf32"):
        super().__init__()
        self.query = nn.Linear(embed_dim, embed_dim)
        self.key = nn.Linear(embed_dim, embed_dim)
        self.value = nn.Linear(embed_dim, embed_dim)
        self.out = nn.Linear(embed_dim, embed_dim)
        self.num_heads = num_heads
        self.dim_head = embed_dim // num_heads
        self.precision = precision

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("MultiHeadSelfAttention (avsum_b200) needs CUDA tensors; there is no CPU fallback")
        batch_size, seq_len, embed = x.size()
        # one GEMM for the three projections (attention.py:17-19): stack the weights as [3E, E]
        w = torch.cat([self.query.weight, self.key.weight, self.value.weight], dim=0)
        b = torch.cat([self.query.bias, self.key.bias, self.value.bias], dim=0)
        qkv = runtime.linear(x.reshape(batch_size * seq_len, embed), w, b, precision=self.precision)
        base = np.arange(batch_size, dtype=np.int32) * seq_len
        ctx = runtime.attention(qkv, embed, self.num_heads, base, np.ones(batch_size, np.int32),
                                np.full(batch_size, seq_len, np.int32), precision=self.precision)
        y = runtime.linear(ctx, self.out.weight, self.out.bias, precision=self.precision)
        return y.reshape(batch_size, seq_len, embed)
