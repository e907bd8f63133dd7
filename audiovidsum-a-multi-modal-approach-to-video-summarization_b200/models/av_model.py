"""Drop-in for the reference's ``models/av_model.py`` (``AVBiLSTMModel``), B200-native.

Constructor, attribute names, ``state_dict`` keys (28 tensors) and ``forward(visual, audio)``
match /root/reference/models/av_model.py:6-46.  The torch sub-modules created here are
*parameter containers only* (so ``.cuda()``, ``.state_dict()``, ``.load_state_dict()``,
``.parameters()`` behave exactly like the reference); the forward pass never calls them -- it
hands the packed weights and the feature tensors to ``libavsum_b200.so`` (C ABI
``avs_forward``), whose kernels are hand-written sm_100a CUDA.  There is no CPU or
PyTorch-operator fallback: a model whose parameters are not on a CUDA device raises.

Extensions the reference lacks (all opt-in, defaults reproduce the reference):
  * ``attn_axis``: "literal" (default; what av_model.py:44 really computes -- attention over
    dim 0, the video axis), "temporal" (frame self-attention inside each video, the intent of
    models/attention.py), "literal_b1" (every video treated as its own B=1 call, i.e. the
    scripts/evaluate.py:12-18 loop as one batch);
  * ``forward(..., lengths=...)``: variable-length masking for padded batches;
  * ``score_videos`` for packed variable-length batches.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from ..runtime import NativeModel
from .attention import MultiHeadSelfAttention  # noqa: F401  (the reference imports it too, av_model.py:3)


class AVBiLSTMModel(nn.Module):
    def __init__(self, visual_dim=4096, audio_dim=296, hidden_dim=512, attn_axis: str = "literal",
                 precision: str = "tf32"):
        super().__init__()
        # parameter containers, created in the reference's order so that a seeded
        # construction yields bit-identical weights (av_model.py:10-31)
        self.visual_fc = nn.Sequential(nn.Linear(visual_dim, hidden_dim), nn.ReLU(), nn.Dropout(0.3))
        self.audio_fc = nn.Sequential(nn.Linear(audio_dim, hidden_dim), nn.ReLU(), nn.Dropout(0.3))
        self.visual_bilstm = nn.LSTM(hidden_dim, hidden_dim // 2, bidirectional=True, batch_first=True)
        self.audio_bilstm = nn.LSTM(hidden_dim, hidden_dim // 2, bidirectional=True, batch_first=True)
        self.attention = nn.MultiheadAttention(embed_dim=hidden_dim * 2, num_heads=4)
        self.scorer = nn.Sequential(nn.Linear(hidden_dim * 2, 64), nn.ReLU(), nn.Linear(64, 1), nn.Sigmoid())
        self.visual_dim, self.audio_dim, self.hidden_dim = visual_dim, audio_dim, hidden_dim
        self.attn_axis = attn_axis
        self.precision = precision
        self._native: Optional[NativeModel] = None
        self._native_key = None

    # ------------------------------------------------------------------ native handle
    def _weights_key(self):
        return tuple((p.data_ptr(), p._version, str(p.device)) for p in self.parameters())

    def native(self) -> NativeModel:
        """The packed device copy of the current parameters (re-packed when they change)."""
        p0 = next(self.parameters())
        if not p0.is_cuda:
            raise RuntimeError(
                "AVBiLSTMModel (avsum_b200) runs only on a CUDA sm_100 device: call .cuda() first; "
                "there is no CPU fallback")
        key = self._weights_key()
        if self._native is None or self._native.device != p0.device.index:
            if self._native is not None:
                self._native.close()
            self._native = NativeModel(self.state_dict(), self.visual_dim, self.audio_dim, self.hidden_dim,
                                       self.attention.num_heads, device=p0.device.index)
            self._native_key = key
        elif key != self._native_key:
            self._native.update(self.state_dict())
            self._native_key = key
        return self._native

    # ------------------------------------------------------------------ forward
    def forward(self, visual, audio, lengths: Optional[Sequence[int]] = None, attn_axis: Optional[str] = None):
        """visual [B, T, Dv], audio [B, T, Da] -> scores, squeezed like av_model.py:46.

        Unbatched [T, Dv] / [T, Da] inputs follow torch's unbatched semantics in the
        reference: the attention then runs over the T frames of the single video.
        """
        if self.training:
            raise NotImplementedError(
                "avsum_b200 implements the inference hot path (model.eval()); the training step of "
                "scripts/train_av_model.py:86-96 (dropout + backward) is outside this build's scope")
        axis = attn_axis or self.attn_axis
        nat = self.native()
        if visual.dim() == 2 and audio.dim() == 2:
            T = visual.shape[0]
            return nat.forward_rows(visual, audio, [0], [T], "temporal", self.precision).squeeze()
        if visual.dim() != 3 or audio.dim() != 3:
            raise ValueError("expected visual [B, T, Dv] and audio [B, T, Da]")
        B, T, _ = visual.shape
        if audio.shape[0] != B or audio.shape[1] != T:
            raise ValueError(f"visual {tuple(visual.shape)} and audio {tuple(audio.shape)} disagree on [B, T]")
        if lengths is None:
            lens = [T] * B
        else:
            lens = [int(x) for x in lengths]
            if len(lens) != B or any(n < 0 or n > T for n in lens):
                raise ValueError("lengths must hold B values in [0, T]")
            if axis == "literal" and B > 1 and any(n != T for n in lens):
                raise ValueError("attn_axis='literal' mixes the videos of a batch and cannot be masked; "
                                 "use 'temporal' or 'literal_b1'")
        rows = nat.forward_rows(visual.reshape(B * T, -1), audio.reshape(B * T, -1),
                                [b * T for b in range(B)], lens, axis, self.precision)
        out = rows.reshape(B, T, 1)
        if lengths is not None:
            mask = torch.arange(T, device=out.device)[None, :] < torch.as_tensor(lens, device=out.device)[:, None]
            out = torch.where(mask[..., None], out, torch.zeros((), dtype=out.dtype, device=out.device))
        return out.squeeze()

    @torch.no_grad()
    def score_videos(self, videos: Sequence[Tuple[torch.Tensor, torch.Tensor]], attn_axis: Optional[str] = None):
        """Packed variable-length batch: [(visual [T_i, Dv], audio [T_i, Da])] -> [scores [T_i]].

        All tensors must be on the model's GPU, or all on the host (then the H2D/D2H
        copies happen inside the native call).  Default axis: "literal_b1" when the model
        is "literal" (each video scored as the reference's B=1 call), else the model's.
        """
        axis = attn_axis or ("literal_b1" if self.attn_axis == "literal" else self.attn_axis)
        lens = [int(v.shape[0]) for v, _ in videos]
        starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32) if lens else np.zeros(0, np.int32)
        visual = torch.cat([v for v, _ in videos], dim=0)
        audio = torch.cat([a for _, a in videos], dim=0)
        rows = self.native().forward_rows(visual, audio, starts, lens, axis, self.precision)
        return list(torch.split(rows, lens))


# names used by BASELINE.json's north_star and by the reference's scripts/train.py:4
AVModel = AVBiLSTMModel
AVSummarizer = AVBiLSTMModel
