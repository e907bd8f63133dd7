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
        self._native_lstm_key = None
        self._train_in_eval = False   # set to differentiate through an eval-mode (dropout-free) forward
        self._side_stream = None      # training forward: the audio branch's stream

    # ------------------------------------------------------------------ native handle
    def _weights_key(self):
        return tuple((p.data_ptr(), p._version, str(p.device)) for p in self.parameters())

    def native(self, for_training: bool = False) -> NativeModel:
        """The packed device copy of the current parameters (re-packed when they change).  ``for_training``: the
        caller is the autograd forward, which reads only the recurrences' tensors through the handle -- those are
        re-packed without a host synchronisation, the rest when an inference call next needs them."""
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
            self._native_key = self._native_lstm_key = key
        elif for_training:
            if key != self._native_lstm_key:
                self._native.update(self.state_dict(), lstm_only=True, sync=False)
                self._native_lstm_key = key
                self._native_key = None      # everything else in the handle is stale now
        elif key != self._native_key:
            self._native.update(self.state_dict())
            self._native_key = self._native_lstm_key = key
        return self._native

    # ------------------------------------------------------------------ forward
    def forward(self, visual, audio, lengths: Optional[Sequence[int]] = None, attn_axis: Optional[str] = None):
        """visual [B, T, Dv], audio [B, T, Da] -> scores, squeezed like av_model.py:46.

        Unbatched [T, Dv] / [T, Da] inputs follow torch's unbatched semantics in the
        reference: the attention then runs over the T frames of the single video.
        """
        axis = attn_axis or self.attn_axis
        if self.training or (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
                             and (visual.requires_grad or self._train_in_eval)):
            return self._forward_train(visual, audio, lengths, axis)
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

    # ------------------------------------------------------------------ training forward (autograd)
    def _forward_train(self, visual, audio, lengths, axis):
        """The forward of scripts/train_av_model.py:88 with a backward: same layer chain as av_model.py:33-46,
        every GEMM and both BiLSTMs as native kernels with native backward kernels (avsum_b200/training.py).

        Attention: the reference's training loop feeds ONE video per step (``unsqueeze(0)``, B = 1), where
        nn.MultiheadAttention over dim 0 degenerates to softmax weight 1 -- context == value projection and the
        q / k thirds of in_proj receive zero gradient.  A batch here is therefore a set of independent B = 1
        samples ("literal_b1"); a literal batch with B > 1 (cross-video mixing) or temporal attention has no
        backward in this build.
        """
        from .. import training as T_
        if visual.dim() == 2:
            raise NotImplementedError("training with unbatched [T, D] inputs (temporal attention) has no backward here")
        B, T, _ = visual.shape
        if axis == "temporal" or (axis == "literal" and B > 1):
            raise NotImplementedError(
                "the training path differentiates the reference's own training semantics (B = 1 per sample: "
                "attn_axis 'literal' with B == 1, or 'literal_b1'); temporal / cross-video attention has no backward")
        if not visual.is_cuda:
            raise RuntimeError("AVBiLSTMModel (avsum_b200) trains only on a CUDA sm_100 device; there is no CPU fallback")
        lens = [T] * B if lengths is None else [int(x) for x in lengths]
        starts = [b * T for b in range(B)]
        nat = self.native(for_training=True)
        xv = visual.reshape(B * T, -1).to(torch.float32)
        xa = audio.reshape(B * T, -1).to(torch.float32)
        relu, drop = torch.relu, torch.nn.functional.dropout
        # the audio branch (audio_fc) is independent of the visual one until the recurrences: it runs on a side stream
        # (autograd runs each node's backward on the stream of its forward, so the two fc backward passes overlap too;
        # fork / join by stream waits -- capturable in the CUDA graph of training.TrainStep)
        cur = torch.cuda.current_stream(visual.device)
        if self._side_stream is None or self._side_stream.device != visual.device:
            self._side_stream = torch.cuda.Stream(device=visual.device)
        side = self._side_stream
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            a = drop(relu(T_.linear(xa, self.audio_fc[0].weight, self.audio_fc[0].bias)), self.audio_fc[2].p, self.training)
        v = drop(relu(T_.linear(xv, self.visual_fc[0].weight, self.visual_fc[0].bias)), self.visual_fc[2].p, self.training)
        cur.wait_stream(side)   # (every later use of the side stream starts by waiting for `cur`: no record_stream needed)
        lw = []
        for mod in (self.visual_bilstm, self.audio_bilstm):
            for suf in ("", "_reverse"):
                lw += [getattr(mod, f"weight_ih_l0{suf}"), getattr(mod, f"weight_hh_l0{suf}"),
                       getattr(mod, f"bias_ih_l0{suf}"), getattr(mod, f"bias_hh_l0{suf}")]
        fused = T_.bilstm_pair(v, a, nat, starts, lens, lw)
        E = self.hidden_dim * 2
        ctx = T_.linear(fused, self.attention.in_proj_weight[2 * E:], self.attention.in_proj_bias[2 * E:])
        y = T_.linear(ctx, self.attention.out_proj.weight, self.attention.out_proj.bias)
        h1 = relu(T_.linear(y, self.scorer[0].weight, self.scorer[0].bias))
        # Linear(64, 1) + Sigmoid: 128 FLOP / frame, left to torch's elementwise autograd
        s = torch.sigmoid(torch.nn.functional.linear(h1, self.scorer[2].weight, self.scorer[2].bias))
        out = s.reshape(B, T, 1)
        if lengths is not None:
            mask = torch.arange(T, device=out.device)[None, :] < torch.as_tensor(lens, device=out.device)[:, None]
            out = out * mask[..., None]
        return out.squeeze()

    @torch.no_grad()
    def score_videos(self, videos: Sequence[Tuple[torch.Tensor, torch.Tensor]], attn_axis: Optional[str] = None):
        """Packed variable-length batch: [(visual [T_i, Dv], audio [T_i, Da])] -> [scores [T_i]].

        All tensors must be on the model's GPU, or all on the host (then the H2D/D2H
        copies happen inside the native call).  Default axis: "literal_b1" when the model
        is "literal" (each video scored as the reference's B=1 call), else the model's.
        """
        axis = attn_axis or ("literal_b1" if self.attn_axis == "literal" else self.attn_axis)
        # packed longest video first (the order data.dataset.packed_batches produces): the recurrence groups then own
        # contiguous row blocks and the native call runs each group's tail behind its own recurrence
        order = sorted(range(len(videos)), key=lambda i: -int(videos[i][0].shape[0]))
        lens = [int(videos[i][0].shape[0]) for i in order]
        starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32) if lens else np.zeros(0, np.int32)
        visual = torch.cat([videos[i][0] for i in order], dim=0)
        audio = torch.cat([videos[i][1] for i in order], dim=0)
        rows = self.native().forward_rows(visual, audio, starts, lens, axis, self.precision)
        out = [None] * len(videos)
        for i, piece in zip(order, torch.split(rows, lens)):
            out[i] = piece
        return out


# names used by BASELINE.json's north_star and by the reference's scripts/train.py:4
AVModel = AVBiLSTMModel
AVSummarizer = AVBiLSTMModel
