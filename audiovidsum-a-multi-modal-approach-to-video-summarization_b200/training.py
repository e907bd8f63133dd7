"""Training-step plumbing: torch.autograd Functions around the native forward / backward kernels.

The reference trains through torch autograd (scripts/train_av_model.py:86-96: ``model.train()``,
``preds = model(visual, audio)``, ``F.mse_loss``, ``loss.backward()``, ``torch.optim.AdamW``).  To stay a drop-in
for that loop the B200 path keeps autograd as the *plumbing* -- graph recording, gradient routing to the 28
parameters, the user's optimiser -- and supplies every heavy op of the step as a native kernel pair:

* ``linear``       forward ``avs_linear`` (tcgen05 kind::tf32 GEMM + bias), backward ``avs_linear_bwd``
                   (dX = dY W, dW = dY^T X as tcgen05 GEMMs on transposed, round-to-nearest operands; db column sums)
* ``bilstm_pair``  forward ``avs_bilstm_pair_train`` (tcgen05 recurrence that also saves the gate pre-activations
                   and cell states), backward ``avs_bilstm_pair_bwd`` (cluster/DSMEM BPTT kernel + tcgen05 GEMMs for
                   dW_ih, dW_hh, d_emb)

ReLU / dropout / sigmoid / MSE are left to torch's elementwise autograd ops (O(rows * 512) work, < 0.1 % of the
step).  ``allreduce_gradients`` is the data-parallel collective of BASELINE configs[4]: one flat fp32 bucket
(8,008,833 elements for the 1024/128 model), one NCCL all-reduce, averaged.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _cabi
from .runtime import _i32, _stream_ptr, linear as _linear_fwd


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return _linear_fwd(x, weight, bias, relu=False, precision="tf32")

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous().to(torch.float32)
        M, K = x.shape
        N = w.shape[0]
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        dx = torch.empty_like(x) if need_x else None
        dw = torch.empty_like(w) if need_w else None
        db = torch.empty(N, dtype=torch.float32, device=x.device) if need_b else None
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib().avs_linear_bwd(_ptr(dy), _ptr(x), _ptr(w.contiguous()), int(M), int(N), int(K),
                                                   _ptr(dx), _ptr(dw), _ptr(db), _stream_ptr(x.device)))
        return dx, dw, db


def linear(x: torch.Tensor, weight: torch.Tensor, bias) -> torch.Tensor:
    """Differentiable nn.Linear on CUDA rows [M, K] through the native GEMMs."""
    if not x.is_cuda:
        raise RuntimeError("avsum_b200 training needs CUDA tensors (no CPU fallback)")
    return _LinearFn.apply(x.contiguous(), weight.contiguous(), bias)


class _BiLSTMPairFn(torch.autograd.Function):
    """fused[R, 1024] = [visual_bilstm(v_emb) | audio_bilstm(a_emb)] on packed rows (av_model.py:39-43)."""

    @staticmethod
    def forward(ctx, v_emb, a_emb, native, row_start, lengths, *weights):
        R = int(v_emb.shape[0])
        rs, ln = _i32(row_start), _i32(lengths)
        dev = v_emb.device
        fused = torch.zeros(R, 1024, dtype=torch.float32, device=dev)
        save_pre = torch.empty(R, 4, 256, 4, dtype=torch.float32, device=dev)
        save_c = torch.empty(R, 4, 256, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(native.lib.avs_bilstm_pair_train(
                native._handle, _ptr(v_emb), _ptr(a_emb), R, int(rs.size), _cabi.np_ptr(rs), _cabi.np_ptr(ln),
                _ptr(fused), _ptr(save_pre), _ptr(save_c), _stream_ptr(dev)))
        ctx.save_for_backward(v_emb, a_emb, fused, save_pre, save_c)
        ctx.native, ctx.rs, ctx.ln = native, rs, ln
        return fused

    @staticmethod
    def backward(ctx, d_fused):
        v_emb, a_emb, fused, save_pre, save_c = ctx.saved_tensors
        native, rs, ln = ctx.native, ctx.rs, ctx.ln
        dev = v_emb.device
        R = int(v_emb.shape[0])
        d_fused = d_fused.contiguous().to(torch.float32)
        d_v, d_a = torch.empty_like(v_emb), torch.empty_like(a_emb)
        dW_ih = [torch.empty(1024, 512, dtype=torch.float32, device=dev) for _ in range(4)]
        dW_hh = [torch.empty(1024, 256, dtype=torch.float32, device=dev) for _ in range(4)]
        db = [torch.empty(1024, dtype=torch.float32, device=dev) for _ in range(4)]
        arr = lambda ts: (C.c_void_p * 4)(*[t.data_ptr() for t in ts])
        with torch.cuda.device(dev):
            _cabi.check(native.lib.avs_bilstm_pair_bwd(
                native._handle, _ptr(d_fused), _ptr(save_pre), _ptr(save_c), _ptr(fused), _ptr(v_emb), _ptr(a_emb), R,
                int(rs.size), _cabi.np_ptr(rs), _cabi.np_ptr(ln), _ptr(d_v), _ptr(d_a), arr(dW_ih), arr(dW_hh), arr(db),
                _stream_ptr(dev)))
        grads = []
        for i in range(4):      # weight_ih, weight_hh, bias_ih, bias_hh per recurrence
            grads += [dW_ih[i], dW_hh[i], db[i], db[i].clone()]
        return (d_v, d_a, None, None, None, *grads)


def bilstm_pair(v_emb, a_emb, native, row_start, lengths, lstm_weights: Sequence[torch.Tensor]) -> torch.Tensor:
    """lstm_weights: 16 tensors -- (weight_ih, weight_hh, bias_ih, bias_hh) for visual fwd, visual reverse, audio
    fwd, audio reverse; they are passed so autograd can route their gradients (the kernels read the packed copy
    inside ``native``, which AVBiLSTMModel.native() keeps in sync with the parameters)."""
    return _BiLSTMPairFn.apply(v_emb.contiguous(), a_emb.contiguous(), native, row_start, lengths, *lstm_weights)


def allreduce_gradients(parameters, world_size: int = None, group=None) -> int:
    """Average the gradients of ``parameters`` over the data-parallel ranks with ONE flat all-reduce
    (NCCL over NVLink on GPUs; any torch.distributed backend works).  Returns the bucket's element count."""
    import torch.distributed as dist
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return 0
    world = dist.get_world_size(group) if world_size is None else world_size
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return int(flat.numel())


class GraphedTrainStep:
    """The single-GPU training step of scripts/train_av_model.py:86-96 (forward, loss, ``backward``,
    ``optimizer.step``) captured ONCE in a CUDA graph and replayed: the eager step is host-bound (82 native launches
    plus torch's elementwise and optimiser kernels, ~3.3 ms for 8 x 320 frames), the replay is one graph launch.

    The batch shape is fixed at capture (``visual [B, T, Dv]``, ``audio [B, T, Da]``, ``target`` as the loss function
    takes it, optional ``lengths``); every call copies the new batch into the captured input buffers.  The
    optimiser must be built with ``capturable=True`` (torch's requirement for optimiser steps inside a graph).
    Launch plans and sequence descriptors travel as kernel parameters (no pageable-memory copies), the recurrences'
    weight re-pack after the optimiser step is part of the graph, and workspaces are grown during the warm-up
    steps, so nothing inside the captured region allocates through cudaMalloc or synchronises.
    """

    def __init__(self, model, optimizer, loss_fn, visual, audio, target, lengths=None, allreduce: bool = False,
                 warmup: int = 3):
        if not visual.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (no CPU fallback)")
        for g in optimizer.param_groups:
            if not g.get("capturable", False):
                raise ValueError("build the optimiser with capturable=True to capture its step in a CUDA graph")
        if allreduce:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                # measured on 2 x B200: capturing the NCCL all-reduce together with the side-stream warm-up steps
                # hangs; data-parallel training therefore uses the eager step (allreduce_gradients + optimizer.step)
                raise NotImplementedError("GraphedTrainStep does not capture the gradient all-reduce (world size > 1): "
                                          "use the eager step with training.allreduce_gradients")
            allreduce = False
        self.model, self.optimizer, self.loss_fn, self.allreduce = model, optimizer, loss_fn, allreduce
        self.lengths = None if lengths is None else [int(x) for x in lengths]
        self.visual, self.audio, self.target = visual.clone(), audio.clone(), target.clone()
        side = torch.cuda.Stream(device=visual.device)
        side.wait_stream(torch.cuda.current_stream(visual.device))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):      # grows the workspaces, sets kernel attributes, creates optimiser state
                optimizer.zero_grad(set_to_none=True)
                self._body()
        torch.cuda.current_stream(visual.device).wait_stream(side)
        torch.cuda.synchronize(visual.device)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        if hasattr(model, "_native_lstm_key"):
            model._native_lstm_key = None      # the captured forward must contain the recurrences' weight re-pack
        with torch.cuda.graph(self.graph):
            self.loss = self._body()

    def _body(self):
        preds = self.model(self.visual, self.audio) if self.lengths is None else \
            self.model(self.visual, self.audio, self.lengths)
        loss = self.loss_fn(preds, self.target)
        loss.backward()
        if self.allreduce:
            allreduce_gradients(self.model.parameters())
        self.optimizer.step()
        return loss

    def __call__(self, visual, audio, target):
        """Run one step on a new batch of the captured shape; returns the (captured) loss tensor."""
        self.visual.copy_(visual, non_blocking=True)
        self.audio.copy_(audio, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        self.graph.replay()
        return self.loss
