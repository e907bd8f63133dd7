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
step).  Data parallelism (BASELINE configs[4]): ``GradBuckets`` keeps ONE flat fp32 gradient buffer (8,008,833
elements = 32 MB for the 1024/128 model) cut into buckets in the order the backward produces the gradients (score
head and attention first, the recurrences next, the fc layers last); a bucket's NCCL all-reduce starts from the
autograd hook of its last gradient and runs beside the rest of the backward.  ``allreduce_gradients`` is the
unbucketed form (one flat bucket after the backward).  ``TrainStep`` runs / captures the whole step.
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
        # zeros: rows no video owns (and everything, when all lengths are 0) must carry no gradient
        d_v, d_a = torch.zeros_like(v_emb), torch.zeros_like(a_emb)
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


class GradBuckets:
    """Bucketed, overlapped gradient averaging for data-parallel training (scripts/train_av_model.py:86-96 on N GPUs).

    All gradients live in one flat fp32 buffer, parameters laid out in REVERSE registration order -- the order in
    which the backward of AVBiLSTMModel finishes them (scorer, attention, the four recurrences, fc).  The buffer is
    cut into buckets of about ``bucket_mb`` at parameter boundaries.  A post-accumulate-grad hook per parameter
    moves the fresh gradient into its slice (``p.grad`` becomes a view of the flat buffer, so the optimiser reads
    the averaged values in place) and, when the last gradient of a bucket has arrived, launches that bucket's
    all-reduce with ``async_op=True``: torch's NCCL process group runs it on its own stream, ordered after the
    producing kernels by an event, so the transfer overlaps the remaining backward kernels.  ``finish()`` launches
    whatever has not been launched (parameters that received no gradient contribute zeros) and makes the current
    stream wait for every bucket.  NCCL averages in the collective (ReduceOp.AVG); other backends (gloo in the CPU
    tests) sum and divide.  Works under CUDA-graph capture: the collectives are captured with the step.
    """

    def __init__(self, parameters, bucket_mb: float = 12.0, group=None, world_size: int = None):
        import torch.distributed as dist
        self.params = [p for p in parameters if p.requires_grad]
        if not self.params:
            raise ValueError("no parameter requires a gradient")
        self.group = group
        self.dist = dist
        self.active = dist.is_available() and dist.is_initialized()
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if self.active else 1)
        dev, dt = self.params[0].device, self.params[0].dtype
        order = list(reversed(self.params))
        self.flat = torch.zeros(sum(p.numel() for p in order), dtype=dt, device=dev)
        limit = max(int(bucket_mb * (1 << 20) / self.flat.element_size()), 1)
        self.buckets = []            # [start, end, n_params]
        self.slot = {}               # id(p) -> (view, bucket index)
        off = start = count = 0
        for p in order:
            n = p.numel()
            if count and off + n - start > limit:
                self.buckets.append([start, off, count])
                start, count = off, 0
            self.slot[id(p)] = (self.flat[off:off + n].view_as(p), len(self.buckets))
            off += n
            count += 1
        self.buckets.append([start, off, count])
        self.avg_in_collective = self.active and dist.get_backend(group) == "nccl"
        self._arrived = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._works = []
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    def close(self):
        for h in self._handles:
            h.remove()
        self._handles = []

    def _launch(self, b):
        self._launched[b] = True
        if self.world <= 1 or not self.active:
            return
        s, e, _ = self.buckets[b]
        op = self.dist.ReduceOp.AVG if self.avg_in_collective else self.dist.ReduceOp.SUM
        self._works.append((self.dist.all_reduce(self.flat[s:e], op=op, group=self.group, async_op=True), b))

    def _hook(self, p):
        view, b = self.slot[id(p)]
        if p.grad is not view:
            view.copy_(p.grad)
            p.grad = view
        self._arrived[b] += 1
        if self._arrived[b] == self.buckets[b][2] and not self._launched[b]:
            self._launch(b)

    def start(self):
        """Call before the backward of a step (after ``zero_grad(set_to_none=True)``)."""
        self._arrived = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._works = []

    def finish(self):
        """Call after the backward: every gradient averaged and visible to the current stream."""
        for p in self.params:                      # a parameter the loss does not depend on: zero gradient
            view, b = self.slot[id(p)]
            if p.grad is None:
                view.zero_()
                p.grad = view
        for b in range(len(self.buckets)):
            if not self._launched[b]:
                self._launch(b)
        for w, b in self._works:
            w.wait()
            if not self.avg_in_collective:
                s, e, _ = self.buckets[b]
                self.flat[s:e].div_(self.world)
        self._works = []
        return int(self.flat.numel())


class TrainStep:
    """The training step of scripts/train_av_model.py:86-96 -- forward, loss, ``backward``, gradient averaging over
    the data-parallel ranks (``GradBuckets``: bucketed NCCL all-reduce overlapped with the backward), ``optimizer
    .step`` -- launched eagerly or, with ``graph=True``, captured ONCE in a CUDA graph (collectives included) and
    replayed: the eager step is host-bound (~100 small launches), the replay is one graph launch.

    The batch shape is fixed at capture (``visual [B, T, Dv]``, ``audio [B, T, Da]``, ``target`` as the loss function
    takes it, optional ``lengths``); every call copies the new batch (device or pinned host tensors) into the
    captured input buffers.  A graphed step needs an optimiser built with ``capturable=True`` (torch's requirement).
    Launch plans and sequence descriptors travel as kernel parameters (no pageable-memory copies), the recurrences'
    weight re-pack after the optimiser step is part of the graph, and workspaces are grown during the warm-up
    steps, so nothing inside the captured region allocates through cudaMalloc or synchronises.
    """

    def __init__(self, model, optimizer, loss_fn, visual, audio, target, lengths=None, world_size: int = None,
                 graph: bool = True, warmup: int = 3, bucket_mb: float = 12.0):
        if not visual.is_cuda:
            raise RuntimeError("TrainStep needs CUDA tensors (no CPU fallback)")
        if graph:
            for g in optimizer.param_groups:
                if not g.get("capturable", False):
                    raise ValueError("build the optimiser with capturable=True to capture its step in a CUDA graph")
        import torch.distributed as dist
        ddp = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if ddp else 1
        if world_size is not None and world_size != self.world:
            raise ValueError(f"world_size={world_size} but torch.distributed reports {self.world} rank(s)")
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.lengths = None if lengths is None else [int(x) for x in lengths]
        self.visual, self.audio, self.target = visual.clone(), audio.clone(), target.clone()
        self.buckets = GradBuckets(model.parameters(), bucket_mb=bucket_mb) if ddp else None
        self.allreduce_mode = (f"{len(self.buckets.buckets)} buckets, async all-reduce launched from autograd hooks "
                               f"(overlaps the backward)" if ddp else "none (1 GPU)")
        self.graphed = False
        self.graph = None
        self.launches_per_replay = 0
        dev = visual.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):      # grows the workspaces, sets kernel attributes, creates optimiser state
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if graph:
            if ddp:
                dist.barrier()
            self.graph = torch.cuda.CUDAGraph()
            if hasattr(model, "_native_lstm_key"):
                model._native_lstm_key = None      # the captured forward must contain the recurrences' weight re-pack
            n0 = _cabi.launch_count()
            with torch.cuda.graph(self.graph):
                self.loss = self._body()
            self.launches_per_replay = _cabi.launch_count() - n0   # native kernels inside one replay
            self.graphed = True

    def _body(self):
        self.optimizer.zero_grad(set_to_none=True)
        if self.buckets is not None:
            self.buckets.start()
        preds = self.model(self.visual, self.audio) if self.lengths is None else \
            self.model(self.visual, self.audio, self.lengths)
        loss = self.loss_fn(preds, self.target)
        loss.backward()
        if self.buckets is not None:
            self.buckets.finish()
        self.optimizer.step()
        return loss

    def __call__(self, visual, audio, target):
        """Run one step on a new batch of the captured shape; returns the loss tensor (of the captured graph)."""
        self.visual.copy_(visual, non_blocking=True)
        self.audio.copy_(audio, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
            return self.loss
        return self._body()


class GraphedTrainStep(TrainStep):
    """Round-1 name of the single-GPU captured step (kept for callers): ``TrainStep(graph=True)``."""

    def __init__(self, model, optimizer, loss_fn, visual, audio, target, lengths=None, allreduce: bool = False,
                 warmup: int = 3):
        super().__init__(model, optimizer, loss_fn, visual, audio, target, lengths=lengths, graph=True, warmup=warmup)
