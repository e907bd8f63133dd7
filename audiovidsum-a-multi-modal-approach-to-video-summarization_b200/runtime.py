"""Host-side plumbing between torch tensors and the C ABI.

PyTorch is used for what it is good at here -- device memory, streams, pinned host
buffers -- while every FLOP of the hot path runs inside ``libavsum_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from fractions import Fraction
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi

_LSTM_ORDER = [("visual_bilstm", ""), ("visual_bilstm", "_reverse"), ("audio_bilstm", ""), ("audio_bilstm", "_reverse")]


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("avsum_b200 needs a CUDA (sm_100) device: there is no CPU fallback")


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


class _HostBlock:
    """Owner of one avs_host_alloc allocation (freed when the last tensor view of it dies)."""

    def __init__(self, nbytes: int, write_combined: bool):
        self.ptr = C.c_void_p()
        _cabi.check(_cabi.lib().avs_host_alloc(C.byref(self.ptr), nbytes, int(write_combined)))
        self.nbytes = nbytes

    def __del__(self):
        try:
            if self.ptr:
                _cabi.lib().avs_host_free(self.ptr)
                self.ptr = C.c_void_p()
        except Exception:
            pass


def pinned_like(t: torch.Tensor, write_combined: bool = False) -> torch.Tensor:
    """A page-locked host copy of ``t`` allocated by the library (``avs_host_alloc``).  ``write_combined`` pages are
    meant for buffers the CPU only writes (a loader packing batches): device reads across PCIe skip the CPU-cache
    snoop, which matters when several GPUs of one box stream from host memory at once."""
    _require_cuda()
    src = t.detach().contiguous().cpu()
    nbytes = src.numel() * src.element_size()
    if nbytes == 0:
        return src.pin_memory()
    blk = _HostBlock(nbytes, write_combined)
    buf = (C.c_byte * nbytes).from_address(blk.ptr.value)
    # the allocation lives as long as the STORAGE: torch.frombuffer keeps a reference to ``buf`` for the lifetime of
    # the storage it creates, and ``buf`` owns the block -- so views (slices, reshapes, torch.split pieces, whatever
    # an in-flight asynchronous step keeps) hold the pinned memory alive after the first tensor object is gone
    buf._avs_block = blk
    out = torch.frombuffer(buf, dtype=src.dtype).reshape(src.shape)
    out.copy_(src)
    return out


class ShotDesc:
    """Host descriptors of the shots of a batch, packed once (by the loader) instead of on every call:
    ``n_frames`` int32 [n], ``cps`` int32 [sum S, 2] (inclusive change points), ``cps_start`` int32 [n + 1],
    ``summary_start`` int64 [n + 1]."""

    __slots__ = ("n_frames", "cps", "cps_start", "summary_start")

    def __init__(self, n_frames, cps_list):
        self.n_frames = _i32(n_frames)
        n = int(self.n_frames.size)
        if len(cps_list) != n:
            raise ValueError("one change-point array per video expected")
        arrs = [np.asarray(c, dtype=np.int32).reshape(-1, 2) for c in cps_list]
        self.cps_start = np.zeros(n + 1, dtype=np.int32)
        if n:
            np.cumsum([a.shape[0] for a in arrs], out=self.cps_start[1:])
        self.cps = _i32(np.concatenate(arrs, axis=0) if n else np.zeros((0, 2), np.int32))
        self.summary_start = np.zeros(n + 1, dtype=np.int64)
        self.summary_start[1:] = np.cumsum(self.n_frames.astype(np.int64))


def _shots(n_frames, cps_list) -> ShotDesc:
    return cps_list if isinstance(cps_list, ShotDesc) else ShotDesc(n_frames, cps_list)


_FRACTIONS = {}


def _fraction(proportion) -> Fraction:
    f = _FRACTIONS.get(proportion)
    if f is None:
        f = Fraction(*proportion) if isinstance(proportion, tuple) else Fraction(proportion).limit_denominator(10000)
        _FRACTIONS[proportion] = f
    return f


class PendingSummary:
    """An asynchronous score_and_summarize_rows step in flight on one of the model's two staging slots."""

    def __init__(self, model, slot, outputs, keep):
        self._model, self.slot, self._outputs, self._keep = model, slot, outputs, keep

    def wait(self):
        """Block until the step has completed; returns what score_and_summarize_rows returns."""
        if self._model is not None:
            with torch.cuda.device(self._model.device):
                _cabi.check(self._model.lib.avs_slot_wait(self._model._handle, self.slot))
            self._model, self._keep = None, None
        return self._outputs

    def __del__(self):   # a dropped handle must not leave its slot marked busy
        try:
            self.wait()
        except Exception:
            pass


class NativeModel:
    """Owns one ``avs_model`` handle built from a reference-format state_dict."""

    def __init__(self, state_dict, visual_dim: int, audio_dim: int, hidden_dim: int = 512, num_heads: int = 4,
                 device: Optional[int] = None):
        _require_cuda()
        self.lib = _cabi.lib()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.dims = (visual_dim, audio_dim, hidden_dim, num_heads)
        self._handle = C.c_void_p()
        self._feat_f16 = False
        w, keep = self._weights_struct(state_dict)
        # avs_model_create packs on its own stream; the parameters were last written on torch's current stream
        torch.cuda.current_stream(self.device).synchronize()
        _cabi.check(self.lib.avs_model_create(C.byref(w), self.device, C.byref(self._handle)))
        del keep

    # -- weights --------------------------------------------------------------
    def _weights_struct(self, sd):
        vd, ad, hd, nh = self.dims
        keep = []

        def ptr(name, shape):
            t = sd[name].detach()
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
            t = t.to(dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        hc, e = hd // 2, hd * 2
        w = _cabi.AvsWeights()
        w.visual_dim, w.audio_dim, w.hidden_dim, w.num_heads = vd, ad, hd, nh
        w.visual_fc_w, w.visual_fc_b = ptr("visual_fc.0.weight", (hd, vd)), ptr("visual_fc.0.bias", (hd,))
        w.audio_fc_w, w.audio_fc_b = ptr("audio_fc.0.weight", (hd, ad)), ptr("audio_fc.0.bias", (hd,))
        for i, (mod, suf) in enumerate(_LSTM_ORDER):
            w.lstm_w_ih[i] = ptr(f"{mod}.weight_ih_l0{suf}", (4 * hc, hd))
            w.lstm_w_hh[i] = ptr(f"{mod}.weight_hh_l0{suf}", (4 * hc, hc))
            w.lstm_b_ih[i] = ptr(f"{mod}.bias_ih_l0{suf}", (4 * hc,))
            w.lstm_b_hh[i] = ptr(f"{mod}.bias_hh_l0{suf}", (4 * hc,))
        w.attn_in_w, w.attn_in_b = ptr("attention.in_proj_weight", (3 * e, e)), ptr("attention.in_proj_bias", (3 * e,))
        w.attn_out_w, w.attn_out_b = ptr("attention.out_proj.weight", (e, e)), ptr("attention.out_proj.bias", (e,))
        w.scorer0_w, w.scorer0_b = ptr("scorer.0.weight", (64, e)), ptr("scorer.0.bias", (64,))
        w.scorer2_w, w.scorer2_b = ptr("scorer.2.weight", (1, 64)), ptr("scorer.2.bias", (1,))
        return w, keep

    def update(self, state_dict, lstm_only: bool = False, sync: bool = True):
        """Re-pack after the parameters changed.  ``sync=False`` queues the packing kernels on the current stream
        and returns (the training loop: the optimiser step that changed the parameters ran on the same stream);
        ``lstm_only`` re-packs just the recurrences' tensors, the only ones a training step reads through the handle."""
        w, keep = self._weights_struct(state_dict)
        # always on torch's CURRENT stream: that is where the optimiser step / load_state_dict that changed the
        # parameters ran, and where earlier packs that share the handle's staging arena were queued
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.avs_model_update_async(self._handle, C.byref(w), int(lstm_only),
                                                        _stream_ptr(self.device)))
            if sync:
                torch.cuda.current_stream(self.device).synchronize()
        del keep   # temporaries (non-fp32 / non-contiguous sources) are freed in stream order by torch's allocator

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self.lib.avs_model_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def range_status(self, synchronize: bool = True) -> bool:
        """fp16 range watch for device-space forwards (host-space calls raise by themselves): True when an fc activation
        of a forward on this handle reached the fp16 limit (65504) and was clamped since the last report -- un-normalised
        features of huge magnitude; the reference computes in fp32.  Reading clears the flag.  ``synchronize`` waits for
        the device first (the flag is written by the kernels of the forward)."""
        if synchronize:
            torch.cuda.synchronize(self.device)
        sat = C.c_int32(0)
        _cabi.check(self.lib.avs_model_range_status(self._handle, C.byref(sat)))
        return bool(sat.value)

    def _features(self, visual: torch.Tensor, audio: torch.Tensor):
        """Feature buffers as the library will read them: both tensors float16 -> the opt-in 16-bit feature format
        (avs_model_set_feature_format, half the bytes per frame); anything else -> float32, as in the reference."""
        f16 = visual.dtype == torch.float16 and audio.dtype == torch.float16
        if f16 != self._feat_f16:
            with torch.cuda.device(self.device):
                _cabi.check(self.lib.avs_model_set_feature_format(
                    self._handle, _cabi.AVS_FEAT_F16 if f16 else _cabi.AVS_FEAT_F32))
            self._feat_f16 = f16
        dt = torch.float16 if f16 else torch.float32
        return visual.to(dt).contiguous(), audio.to(dt).contiguous()

    # -- forward ----------------------------------------------------------------
    def forward_rows(self, visual: torch.Tensor, audio: torch.Tensor, row_start, lengths, attn_axis: str = "literal",
                     precision: str = "tf32", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """visual [R, Dv], audio [R, Da] (both on the model's GPU, or both on the host) -> scores [R]."""
        vd, ad, _, _ = self.dims
        if visual.dim() != 2 or audio.dim() != 2 or visual.shape[1] != vd or audio.shape[1] != ad:
            raise ValueError(f"expected visual [R, {vd}] and audio [R, {ad}], got {tuple(visual.shape)} / {tuple(audio.shape)}")
        if visual.shape[0] != audio.shape[0]:
            raise ValueError("visual and audio must have the same number of frames")
        if visual.device != audio.device:
            raise ValueError("visual and audio must live on the same device")
        R = int(visual.shape[0])
        visual, audio = self._features(visual, audio)
        rs, ln = _i32(row_start), _i32(lengths)
        if rs.shape != ln.shape or rs.ndim != 1:
            raise ValueError("row_start and lengths must be 1-D arrays of equal size")
        on_gpu = visual.is_cuda
        if on_gpu and visual.device.index != self.device:
            raise ValueError(f"inputs are on cuda:{visual.device.index}, the model on cuda:{self.device}")
        if out is None:
            out = torch.empty(R, dtype=torch.float32, device=visual.device, pin_memory=(not on_gpu))
        space = _cabi.AVS_DEVICE if on_gpu else _cabi.AVS_HOST
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.avs_forward(
                self._handle, C.c_void_p(visual.data_ptr()), C.c_void_p(audio.data_ptr()), R, int(rs.size),
                _cabi.np_ptr(rs), _cabi.np_ptr(ln), _cabi.ATTN_AXES[attn_axis], _cabi.PRECISIONS[precision],
                C.c_void_p(out.data_ptr()), space, _stream_ptr(self.device)))
        return out

    # -- summary ------------------------------------------------------------------
    def summarize_rows(self, scores: torch.Tensor, positions: torch.Tensor, row_start, lengths, n_frames, cps_list,
                       proportion=0.15, want_summary: bool = True):
        """Batched shot pooling + knapsack.  Returns (picks uint8[sum S], seg_mean int64[sum S],
        summary uint8[sum n_frames] | None, cps_start, summary_start) as torch tensors in the
        memory space of ``scores``."""
        frac = _fraction(proportion)
        rs, ln = _i32(row_start), _i32(lengths)
        n = int(rs.size)
        sd = _shots(n_frames, cps_list)
        nf, cps, cps_start, summary_start = sd.n_frames, sd.cps, sd.cps_start, sd.summary_start
        if int(nf.size) != n:
            raise ValueError("n_frames / change points do not match the number of videos")
        total_S = int(cps_start[-1])
        on_gpu = scores.is_cuda
        dev = scores.device
        scores = scores.to(torch.float32).contiguous()
        positions = positions.to(device=dev, dtype=torch.int32).contiguous()
        pin = not on_gpu
        picks = torch.empty(total_S, dtype=torch.uint8, device=dev, pin_memory=pin)
        seg_mean = torch.empty(total_S, dtype=torch.int64, device=dev, pin_memory=pin)
        summary = torch.empty(int(summary_start[-1]), dtype=torch.uint8, device=dev, pin_memory=pin) if want_summary else None
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.avs_summarize(
                self._handle, C.c_void_p(scores.data_ptr()), C.c_void_p(positions.data_ptr()), n, _cabi.np_ptr(rs),
                _cabi.np_ptr(ln), _cabi.np_ptr(nf), _cabi.np_ptr(cps), _cabi.np_ptr(cps_start),
                int(frac.numerator), int(frac.denominator), C.c_void_p(picks.data_ptr()),
                C.c_void_p(seg_mean.data_ptr()), C.c_void_p(summary.data_ptr()) if want_summary else None,
                _cabi.np_ptr(summary_start) if want_summary else None,
                _cabi.AVS_DEVICE if on_gpu else _cabi.AVS_HOST, _stream_ptr(self.device)))
        return picks, seg_mean, summary, cps_start, summary_start


    # -- forward + summary in one native call -----------------------------------------------
    def score_and_summarize_rows(self, visual: torch.Tensor, audio: torch.Tensor, positions: torch.Tensor, row_start,
                                 lengths, n_frames, cps_list, proportion=0.15, attn_axis: str = "literal_b1",
                                 precision: str = "tf32", want_summary: bool = True, slot: Optional[int] = None):
        """avs_forward_summarize: the whole "scored + summarised" step.  All tensors on the model's GPU, or all on
        the host (pinned recommended): then the features are pipelined across PCIe by video group, the scores
        never leave the device between the two halves and the call returns after one synchronisation.
        ``cps_list`` may be a ShotDesc packed once by the loader (``n_frames`` is then ignored).
        Returns (scores fp32 [R], picks uint8 [sum S], seg_mean int64 [sum S], summary uint8 | None, cps_start,
        summary_start).

        ``slot`` (0 or 1, pinned host tensors only) makes the call asynchronous (avs_forward_summarize_async): it
        returns a PendingSummary at once and the next batch can be submitted on the other slot, so that its
        features cross PCIe while this batch is still being computed; ``.wait()`` returns the tuple above."""
        vd, ad, _, _ = self.dims
        if visual.dim() != 2 or audio.dim() != 2 or visual.shape[1] != vd or audio.shape[1] != ad:
            raise ValueError(f"expected visual [R, {vd}] and audio [R, {ad}], got {tuple(visual.shape)} / {tuple(audio.shape)}")
        if not (visual.device == audio.device == positions.device):
            raise ValueError("visual, audio and positions must live on the same device")
        R = int(visual.shape[0])
        visual, audio = self._features(visual, audio)
        positions = positions.to(torch.int32).contiguous()
        frac = _fraction(proportion)
        rs, ln = _i32(row_start), _i32(lengths)
        n = int(rs.size)
        sd = _shots(n_frames, cps_list)
        nf, cps, cps_start, summary_start = sd.n_frames, sd.cps, sd.cps_start, sd.summary_start
        if int(nf.size) != n:
            raise ValueError("n_frames / change points do not match the number of videos")
        total_S = int(cps_start[-1])
        on_gpu = visual.is_cuda
        dev, pin = visual.device, not on_gpu
        scores = torch.empty(R, dtype=torch.float32, device=dev, pin_memory=pin)
        picks = torch.empty(total_S, dtype=torch.uint8, device=dev, pin_memory=pin)
        seg_mean = torch.empty(total_S, dtype=torch.int64, device=dev, pin_memory=pin)
        summary = torch.empty(int(summary_start[-1]), dtype=torch.uint8, device=dev, pin_memory=pin) if want_summary else None
        outputs = (scores, picks, seg_mean, summary, cps_start, summary_start)
        if slot is not None:
            if on_gpu or not (visual.is_pinned() and audio.is_pinned() and positions.is_pinned()):
                raise ValueError("asynchronous steps need pinned host tensors")
            with torch.cuda.device(self.device):
                _cabi.check(self.lib.avs_forward_summarize_async(
                    self._handle, C.c_void_p(visual.data_ptr()), C.c_void_p(audio.data_ptr()),
                    C.c_void_p(positions.data_ptr()), R, n, _cabi.np_ptr(rs), _cabi.np_ptr(ln),
                    _cabi.ATTN_AXES[attn_axis], _cabi.PRECISIONS[precision], _cabi.np_ptr(nf), _cabi.np_ptr(cps),
                    _cabi.np_ptr(cps_start), int(frac.numerator), int(frac.denominator),
                    C.c_void_p(scores.data_ptr()), C.c_void_p(picks.data_ptr()), C.c_void_p(seg_mean.data_ptr()),
                    C.c_void_p(summary.data_ptr()) if want_summary else None,
                    _cabi.np_ptr(summary_start) if want_summary else None, int(slot), _stream_ptr(self.device)))
            return PendingSummary(self, int(slot), outputs, (visual, audio, positions))
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.avs_forward_summarize(
                self._handle, C.c_void_p(visual.data_ptr()), C.c_void_p(audio.data_ptr()), C.c_void_p(positions.data_ptr()),
                R, n, _cabi.np_ptr(rs), _cabi.np_ptr(ln), _cabi.ATTN_AXES[attn_axis], _cabi.PRECISIONS[precision],
                _cabi.np_ptr(nf), _cabi.np_ptr(cps), _cabi.np_ptr(cps_start), int(frac.numerator), int(frac.denominator),
                C.c_void_p(scores.data_ptr()), C.c_void_p(picks.data_ptr()), C.c_void_p(seg_mean.data_ptr()),
                C.c_void_p(summary.data_ptr()) if want_summary else None,
                _cabi.np_ptr(summary_start) if want_summary else None,
                _cabi.AVS_DEVICE if on_gpu else _cabi.AVS_HOST, _stream_ptr(self.device)))
        return outputs


# ---- stateless building blocks (device tensors) ------------------------------------

def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = False,
           precision: str = "tf32") -> torch.Tensor:
    """nn.Linear(+ReLU) on a CUDA tensor [..., K] through avs_linear."""
    _require_cuda()
    if not x.is_cuda:
        raise RuntimeError("avsum_b200.linear needs CUDA tensors (no CPU fallback)")
    K = x.shape[-1]
    N = weight.shape[0]
    x2 = x.reshape(-1, K).to(torch.float32).contiguous()
    w = weight.detach().to(device=x.device, dtype=torch.float32).contiguous()
    b = None if bias is None else bias.detach().to(device=x.device, dtype=torch.float32).contiguous()
    out = torch.empty(x2.shape[0], N, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().avs_linear(
            C.c_void_p(x2.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()) if b is not None else None,
            int(x2.shape[0]), int(N), int(K), int(relu), _cabi.PRECISIONS[precision], C.c_void_p(out.data_ptr()),
            _stream_ptr(x.device)))
    return out.reshape(*x.shape[:-1], N)


def attention(qkv: torch.Tensor, embed_dim: int, num_heads: int, seq_base, seq_stride, seq_len,
              precision: str = "tf32") -> torch.Tensor:
    """Multi-head softmax attention over row sequences of a packed [rows, 3E] q|k|v tensor."""
    _require_cuda()
    rows = int(qkv.shape[0])
    qkv = qkv.to(torch.float32).contiguous()
    ctx = torch.zeros(rows, embed_dim, dtype=torch.float32, device=qkv.device)
    sb, ss, sl = _i32(seq_base), _i32(seq_stride), _i32(seq_len)
    with torch.cuda.device(qkv.device):
        _cabi.check(_cabi.lib().avs_attention(
            C.c_void_p(qkv.data_ptr()), rows, int(embed_dim), int(num_heads), int(sb.size), _cabi.np_ptr(sb),
            _cabi.np_ptr(ss), _cabi.np_ptr(sl), _cabi.PRECISIONS[precision], C.c_void_p(ctx.data_ptr()),
            _stream_ptr(qkv.device)))
    return ctx


def temporal_f1_batch(pred_lists: Sequence[Sequence[Tuple[int, int]]], gt_lists: Sequence[Sequence[Tuple[int, int]]]) -> np.ndarray:
    """Batched overlap-F1 (evaluation/metrics.py:1-9) on the GPU; returns float64[n]."""
    _require_cuda()
    n = len(pred_lists)
    ps = np.zeros(n + 1, np.int32)
    gs = np.zeros(n + 1, np.int32)
    for i in range(n):
        ps[i + 1] = ps[i] + len(pred_lists[i])
        gs[i + 1] = gs[i] + len(gt_lists[i])
    flat = lambda lists: _i32(np.asarray([p for l in lists for p in l], dtype=np.int32).reshape(-1, 2))
    pred, gt = flat(pred_lists), flat(gt_lists)
    out = np.empty(n, dtype=np.float64)
    dev = torch.cuda.current_device()
    _cabi.check(_cabi.lib().avs_temporal_f1(_cabi.np_ptr(pred), _cabi.np_ptr(ps), _cabi.np_ptr(gt), _cabi.np_ptr(gs),
                                            n, _cabi.np_ptr(out), _stream_ptr(dev)))
    return out


# ---- evaluate() metric block and features/fusion.py helpers -------------------------------------

def eval_metrics_rows(pred: torch.Tensor, target: torch.Tensor, row_start, lengths):
    """Batched scripts/evaluate.py:25-36 on the GPU.

    pred fp32 [rows], target fp32 / fp64 [rows] -- both on the GPU or both on the host -- and the usual
    (row_start, lengths) descriptors.  Returns (metrics float64 [n, 4] = f1, spearman, kendall, mean(pred);
    counts int64 [n, 8]) as numpy arrays.
    """
    _require_cuda()
    rs, ln = _i32(row_start), _i32(lengths)
    n = int(rs.size)
    if pred.device != target.device:
        raise ValueError("pred and target must live on the same device")
    pred = pred.to(torch.float32).contiguous()
    if target.dtype not in (torch.float32, torch.float64):
        target = target.to(torch.float64)       # integer annotations compare like float64 in numpy
    target = target.contiguous()
    if pred.numel() != target.numel():
        raise ValueError("pred and target must have the same number of rows")
    on_gpu = pred.is_cuda
    dev = pred.device if on_gpu else torch.device("cuda", torch.cuda.current_device())
    metrics = torch.empty(n, 4, dtype=torch.float64, device=pred.device)
    counts = torch.empty(n, 8, dtype=torch.int64, device=pred.device)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().avs_eval_metrics(
            C.c_void_p(pred.data_ptr()), C.c_void_p(target.data_ptr()), int(target.dtype == torch.float64), n,
            _cabi.np_ptr(rs), _cabi.np_ptr(ln), C.c_void_p(metrics.data_ptr()), C.c_void_p(counts.data_ptr()),
            _cabi.AVS_DEVICE if on_gpu else _cabi.AVS_HOST, _stream_ptr(dev)))
    return metrics.cpu().numpy(), counts.cpu().numpy()


def cdist_euclidean(a: torch.Tensor, b: torch.Tensor) -> np.ndarray:
    """scipy.spatial.distance.cdist(a, b, 'euclidean') -> float64 numpy [na, nb] (bit-exact), on the GPU."""
    _require_cuda()
    a = a.detach().to(torch.float32).contiguous()
    b = b.detach().to(torch.float32).contiguous()
    on_gpu = a.is_cuda
    if b.is_cuda != on_gpu:
        raise ValueError("both inputs must live on the same device")
    dev = a.device if on_gpu else torch.device("cuda", torch.cuda.current_device())
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float64, device=a.device)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().avs_cdist(
            C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), int(a.shape[0]), int(b.shape[0]), int(a.shape[1]),
            C.c_void_p(out.data_ptr()), _cabi.AVS_DEVICE if on_gpu else _cabi.AVS_HOST, _stream_ptr(dev)))
    return out.cpu().numpy()


def gather_scale(features: torch.Tensor, idx, weights) -> torch.Tensor:
    """out[k, :] = features[idx[k], :] * float32(weights[k]) on the GPU; result in the memory space of features."""
    _require_cuda()
    f = features.detach().to(torch.float32).contiguous()
    ix = _i32(idx)
    w = np.ascontiguousarray(np.asarray(weights, dtype=np.float32))
    on_gpu = f.is_cuda
    dev = f.device if on_gpu else torch.device("cuda", torch.cuda.current_device())
    out = torch.empty(int(ix.size), f.shape[1], dtype=torch.float32, device=f.device)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().avs_interpolate(
            C.c_void_p(f.data_ptr()), int(f.shape[0]), int(f.shape[1]), _cabi.np_ptr(ix), _cabi.np_ptr(w),
            int(ix.size), C.c_void_p(out.data_ptr()), _cabi.AVS_DEVICE if on_gpu else _cabi.AVS_HOST, _stream_ptr(dev)))
    return out


def dtw_path(cost) -> Tuple[float, np.ndarray]:
    """Exact DTW through a float64 cost matrix (numpy or torch, host or device) -> (total cost, path int [P, 2])."""
    _require_cuda()
    if isinstance(cost, torch.Tensor):
        c = cost.detach().to(torch.float64).contiguous()
        ptr, keep = c.data_ptr(), c
        n, m = int(c.shape[0]), int(c.shape[1])
    else:
        c = np.ascontiguousarray(np.asarray(cost, dtype=np.float64))
        ptr, keep = c.ctypes.data, c
        n, m = int(c.shape[0]), int(c.shape[1])
    path = np.zeros((max(n + m - 1, 1), 2), dtype=np.int32)
    plen = np.zeros(1, dtype=np.int32)
    total = np.zeros(1, dtype=np.float64)
    dev = torch.cuda.current_device()
    _cabi.check(_cabi.lib().avs_dtw_path(C.c_void_p(ptr), n, m, _cabi.np_ptr(path), _cabi.np_ptr(plen),
                                         _cabi.np_ptr(total), _stream_ptr(dev)))
    del keep
    return float(total[0]), path[:int(plen[0])].astype(np.int64)
