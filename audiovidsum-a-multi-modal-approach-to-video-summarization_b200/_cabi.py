"""ctypes binding of ``libavsum_b200.so`` (the C ABI in ``include/avsum_b200.h``).

This module is the only place Python touches the native library.  It never falls
back to another implementation: if the library is missing or a call fails, an
exception carrying ``avs_last_error()`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# AVS_LIB_PATH: load another build of the same library (A/B timing of kernel changes, tools/ab_lib.py)
LIB_PATH = os.environ.get("AVS_LIB_PATH") or os.path.join(_HERE, "libavsum_b200.so")

AVS_OK, AVS_ERR_INVALID, AVS_ERR_UNSUPPORTED, AVS_ERR_CUDA, AVS_ERR_OOM = range(5)
AVS_HOST, AVS_DEVICE = 0, 1
AVS_ATTN_LITERAL, AVS_ATTN_TEMPORAL, AVS_ATTN_LITERAL_B1 = 0, 1, 2
AVS_PREC_TF32, AVS_PREC_BF16, AVS_PREC_FP32_SIMT = 0, 1, 2
AVS_FEAT_F32, AVS_FEAT_F16 = 0, 1

ATTN_AXES = {"literal": AVS_ATTN_LITERAL, "temporal": AVS_ATTN_TEMPORAL, "literal_b1": AVS_ATTN_LITERAL_B1}
PRECISIONS = {"tf32": AVS_PREC_TF32, "bf16": AVS_PREC_BF16, "fp32_simt": AVS_PREC_FP32_SIMT}

# every symbol include/avsum_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "avs_last_error", "avs_version", "avs_device_ok", "avs_model_create", "avs_model_update",
    "avs_model_destroy", "avs_forward", "avs_summarize", "avs_linear", "avs_bilstm_pair",
    "avs_attention", "avs_temporal_f1", "avs_launch_count", "avs_profile", "avs_profile_stages",
    "avs_profile_stage_name", "avs_profile_read", "avs_debug_lstm_trace", "avs_debug_bptt_trace", "avs_model_range_status",
    "avs_eval_metrics", "avs_cdist", "avs_interpolate", "avs_dtw_path",
    "avs_bilstm_pair_train", "avs_bilstm_pair_bwd", "avs_linear_bwd", "avs_forward_summarize",
    "avs_debug_e2e_trace", "avs_forward_summarize_async", "avs_slot_wait",
    "avs_debug_gemm_trace", "avs_model_update_async", "avs_host_alloc", "avs_host_free",
    "avs_model_set_feature_format", "avs_debug_plan",
]


class AvsError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"avsum_b200 native call failed (status {status}): {message}")
        self.status = status


class AvsUnsupported(AvsError, NotImplementedError):
    pass


class AvsWeights(C.Structure):
    _fields_ = [
        ("visual_dim", C.c_int32), ("audio_dim", C.c_int32), ("hidden_dim", C.c_int32), ("num_heads", C.c_int32),
        ("visual_fc_w", C.c_void_p), ("visual_fc_b", C.c_void_p),
        ("audio_fc_w", C.c_void_p), ("audio_fc_b", C.c_void_p),
        ("lstm_w_ih", C.c_void_p * 4), ("lstm_w_hh", C.c_void_p * 4),
        ("lstm_b_ih", C.c_void_p * 4), ("lstm_b_hh", C.c_void_p * 4),
        ("attn_in_w", C.c_void_p), ("attn_in_b", C.c_void_p),
        ("attn_out_w", C.c_void_p), ("attn_out_b", C.c_void_p),
        ("scorer0_w", C.c_void_p), ("scorer0_b", C.c_void_p),
        ("scorer2_w", C.c_void_p), ("scorer2_b", C.c_void_p),
    ]


_lib = None


def build(verbose: bool = False) -> str:
    """Compile csrc/ -> libavsum_b200.so with nvcc for sm_100a (no GPU needed)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libavsum_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the native library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C <package>/csrc`).  avsum_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.avs_last_error.restype = C.c_char_p
    L.avs_last_error.argtypes = []
    L.avs_version.restype = C.c_int
    L.avs_device_ok.restype = C.c_int
    L.avs_launch_count.restype = i64
    L.avs_model_create.restype = C.c_int
    L.avs_model_create.argtypes = [C.POINTER(AvsWeights), C.c_int, C.POINTER(vp)]
    L.avs_model_update.restype = C.c_int
    L.avs_model_update.argtypes = [vp, C.POINTER(AvsWeights)]
    L.avs_model_update_async.restype = C.c_int
    L.avs_model_update_async.argtypes = [vp, C.POINTER(AvsWeights), C.c_int, vp]
    L.avs_model_set_feature_format.restype = C.c_int
    L.avs_model_set_feature_format.argtypes = [vp, C.c_int]
    L.avs_host_alloc.restype = C.c_int
    L.avs_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t, C.c_int]
    L.avs_host_free.restype = C.c_int
    L.avs_host_free.argtypes = [vp]
    L.avs_debug_plan.restype = C.c_int
    L.avs_debug_plan.argtypes = [i32, vp, vp, i64, vp, vp, vp, i32]
    L.avs_model_destroy.restype = None
    L.avs_model_destroy.argtypes = [vp]
    L.avs_forward.restype = C.c_int
    L.avs_forward.argtypes = [vp, vp, vp, i64, i32, vp, vp, C.c_int, C.c_int, vp, C.c_int, vp]
    L.avs_summarize.restype = C.c_int
    L.avs_summarize.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, C.c_int, vp]
    L.avs_forward_summarize.restype = C.c_int
    L.avs_forward_summarize.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp, C.c_int, C.c_int, vp, vp, vp, i32, i32,
                                        vp, vp, vp, vp, vp, C.c_int, vp]
    L.avs_forward_summarize_async.restype = C.c_int
    L.avs_forward_summarize_async.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp, C.c_int, C.c_int, vp, vp, vp, i32, i32,
                                              vp, vp, vp, vp, vp, C.c_int, vp]
    L.avs_slot_wait.restype = C.c_int
    L.avs_slot_wait.argtypes = [vp, C.c_int]
    L.avs_linear.restype = C.c_int
    L.avs_linear.argtypes = [vp, vp, vp, i64, i32, i32, C.c_int, C.c_int, vp, vp]
    L.avs_bilstm_pair.restype = C.c_int
    L.avs_bilstm_pair.argtypes = [vp, vp, vp, i64, i32, vp, vp, C.c_int, vp, vp]
    L.avs_attention.restype = C.c_int
    L.avs_attention.argtypes = [vp, i64, i32, i32, i32, vp, vp, vp, C.c_int, vp, vp]
    L.avs_temporal_f1.restype = C.c_int
    L.avs_temporal_f1.argtypes = [vp, vp, vp, vp, i32, vp, vp]
    L.avs_eval_metrics.restype = C.c_int
    L.avs_eval_metrics.argtypes = [vp, vp, C.c_int, i32, vp, vp, vp, vp, C.c_int, vp]
    L.avs_cdist.restype = C.c_int
    L.avs_cdist.argtypes = [vp, vp, i32, i32, i32, vp, C.c_int, vp]
    L.avs_interpolate.restype = C.c_int
    L.avs_interpolate.argtypes = [vp, i64, i32, vp, vp, i32, vp, C.c_int, vp]
    L.avs_dtw_path.restype = C.c_int
    L.avs_dtw_path.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    L.avs_bilstm_pair_train.restype = C.c_int
    L.avs_bilstm_pair_train.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp]
    L.avs_bilstm_pair_bwd.restype = C.c_int
    L.avs_bilstm_pair_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.avs_linear_bwd.restype = C.c_int
    L.avs_linear_bwd.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp, vp, vp]
    L.avs_debug_lstm_trace.restype = C.c_int
    L.avs_debug_lstm_trace.argtypes = [vp]
    L.avs_model_range_status.restype = C.c_int
    L.avs_model_range_status.argtypes = [vp, vp]
    L.avs_debug_bptt_trace.restype = C.c_int
    L.avs_debug_bptt_trace.argtypes = [vp]
    L.avs_debug_gemm_trace.restype = C.c_int
    L.avs_debug_gemm_trace.argtypes = [vp]
    L.avs_debug_e2e_trace.restype = C.c_int
    L.avs_debug_e2e_trace.argtypes = [vp]
    L.avs_profile.restype = None
    L.avs_profile.argtypes = [C.c_int]
    L.avs_profile_stages.restype = C.c_int
    L.avs_profile_stage_name.restype = C.c_char_p
    L.avs_profile_stage_name.argtypes = [C.c_int]
    L.avs_profile_read.restype = None
    L.avs_profile_read.argtypes = [vp, vp]
    _lib = L
    return L


def check(status: int) -> None:
    if status == AVS_OK:
        return
    msg = lib().avs_last_error().decode("utf-8", "replace")
    if status == AVS_ERR_UNSUPPORTED:
        raise AvsUnsupported(status, msg)
    if status == AVS_ERR_INVALID:
        raise ValueError(f"avsum_b200: {msg}")
    if status == AVS_ERR_OOM:
        raise MemoryError(f"avsum_b200: {msg}")
    raise AvsError(status, msg)


def np_ptr(a):
    """void* of a C-contiguous numpy array (host descriptor arrays)."""
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def recurrence_plan(row_start, lengths, total_rows=None):
    """avs_debug_plan: the recurrence groups avs_forward cuts a batch into (host logic, no GPU needed).
    Returns a dict: group_of [n] (-1 for empty videos), n_groups, slots_per_group, rows_ordered_by_group (what the
    per-group schedule needs), groups_end_apart, group_rows [n_groups, 2] or None."""
    import numpy as np
    rs = np.ascontiguousarray(row_start, dtype=np.int32)
    ln = np.ascontiguousarray(lengths, dtype=np.int32)
    n = int(rs.size)
    total = int((rs.astype(np.int64) + ln).max()) if total_rows is None and n else int(total_rows or 0)
    group_of = np.full(max(n, 1), -1, dtype=np.int32)
    info = np.zeros(4, dtype=np.int32)
    rows = np.zeros((64, 2), dtype=np.int64)
    check(lib().avs_debug_plan(n, np_ptr(rs), np_ptr(ln), total, np_ptr(group_of), np_ptr(info), np_ptr(rows), 64))
    g = int(info[0])
    return {"group_of": group_of[:n].copy(), "n_groups": g, "slots_per_group": int(info[1]),
            "rows_ordered_by_group": bool(info[2]), "groups_end_apart": bool(info[3]),
            "group_rows": rows[:g].copy() if info[2] and g <= 64 else None}


def launch_count() -> int:
    return int(lib().avs_launch_count())


def profile(enable: int) -> None:
    lib().avs_profile(int(enable))


def profile_read():
    """{stage name: (total ms, calls)} accumulated since the last reset."""
    import numpy as np
    L = lib()
    n = L.avs_profile_stages()
    ms = np.zeros(n, dtype=np.float64)
    calls = np.zeros(n, dtype=np.int64)
    L.avs_profile_read(C.c_void_p(ms.ctypes.data), C.c_void_p(calls.ctypes.data))
    return {L.avs_profile_stage_name(i).decode(): (float(ms[i]), int(calls[i])) for i in range(n)}
