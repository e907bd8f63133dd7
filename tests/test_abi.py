"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/avsum_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from avsum_b200 import _cabi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "avsum_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(avs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(native_lib):
    declared = _declared_symbols()
    assert declared, "no declarations parsed from the header"
    assert sorted(_cabi.EXPORTS) == declared
    for name in declared:
        assert getattr(native_lib, name) is not None


def test_version_and_error_string(native_lib):
    assert native_lib.avs_version() >= 100
    assert isinstance(native_lib.avs_last_error(), bytes)
    assert native_lib.avs_launch_count() >= 0


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "avsum_b200.h"\nint main(void){ avs_weights w; (void)w; return AVS_OK; }\n')
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                        "-o", str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_weights_struct_layout_matches_header():
    # 4 int32 + 28 pointers, no padding surprises
    assert C.sizeof(_cabi.AvsWeights) == 16 + 28 * C.sizeof(C.c_void_p)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(native_lib):
    assert native_lib.avs_device_ok() == 0
    from avsum_b200.models.av_model import AVBiLSTMModel
    from avsum_b200.models.attention import MultiHeadSelfAttention
    m = AVBiLSTMModel(1024, 128, 512).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 1024), torch.zeros(1, 4, 128))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MultiHeadSelfAttention(1024, 4)(torch.zeros(1, 4, 1024))
    # the raw ABI reports a CUDA error instead of computing on the host
    sd = synth.seeded_state_dict()
    keep = {k: v.contiguous() for k, v in sd.items()}
    w = _cabi.AvsWeights()
    w.visual_dim, w.audio_dim, w.hidden_dim, w.num_heads = 1024, 128, 512, 4
    for f, k in [("visual_fc_w", "visual_fc.0.weight"), ("visual_fc_b", "visual_fc.0.bias"), ("audio_fc_w", "audio_fc.0.weight"),
                 ("audio_fc_b", "audio_fc.0.bias"), ("attn_in_w", "attention.in_proj_weight"), ("attn_in_b", "attention.in_proj_bias"),
                 ("attn_out_w", "attention.out_proj.weight"), ("attn_out_b", "attention.out_proj.bias"),
                 ("scorer0_w", "scorer.0.weight"), ("scorer0_b", "scorer.0.bias"), ("scorer2_w", "scorer.2.weight"),
                 ("scorer2_b", "scorer.2.bias")]:
        setattr(w, f, keep[k].data_ptr())
    for i, (mod, suf) in enumerate([("visual_bilstm", ""), ("visual_bilstm", "_reverse"), ("audio_bilstm", ""), ("audio_bilstm", "_reverse")]):
        w.lstm_w_ih[i] = keep[f"{mod}.weight_ih_l0{suf}"].data_ptr()
        w.lstm_w_hh[i] = keep[f"{mod}.weight_hh_l0{suf}"].data_ptr()
        w.lstm_b_ih[i] = keep[f"{mod}.bias_ih_l0{suf}"].data_ptr()
        w.lstm_b_hh[i] = keep[f"{mod}.bias_hh_l0{suf}"].data_ptr()
    h = C.c_void_p()
    st = native_lib.avs_model_create(C.byref(w), 0, C.byref(h))
    assert st == _cabi.AVS_ERR_CUDA and not h.value
    assert b"cuda" in native_lib.avs_last_error().lower()


def test_unsupported_dims_are_reported(native_lib):
    w = _cabi.AvsWeights()
    w.visual_dim, w.audio_dim, w.hidden_dim, w.num_heads = 1024, 128, 256, 4
    h = C.c_void_p()
    assert native_lib.avs_model_create(C.byref(w), 0, C.byref(h)) == _cabi.AVS_ERR_UNSUPPORTED
    assert b"hidden_dim" in native_lib.avs_last_error()
    with pytest.raises(_cabi.AvsUnsupported):
        _cabi.check(_cabi.AVS_ERR_UNSUPPORTED)
