"""GPU parity tests (run with ``-m gpu`` on a B200).  Every check goes through the C ABI
(``libavsum_b200.so``) and compares with the oracle / golden vectors on identical seeded
inputs.  Tolerances: fp32 frame scores <= 1e-3 relative in the tf32 tensor-core mode
(BASELINE.json north_star), <= 2e-5 in the CUDA-core fp32 mode; integer outputs bit-exact.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from avsum_b200 import _cabi, runtime, synth
from oracle import av_oracle, av_oracle_torch

pytestmark = pytest.mark.gpu

# "tf32" = the default mixed mode: tf32 operands for the user features, fp16 (same 11-bit significand) for
# the internal activations, fp32 accumulation.  "bf16" = BASELINE configs[2]: bf16 operands everywhere
# (8-bit significand -> stated looser tolerance, north_star: "looser, stated, for bf16").
REL_TOL = {"tf32": 1e-3, "fp32_simt": 2e-5, "bf16": 2e-2}


def rel(got, want):
    return float(np.max(np.abs(np.asarray(got, np.float64) - want) / np.abs(want)))


def make_model(vd=1024, ad=128, spread=False, **kw):
    from avsum_b200.models.av_model import AVBiLSTMModel
    m = AVBiLSTMModel(vd, ad, 512, **kw).eval()
    m.load_state_dict(synth.seeded_state_dict(vd, ad, 512, 0, spread))
    return m.cuda()


@pytest.fixture(scope="module")
def native(cuda_ready):
    sd = synth.seeded_state_dict()
    return runtime.NativeModel({k: v.cuda() for k, v in sd.items()}, 1024, 128)


# ------------------------------------------------------------------ K1/K3/K5: nn.Linear
@pytest.mark.parametrize("prec", ["tf32", "fp32_simt", "bf16"])
@pytest.mark.parametrize("shape", [(128, 128, 32), (300, 512, 128), (1000, 2048, 512), (500, 64, 1024),
                                   (320, 512, 296), (77, 1024, 1024), (1, 512, 4096), (129, 3072, 1024),
                                   (200, 16, 64), (333, 48, 256), (4097, 192, 64)])
def test_linear_matches_torch(cuda_ready, prec, shape):
    M, N, K = shape
    g = torch.Generator().manual_seed(M * 7 + N)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    for relu in (False, True):
        want = torch.nn.functional.linear(x.double(), w.double(), b.double())
        want = want.relu() if relu else want
        got = runtime.linear(x, w, b, relu=relu, precision=prec)
        err = float((got.double() - want).abs().max() / want.abs().max())
        assert err < {"tf32": 2e-3, "bf16": 1.5e-2, "fp32_simt": 2e-6}[prec], (shape, relu, err)


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
@pytest.mark.parametrize("shape", [(38000, 512, 1024), (37999, 1024, 512), (20001, 2048, 128), (75777, 256, 96)])
def test_linear_cta_pair_tiles(cuda_ready, prec, shape):
    """Shapes large enough for the 2-CTA GEMM (256 x 256 tiles, cta_group::2): ragged M (the pair's second CTA partly
    or wholly past the last row), several tiles per pair, short and long K."""
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    want = torch.nn.functional.linear(x, w, b).relu()          # fp32 cuBLAS reference, TF32 disabled by default
    got = runtime.linear(x, w, b, relu=True, precision=prec)
    err = float((got - want).abs().max() / want.abs().max())
    assert err < {"tf32": 2e-3, "bf16": 1.5e-2}[prec], (shape, err)
    # every row block individually (a swapped or dropped tile would pass a max-norm check only by accident)
    blk = (got - want).abs().reshape(-1)[: (M // 128) * 128 * N].reshape(M // 128, -1).max(dim=1).values
    assert float(blk.max()) < {"tf32": 2e-3, "bf16": 1.5e-2}[prec] * float(want.abs().max())


def test_linear_without_bias_and_bad_shapes(cuda_ready):
    x = torch.randn(40, 64, device="cuda")
    w = torch.randn(32, 64, device="cuda")
    got = runtime.linear(x, w, None, precision="fp32_simt")
    assert float((got - x @ w.T).abs().max()) < 1e-4
    with pytest.raises(_cabi.AvsUnsupported):   # N % 16 != 0 on the tensor-core path
        runtime.linear(x, torch.randn(30, 64, device="cuda"), None, precision="tf32")
    with pytest.raises(ValueError):             # K % 4 != 0 breaks the 16-byte TMA pitch
        runtime.linear(torch.randn(8, 30, device="cuda"), torch.randn(32, 30, device="cuda"), None, precision="tf32")


# ------------------------------------------------------------------ K2: both BiLSTMs
@pytest.mark.parametrize("lens", [[37], [5, 64, 1, 33], list(range(20, 39)), [3] * 70])
def test_bilstm_pair_matches_torch_lstm(native, lens):
    sd = synth.seeded_state_dict()
    R = sum(lens)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    g = torch.Generator().manual_seed(len(lens))
    v, a = torch.randn(R, 512, generator=g), torch.randn(R, 512, generator=g)
    lv = torch.nn.LSTM(512, 256, bidirectional=True, batch_first=True)
    la = torch.nn.LSTM(512, 256, bidirectional=True, batch_first=True)
    lv.load_state_dict({k.split(".", 1)[1]: t for k, t in sd.items() if k.startswith("visual_bilstm")})
    la.load_state_dict({k.split(".", 1)[1]: t for k, t in sd.items() if k.startswith("audio_bilstm")})
    want = torch.empty(R, 1024)
    with torch.no_grad():
        for s, n in zip(starts, lens):
            want[s:s + n, :512] = lv(v[None, s:s + n])[0][0]
            want[s:s + n, 512:] = la(a[None, s:s + n])[0][0]
    ln = np.asarray(lens, dtype=np.int32)
    for prec, tol in (("fp32_simt", 2e-5), ("tf32", 3e-3), ("bf16", 3e-2)):
        fused = torch.zeros(R, 1024, device="cuda")
        vc, ac = v.cuda(), a.cuda()
        _cabi.check(native.lib.avs_bilstm_pair(native._handle, C.c_void_p(vc.data_ptr()), C.c_void_p(ac.data_ptr()), R,
                                               len(lens), _cabi.np_ptr(starts), _cabi.np_ptr(ln), _cabi.PRECISIONS[prec],
                                               C.c_void_p(fused.data_ptr()),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        print(f"bilstm pair {prec} lens={lens}: max |h - torch| {float((fused.cpu() - want).abs().max()):.2e} (tolerance {tol})")
        assert float((fused.cpu() - want).abs().max()) < tol


# ------------------------------------------------------------------ full forward vs reference goldens
@pytest.mark.parametrize("prec", ["tf32", "fp32_simt", "bf16"])
@pytest.mark.parametrize("spread", [0, 1])
def test_forward_config1_matches_reference(cuda_ready, golden_dir, prec, spread):
    g = np.load(os.path.join(golden_dir, f"config1_spread{spread}.npz"))
    m = make_model(spread=bool(spread), precision=prec)
    vid = synth.config1()
    for axis in ("literal", "temporal"):
        got = m(vid.visual[None].cuda(), vid.audio[None].cuda(), attn_axis=axis)
        assert got.shape == (320,) and got.is_cuda
        assert rel(got.cpu().numpy(), g["scores_" + axis]) < REL_TOL[prec]


@pytest.mark.parametrize("prec", ["tf32", "fp32_simt", "bf16"])
@pytest.mark.parametrize("name", ["batch3_T17", "batch2_T1", "batch1_T1", "default_dims_T40", "batch2_T130_spread"])
def test_forward_cases_match_reference(cuda_ready, golden_dir, prec, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    vd, ad, B, T = int(g["visual_dim"]), int(g["audio_dim"]), int(g["B"]), int(g["T"])
    m = make_model(vd, ad, bool(int(g["spread"])), precision=prec)
    gen = torch.Generator().manual_seed(int(g["seed_in"]))
    visual, audio = torch.randn(B, T, vd, generator=gen), torch.randn(B, T, ad, generator=gen)
    for axis in ("literal", "temporal"):
        got = m(visual.cuda(), audio.cuda(), attn_axis=axis).cpu().numpy()
        want = g["scores_" + axis]
        assert got.shape == want.shape          # .squeeze() semantics of av_model.py:46
        assert rel(got, want) < REL_TOL[prec]


def test_forward_host_tensors_and_state_dict_reload(cuda_ready, golden_dir):
    g0 = np.load(os.path.join(golden_dir, "config1_spread0.npz"))
    g1 = np.load(os.path.join(golden_dir, "config1_spread1.npz"))
    m = make_model()
    vid = synth.config1()
    got = m(vid.visual[None], vid.audio[None])            # host inputs: H2D/D2H inside the native call
    assert not got.is_cuda and rel(got.numpy(), g0["scores_literal"]) < 1e-3
    m.load_state_dict(synth.seeded_state_dict(spread=True))   # in-place parameter change -> re-pack
    got = m(vid.visual[None].cuda(), vid.audio[None].cuda()).cpu().numpy()
    assert rel(got, g1["scores_literal"]) < 1e-3


def test_forward_with_fp16_exact_features(cuda_ready):
    """Operand policy of the kind::tf32 fc GEMMs (csrc/api.cu, avs_linear / avs_forward): the activation operand is
    truncated by the tensor core and the -2^-11 MEAN of that truncation is scaled out of the accumulator, which
    assumes uniformly distributed low mantissa bits.  Features that are already exact in tf32 (fp16 values upcast to
    fp32 -- e.g. .npy files stored in half precision) are not truncated at all and see the +4.9e-4 scale as a bias;
    the scores must stay inside the 1e-3 budget for them too."""
    vid = synth.config1()
    visual = vid.visual.to(torch.float16).to(torch.float32)
    audio = vid.audio.to(torch.float16).to(torch.float32)
    port = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
    for spread, peaked in ((True, False), (True, True)):
        sd = synth.seeded_state_dict(spread=spread, peaked=peaked)
        port.load_state_dict(sd)
        m = make_model(spread=spread)
        m.load_state_dict(sd)
        for axis in ("literal", "temporal"):
            want = av_oracle_torch.run_videos(port, [(visual, audio)], axis)[0].numpy()
            got = m(visual[None].cuda(), audio[None].cuda(), attn_axis=axis).cpu().numpy()
            assert rel(got, want) < 1e-3, (peaked, axis, rel(got, want))


def test_fp16_feature_format_matches_reference(cuda_ready):
    """Opt-in 16-bit host feature cache (avs_model_set_feature_format / packed_batches(feature_dtype="fp16")): the
    features travel and are read as IEEE half; scores vs the CPU port ON THE ORIGINAL fp32 FEATURES within 1e-3
    (config 2, first 12 videos, both attention axes), through host (pinned) and device buffers, streamed and not;
    switching back to fp32 buffers restores the default path bit for bit."""
    from avsum_b200.evaluation.summary import summarize_videos
    vids = sorted(synth.config2()[:12], key=lambda v: -v.T)
    port = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
    port.load_state_dict(synth.seeded_state_dict(spread=True))
    for axis, cpu_axis in (("literal_b1", "literal"), ("temporal", "temporal")):
        want = av_oracle_torch.run_videos(port, [(v.visual, v.audio) for v in vids], cpu_axis)
        m = make_model(spread=True, attn_axis=axis)
        base = m.score_videos([(v.visual.cuda(), v.audio.cuda()) for v in vids])
        for space in ("cuda", "cpu"):
            got = m.score_videos([(v.visual.half().to(space), v.audio.half().to(space)) for v in vids])
            worst = max(rel(a.cpu().numpy(), b.numpy()) for a, b in zip(got, want))
            assert worst < 1e-3, (axis, space, worst)
        half_vids = [synth.Video(v.visual.half().pin_memory(), v.audio.half().pin_memory(), v.n_frames, v.positions, v.cps)
                     for v in vids]
        res = summarize_videos(m, half_vids)
        for v, r, w in zip(vids, res, want):
            assert rel(r.scores.numpy(), w.numpy()) < 1e-3
            wp, ws, _ = av_oracle.generate_summary(r.scores.numpy(), v.cps, v.n_frames, v.positions)
            assert np.array_equal(wp, r.picks) and np.array_equal(ws, r.summary)
        again = m.score_videos([(v.visual.cuda(), v.audio.cuda()) for v in vids])
        assert all(torch.equal(a, b) for a, b in zip(again, base))
    with pytest.raises(_cabi.AvsUnsupported):      # bf16 operand mode reads bf16 copies of fp32 features only
        mb = make_model(precision="bf16")
        mb(vids[0].visual.half()[None].cuda(), vids[0].audio.half()[None].cuda())


def test_padded_batch_rows_are_packed_before_the_gemms(cuda_ready):
    """BASELINE configs[2] layout ([B, Tmax] + lengths): the library packs the valid rows densely before the GEMMs
    (csrc/api.cu forward_entry); the result must equal the unpacked path (AVS_NO_ROW_PACKING) bit for bit, in device
    and host space, and padding (NaN) must never reach a valid frame."""
    lens = [300, 12, 700, 129, 0, 64]
    T = max(lens)
    g = torch.Generator().manual_seed(21)
    visual, audio = torch.randn(len(lens), T, 1024, generator=g), torch.randn(len(lens), T, 128, generator=g)
    for b, n in enumerate(lens):
        visual[b, n:] = float("nan")
    m = make_model(spread=True, attn_axis="temporal")
    nat = m.native()
    starts = [b * T for b in range(len(lens))]
    v2, a2 = visual.reshape(-1, 1024), audio.reshape(-1, 128)
    outs = {}
    for space in ("cuda", "cpu"):
        for packing in (True, False):
            if packing:
                os.environ.pop("AVS_NO_ROW_PACKING", None)
            else:
                os.environ["AVS_NO_ROW_PACKING"] = "1"
            try:
                out = nat.forward_rows(v2.to(space), a2.to(space), starts, lens, "temporal", "tf32",
                                       out=torch.zeros(len(lens) * T, device=space))
                torch.cuda.synchronize()
            finally:
                os.environ.pop("AVS_NO_ROW_PACKING", None)
            outs[(space, packing)] = out.cpu().reshape(len(lens), T)
    ref = outs[("cuda", False)]
    for key, out in outs.items():
        for b, n in enumerate(lens):
            assert torch.equal(out[b, :n], ref[b, :n]), (key, b)
            assert bool(torch.isfinite(out[b, :n]).all())
    for b, n in enumerate(lens):       # and each video equals its own B = 1 run
        if n:
            alone = nat.forward_rows(visual[b, :n].cuda(), audio[b, :n].cuda(), [0], [n], "temporal", "tf32")
            assert torch.equal(alone.cpu(), ref[b, :n])


def test_range_check_flags_saturating_features(cuda_ready):
    """AVS_CHECK_RANGE=1: the fp16 activations of the default mode saturate at 65504 instead of overflowing; with the
    switch on, features large enough to saturate the fc outputs fail the call loudly (and bf16, with its fp32
    exponent range, still scores them), normal features pass unchanged."""
    vid = synth.config1()
    m = make_model(spread=True)
    want = m(vid.visual[None].cuda(), vid.audio[None].cuda())
    os.environ["AVS_CHECK_RANGE"] = "1"
    try:
        got = m(vid.visual[None].cuda(), vid.audio[None].cuda())
        assert torch.equal(got, want)
        with pytest.raises(_cabi.AvsUnsupported, match="saturate"):
            m((vid.visual[None] * 3e6).cuda(), vid.audio[None].cuda())
        mb = make_model(spread=True, precision="bf16")
        assert bool(torch.isfinite(mb((vid.visual[None] * 3e6).cuda(), vid.audio[None].cuda())).all())
    finally:
        os.environ.pop("AVS_CHECK_RANGE", None)


def test_training_forward_equals_eval_forward(cuda_ready):
    """One operand policy on both paths: the autograd forward (avs_linear per layer) and the inference forward
    (avs_forward) of the same weights agree to the tolerance of their operand formats (fp32 vs fp16 activations
    between layers), without the systematic shrink a doubly truncated tf32 product would show."""
    vids = [synth.make_video(64, 1024, 128, 300 + i) for i in range(3)]
    visual, audio = torch.stack([v.visual for v in vids]).cuda(), torch.stack([v.audio for v in vids]).cuda()
    m = make_model(spread=True, attn_axis="literal_b1")
    with torch.no_grad():
        ev = m(visual, audio)
    m.train()
    m.visual_fc[2].p = 0.0
    m.audio_fc[2].p = 0.0
    tr = m(visual, audio).detach()
    port = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
    port.load_state_dict(synth.seeded_state_dict(spread=True))
    want = torch.stack(av_oracle_torch.run_videos(port, [(v.visual, v.audio) for v in vids], "literal"))
    assert rel(ev.cpu().numpy(), want.numpy()) < 1e-3 and rel(tr.cpu().numpy(), want.numpy()) < 1e-3
    # signed mean deviation from the reference: no common bias between the two paths beyond noise
    assert abs(float(((tr.cpu() - want) / want).mean())) < 2e-4


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_handles_on_two_devices_in_one_process(cuda_ready, golden_dir):
    """Kernel attributes (opt-in shared memory), SM counts and helper streams are cached PER DEVICE: a second handle on
    another GPU of the same process must run every kernel (process-global flags made it fail)."""
    g = np.load(os.path.join(golden_dir, "config1_spread1.npz"))
    vid = synth.config1()
    vids = synth.config2()[:9]
    outs = []
    for d in (0, 1, 0):
        with torch.cuda.device(d):
            m = make_model(spread=True).to(f"cuda:{d}")
            for axis in ("literal", "temporal"):
                got = m(vid.visual[None].to(f"cuda:{d}"), vid.audio[None].to(f"cuda:{d}"), attn_axis=axis)
                assert got.device.index == d and rel(got.cpu().numpy(), g["scores_" + axis]) < 1e-3
            from avsum_b200.evaluation.summary import summarize_videos
            res = summarize_videos(m, [synth.Video(v.visual.to(f"cuda:{d}"), v.audio.to(f"cuda:{d}"), v.n_frames,
                                                   v.positions, v.cps) for v in vids])
            outs.append(res)
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a.scores, b.scores) and np.array_equal(a.picks, b.picks)


def test_unbatched_input_is_temporal_attention(cuda_ready, golden_dir):
    g = np.load(os.path.join(golden_dir, "config1_spread1.npz"))
    m = make_model(spread=True)
    vid = synth.config1()
    got = m(vid.visual.cuda(), vid.audio.cuda()).cpu().numpy()   # [T, D] -> torch unbatched semantics
    assert rel(got, g["scores_temporal"]) < 1e-3


def test_masked_batch_equals_unbatched(cuda_ready):
    """SURVEY section 4 'masking parity': batched+masked output for video i == B=1 run of video i."""
    m = make_model(spread=True, attn_axis="temporal")
    lens = [57, 130, 1, 88]
    T = max(lens)
    g = torch.Generator().manual_seed(9)
    visual, audio = torch.randn(4, T, 1024, generator=g).cuda(), torch.randn(4, T, 128, generator=g).cuda()
    visual[1, 100:] = float("nan")   # garbage in padding of shorter videos must not leak
    visual[0, 57:] = float("inf")
    lens[1] = 100
    batched = m(visual, audio, lengths=lens)
    assert batched.shape == (4, T)
    for b, n in enumerate(lens):
        alone = m(visual[b:b + 1, :n], audio[b:b + 1, :n]).reshape(-1)
        assert torch.equal(batched[b, :n], alone), b      # same kernels, same data -> bit identical
        assert float(batched[b, n:].abs().sum()) == 0.0
    # literal_b1 accepts ragged lengths too and equals the per-video literal call
    b1 = m(visual, audio, lengths=lens, attn_axis="literal_b1")
    for b, n in enumerate(lens):
        alone = m(visual[b:b + 1, :n], audio[b:b + 1, :n], attn_axis="literal").reshape(-1)
        assert torch.equal(b1[b, :n], alone)
    with pytest.raises(ValueError):
        m(visual, audio, lengths=lens, attn_axis="literal")


def test_config2_full_batch_matches_cpu_port(cuda_ready, golden_dir):
    """BASELINE configs[1] at full size: 50 videos, 21,477 frames, scores vs the torch CPU port
    (bit-identical to the reference, tests/test_oracle_golden.py) and vs the reference goldens."""
    g = np.load(os.path.join(golden_dir, "config2_first4_spread1.npz"))
    vids = synth.config2()
    sd = synth.seeded_state_dict(spread=True)
    port = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
    port.load_state_dict(sd)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for axis, key in (("literal_b1", "scores_literal"), ("temporal", "scores_temporal")):
        m = make_model(spread=True, attn_axis=axis)
        got = m.score_videos([(v.visual, v.audio) for v in vids])            # host in / host out
        got_dev = m.score_videos([(v.visual.cuda(), v.audio.cuda()) for v in vids])
        assert all(torch.equal(a, b.cpu()) for a, b in zip(got, got_dev))
        assert rel(torch.cat(got[:4]).numpy(), g[key]) < 1e-3
        want = av_oracle_torch.run_videos(port, [(v.visual, v.audio) for v in vids],
                                          "literal" if axis == "literal_b1" else "temporal")
        worst = max(rel(a.numpy(), b.numpy()) for a, b in zip(got, want))
        print(f"config 2, spread weights, {axis}: worst relative score error {worst:.2e} (tolerance 1e-3)")
        assert worst < 1e-3, (axis, worst)


def test_config3_bf16_padded_long_videos(cuda_ready):
    """BASELINE configs[2]: SumMe-shaped videos (T in [100, 1000]), bf16 operands with fp32 accumulation,
    padded to Tmax with lengths[B] (variable-length masking), temporal attention -- vs the fp32 CPU port."""
    vids = synth.config3()
    lens = [v.T for v in vids]
    T = max(lens)
    visual, audio = torch.zeros(len(vids), T, 1024), torch.zeros(len(vids), T, 128)
    for b, v in enumerate(vids):
        visual[b, :v.T], audio[b, :v.T] = v.visual, v.audio
        visual[b, v.T:] = float("nan")      # padding must never reach a valid frame
    sd = synth.seeded_state_dict(spread=True)
    port = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
    port.load_state_dict(sd)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    want = av_oracle_torch.run_videos(port, [(v.visual, v.audio) for v in vids], "temporal")
    for prec in ("bf16", "tf32"):
        m = make_model(spread=True, attn_axis="temporal", precision=prec)
        got = m(visual.cuda(), audio.cuda(), lengths=lens).cpu()
        assert got.shape == (len(vids), T)
        worst = max(rel(got[b, :n].numpy(), want[b].numpy()) for b, n in enumerate(lens))
        assert worst < REL_TOL[prec], (prec, worst)
        assert all(float(got[b, n:].abs().sum()) == 0.0 for b, n in enumerate(lens))


# ------------------------------------------------------------------ MultiHeadSelfAttention drop-in
def test_mhsa_matches_reference(cuda_ready, golden_dir):
    from avsum_b200.models.attention import MultiHeadSelfAttention
    g = np.load(os.path.join(golden_dir, "mhsa_E1024_H4.npz"))
    torch.manual_seed(int(g["seed_w"]))
    att = MultiHeadSelfAttention(1024, 4).eval()
    assert abs(synth.state_dict_checksum(att.state_dict()) - float(g["weights_checksum"])) < 1e-6
    x = torch.randn(2, 33, 1024, generator=torch.Generator().manual_seed(int(g["seed_in"])))
    for prec, tol in (("fp32_simt", 5e-6), ("tf32", 2e-3), ("bf16", 2e-2)):
        att.precision = prec
        y = att.cuda()(x.cuda()).cpu().numpy()
        assert y.shape == (2, 33, 1024)
        assert float(np.max(np.abs(y[:, :, ::8] - g["out"]))) < tol * float(np.max(np.abs(g["out"])))


# ------------------------------------------------------------------ K7/K8: pooling + knapsack (bit-exact)
def _summarize_and_compare(native, vids, scores, space):
    lens = [v.T for v in vids]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    pos = np.concatenate([v.positions for v in vids]).astype(np.int32)
    picks, seg_mean, summary, cps_start, sum_start = native.summarize_rows(
        torch.from_numpy(scores).to(space), torch.from_numpy(pos).to(space), starts, lens,
        [v.n_frames for v in vids], [v.cps for v in vids], 0.15)
    torch.cuda.synchronize()
    picks, seg_mean, summary = picks.cpu().numpy(), seg_mean.cpu().numpy(), summary.cpu().numpy()
    for i, v in enumerate(vids):
        wp, ws, wm = av_oracle.generate_summary(scores[starts[i]:starts[i] + lens[i]], v.cps, v.n_frames, v.positions)
        assert np.array_equal(wm, seg_mean[cps_start[i]:cps_start[i + 1]]), i
        assert np.array_equal(wp, picks[cps_start[i]:cps_start[i + 1]]), i
        assert np.array_equal(ws, summary[sum_start[i]:sum_start[i + 1]]), i
        nf = v.cps[:, 1] - v.cps[:, 0] + 1
        assert int(ws.sum()) == int((wp * nf).sum()) <= (v.n_frames * 15) // 100
    return picks


@pytest.mark.parametrize("kind", ["uniform", "near_tie", "spiky", "constant", "zeros"])
@pytest.mark.parametrize("space", ["cuda", "cpu"])
def test_summarize_config2_bit_exact(native, kind, space):
    vids = synth.config2()
    n = sum(v.T for v in vids)
    rng = np.random.default_rng(11)
    scores = {"uniform": rng.random(n), "near_tie": 0.52 + 1e-4 * rng.standard_normal(n),
              "spiky": rng.random(n) ** 8, "constant": np.full(n, 0.5), "zeros": np.zeros(n)}[kind].astype(np.float32)
    _summarize_and_compare(native, vids, scores, space)


def test_summarize_edge_cases(native):
    rng = np.random.default_rng(5)
    # single-shot video, a one-frame video, irregular positions, shots not covering the tail
    v1 = synth.Video(torch.zeros(1, 1), torch.zeros(1, 1), 40, np.array([0], np.int32), np.array([[0, 39]], np.int32))
    v2 = synth.Video(torch.zeros(7, 1), torch.zeros(7, 1), 100, np.array([3, 10, 11, 40, 41, 90, 99], np.int32),
                     np.array([[0, 4], [5, 5], [6, 50], [60, 98]], np.int32))
    v3 = synth.Video(torch.zeros(3, 1), torch.zeros(3, 1), 6, np.array([0, 2, 4], np.int32),
                     np.array([[0, 0], [1, 1], [2, 2], [3, 3], [4, 4], [5, 5]], np.int32))
    vids = [v1, v2, v3]
    scores = rng.random(sum(v.T for v in vids)).astype(np.float32)
    scores[0] = np.nan
    scores[3] = 7.0      # out-of-range scores are clamped by the quantiser
    scores[4] = -1.0
    _summarize_and_compare(native, vids, scores, "cuda")
    # invalid change points are rejected, not mis-computed
    bad = synth.Video(torch.zeros(2, 1), torch.zeros(2, 1), 10, np.array([0, 5], np.int32), np.array([[0, 6], [5, 9]], np.int32))
    with pytest.raises(ValueError):
        native.summarize_rows(torch.zeros(2, device="cuda"), torch.tensor([0, 5], dtype=torch.int32, device="cuda"),
                              [0], [2], [10], [bad.cps], 0.15)


def test_summarize_long_video_uses_global_dp_rows(native):
    """T = 8192 (BASELINE configs[3]): capacity 18,432 exceeds the shared-memory DP rows."""
    vids = [synth.make_video(8192, 4, 4, 9000), synth.make_video(300, 4, 4, 9001)]
    scores = np.random.default_rng(2).random(sum(v.T for v in vids)).astype(np.float32)
    picks = _summarize_and_compare(native, vids, scores, "cuda")
    assert picks.sum() > 0


def test_summarize_capacity_beyond_every_on_chip_path(native):
    """A capacity no on-chip variant covers (60,000 cells: the cluster kernel stops at 24,575, the one-SM register
    kernel at 20,480): DP rows and keep bits in the global workspace.  Bit-exact against the oracle, next to a short
    video in the same batch."""
    rng = np.random.default_rng(77)
    vids, scores = [], []
    for nf, n_shots in ((400000, 300), (900, 12)):
        shots, f = [], 0
        while f < nf and len(shots) < n_shots:
            end = min(f + int(rng.integers(1, 3000 if nf > 1000 else 120)) - 1, nf - 1)
            shots.append((f, end))
            f = end + 1 + int(rng.integers(0, 50))
        T = min(nf, 500)
        pos = np.sort(rng.choice(nf, size=T, replace=False)).astype(np.int32)
        vids.append(synth.Video(torch.zeros(T, 1), torch.zeros(T, 1), nf, pos, np.asarray(shots, np.int32)))
        scores.append(rng.random(T).astype(np.float32))
    picks = _summarize_and_compare(native, vids, np.concatenate(scores), "cuda")
    assert picks.sum() > 0


def test_generate_summary_single_video_api(cuda_ready):
    from avsum_b200.evaluation.summary import generate_summary
    m = make_model()
    v = synth.config2()[3]
    sc = np.random.default_rng(1).random(v.T).astype(np.float32)
    got = generate_summary(m, sc, v.cps, v.n_frames, v.positions, 0.15)
    assert np.array_equal(got, av_oracle.generate_summary(sc, v.cps, v.n_frames, v.positions)[1])


def test_end_to_end_keyshots_match_reference_scores(cuda_ready):
    """Identical keyshot summaries to the reference pipeline: reference scores (CPU port) ->
    oracle summary  vs  GPU scores -> GPU summary, on the spread weight set."""
    from avsum_b200.evaluation.summary import summarize_videos
    vids = synth.config2()[:12]
    sd = synth.seeded_state_dict(spread=True)
    port = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
    port.load_state_dict(sd)
    ref_scores = av_oracle_torch.run_videos(port, [(v.visual, v.audio) for v in vids])
    m = make_model(spread=True)
    res = summarize_videos(m, vids)
    same = 0
    for v, r, rs in zip(vids, res, ref_scores):
        # on identical inputs (the GPU's own scores) the selection is bit-exact ...
        wp, ws, _ = av_oracle.generate_summary(r.scores.numpy(), v.cps, v.n_frames, v.positions)
        assert np.array_equal(wp, r.picks) and np.array_equal(ws, r.summary)
        # ... and it agrees with the selection made from the reference's fp32 scores
        same += int(np.array_equal(av_oracle.generate_summary(rs.numpy(), v.cps, v.n_frames, v.positions)[0], r.picks))
    assert same == len(vids)


# ------------------------------------------------------------------ overlap F1 (bit-exact doubles)
def test_temporal_f1_bit_exact(cuda_ready, golden_dir):
    from avsum_b200.evaluation.metrics import compute_temporal_f1, compute_temporal_f1_batch
    from avsum_b200.utils.shot_metrics import compute_f1
    g = np.load(os.path.join(golden_dir, "helpers.npz"))
    pred = [tuple(int(x) for x in p) for p in g["pred"]]
    gt = [tuple(int(x) for x in p) for p in g["gt"]]
    assert compute_temporal_f1(pred, gt, 1000) == float(g["f1_metrics"])
    assert compute_f1(pred, gt, 1000) == float(g["f1_shot"])
    rng = np.random.default_rng(3)
    preds, gts = [], []
    for _ in range(60):
        def shots():
            e = np.sort(rng.choice(5000, size=2 * int(rng.integers(1, 30)), replace=False))
            return [(int(e[2 * i]), int(e[2 * i + 1])) for i in range(len(e) // 2)]
        preds.append(shots())
        gts.append(shots())
    got = compute_temporal_f1_batch(preds, gts)
    want = np.asarray([av_oracle.temporal_f1(p, q) for p, q in zip(preds, gts)])
    assert np.array_equal(got, want)
    with pytest.raises(ZeroDivisionError):
        compute_temporal_f1([], gt, 1000)


# ------------------------------------------------------------------ ragged / empty inputs
def test_empty_and_zero_length_videos(native):
    out = native.forward_rows(torch.zeros(0, 1024, device="cuda"), torch.zeros(0, 128, device="cuda"), [], [])
    assert out.numel() == 0
    v = torch.randn(10, 1024, device="cuda")
    a = torch.randn(10, 128, device="cuda")
    got = native.forward_rows(v, a, [0, 4, 4], [4, 0, 6], "temporal")
    a1 = native.forward_rows(v[:4], a[:4], [0], [4], "temporal")
    a2 = native.forward_rows(v[4:], a[4:], [0], [6], "temporal")
    assert torch.equal(got[:4], a1) and torch.equal(got[4:], a2)
    with pytest.raises(ValueError):
        native.forward_rows(v, a, [0, 8], [4, 6], "temporal")      # rows outside the buffer


# ------------------------------------------------------------------ evaluate() metric block (a13)
def _metric_case(rng, n, kind, tgt_dtype):
    if kind == "ties":
        pred = rng.integers(0, 6, n).astype(np.float32) / 5
        target = rng.integers(0, 4, n).astype(tgt_dtype) / 3
    elif kind == "constant":
        pred = np.full(n, 0.5, np.float32)
        target = rng.random(n).astype(tgt_dtype)
    elif kind == "nan":
        pred = rng.random(n).astype(np.float32)
        target = rng.random(n).astype(tgt_dtype)
        if n:
            pred[n // 2] = np.nan
    else:
        pred = rng.random(n).astype(np.float32)
        target = (0.3 * pred + rng.random(n)).astype(tgt_dtype)
    return pred, target


@pytest.mark.parametrize("space", ["cuda", "cpu"])
@pytest.mark.parametrize("tgt_dtype", [np.float32, np.float64])
def test_eval_metrics_match_numpy_scipy(cuda_ready, space, tgt_dtype):
    """scripts/evaluate.py:25-36 per video: F1 and Kendall's tau bit-exact against numpy / scipy (the calls the
    reference makes), Spearman's rho to 1e-13, integer pair counts exact, np.mean(pred) bit-exact."""
    rng = np.random.default_rng(7)
    cases = [(320, "random"), (700, "random"), (57, "ties"), (129, "ties"), (1, "random"), (2, "random"),
             (40, "constant"), (33, "nan"), (2049, "random"), (0, "random")]
    preds, targets = zip(*[_metric_case(rng, n, k, tgt_dtype) for n, k in cases])
    lens = [len(p) for p in preds]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    metrics, counts = runtime.eval_metrics_rows(torch.from_numpy(np.concatenate(preds)).to(space),
                                                torch.from_numpy(np.concatenate(targets)).to(space), starts, lens)
    same = lambda a, b: (a == b) or (np.isnan(a) and np.isnan(b))
    for i, (p, t) in enumerate(zip(preds, targets)):
        if len(p) == 0:
            assert np.isnan(metrics[i, 0]) and counts[i, 7] == 0
            continue
        f1, rho, tau = av_oracle.eval_metrics(p, t)
        assert same(metrics[i, 0], f1), (i, metrics[i, 0], f1)
        tau_gpu = np.float32(metrics[i, 2]) if tgt_dtype == np.float32 else metrics[i, 2]
        assert same(tau_gpu, tau), (i, tau_gpu, tau)     # scipy rounds once to float32 for float32 inputs
        assert same(metrics[i, 1], rho) or abs(metrics[i, 1] - rho) < 1e-13, (i, metrics[i, 1], rho)
        assert np.float32(metrics[i, 3]) == np.mean(p) or np.isnan(np.mean(p))
        if not np.isnan(p).any():
            dis, xt, yt, nt, tot = av_oracle.kendall_counts(p, t)
            assert list(counts[i, 3:8]) == [dis, xt, yt, nt, len(p)]
            assert counts[i, 1] == int((p > np.mean(p)).sum()) and counts[i, 2] == int((t > np.mean(t)).sum())


def test_evaluate_drop_in_matches_reference_loop(cuda_ready):
    """evaluate(model, dataset) (scripts/evaluate.py:6-42) as one packed batch == the reference's per-video loop
    applied to the same scores."""
    from avsum_b200.scripts.evaluate import evaluate
    vids = synth.config2()[:9]
    rng = np.random.default_rng(5)
    dataset = [({"visual": v.visual, "audio": v.audio}, torch.from_numpy(rng.random(v.T).astype(np.float32)))
               for v in vids]
    m = make_model(spread=True)
    got, metrics, _ = evaluate(m, dataset, return_per_video=True)
    scores = m.score_videos([(v.visual.cuda(), v.audio.cuda()) for v in vids])
    f1s, rhos, taus = zip(*[av_oracle.eval_metrics(s.cpu().numpy(), t.numpy()) for s, (_, t) in zip(scores, dataset)])
    from scipy.stats import kendalltau
    taus32 = [kendalltau(s.cpu().numpy(), t.numpy()).correlation for s, (_, t) in zip(scores, dataset)]
    assert got["f1"] == np.mean(f1s)
    assert got["kendall"] == np.mean(taus32) and got["kendall"].dtype == np.float32
    assert abs(got["spearman"] - np.mean(rhos)) < 1e-13
    assert set(got) == {"f1", "spearman", "kendall"}
    # and against the reference scores (CPU port): same metrics within the score tolerance
    port = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
    port.load_state_dict(synth.seeded_state_dict(spread=True))
    ref_scores = av_oracle_torch.run_videos(port, [(v.visual, v.audio) for v in vids])
    ref = [av_oracle.eval_metrics(s.numpy(), t.numpy()) for s, (_, t) in zip(ref_scores, dataset)]
    assert abs(got["spearman"] - np.mean([r[1] for r in ref])) < 5e-3
    assert abs(float(got["kendall"]) - np.mean([r[2] for r in ref])) < 5e-3


def test_evaluate_on_a_device_resident_dataset(cuda_ready):
    """data.dataset.DeviceDataset: the dataset packed once into device memory gives the same metrics as the per-call
    upload path, bit for bit (same kernels, same data); the fp16 feature cache stays within the score tolerance."""
    from avsum_b200.scripts.evaluate import evaluate
    from avsum_b200.data.dataset import DeviceDataset
    vids = synth.config2()[:7]
    rng = np.random.default_rng(6)
    dataset = [({"visual": v.visual, "audio": v.audio}, torch.from_numpy(rng.random(v.T).astype(np.float32)))
               for v in vids]
    m = make_model(spread=True)
    want = evaluate(m, dataset)
    cached = DeviceDataset(dataset)
    assert len(cached) == 7 and cached.visual.is_cuda and cached[2][0]["visual"].shape == (vids[2].T, 1024)
    for _ in range(2):
        got = evaluate(m, cached)
        assert got["f1"] == want["f1"] and got["spearman"] == want["spearman"] and got["kendall"] == want["kendall"]
    half = evaluate(m, DeviceDataset(dataset, feature_dtype="fp16"))
    assert abs(half["spearman"] - want["spearman"]) < 5e-3 and abs(float(half["kendall"]) - float(want["kendall"])) < 5e-3


# ------------------------------------------------------------------ features/fusion.py helpers (a9-a11)
def test_fusion_helpers_match_reference(cuda_ready, golden_dir):
    from avsum_b200.features import fusion
    g = np.load(os.path.join(golden_dir, "helpers.npz"))
    gen = torch.Generator().manual_seed(11)
    fv, fa = torch.randn(23, 64, generator=gen), torch.randn(31, 64, generator=gen)
    for dev in ("cpu", "cuda"):
        dtw = fusion.compute_dtw(fv.to(dev), fa.to(dev)) if dev == "cpu" else runtime.cdist_euclidean(fv.cuda(), fa.cuda())
        assert dtw.dtype == np.float64 and np.array_equal(dtw, g["dtw"])          # bit-exact vs the reference
        interp = fusion.interpolate_features(fv.to(dev), g["path"], 20)
        assert isinstance(interp, torch.Tensor) and np.array_equal(interp.cpu().numpy(), g["interp"])
    with pytest.raises(ValueError):                                               # 1024-d vs 128-d
        fusion.compute_dtw(torch.randn(5, 1024), torch.randn(5, 128))
    with pytest.raises(TypeError):                                                # the reference's broken call (a10)
        fusion.compute_optimal_path(g["dtw"])
    # larger, odd-sized case against the oracle restatement (summation order matters at D = 1000)
    a, b = torch.randn(70, 1000, generator=gen), torch.randn(33, 1000, generator=gen)
    assert np.array_equal(fusion.compute_dtw(a, b), av_oracle.cdist_euclidean(a.numpy(), b.numpy()))


def test_exact_dtw_path_matches_oracle(cuda_ready):
    from avsum_b200.features import fusion
    rng = np.random.default_rng(4)
    for n, m, ties in [(1, 1, False), (1, 9, False), (8, 1, False), (23, 31, False), (40, 40, True), (130, 77, True)]:
        cost = rng.integers(0, 5, (n, m)).astype(np.float64) if ties else rng.random((n, m))
        total, path = av_oracle.dtw_path(cost)
        got = fusion.compute_optimal_path(cost, exact=True)
        assert np.array_equal(got, path), (n, m)
        t2, _ = runtime.dtw_path(torch.from_numpy(cost).cuda())
        assert t2 == total


# ------------------------------------------------------------------ training step (BASELINE configs[4])
def _grad_err(got, want):
    return float((got.double().cpu() - want.double()).abs().max() / want.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape", [(300, 512, 128), (2560, 1024, 1024), (77, 64, 1024), (129, 2048, 512), (40, 512, 296)])
def test_linear_backward_matches_autograd(cuda_ready, shape):
    from avsum_b200 import training
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(N, K, generator=g, dtype=torch.float64) / K ** 0.5).requires_grad_()
    b = torch.randn(N, generator=g, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(M, N, generator=g, dtype=torch.float64)
    torch.nn.functional.linear(x, w, b).backward(dy)
    xc, wc, bc = (t.detach().float().cuda().requires_grad_() for t in (x, w, b))
    y = training.linear(xc, wc, bc)
    y.backward(dy.float().cuda())
    assert _grad_err(xc.grad, x.grad) < 3e-3 and _grad_err(wc.grad, w.grad) < 3e-3 and _grad_err(bc.grad, b.grad) < 1e-5


# videos per cluster of the tcgen05 BPTT kernel: 1 ([37]), 2 (four ragged videos), 4 (11 videos, 3 groups; 8 x 320 = the
# config-5 shape), 8 (13 ragged videos: the second group has empty slots)
@pytest.mark.parametrize("lens", [[37], [5, 64, 1, 33], [20] * 11, [50, 3, 41, 17, 29, 8, 33, 21, 12, 45, 6, 38, 27],
                                  [320] * 8])
def test_bilstm_pair_backward_matches_torch_lstm(native, lens):
    """BPTT kernel + GEMMs vs torch.nn.LSTM autograd on the CPU (fp32): d_emb, dW_ih, dW_hh, db per recurrence."""
    from avsum_b200 import training
    sd = synth.seeded_state_dict()
    R = sum(lens)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    g = torch.Generator().manual_seed(len(lens) + 100)
    v = torch.randn(R, 512, generator=g).requires_grad_()
    a = torch.randn(R, 512, generator=g).requires_grad_()
    dfused = torch.randn(R, 1024, generator=g)
    lv = torch.nn.LSTM(512, 256, bidirectional=True, batch_first=True)
    la = torch.nn.LSTM(512, 256, bidirectional=True, batch_first=True)
    lv.load_state_dict({k.split(".", 1)[1]: t for k, t in sd.items() if k.startswith("visual_bilstm")})
    la.load_state_dict({k.split(".", 1)[1]: t for k, t in sd.items() if k.startswith("audio_bilstm")})
    outs = []
    for s, n in zip(starts, lens):
        outs.append(torch.cat([lv(v[None, s:s + n])[0][0], la(a[None, s:s + n])[0][0]], dim=1))
    torch.cat(outs).backward(dfused)
    vc, ac = v.detach().cuda().requires_grad_(), a.detach().cuda().requires_grad_()
    weights = []
    for mod in (lv, la):
        for suf in ("", "_reverse"):
            weights += [getattr(mod, f"{n}_l0{suf}").detach().cuda().requires_grad_()
                        for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    fused = training.bilstm_pair(vc, ac, native, starts, lens, weights)
    assert float((fused.cpu() - torch.cat(outs).detach()).abs().max()) < 3e-3
    fused.backward(dfused.cuda())
    assert _grad_err(vc.grad, v.grad) < 2e-3 and _grad_err(ac.grad, a.grad) < 2e-3   # measured 3.5e-4 .. 4.7e-4
    i = 0
    worst = 0.0
    for mod in (lv, la):
        for suf in ("", "_reverse"):
            for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                want = getattr(mod, f"{n}_l0{suf}").grad
                assert _grad_err(weights[i].grad, want) < 2e-3, (n, suf, _grad_err(weights[i].grad, want))   # measured <= 6.7e-4
                worst = max(worst, _grad_err(weights[i].grad, want))
                i += 1
    print(f"bptt parity lens={lens[:4]}..x{len(lens)}: d_emb {_grad_err(vc.grad, v.grad):.2e} / "
          f"{_grad_err(ac.grad, a.grad):.2e}, worst weight gradient {worst:.2e} (max-norm relative)")
    # the reduce-scatter of the partial dh adds the eight partials in a fixed order: a second pass gives the same bits
    first = vc.grad.clone()
    vc.grad = None
    training.bilstm_pair(vc, ac, native, starts, lens, weights).backward(dfused.cuda())
    assert torch.equal(vc.grad, first)


def test_training_step_matches_reference_autograd(cuda_ready):
    """One step of scripts/train_av_model.py:86-96 on a batch of B = 1 samples: loss, all 28 gradients and the
    AdamW update vs the reference's torch CPU path (dropout p = 0 so both sides are deterministic)."""
    vids = [synth.make_video(t, 1024, 128, 4000 + i) for i, t in enumerate([48, 48, 48])]
    visual = torch.stack([v.visual for v in vids])
    audio = torch.stack([v.audio for v in vids])
    target = torch.rand(3, 48, generator=torch.Generator().manual_seed(1))
    sd = synth.seeded_state_dict(spread=True)
    port = av_oracle_torch.RefPortModel(1024, 128, 512).train()
    port.load_state_dict(sd)
    for seq in (port.visual_fc, port.audio_fc):
        seq[2].p = 0.0
    opt_ref = torch.optim.AdamW(port.parameters(), lr=1e-4)
    preds = torch.stack([port(visual[b:b + 1], audio[b:b + 1], "literal") for b in range(3)])   # B = 1 per sample
    loss_ref = torch.nn.functional.mse_loss(preds, target)
    opt_ref.zero_grad()
    loss_ref.backward()

    m = make_model(spread=True, attn_axis="literal_b1").train()
    m.visual_fc[2].p = 0.0
    m.audio_fc[2].p = 0.0
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
    out = m(visual.cuda(), audio.cuda())
    assert out.shape == (3, 48) and out.requires_grad
    loss = torch.nn.functional.mse_loss(out, target.cuda())
    opt.zero_grad()
    loss.backward()
    assert abs(float(loss) - float(loss_ref)) < 1e-3 * abs(float(loss_ref))
    ref_grads = dict(port.named_parameters())
    worst = {}
    for name, p in m.named_parameters():
        want = ref_grads[name].grad
        assert p.grad is not None and p.grad.shape == want.shape, name
        if float(want.abs().max()) == 0.0:
            assert float(p.grad.abs().max()) == 0.0, name      # q / k thirds of in_proj get exactly zero
            continue
        if name.startswith("attention.in_proj"):
            E = 1024
            assert float(p.grad[:2 * E].abs().max()) == 0.0
        l2 = float((p.grad.double().cpu() - want.double()).norm() / want.double().norm())
        worst[name] = (round(_grad_err(p.grad, want), 4), round(l2, 4))
    print(worst)
    # Stated tolerance: the forward runs on 11-bit-significand operands (tf32 / fp16), so a handful of ReLU
    # pre-activations that the fp32 reference has within ~1e-3 of zero land on the other side; each such unit
    # switches a whole gradient path on or off (the x50 "spread" head makes those paths large).  The per-layer
    # backward kernels are checked tightly above (3e-3 linear, 1e-2 BiLSTM); end to end we require 6 % in the
    # max norm and 4 % in the L2 norm for every one of the 28 gradients.
    assert max(v[0] for v in worst.values()) < 6e-2 and max(v[1] for v in worst.values()) < 4e-2, worst
    opt.step()
    opt_ref.step()
    for name, p in m.named_parameters():
        assert float((p.detach().cpu() - ref_grads[name].detach()).abs().max()) < 2.5e-4, name   # lr-sized updates
    # the updated parameters are re-packed for the next forward
    out2 = m(visual.cuda(), audio.cuda())
    assert float((out2 - out).abs().max()) > 0


def test_training_gradients_match_reference_with_shared_relu_masks(cuda_ready):
    """The end-to-end gradient check above has to allow for ReLU units that the 11-bit-significand forward puts on the
    other side of zero (whole gradient paths switch), which makes it too loose to catch a wrong small term.  Here the
    reference (fp64 torch autograd over the reference's layers, B = 1 per sample) is evaluated with the SAME three ReLU
    masks the GPU forward produced, so both sides differentiate the same piecewise-linear function and only rounding is
    left: every one of the 28 gradients must agree to 3e-3 in the max norm and 2e-3 in the L2 norm (measured: <= 1.0e-3 and
    <= 7.2e-4)."""
    B, T = 3, 64
    vids = [synth.make_video(T, 1024, 128, 5000 + i) for i in range(B)]
    visual = torch.stack([v.visual for v in vids])
    audio = torch.stack([v.audio for v in vids])
    target = torch.rand(B, T, generator=torch.Generator().manual_seed(2))
    sd = synth.seeded_state_dict(spread=True)
    m = make_model(spread=True, attn_axis="literal_b1").train()
    m.visual_fc[2].p = 0.0
    m.audio_fc[2].p = 0.0
    masks, orig_relu = [], torch.relu

    def spy(x):
        y = orig_relu(x)
        masks.append((y.detach() > 0).cpu())
        return y

    torch.relu = spy          # models/av_model.py::_forward_train calls torch.relu: audio fc, visual fc, scorer.0
    try:
        out = m(visual.cuda(), audio.cuda())
    finally:
        torch.relu = orig_relu
    assert len(masks) == 3 and masks[0].shape == (B * T, 512) and masks[2].shape == (B * T, 64)
    torch.nn.functional.mse_loss(out, target.cuda()).backward()

    port = av_oracle_torch.RefPortModel(1024, 128, 512).double()
    port.load_state_dict({k: v.double() for k, v in sd.items()})
    ma, mv, mh = (x.double() for x in masks)
    E = 1024
    preds = []
    for b in range(B):                      # B = 1 per sample, as scripts/train_av_model.py:86-88
        rows = slice(b * T, (b + 1) * T)
        v = port.visual_fc[0](visual[b:b + 1].double()) * mv[rows]
        a = port.audio_fc[0](audio[b:b + 1].double()) * ma[rows]
        fused = torch.cat([port.visual_bilstm(v)[0], port.audio_bilstm(a)[0]], dim=-1)
        # L = 1: softmax weight 1, attention == out_proj(value projection)  (av_model.py:44 with B = 1)
        ctx = torch.nn.functional.linear(fused, port.attention.in_proj_weight[2 * E:], port.attention.in_proj_bias[2 * E:])
        y = port.attention.out_proj(ctx)
        h1 = port.scorer[0](y) * mh[rows]
        preds.append(torch.sigmoid(port.scorer[2](h1)).reshape(T))
    torch.nn.functional.mse_loss(torch.stack(preds), target.double()).backward()
    ref = dict(port.named_parameters())
    worst = {}
    for name, p in m.named_parameters():
        want = ref[name].grad
        if name.startswith("attention.in_proj"):
            assert float(p.grad[:2 * E].abs().max()) == 0.0          # q / k thirds: exactly zero
            got, want = p.grad[2 * E:], (want[2 * E:] if want is not None else None)
        else:
            got = p.grad
        assert want is not None and float(want.abs().max()) > 0.0, name
        l2 = float((got.double().cpu() - want).norm() / want.norm())
        worst[name] = (_grad_err(got, want), l2)
    print({k: (round(a, 5), round(b, 5)) for k, (a, b) in worst.items()})
    assert max(v[0] for v in worst.values()) < 3e-3 and max(v[1] for v in worst.values()) < 2e-3, worst


def test_graphed_training_step_equals_eager_steps(cuda_ready):
    """training.GraphedTrainStep (forward + loss + backward + AdamW captured in one CUDA graph, replayed per batch)
    follows the same parameter trajectory as the eager loop of scripts/train_av_model.py:86-96: the losses of five
    consecutive steps on changing batches and the final parameters agree (dropout p = 0 on both sides)."""
    from avsum_b200 import training
    g = torch.Generator().manual_seed(7)
    batches = [(torch.randn(4, 64, 1024, generator=g).cuda(), torch.randn(4, 64, 128, generator=g).cuda(),
                torch.rand(4, 64, generator=g).cuda()) for _ in range(5)]
    losses, finals = {}, {}
    for mode in ("eager", "graph"):
        m = make_model(attn_axis="literal_b1").train()   # default head: scores near 0.52, the loss follows every update
        m.visual_fc[2].p = 0.0
        m.audio_fc[2].p = 0.0
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3, capturable=True)
        out = []
        if mode == "graph":
            # the constructor runs 3 eager warm-up steps on the first batch (capturing itself executes nothing)
            step = training.GraphedTrainStep(m, opt, torch.nn.functional.mse_loss, *batches[0], warmup=3)
            for v, a, t in batches:
                out.append(float(step(v, a, t)))
        else:
            def eager(v, a, t):
                opt.zero_grad(set_to_none=True)
                loss = torch.nn.functional.mse_loss(m(v, a), t)
                loss.backward()
                opt.step()
                return float(loss.detach())
            for _ in range(3):
                eager(*batches[0])
            for v, a, t in batches:
                out.append(eager(v, a, t))
        losses[mode] = out
        finals[mode] = {k: p.detach().clone() for k, p in m.named_parameters()}
    print(losses)
    assert len(set(losses["eager"])) == 5, "the loss must move with the updates for this comparison to mean anything"
    for le, lg in zip(losses["eager"], losses["graph"]):
        assert abs(le - lg) <= 1e-5 * abs(le), losses
    for k in finals["eager"]:
        assert float((finals["eager"][k] - finals["graph"][k]).abs().max()) <= 1e-5, k
    # and the handle follows the graph-updated parameters back to eval
    m.eval()
    with torch.no_grad():
        s = m(batches[0][0], batches[0][1])
    assert s.shape == (4, 64) and bool(torch.isfinite(s).all())


def test_training_rejects_unsupported_attention(cuda_ready):
    m = make_model(attn_axis="temporal").train()
    with pytest.raises(NotImplementedError):
        m(torch.randn(2, 8, 1024).cuda(), torch.randn(2, 8, 128).cuda())
    m = make_model(attn_axis="literal").train()
    with pytest.raises(NotImplementedError):
        m(torch.randn(2, 8, 1024).cuda(), torch.randn(2, 8, 128).cuda())
    out = m(torch.randn(1, 8, 1024).cuda(), torch.randn(1, 8, 128).cuda())      # the reference's B = 1 step
    assert out.shape == (8,) and out.requires_grad


def test_fused_score_and_summarize_equals_two_calls(native):
    """avs_forward_summarize (host space: pipelined by video group, one sync) == avs_forward + avs_summarize."""
    vids = sorted(synth.config2()[:30], key=lambda v: -v.T)
    lens = [v.T for v in vids]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    visual, audio = torch.cat([v.visual for v in vids]), torch.cat([v.audio for v in vids])
    pos = torch.from_numpy(np.concatenate([v.positions for v in vids]).astype(np.int32))
    nf, cps = [v.n_frames for v in vids], [v.cps for v in vids]
    for axis in ("literal_b1", "temporal"):
        want_s = native.forward_rows(visual.cuda(), audio.cuda(), starts, lens, axis)
        want = native.summarize_rows(want_s, pos.cuda(), starts, lens, nf, cps, 0.15)
        for dev in ("cpu", "cuda"):
            v, a, p = (t.pin_memory() if dev == "cpu" else t.cuda() for t in (visual, audio, pos))
            got = native.score_and_summarize_rows(v, a, p, starts, lens, nf, cps, 0.15, axis)
            torch.cuda.synchronize()
            assert torch.equal(got[0].cpu(), want_s.cpu()), (axis, dev)        # same kernels, same data
            for g, w in zip(got[1:4], want[:3]):
                assert torch.equal(g.cpu(), w.cpu()), (axis, dev)


def test_sharded_result_equals_single_gpu_result(native):
    """SURVEY section 4 'multi-GPU': the N-shard result equals the 1-GPU result EXACTLY (same kernels, sharding
    only).  The shards of sharding.shard_videos are run one after the other on this GPU and re-assembled."""
    from avsum_b200 import sharding
    vids = synth.config2()[:18]
    lens_all = [v.T for v in vids]

    def run(idx):
        sub = [vids[i] for i in idx]
        lens = [v.T for v in sub]
        starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
        pos = torch.from_numpy(np.concatenate([v.positions for v in sub]).astype(np.int32)).cuda()
        sc, picks, segm, summ, cps_start, sum_start = native.score_and_summarize_rows(
            torch.cat([v.visual for v in sub]).cuda(), torch.cat([v.audio for v in sub]).cuda(), pos, starts, lens,
            [v.n_frames for v in sub], [v.cps for v in sub], 0.15, "temporal")
        torch.cuda.synchronize()
        out = {}
        for k, i in enumerate(idx):
            out[i] = (sc[starts[k]:starts[k] + lens[k]].cpu(), picks[cps_start[k]:cps_start[k + 1]].cpu(),
                      summ[sum_start[k]:sum_start[k + 1]].cpu())
        return out

    whole = run(list(range(len(vids))))
    for world in (2, 4):
        merged = {}
        for shard in sharding.shard_videos(lens_all, world):
            merged.update(run(shard))
        assert sorted(merged) == list(range(len(vids)))
        for i in range(len(vids)):
            for a, b in zip(merged[i], whole[i]):
                assert torch.equal(a, b), (world, i)


def test_unaligned_device_features_are_staged(native):
    """A raw C-ABI caller may hand over device buffers that are only 4-byte aligned; the TMA-fed GEMMs need 16."""
    v = synth.config1()
    big_v = torch.zeros(320 * 1024 + 1, device="cuda")
    big_a = torch.zeros(320 * 128 + 1, device="cuda")
    big_v[1:] = v.visual.cuda().reshape(-1)
    big_a[1:] = v.audio.cuda().reshape(-1)
    want = native.forward_rows(v.visual.cuda(), v.audio.cuda(), [0], [320], "temporal")
    out = torch.empty(320, device="cuda")
    rs, ln = np.zeros(1, np.int32), np.full(1, 320, np.int32)
    _cabi.check(native.lib.avs_forward(native._handle, C.c_void_p(big_v.data_ptr() + 4), C.c_void_p(big_a.data_ptr() + 4),
                                       320, 1, _cabi.np_ptr(rs), _cabi.np_ptr(ln), _cabi.AVS_ATTN_TEMPORAL, _cabi.AVS_PREC_TF32,
                                       C.c_void_p(out.data_ptr()), _cabi.AVS_DEVICE,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert torch.equal(out, want)


@pytest.mark.parametrize("prop", [(15, 100), (1, 2), (0, 1), (1, 1), (3, 7)])
def test_summarize_randomised_batches_bit_exact(native, prop):
    """Randomised change points (gaps, one-frame shots, shots longer than the capacity), irregular positions,
    degenerate scores and budgets: picks / shot means / bitmap bit-exact against the oracle, on the fast path
    (everything in shared memory) and, with one oversized video in the batch, on the general path."""
    rng = np.random.default_rng(100 + prop[0] * 7 + prop[1])
    for oversized in (False, True):
        vids, scores = [], []
        n_vid = 40
        for i in range(n_vid):
            nf = int(rng.integers(1, 3000)) if not (oversized and i == 0) else 40000
            # sorted, disjoint, inclusive shots with random gaps
            cuts = np.sort(rng.choice(nf + 1, size=min(nf + 1, 2 * int(rng.integers(1, 40))), replace=False))
            shots = []
            for a, b in zip(cuts[0::2], cuts[1::2]):
                if b > a:
                    shots.append((int(a), int(b) - 1))
            if not shots:
                shots = [(0, nf - 1)]
            T = int(rng.integers(1, min(nf, 400) + 1))
            pos = np.sort(rng.choice(nf, size=T, replace=False)).astype(np.int32)
            sc = rng.random(T).astype(np.float32)
            sc[rng.random(T) < 0.05] = 0.0
            sc[rng.random(T) < 0.05] = 1.0
            if T > 3:
                sc[1] = np.nan
            vids.append(synth.Video(torch.zeros(T, 1), torch.zeros(T, 1), nf, pos, np.asarray(shots, np.int32)))
            scores.append(sc)
        lens = [v.T for v in vids]
        starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
        allsc = np.concatenate(scores)
        allpos = np.concatenate([v.positions for v in vids]).astype(np.int32)
        picks, seg_mean, summary, cps_start, sum_start = native.summarize_rows(
            torch.from_numpy(allsc).cuda(), torch.from_numpy(allpos).cuda(), starts, lens, [v.n_frames for v in vids],
            [v.cps for v in vids], prop)
        torch.cuda.synchronize()
        picks, seg_mean, summary = picks.cpu().numpy(), seg_mean.cpu().numpy(), summary.cpu().numpy()
        for i, v in enumerate(vids):
            wp, ws, wm = av_oracle.generate_summary(scores[i], v.cps, v.n_frames, v.positions, prop[0], prop[1])
            assert np.array_equal(wm, seg_mean[cps_start[i]:cps_start[i + 1]]), (oversized, i)
            assert np.array_equal(wp, picks[cps_start[i]:cps_start[i + 1]]), (oversized, i)
            assert np.array_equal(ws, summary[sum_start[i]:sum_start[i + 1]]), (oversized, i)


@pytest.mark.parametrize("n_vid,prop", [(3, (15, 100)), (9, (1, 2)), (20, (15, 100)), (30, (3, 7))])
def test_summarize_long_videos_on_clusters_bit_exact(native, n_vid, prop):
    """Batches whose largest capacity no longer fits one SM's fast path and that leave room for a cluster per video:
    8 CTAs per video up to 18 videos, 4 CTAs up to 37 (csrc/summarize.cu knapsack_cluster_kernel).  Ragged on purpose:
    capacities from a few cells to ~20,000, shots from one frame to longer than a CTA's slice, gaps between shots,
    short videos next to long ones.  Picks / shot means / bitmap bit-exact against the oracle."""
    rng = np.random.default_rng(1000 + n_vid)
    vids, scores = [], []
    for i in range(n_vid):
        # capacity = nf * num // den: 20,000 cells for the first video (the batch's largest; the cluster kernel covers
        # up to 24,575), a few cells or several thousand for the others
        cap = 20000 if i == 0 else int(rng.choice([int(rng.integers(3, 300)), int(rng.integers(3000, 19000))]))
        nf = cap * prop[1] // prop[0] + int(rng.integers(0, max(prop[1] // prop[0], 1)))
        # shots of 1 .. 600 frames (TVSum / SumMe shots are 30 - 300) with random gaps, up to ~400 of them
        shots, f, n_max = [], int(rng.integers(0, 50)), int(rng.integers(1, 400))
        while f < nf and len(shots) < n_max:
            end = min(f + int(rng.integers(1, 601)) - 1, nf - 1)
            shots.append((f, end))
            f = end + 1 + int(rng.integers(0, 200))
        if i == 1 and nf > 20000:
            # one long shot: 900 frames, whose halo reaches far into the left neighbour's slice (the kernel handles shots
            # up to 1,024 frames); in the 9-video case a third of the video, which sends the batch to the one-SM path
            end = nf // 3 if n_vid == 9 else 899
            shots = [(0, end)] + [(a, b) for a, b in shots if a > end]
        T = int(rng.integers(1, min(nf, 600) + 1))
        pos = np.sort(rng.choice(nf, size=T, replace=False)).astype(np.int32)
        sc = rng.random(T).astype(np.float32)
        sc[rng.random(T) < 0.05] = 0.0
        vids.append(synth.Video(torch.zeros(T, 1), torch.zeros(T, 1), nf, pos, np.asarray(shots, np.int32)))
        scores.append(sc)
    lens = [v.T for v in vids]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    from torch.profiler import ProfilerActivity, profile

    def run():
        out = native.summarize_rows(
            torch.from_numpy(np.concatenate(scores)).cuda(),
            torch.from_numpy(np.concatenate([v.positions for v in vids]).astype(np.int32)).cuda(), starts, lens,
            [v.n_frames for v in vids], [v.cps for v in vids], prop)
        torch.cuda.synchronize()
        return out

    res, kernels = None, ""
    try:
        with profile(activities=[ProfilerActivity.CUDA]) as prof:      # which kernel made the selection?
            res = run()
        kernels = " ".join(e.key for e in prof.key_averages())
    except Exception:       # no CUPTI (another subscriber holds it): the parity check below does not depend on it
        if res is None:
            res = run()
    picks, seg_mean, summary, cps_start, sum_start = res
    if "knapsack" in kernels:          # (CUPTI records available)
        want_kernel = {3: "knapsack_cluster_kernel<8>", 20: "knapsack_cluster_kernel<4>", 30: "knapsack_cluster_kernel<4>",
                       9: "knapsack_fast_kernel"}[n_vid]
        if n_vid == 9 and not any(v.n_frames > 20000 for v in vids[1:2]):
            want_kernel = "knapsack_cluster_kernel<8>"       # (no long shot drawn: the cluster kernel takes the batch)
        assert want_kernel in kernels, kernels
    picks, seg_mean, summary = picks.cpu().numpy(), seg_mean.cpu().numpy(), summary.cpu().numpy()
    for i, v in enumerate(vids):
        wp, ws, wm = av_oracle.generate_summary(scores[i], v.cps, v.n_frames, v.positions, prop[0], prop[1])
        assert np.array_equal(wm, seg_mean[cps_start[i]:cps_start[i + 1]]), i
        assert np.array_equal(wp, picks[cps_start[i]:cps_start[i + 1]]), i
        assert np.array_equal(ws, summary[sum_start[i]:sum_start[i + 1]]), i


@pytest.mark.parametrize("n_videos", [20, 70, 150, 300])
def test_forward_many_short_videos_all_lstm_variants(cuda_ready, n_videos):
    """Batches of 20 / 70 / 150 / 300 videos select the 16- / 32- / 64-slot recurrence kernels (and, beyond one
    wave of clusters, several waves); every video must equal its own B = 1 run bit for bit, and the reference."""
    rng = np.random.default_rng(n_videos)
    lens = [int(x) for x in rng.integers(1, 12, n_videos)]
    g = torch.Generator().manual_seed(n_videos)
    vis = [torch.randn(t, 1024, generator=g) for t in lens]
    aud = [torch.randn(t, 128, generator=g) for t in lens]
    m = make_model(spread=True, attn_axis="temporal")
    got = m.score_videos([(v.cuda(), a.cuda()) for v, a in zip(vis, aud)])
    sd = synth.seeded_state_dict(spread=True)
    port = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
    port.load_state_dict(sd)
    for i in list(range(0, n_videos, max(1, n_videos // 12))) + [n_videos - 1]:
        alone = m.score_videos([(vis[i].cuda(), aud[i].cuda())])[0]
        assert torch.equal(got[i], alone), i
        want = av_oracle_torch.run_videos(port, [(vis[i], aud[i])], "temporal")[0]
        assert rel(got[i].cpu().numpy(), want.numpy()) < 1e-3, i


def test_streamed_batches_equal_synchronous_calls(cuda_ready):
    """avs_forward_summarize_async (two batches in flight on the two staging slots, evaluation.summary.
    summarize_stream) returns, batch by batch, exactly what the synchronous call returns -- including when the
    batches differ in size (arena growth while the other slot is in flight) and when a slot is reused."""
    from avsum_b200.evaluation.summary import summarize_stream
    from avsum_b200.runtime import ShotDesc
    model = make_model(spread=True, attn_axis="literal_b1")
    nat = model.native()
    vids = sorted(synth.config2(), key=lambda v: -v.T)
    subsets = [vids[10:22], vids[0:50], vids[30:50], vids[0:50], vids[5:9], vids[20:45]]
    batches = []
    for sub in subsets:
        lens = [v.T for v in sub]
        starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
        batches.append((torch.cat([v.visual for v in sub]).pin_memory(), torch.cat([v.audio for v in sub]).pin_memory(),
                        torch.from_numpy(np.concatenate([v.positions for v in sub]).astype(np.int32)).pin_memory(),
                        starts, lens, ShotDesc([v.n_frames for v in sub], [v.cps for v in sub])))
    want = [nat.score_and_summarize_rows(b[0], b[1], b[2], b[3], b[4], None, b[5], 0.15, "literal_b1") for b in batches]
    for depth in (2, 1):
        got = list(summarize_stream(model, iter(batches), 0.15, depth=depth))
        assert len(got) == len(want)
        for k, (g, w) in enumerate(zip(got, want)):
            for a, b in zip(g[:4], w[:4]):
                assert torch.equal(a, b), (depth, k)
    # protocol: a busy slot is refused until it has been waited for
    b = batches[0]
    p = nat.score_and_summarize_rows(b[0], b[1], b[2], b[3], b[4], None, b[5], 0.15, "literal_b1", slot=0)
    with pytest.raises(ValueError):
        nat.score_and_summarize_rows(b[0], b[1], b[2], b[3], b[4], None, b[5], 0.15, "literal_b1", slot=0)
    out = p.wait()
    assert torch.equal(out[1], want[0][1])


def test_device_space_video_groups_are_bit_identical(native):
    """The video-group pipeline (used for host-space calls; a tuning switch for device-resident inputs) only changes
    WHEN kernels run: scores must equal the single-group run bit for bit, whatever the split."""
    vids = sorted(synth.config2(), key=lambda v: -v.T)
    lens = [v.T for v in vids]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    vd = torch.cat([v.visual for v in vids]).cuda()
    ad = torch.cat([v.audio for v in vids]).cuda()
    os.environ.pop("AVS_DEV_GROUPS", None)
    want = native.forward_rows(vd, ad, starts, lens, "temporal", "tf32").clone()
    try:
        for groups, shares in (("2", None), ("3", "20,70,100"), ("6", None)):
            os.environ["AVS_DEV_GROUPS"] = groups
            if shares:
                os.environ["AVS_DEV_SHARES"] = shares
            else:
                os.environ.pop("AVS_DEV_SHARES", None)
            got = native.forward_rows(vd, ad, starts, lens, "temporal", "tf32")
            torch.cuda.synchronize()
            assert torch.equal(got, want), (groups, shares)
    finally:
        os.environ.pop("AVS_DEV_GROUPS", None)
        os.environ.pop("AVS_DEV_SHARES", None)


def test_handle_follows_the_parameters_through_training_and_back_to_eval(cuda_ready):
    """A training step re-packs only the recurrences' tensors in the native handle (without synchronising); the next
    eval-mode forward must see EVERY updated parameter -- and a second training step the updated LSTM weights."""
    vid = synth.make_video(64, 1024, 128, 777)
    xv, xa = vid.visual[None].cuda(), vid.audio[None].cuda()
    m = make_model(spread=True, attn_axis="literal_b1")
    opt = torch.optim.SGD(m.parameters(), lr=0.5)     # a step large enough to move the scores well beyond 1e-3
    with torch.no_grad():
        before = m.eval()(xv, xa).clone()
    for _ in range(2):
        m.train()
        for seq in (m.visual_fc, m.audio_fc):
            seq[2].p = 0.0
        loss = ((m(xv, xa) - 0.9) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
    m.eval()
    with torch.no_grad():
        got = m(xv, xa)
    port = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
    port.load_state_dict({k: v.detach().cpu() for k, v in m.state_dict().items()})
    want = av_oracle_torch.run_videos(port, [(vid.visual, vid.audio)], "literal")[0]
    assert float((got.cpu() - before.cpu()).abs().max()) > 5e-3, "the optimiser steps did not move the scores"
    assert rel(got.cpu().numpy(), want.numpy()) < REL_TOL["tf32"]


def test_attention_core_is_repeatable_under_concurrent_load(cuda_ready):
    """Regression test for a barrier-phase aliasing bug: with another kernel sharing the SMs, a softmax warp could run
    a whole key block ahead of the slowest warp and pass the final "O complete" parity wait two phases early
    (32-row blocks of wrong context, ~8 % of launches under load).  Every launch must reproduce the first bit for bit."""
    vids = sorted(synth.config2(), key=lambda v: -v.T)[:30]
    lens = [v.T for v in vids]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    g = torch.Generator().manual_seed(5)
    qkv = (torch.randn(int(sum(lens)), 3072, generator=g) * 0.5).cuda()
    side = torch.cuda.Stream()
    a = torch.randn(4096, 4096, device="cuda")
    ones = np.ones_like(starts)
    ref = runtime.attention(qkv, 1024, 4, starts, ones, lens).clone()
    for i in range(120):
        with torch.cuda.stream(side):
            for _ in range(3):
                torch.mm(a, a)
        out = runtime.attention(qkv, 1024, 4, starts, ones, lens)
        assert torch.equal(out, ref), f"launch {i} differs from the first"
    torch.cuda.synchronize()


def test_fp16_range_watch_reports_clamped_activations(native):
    """The default precision keeps the fc activations in fp16, whose cast clamps at 65504: features of huge magnitude
    must not produce silently wrong scores.  Host-space calls fail loudly, device-space callers ask range_status();
    ordinary features and the bf16 mode (fp32 exponent range) do not trip it."""
    g = torch.Generator().manual_seed(5)
    R = 300
    visual, audio = torch.randn(R, 1024, generator=g), torch.randn(R, 128, generator=g)
    starts, lens = [0, 200], [200, 100]
    native.forward_rows(visual.cuda(), audio.cuda(), starts, lens, "literal_b1", "tf32")
    assert native.range_status() is False
    huge = visual * 3e6                                   # |fc output| ~ 3e6 * sqrt(1024) * |w| >> 65504
    native.forward_rows(huge.cuda(), audio.cuda(), starts, lens, "literal_b1", "tf32")
    assert native.range_status() is True
    assert native.range_status() is False                 # reading cleared it
    with pytest.raises(RuntimeError, match="fp16 range limit"):
        native.forward_rows(huge.pin_memory(), audio.pin_memory(), starts, lens, "literal_b1", "tf32")
    out = native.forward_rows(visual.pin_memory(), audio.pin_memory(), starts, lens, "literal_b1", "tf32")   # recovered
    assert torch.isfinite(out).all()
    native.forward_rows(huge.cuda(), audio.cuda(), starts, lens, "literal_b1", "bf16")
    assert native.range_status() is False


@pytest.mark.gpu
@pytest.mark.parametrize("axis", ["literal_b1", "temporal"])
@pytest.mark.parametrize("n_videos", [17, 50, 64])
def test_pipelined_tail_is_bit_identical(native, n_videos, axis):
    """Device-resident batches ordered longest video first run the tail (value projection -- or q | k | v projection and
    the attention core over the group's videos --, out_proj, score head) group by group behind each group's recurrence (AVS_PIPE_TAIL, on by default).  That only changes WHEN
    and over which row range the GEMMs are launched: the scores must equal the one-launch schedule bit for bit, for 3
    to 8 recurrence groups, and an unordered batch (which falls back to the one-launch schedule) must agree per video."""
    rng = np.random.default_rng(n_videos)
    base = sorted(synth.config2(), key=lambda v: -v.T)
    vids = [base[i % len(base)] for i in range(n_videos)]
    vids.sort(key=lambda v: -v.T)
    lens = [v.T for v in vids]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    vd = torch.cat([v.visual for v in vids]).cuda()
    ad = torch.cat([v.audio for v in vids]).cuda()
    try:
        os.environ["AVS_PIPE_TAIL"] = "0"
        want = native.forward_rows(vd, ad, starts, lens, axis, "tf32").clone()
        os.environ["AVS_PIPE_TAIL"] = "2"   # forced (1 = the default: only when the groups' lengths differ or the batch is long)
        for _ in range(3):
            got = native.forward_rows(vd, ad, starts, lens, axis, "tf32")
            torch.cuda.synchronize()
            assert torch.equal(got, want)
        # equally long videos (the default schedule leaves these alone): forced, 12 videos = 3 groups of 4
        eq_lens = [200] * 12
        eq_starts = (np.arange(12) * 200).astype(np.int32)
        os.environ["AVS_PIPE_TAIL"] = "0"
        want_eq = native.forward_rows(vd[:2400], ad[:2400], eq_starts, eq_lens, axis, "tf32").clone()
        os.environ["AVS_PIPE_TAIL"] = "2"
        got_eq = native.forward_rows(vd[:2400], ad[:2400], eq_starts, eq_lens, axis, "tf32")
        torch.cuda.synchronize()
        assert torch.equal(got_eq, want_eq)
        os.environ["AVS_PIPE_TAIL"] = "1"
        # the same videos in a shuffled order: groups are no longer contiguous row blocks -> one-launch schedule
        perm = rng.permutation(n_videos)
        vd2 = torch.cat([vids[i].visual for i in perm]).cuda()
        ad2 = torch.cat([vids[i].audio for i in perm]).cuda()
        lens2 = [lens[i] for i in perm]
        starts2 = np.concatenate([[0], np.cumsum(lens2)[:-1]]).astype(np.int32)
        got2 = native.forward_rows(vd2, ad2, starts2, lens2, axis, "tf32")
        torch.cuda.synchronize()
        for j, i in enumerate(perm):
            assert torch.equal(got2[starts2[j]:starts2[j] + lens2[j]], want[starts[i]:starts[i] + lens[i]])
    finally:
        os.environ.pop("AVS_PIPE_TAIL", None)
