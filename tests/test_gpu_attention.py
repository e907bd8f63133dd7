"""GPU parity tests of the tcgen05 attention core (csrc/attention_tc.cu) against the oracle, through the C ABI
(``avs_attention``), and of the full forward with attention weights that are far from uniform.

Round-1 gap (VERDICT.md "What's weak" #1): with torch's default init every temporal-attention weight of the model is
within 2 % of 1/T, so a forward-level parity test only pins a masked mean of V; and the one direct test of the core
used T = 33, less than one 64-key block.  Here the core is compared with ``oracle.av_oracle.mha_core``
(/root/reference/models/attention.py:17-23 restated) on ALL heads and rows for

* T in {65, 200, 700, 1000, 8192}: 2 ... 128 key blocks, ragged last block, several 128-query tiles;
* logits built so that the running maximum of a row GROWS by 12 log2 units in late key blocks (the O-rescale branch
  of the lazy online softmax, ``bmax > m_used + 8`` at j > 0), SHRINKS (later blocks underflow), grows by less than
  the lazy threshold (no rescale, P up to 2^8), or stays flat -- all four kinds of row inside every 32-row warp tile;
* packed variable-length batches (the K / V tiles of a video's last block reach into the next video's rows);
* fp16 (AVS_PREC_TF32 mode) and bf16 operands.

The inputs are pre-rounded to the operand grid, so the oracle (fp32 arithmetic on the same numbers) and the kernel
see identical q, k, v; what remains is P rounded to 16 bits, ex2.approx and fp32 accumulation order.
Stated tolerances: max |err| <= 1.5e-3 * max |out| (fp16 operands), 1.0e-2 (bf16).
"""
import os

import numpy as np
import pytest
import torch

from avsum_b200 import runtime, synth
from oracle import av_oracle, av_oracle_torch

pytestmark = pytest.mark.gpu

E, H, DH = 1024, 4, 256
TOL = {"tf32": 1.5e-3, "bf16": 1.0e-2}
LOG2E = 1.4426950408889634


def _round_to(x: torch.Tensor, prec: str) -> torch.Tensor:
    return x.to(torch.float16 if prec == "tf32" else torch.bfloat16).to(torch.float32)


def make_qkv(T: int, kind: str, prec: str, seed: int):
    """q | k | v [T, 3E] fp32 on the operand grid.  kind:
    "random"    logits ~ N(0, 1 nat)
    "staircase" every head has a direction u; key block structure step(j) in {0..4} jumps at four key blocks (the
                second, one in the middle, the second to last and the last), query row i has gain g_i in
                {+12, -12, +3, 0} log2 units per step (i % 4), so inside every warp tile there are rows whose running
                max grows by 12 log2 units at late blocks (rescale), rows whose max is in block 0, rows that grow by
                3 (below the lazy threshold of 8) and flat rows;
    "mid_peak"  like staircase but the steps go up to the middle of the sequence and down again."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(T, E, generator=g)
    k = torch.randn(T, E, generator=g)
    v = torch.randn(T, E, generator=g)
    if kind != "random":
        nblk = (T + 63) // 64
        blk = torch.arange(T) // 64
        if kind == "staircase":
            jumps = sorted({min(1, nblk - 1), nblk // 2, max(nblk - 2, 0), nblk - 1} - {0})
            step = sum(((blk >= jb).to(torch.float32) for jb in jumps), torch.zeros(T))
        else:
            mid = max(nblk // 2, 1)
            up = sorted({min(1, mid), mid // 2, mid} - {0})
            down = sorted({min(mid + 1, nblk - 1), nblk - 1} - {0} - set(up))
            step = sum(((blk >= jb).to(torch.float32) for jb in up), torch.zeros(T)) - \
                sum(((blk >= jb).to(torch.float32) for jb in down), torch.zeros(T))
        gain = torch.tensor([12.0, -12.0, 3.0, 0.0])[torch.arange(T) % 4]
        ab = (16.0 / LOG2E) ** 0.5          # alpha * beta / 16 * log2(e) == 1 log2 unit per (gain x step)
        for h in range(H):
            u = torch.randn(DH, generator=g)
            u = u / u.norm()
            sl = slice(h * DH, (h + 1) * DH)
            # remove the random component along u so the staircase is what decides the maxima
            q[:, sl] -= (q[:, sl] @ u)[:, None] * u
            k[:, sl] -= (k[:, sl] @ u)[:, None] * u
            q[:, sl] += (gain * ab)[:, None] * u
            k[:, sl] += (step * ab)[:, None] * u
    return _round_to(torch.cat([q, k, v], dim=1), prec)


def oracle_ctx(qkv: torch.Tensor, lens, starts):
    out = np.zeros((qkv.shape[0], E), np.float32)
    x = qkv.numpy()
    for s, n in zip(starts, lens):
        if n:
            out[s:s + n] = av_oracle.mha_core(x[s:s + n, :E], x[s:s + n, E:2 * E], x[s:s + n, 2 * E:], H)
    return out


def check(got: torch.Tensor, want: np.ndarray, prec: str, what):
    got = got.cpu().numpy()
    assert np.isfinite(got).all(), what
    scale = float(np.max(np.abs(want)))
    err = np.abs(got - want)
    worst = float(err.max()) / scale
    assert worst < TOL[prec], (what, worst, np.unravel_index(int(err.argmax()), err.shape))
    # every (128-query tile, head) individually: a wrong tile must not hide behind a max-norm over the whole output
    T = want.shape[0]
    for h in range(H):
        e = err[:, h * DH:(h + 1) * DH].max(axis=1)
        for t0 in range(0, T, 128):
            assert float(e[t0:t0 + 128].max()) / scale < TOL[prec], (what, h, t0)
    return worst


@pytest.fixture(params=["auto", "tmem", "smem"])
def q_mode(request):
    """The kernel has two variants -- Q in tensor memory (TS product, chosen for max length >= 1024) and Q in shared
    memory (SS product, short sequences); AVS_ATTN_Q forces one, so every case below runs through both."""
    if request.param == "auto":
        os.environ.pop("AVS_ATTN_Q", None)
    else:
        os.environ["AVS_ATTN_Q"] = request.param
    yield request.param
    os.environ.pop("AVS_ATTN_Q", None)


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
@pytest.mark.parametrize("kind", ["random", "staircase", "mid_peak"])
@pytest.mark.parametrize("T", [65, 200, 700, 1000])
def test_attention_core_matches_oracle_all_heads(cuda_ready, q_mode, T, kind, prec):
    qkv = make_qkv(T, kind, prec, seed=T * 3 + len(kind))
    want = oracle_ctx(qkv, [T], [0])
    if kind == "staircase":   # the construction really forces what it claims: late maxima, wide spans
        x = qkv.numpy()
        s = (x[:, :DH] @ x[:, E:E + DH].T) / 16.0 * LOG2E
        grow_rows = np.arange(T) % 4 == 0
        assert (s[grow_rows].argmax(axis=1) >= T - 64 - (T % 64 or 64)).all()
        assert float((s[grow_rows].max(axis=1) - s[grow_rows].min(axis=1)).min()) > (40.0 if T > 256 else 30.0 if T > 128 else 10.0)
    got = runtime.attention(qkv.cuda(), E, H, [0], [1], [T], precision=prec)
    check(got, want, prec, (T, kind, prec))


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
def test_attention_core_long_sequence_T8192(cuda_ready, q_mode, prec):
    """BASELINE configs[3]: 128 key blocks, 64 query tiles per head; rescaling rows at blocks 1, 64, 126 and 127."""
    T = 8192
    qkv = make_qkv(T, "staircase", prec, seed=8192)
    want = oracle_ctx(qkv, [T], [0])
    got = runtime.attention(qkv.cuda(), E, H, [0], [1], [T], precision=prec)
    check(got, want, prec, (T, prec))


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
def test_attention_core_packed_varlen_batch(cuda_ready, q_mode, prec):
    """Packed rows, ragged lengths: the 64-key K / V tiles of a video's last block cover rows of the NEXT video
    (masked), the 128-query Q tile of its last block too (not stored); a zero-length video; lengths of 1, 64, 65."""
    lens = [130, 65, 700, 1, 0, 64, 257, 1000, 63]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    parts = [make_qkv(max(n, 1), ["staircase", "mid_peak", "random"][i % 3], prec, seed=100 + i)[:n]
             for i, n in enumerate(lens)]
    qkv = torch.cat(parts, dim=0)
    want = oracle_ctx(qkv, lens, starts)
    got = runtime.attention(qkv.cuda(), E, H, starts, np.ones_like(starts), lens, precision=prec)
    for i, (s, n) in enumerate(zip(starts, lens)):
        if n:
            check(got[s:s + n], want[s:s + n], prec, ("video", i, n))
    # the same rows in another batch composition (different neighbours) give the same bits
    order = [7, 2, 0, 6, 1, 5, 8, 3]
    qkv2 = torch.cat([parts[i] for i in order], dim=0)
    lens2 = [lens[i] for i in order]
    starts2 = np.concatenate([[0], np.cumsum(lens2)[:-1]]).astype(np.int32)
    got2 = runtime.attention(qkv2.cuda(), E, H, starts2, np.ones_like(starts2), lens2, precision=prec)
    for s2, i in zip(starts2, order):
        assert torch.equal(got2[s2:s2 + lens[i]], got[starts[i]:starts[i] + lens[i]]), i


@pytest.mark.parametrize("q_force", ["tmem", "smem"])
def test_attention_rescale_branch_is_repeatable_under_concurrent_load(cuda_ready, q_force):
    """The O-rescale branch waits for "P_{j-1} V_{j-1} has landed" on a per-key-block barrier by phase parity
    (attention_tc.cu, softmax warps, j > 0).  A parity wait only separates adjacent phases; the argument why it is
    safe there (S_j can only be ready after P_{j-2} V_{j-2} was committed, so the barrier is at most one phase behind)
    is checked here the way the bar_done bug of round 1 was found: another kernel shares the SMs, the staircase input
    makes every fourth row rescale at four late blocks, and every launch must reproduce the oracle-checked first one
    bit for bit."""
    os.environ["AVS_ATTN_Q"] = q_force
    try:
        _rescale_repeatability()
    finally:
        os.environ.pop("AVS_ATTN_Q", None)


def _rescale_repeatability():
    lens = [700, 650, 1000, 333, 512, 200]
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    qkv = torch.cat([make_qkv(n, "staircase", "tf32", seed=500 + i) for i, n in enumerate(lens)], dim=0)
    want = oracle_ctx(qkv, lens, starts)
    qkv_d = qkv.cuda()
    ones = np.ones_like(starts)
    ref = runtime.attention(qkv_d, E, H, starts, ones, lens).clone()
    check(ref, want, "tf32", "first launch")
    side = torch.cuda.Stream()
    a = torch.randn(4096, 4096, device="cuda")
    for i in range(150):
        with torch.cuda.stream(side):
            for _ in range(3):
                torch.mm(a, a)
        out = runtime.attention(qkv_d, E, H, starts, ones, lens)
        assert torch.equal(out, ref), f"launch {i} differs from the first"
    torch.cuda.synchronize()


# ------------------------------------------------------------------ full forward, attention far from uniform
REL_TOL = {"tf32": 1e-3, "fp32_simt": 1e-4, "bf16": 2e-2}


def rel(got, want):
    return float(np.max(np.abs(np.asarray(got, np.float64) - want) / np.abs(want)))


def make_model(spread=True, peaked=True, **kw):
    from avsum_b200.models.av_model import AVBiLSTMModel
    m = AVBiLSTMModel(1024, 128, 512, **kw).eval()
    m.load_state_dict(synth.seeded_state_dict(spread=spread, peaked=peaked))
    return m.cuda()


@pytest.mark.parametrize("prec", ["tf32", "fp32_simt", "bf16"])
def test_forward_peaked_attention_matches_reference(cuda_ready, golden_dir, prec):
    """Reference goldens (tests/golden/make_golden.py, imported reference) for the weight set whose temporal
    attention weights span [5e-13, 0.94]: config 1, a B = 2 batch, the first four videos of config 2."""
    g = np.load(os.path.join(golden_dir, "config1_peaked.npz"))
    m = make_model(precision=prec)
    vid = synth.config1()
    for axis in ("literal", "temporal"):
        got = m(vid.visual[None].cuda(), vid.audio[None].cuda(), attn_axis=axis).cpu().numpy()
        assert rel(got, g["scores_" + axis]) < REL_TOL[prec], (axis, rel(got, g["scores_" + axis]))
    g = np.load(os.path.join(golden_dir, "batch2_T130_peaked.npz"))
    gen = torch.Generator().manual_seed(int(g["seed_in"]))
    visual, audio = torch.randn(2, 130, 1024, generator=gen), torch.randn(2, 130, 128, generator=gen)
    for axis in ("literal", "temporal"):
        got = m(visual.cuda(), audio.cuda(), attn_axis=axis).cpu().numpy()
        assert got.shape == (2, 130) and rel(got, g["scores_" + axis]) < REL_TOL[prec], axis
    g = np.load(os.path.join(golden_dir, "config2_first4_peaked.npz"))
    vids = synth.config2()[:4]
    got = m.score_videos([(v.visual.cuda(), v.audio.cuda()) for v in vids], attn_axis="temporal")
    assert rel(torch.cat(got).cpu().numpy(), g["scores_temporal"]) < REL_TOL[prec]


def test_config2_full_batch_peaked_attention_matches_cpu_port(cuda_ready):
    """BASELINE configs[1] at full size (50 videos, 21,477 frames), temporal attention with peaked weights, vs the
    torch CPU port of the reference (bit-identical to it, tests/test_oracle_golden.py); 1e-3 relative."""
    vids = synth.config2()
    port = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
    port.load_state_dict(synth.seeded_state_dict(spread=True, peaked=True))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    want = av_oracle_torch.run_videos(port, [(v.visual, v.audio) for v in vids], "temporal")
    m = make_model(attn_axis="temporal")
    got = m.score_videos([(v.visual, v.audio) for v in vids])
    worst = max(rel(a.numpy(), b.numpy()) for a, b in zip(got, want))
    spread = max(float(w.max() - w.min()) for w in want)
    assert spread > 0.2, "the peaked weight set must make the frames of a video differ"
    assert worst < 1e-3, worst


@pytest.mark.parametrize("peaked", [False, True])
def test_config4_long_video_full_forward_matches_cpu_port(cuda_ready, peaked):
    """BASELINE configs[3]: T = 8192 frames, temporal attention (128 key blocks per query tile), full forward of two
    videos in one batch vs the CPU port; fp32 frame scores within 1e-3 relative."""
    vids = synth.config4(2, 8192)
    port = av_oracle_torch.RefPortModel(1024, 128, 512).eval()
    port.load_state_dict(synth.seeded_state_dict(spread=True, peaked=peaked))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    want = av_oracle_torch.run_videos(port, [(v.visual, v.audio) for v in vids], "temporal")
    m = make_model(peaked=peaked, attn_axis="temporal")
    got = m.score_videos([(v.visual.cuda(), v.audio.cuda()) for v in vids])
    for a, b in zip(got, want):
        assert a.shape == (8192,)
        r = rel(a.cpu().numpy(), b.numpy())
        assert r < 1e-3, (peaked, r)
    if peaked:
        assert float(want[0].max() - want[0].min()) > 0.1
