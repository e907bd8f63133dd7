"""CPU oracle for the AudioVidSum hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product path
(``avsum_b200``) never does and fails loudly if its CUDA library is missing.

Two halves:

1. **Model forward** -- a plain-numpy fp32 restatement of the reference's
   ``AVBiLSTMModel.forward`` (``/root/reference/models/av_model.py:33-46``) and of
   ``MultiHeadSelfAttention.forward`` (``/root/reference/models/attention.py:15-25``).
   The arithmetic the reference delegates to PyTorch (``nn.Linear``, ``nn.LSTM``,
   ``nn.MultiheadAttention``, ``nn.Sigmoid``; unpinned version, torch 2.11.0 in
   this image) is restated from its published definitions.  PARITY PINNED: the
   restatement is checked against ``tests/golden/*.npz``, which were produced by
   importing the reference itself (``tests/golden/make_golden.py``).

2. **Summary generation** (shot pooling over change points, 0/1 knapsack at a
   length budget, keyshot bitmap).  The reference contains NO such code
    -- PARITY UNPINNED BY THE REFERENCE for this half.  The
   algorithm is specified here, in integers, so that the CUDA kernels can be
   bit-exact; it follows the de-facto TVSum/SumMe protocol (per-shot mean of
   upsampled frame scores, capacity floor(0.15 * n_frames), shot lengths as
   weights) and stays consistent with the reference's nearest analogues
   (mean pooling per shot ``utils/alignments.py:19-20``; half-open (start, end)
   shot lists ``evaluation/metrics.py:3,7``).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32

# ----------------------------------------------------------------------------
# 1. model forward
# ----------------------------------------------------------------------------


def _sigmoid(x):
    x = np.asarray(x, dtype=F32)
    return (F32(1.0) / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


def linear(x, w, b):
    """nn.Linear: x @ w.T + b (reference call sites av_model.py:10-15, 29-31)."""
    return (x.astype(F32) @ w.astype(F32).T + b.astype(F32)).astype(F32)


def fc_relu(x, w, b):
    """visual_fc / audio_fc in eval mode (av_model.py:10-15, 35-36): Dropout(0.3) is identity."""
    return np.maximum(linear(x, w, b), F32(0))


def lstm_direction(x, w_ih, w_hh, b_ih, b_hh, reverse=False):
    """One direction of nn.LSTM(batch_first) for ONE sequence x[T, I] -> h[T, Hc].

    PyTorch semantics (av_model.py:18-23): gate order i, f, g, o in the stacked
    [4*Hc, *] weights, both bias vectors added, zero initial (h, c); the reverse
    direction walks T-1 -> 0 and its output is stored at the frame it consumed.
    """
    T = x.shape[0]
    hc = w_hh.shape[1]
    xg = (x.astype(F32) @ w_ih.astype(F32).T + (b_ih + b_hh).astype(F32)).astype(F32)  # [T, 4Hc]
    whh_t = np.ascontiguousarray(w_hh.astype(F32).T)  # [Hc, 4Hc]
    h = np.zeros(hc, dtype=F32)
    c = np.zeros(hc, dtype=F32)
    out = np.empty((T, hc), dtype=F32)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        g = xg[t] + h @ whh_t
        i = _sigmoid(g[0:hc])
        f = _sigmoid(g[hc:2 * hc])
        gg = np.tanh(g[2 * hc:3 * hc], dtype=F32)
        o = _sigmoid(g[3 * hc:4 * hc])
        c = (f * c + i * gg).astype(F32)
        h = (o * np.tanh(c, dtype=F32)).astype(F32)
        out[t] = h
    return out


def bilstm(x, p, prefix):
    """Bidirectional single-layer LSTM on one sequence: [T, I] -> [T, 2*Hc] = [h_fwd | h_bwd]."""
    fwd = lstm_direction(x, p[prefix + ".weight_ih_l0"], p[prefix + ".weight_hh_l0"],
                         p[prefix + ".bias_ih_l0"], p[prefix + ".bias_hh_l0"], reverse=False)
    bwd = lstm_direction(x, p[prefix + ".weight_ih_l0_reverse"], p[prefix + ".weight_hh_l0_reverse"],
                         p[prefix + ".bias_ih_l0_reverse"], p[prefix + ".bias_hh_l0_reverse"], reverse=True)
    return np.concatenate([fwd, bwd], axis=-1)


def mha_core(q, k, v, num_heads):
    """Scaled-dot-product attention for one sequence.  q, k, v: [L, E] -> [L, E].

    Head split is the contiguous view(L, H, dh) (attention.py:17-19), scale
    1/sqrt(dh) (attention.py:21), softmax over keys (attention.py:22).
    """
    L, E = q.shape
    dh = E // num_heads
    out = np.empty((L, E), dtype=F32)
    scale = F32(1.0 / math.sqrt(dh))
    for h in range(num_heads):
        sl = slice(h * dh, (h + 1) * dh)
        s = (q[:, sl] @ k[:, sl].T).astype(F32) * scale
        s = s - s.max(axis=-1, keepdims=True)
        e = np.exp(s, dtype=F32)
        a = e / e.sum(axis=-1, keepdims=True, dtype=F32)
        out[:, sl] = a.astype(F32) @ v[:, sl]
    return out


def mha_packed(x_seq, in_w, in_b, out_w, out_b, num_heads):
    """nn.MultiheadAttention(q=k=v=x) for one sequence x_seq[L, E] (av_model.py:26, 44)."""
    E = x_seq.shape[1]
    qkv = linear(x_seq, in_w, in_b)
    ctx = mha_core(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], num_heads)
    return linear(ctx, out_w, out_b)


def attention_block(fused, p, num_heads, attn_axis):
    """fused[B, T, E] -> attn_out[B, T, E].

    attn_axis == "literal": exactly what av_model.py:44 does -- the module was built
        without batch_first (av_model.py:26) so dim 0 (the video axis) is the sequence
        and dim 1 (frames) the batch: frame t of video b attends over the B videos.
    attn_axis == "temporal": frame self-attention within each video (what
        models/attention.py computes; identical to feeding fused.transpose(0, 1)).
    """
    B, T, E = fused.shape
    args = (p["attention.in_proj_weight"], p["attention.in_proj_bias"],
            p["attention.out_proj.weight"], p["attention.out_proj.bias"], num_heads)
    out = np.empty_like(fused, dtype=F32)
    if attn_axis == "literal":
        for t in range(T):
            out[:, t, :] = mha_packed(fused[:, t, :], *args)
    elif attn_axis == "temporal":
        for b in range(B):
            out[b] = mha_packed(fused[b], *args)
    else:
        raise ValueError(attn_axis)
    return out


def scorer(x, p):
    """scorer (av_model.py:29-31): Linear(E,64)+ReLU+Linear(64,1)+Sigmoid -> [..., 1]."""
    h = np.maximum(linear(x, p["scorer.0.weight"], p["scorer.0.bias"]), F32(0))
    return _sigmoid(linear(h, p["scorer.2.weight"], p["scorer.2.bias"]))


def forward(p, visual, audio, num_heads=4, attn_axis="literal", lengths=None):
    """AVBiLSTMModel.forward (av_model.py:33-46) on visual[B,T,Dv], audio[B,T,Da].

    With ``lengths`` (an extension the reference lacks) every video b only uses its
    first lengths[b] frames (packed-sequence semantics; temporal axis only) and a
    list of per-video score vectors is returned.  Without it the result is
    squeezed exactly like the reference's ``.squeeze()`` (av_model.py:46).
    """
    p = {k: np.asarray(v, dtype=F32) for k, v in p.items()}
    visual = np.asarray(visual, dtype=F32)
    audio = np.asarray(audio, dtype=F32)
    B, T, _ = visual.shape
    if lengths is not None:
        if attn_axis != "temporal":
            raise ValueError("lengths require attn_axis='temporal'")
        return [forward(p, visual[b:b + 1, :n], audio[b:b + 1, :n], num_heads, "temporal").reshape(-1)
                for b, n in enumerate(lengths)]
    v_emb = fc_relu(visual, p["visual_fc.0.weight"], p["visual_fc.0.bias"])
    a_emb = fc_relu(audio, p["audio_fc.0.weight"], p["audio_fc.0.bias"])
    fused = np.empty((B, T, v_emb.shape[-1] * 2), dtype=F32)
    for b in range(B):
        fused[b] = np.concatenate([bilstm(v_emb[b], p, "visual_bilstm"),
                                   bilstm(a_emb[b], p, "audio_bilstm")], axis=-1)
    attn = attention_block(fused, p, num_heads, attn_axis)
    return np.squeeze(scorer(attn, p))


def mhsa_forward(p, x, num_heads):
    """MultiHeadSelfAttention.forward (attention.py:15-25); p has query/key/value/out .weight/.bias."""
    x = np.asarray(x, dtype=F32)
    out = np.empty_like(x)
    for b in range(x.shape[0]):
        q = linear(x[b], p["query.weight"], p["query.bias"])
        k = linear(x[b], p["key.weight"], p["key.bias"])
        v = linear(x[b], p["value.weight"], p["value.bias"])
        out[b] = linear(mha_core(q, k, v, num_heads), p["out.weight"], p["out.bias"])
    return out


# ----------------------------------------------------------------------------
# 2. summary generation (repo-specified; parity unpinned by the reference)
# ----------------------------------------------------------------------------

SCORE_FRAC_BITS = 24  # scores in [0, 1] are quantised to q = rint(s * 2**24)


def quantize_scores(scores):
    """fp32 score -> int64 fixed point; the ONLY floating-point step of the summary path.

    s * 2**24 is exact in fp32 (power-of-two scaling), rint is round-half-even;
    NaN -> 0 and the result is clamped to [0, 2**24].
    """
    s = np.asarray(scores, dtype=F32)
    q = np.rint(s * F32(1 << SCORE_FRAC_BITS))
    q = np.where(np.isnan(q), 0, q)
    return np.clip(q, 0, 1 << SCORE_FRAC_BITS).astype(np.int64)


def shot_pool(scores, positions, n_frames, cps):
    """Mean of the upsampled frame scores over each shot, in fixed point.

    Sampled frame i covers original frames [positions[i], positions[i+1]) with
    positions[T] := n_frames; shot s covers the inclusive range cps[s] = [a, b].
    Returns (seg_sum int64[S], nfps int64[S], seg_mean int64[S]) where
    seg_mean = floor((2*seg_sum + nfps) / (2*nfps)) (mean rounded half up).
    """
    q = quantize_scores(scores)
    pos = np.asarray(positions, dtype=np.int64)
    T = q.shape[0]
    edges = np.concatenate([pos[:T], [int(n_frames)]]).astype(np.int64)
    cps = np.asarray(cps, dtype=np.int64).reshape(-1, 2)
    S = cps.shape[0]
    seg_sum = np.zeros(S, dtype=np.int64)
    nfps = (cps[:, 1] - cps[:, 0] + 1).astype(np.int64)
    for s in range(S):
        a, b1 = cps[s, 0], cps[s, 1] + 1
        ov = np.minimum(edges[1:], b1) - np.maximum(edges[:-1], a)
        seg_sum[s] = int(np.sum(np.maximum(ov, 0) * q))
    seg_mean = np.where(nfps > 0, (2 * seg_sum + nfps) // np.maximum(2 * nfps, 1), 0).astype(np.int64)
    return seg_sum, nfps, seg_mean


def knapsack(values, weights, capacity):
    """0/1 knapsack, integer DP with a fixed tie rule.

    dp_s[w] = max(dp_{s-1}[w], dp_{s-1}[w - wt_s] + v_s); item s is marked
    'kept at w' only on a STRICT improvement.  Back-trace from the last item at
    w = capacity.  Returns picks uint8[S].
    """
    values = np.asarray(values, dtype=np.int64)
    weights = np.asarray(weights, dtype=np.int64)
    S = values.shape[0]
    cap = int(max(capacity, 0))
    dp = np.zeros(cap + 1, dtype=np.int64)
    keep = np.zeros((S, cap + 1), dtype=bool)
    for s in range(S):
        wt = int(weights[s])
        if wt <= cap and wt > 0:
            cand = dp[:cap + 1 - wt] + values[s]
            better = cand > dp[wt:]
            keep[s, wt:] = better
            dp[wt:] = np.where(better, cand, dp[wt:])
        elif wt == 0 and values[s] > 0:  # degenerate empty shot: free value
            keep[s, :] = True
            dp = dp + values[s]
    picks = np.zeros(S, dtype=np.uint8)
    w = cap
    for s in range(S - 1, -1, -1):
        if keep[s, w]:
            picks[s] = 1
            w -= int(weights[s])
    return picks


def generate_summary(scores, cps, n_frames, positions, proportion_num=15, proportion_den=100):
    """scores[T] -> (picks uint8[S], summary uint8[n_frames], seg_mean int64[S]).

    capacity = floor(n_frames * 15 / 100) computed in integers.
    """
    _, nfps, seg_mean = shot_pool(scores, positions, n_frames, cps)
    cap = (int(n_frames) * proportion_num) // proportion_den
    picks = knapsack(seg_mean, nfps, cap)
    cps = np.asarray(cps, dtype=np.int64).reshape(-1, 2)
    summary = np.zeros(int(n_frames), dtype=np.uint8)
    for s in range(cps.shape[0]):
        if picks[s]:
            summary[max(cps[s, 0], 0):min(cps[s, 1] + 1, int(n_frames))] = 1
    return picks, summary, seg_mean


# ----------------------------------------------------------------------------
# 3. evaluation helpers that follow reference files line by line
# ----------------------------------------------------------------------------


def temporal_f1(pred_shots, gt_shots):
    """compute_temporal_f1 (evaluation/metrics.py:1-9) == compute_f1 (utils/shot_metrics.py:12-16)."""
    overlap = 0
    for ps, pe in pred_shots:
        for gs, ge in gt_shots:
            overlap += max(0, min(pe, ge) - max(ps, gs))
    precision = overlap / sum(pe - ps for ps, pe in pred_shots)
    recall = overlap / sum(ge - gs for gs, ge in gt_shots)
    return 2 * (precision * recall) / (precision + recall + 1e-8)


def align_shots_to_annotations(shot_boundaries, annotations, fps):
    """utils/alignments.py:4-22: mean of annotations over 2-second bins per shot."""
    out = []
    for start, end in shot_boundaries:
        a = int((start / fps) // 2)
        b = int((end / fps) // 2) + 1
        out.append(np.asarray(annotations)[a:b].mean())
    return np.asarray(out)


def threshold_f1(pred, target):
    """Metric block of scripts/evaluate.py:26-33 for one video."""
    bp = (pred > np.mean(pred)).astype(int)
    bt = (target > np.mean(target)).astype(int)
    tp = np.logical_and(bp, bt).sum()
    precision = tp / bp.sum()
    recall = tp / bt.sum()
    return 2 * (precision * recall) / (precision + recall + 1e-8)


def eval_metrics(pred, target):
    """Metric block of scripts/evaluate.py:25-36 for one video, with the SAME library calls the reference
    makes (numpy + scipy.stats): returns (f1, spearman, kendall)."""
    from scipy.stats import kendalltau, spearmanr
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            f1 = threshold_f1(pred, target)
            return float(f1), float(spearmanr(pred, target).correlation), float(kendalltau(pred, target).correlation)


def np_mean_pairwise(a):
    """numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src) restated, then / n in the array's
    dtype -- what np.mean(a) computes for a contiguous 1-D float array.  Used to pin the CUDA restatement."""
    a = np.ascontiguousarray(a)
    T = a.dtype.type

    def block(x):
        n = x.shape[0]
        if n < 8:
            res = T(0)
            for v in x:
                res = T(res + v)
            return res
        r = [T(v) for v in x[:8]]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = T(r[j] + x[i + j])
            i += 8
        res = T(T(T(r[0] + r[1]) + T(r[2] + r[3])) + T(T(r[4] + r[5]) + T(r[6] + r[7])))
        while i < n:
            res = T(res + x[i])
            i += 1
        return res

    def rec(x):
        n = x.shape[0]
        if n <= 128:
            return block(x)
        n2 = n // 2
        n2 -= n2 % 8
        return T(rec(x[:n2]) + rec(x[n2:]))

    return T(rec(a) / T(a.shape[0]))


def kendall_counts(x, y):
    """Integer ingredients of scipy.stats.kendalltau (tau-b) by brute force: (dis, xtie, ytie, ntie, tot)."""
    x = np.asarray(x)
    y = np.asarray(y)
    n = x.shape[0]
    dx = np.sign(x[:, None] - x[None, :])
    dy = np.sign(y[:, None] - y[None, :])
    iu = np.triu_indices(n, 1)
    dis = int(np.sum((dx * dy)[iu] < 0))
    xtie = int(np.sum(dx[iu] == 0))
    ytie = int(np.sum(dy[iu] == 0))
    ntie = int(np.sum((dx[iu] == 0) & (dy[iu] == 0)))
    return dis, xtie, ytie, ntie, n * (n - 1) // 2


def cdist_euclidean(a, b):
    """scipy cdist(a, b, 'euclidean') restated (features/fusion.py:11): float64, the feature loop summed
    sequentially (np.cumsum accumulates left to right), square root at the end."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = a[:, None, :] - b[None, :, :]
    return np.sqrt(np.cumsum(d * d, axis=-1)[..., -1])


def interpolate_features(features, path, target_length):
    """features/fusion.py:21-32 on numpy arrays (float32 rows scaled by float32(count / total))."""
    path = np.asarray(path)
    u, c = np.unique(path[:, 0], return_counts=True)
    w = c / c.sum()
    f = np.asarray(features, dtype=F32)
    return np.stack([f[i] * F32(wt) for i, wt in zip(u, w)])[:target_length]


def dtw_path(cost):
    """Exact DTW through a cost matrix: fastdtw's published `__dtw` (fastdtw 0.3.x, fastdtw.py) with the full
    window -- D[i, j] = min((D[i-1, j] + c, ...), (D[i, j-1] + c, ...), (D[i-1, j-1] + c, ...)) keyed on the sum,
    first minimum wins; D = inf outside, D[0, 0] = 0 (1-based).  PARITY UNPINNED: fastdtw is absent from this
    image and the reference's own call site (features/fusion.py:17) raises TypeError."""
    cost = np.asarray(cost, dtype=np.float64)
    n, m = cost.shape
    inf = float("inf")
    D = {(0, 0): (0.0, 0, 0)}
    get = lambda i, j: D.get((i, j), (inf,))
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            dt = float(cost[i - 1, j - 1])
            D[i, j] = min((get(i - 1, j)[0] + dt, i - 1, j), (get(i, j - 1)[0] + dt, i, j - 1),
                          (get(i - 1, j - 1)[0] + dt, i - 1, j - 1), key=lambda a: a[0])
    path = []
    i, j = n, m
    while not (i == j == 0):
        path.append((i - 1, j - 1))
        i, j = D[i, j][1], D[i, j][2]
    path.reverse()
    return D[n, m][0], np.asarray(path, dtype=np.int64)
