// Callers and helpers on either side of the forward pass:
//
//   eval_metrics_kernel   the metric block of /root/reference/scripts/evaluate.py:25-36 for a batch of videos:
//                         mean-threshold F1 (np.mean pairwise summation restated exactly), Spearman's rho on
//                         average ranks and Kendall's tau-b (scipy.stats, integer pair counts -> bit-exact tau)
//   cdist_kernel          features/fusion.py:7-12  compute_dtw == scipy cdist(..., "euclidean"), float64, the
//                         k-loop in scipy's summation order (sequential, no FMA contraction) -> bit-exact
//   gather_scale_kernel   features/fusion.py:21-32 interpolate_features: out[k, :] = features[u_k, :] * w_k
//   dtw_kernel            exact DTW over a cost matrix (the evident intent of features/fusion.py:15-18, whose
//                         fastdtw call is broken in the reference): anti-diagonal wavefront, fastdtw's
//                         published recurrence and tie order, float64 -> bit-exact against the oracle
//
// All of this is integer / fp64 / HBM-bound work: no tensor cores, block-per-video or tile-per-block grids.
#include <cmath>

#include "common.cuh"

namespace avs {

namespace {

// ---------------------------------------------------------------------------------------------
// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum_@TYPE@), which is
// what np.mean(a) runs for a contiguous 1-D array: blocks of <= 128 elements are summed with 8
// interleaved accumulators, larger ranges are split in halves (first half rounded down to a multiple of 8).
template <typename T>
__device__ T pairwise_block(const T* a, int n) {   // n <= 128
    if (n < 8) {
        T res = 0;
        for (int i = 0; i < n; ++i) res = res + a[i];
        return res;
    }
    T r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] = r[j] + a[i + j];
    T res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res = res + a[i];
    return res;
}
template <typename T>
__device__ T pairwise_sum(const T* a, int n) {
    // iterative form of the recursion: post-order traversal with an explicit stack of (offset, length, state)
    struct Frame { int off, n, state; T left; };
    Frame st[40];
    int sp = 0;
    st[sp++] = {0, n, 0, T(0)};
    T ret = 0;
    while (sp > 0) {
        Frame& f = st[sp - 1];
        if (f.n <= 128) {
            ret = pairwise_block(a + f.off, f.n);
            --sp;
            continue;
        }
        int n2 = f.n / 2;
        n2 -= n2 % 8;
        if (f.state == 0) {
            f.state = 1;
            st[sp++] = {f.off, n2, 0, T(0)};
        } else if (f.state == 1) {
            f.left = ret;
            f.state = 2;
            st[sp++] = {f.off + n2, f.n - n2, 0, T(0)};
        } else {
            ret = f.left + ret;
            --sp;
        }
    }
    return ret;
}

__device__ __forceinline__ long long block_reduce_ll(long long v, long long* scratch) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    long long tot = 0;
    for (int w = 0; w < nw; ++w) tot += scratch[w];
    return tot;
}

// One block per video.  pred: fp32 scores; target: fp32 or fp64 annotations (tgt_f64), both indexed by row.
// out_f[v*4 + {0,1,2,3}] = f1, spearman, kendall, np.mean(pred);  out_i[v*8 + ...] = tp, n_pred, n_target, dis, xtie, ytie, ntie, n.
__global__ void __launch_bounds__(256) eval_metrics_kernel(const float* __restrict__ pred, const void* __restrict__ target,
                                                           int tgt_f64, const int32_t* __restrict__ row_start,
                                                           const int32_t* __restrict__ lengths, double* __restrict__ out_f,
                                                           long long* __restrict__ out_i) {
    const int v = blockIdx.x;
    const int n = lengths[v];
    const size_t r0 = static_cast<size_t>(row_start[v]);
    const float* x = pred + r0;
    const float* yf = static_cast<const float*>(target) + r0;
    const double* yd = static_cast<const double*>(target) + r0;
    __shared__ float s_mean_x;
    __shared__ double s_mean_y;
    __shared__ long long scratch[8];
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);

    // ---- np.mean in the array's own dtype (evaluate.py:26-27)
    if (threadIdx.x == 0) s_mean_x = n > 0 ? __fdiv_rn(pairwise_sum<float>(x, n), static_cast<float>(n)) : nanf("");
    if (threadIdx.x == 32) {
        if (n == 0) s_mean_y = qnan;
        else if (tgt_f64) s_mean_y = __ddiv_rn(pairwise_sum<double>(yd, n), static_cast<double>(n));
        else s_mean_y = static_cast<double>(__fdiv_rn(pairwise_sum<float>(yf, n), static_cast<float>(n)));
    }
    __syncthreads();
    const float mx = s_mean_x;
    const double my = s_mean_y;   // a float mean widened exactly; comparisons below are equivalent

    long long tp = 0, np_ = 0, nt = 0, dis2 = 0, xt2 = 0, yt2 = 0, jt2 = 0;
    long long sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
    int has_nan = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float xi = x[i];
        const double yi = tgt_f64 ? yd[i] : static_cast<double>(yf[i]);
        const int bp = xi > mx, bt = yi > my;
        tp += bp & bt;
        np_ += bp;
        nt += bt;
        has_nan |= (xi != xi) | (yi != yi);
        int lx = 0, ex = 0, ly = 0, ey = 0, joint = 0, disc = 0;
        for (int j = 0; j < n; ++j) {
            const float xj = x[j];
            const double yj = tgt_f64 ? yd[j] : static_cast<double>(yf[j]);
            const int xl = xj < xi, xe = xj == xi, xg = xj > xi;
            const int yl = yj < yi, ye = yj == yi, yg = yj > yi;
            lx += xl;
            ex += xe;
            ly += yl;
            ey += ye;
            joint += xe & ye;
            disc += (xl & yg) | (xg & yl);
        }
        // doubled average ranks (scipy rankdata 'average'): 2 * rank = 2 * #less + #equal + 1
        const long long a = 2ll * lx + ex + 1, b = 2ll * ly + ey + 1;
        sa += a;
        sb += b;
        saa += a * a;
        sbb += b * b;
        sab += a * b;
        dis2 += disc;
        xt2 += ex - 1;
        yt2 += ey - 1;
        jt2 += joint - 1;
    }
    tp = block_reduce_ll(tp, scratch);
    np_ = block_reduce_ll(np_, scratch);
    nt = block_reduce_ll(nt, scratch);
    dis2 = block_reduce_ll(dis2, scratch);
    xt2 = block_reduce_ll(xt2, scratch);
    yt2 = block_reduce_ll(yt2, scratch);
    jt2 = block_reduce_ll(jt2, scratch);
    sa = block_reduce_ll(sa, scratch);
    sb = block_reduce_ll(sb, scratch);
    saa = block_reduce_ll(saa, scratch);
    sbb = block_reduce_ll(sbb, scratch);
    sab = block_reduce_ll(sab, scratch);
    const long long any_nan = block_reduce_ll(has_nan, scratch);
    if (threadIdx.x != 0) return;

    // ---- evaluate.py:29-33 (numpy int / int -> float64 true division; 0/0 = nan, k/0 = inf)
    const double precision = __ddiv_rn(static_cast<double>(tp), static_cast<double>(np_));
    const double recall = __ddiv_rn(static_cast<double>(tp), static_cast<double>(nt));
    const double f1 = __ddiv_rn(__dmul_rn(2.0, __dmul_rn(precision, recall)),
                                __dadd_rn(__dadd_rn(precision, recall), 1e-8));
    // ---- Spearman (evaluate.py:35): Pearson correlation of the average ranks, from exact integer moments
    const long long nn = n;
    double rho = qnan, tau = qnan;
    const long long dis = dis2 / 2, xtie = xt2 / 2, ytie = yt2 / 2, ntie = jt2 / 2;
    if (!any_nan && n >= 2) {
        const long long cov = nn * sab - sa * sb, vx = nn * saa - sa * sa, vy = nn * sbb - sb * sb;
        if (vx > 0 && vy > 0)
            rho = __ddiv_rn(static_cast<double>(cov), __dmul_rn(sqrt(static_cast<double>(vx)), sqrt(static_cast<double>(vy))));
        // ---- Kendall tau-b (evaluate.py:36; scipy.stats.kendalltau variant 'b')
        const long long tot = nn * (nn - 1) / 2;
        if (xtie != tot && ytie != tot) {
            const long long cmd = tot - xtie - ytie + ntie - 2 * dis;
            tau = __ddiv_rn(__ddiv_rn(static_cast<double>(cmd), sqrt(static_cast<double>(tot - xtie))),
                            sqrt(static_cast<double>(tot - ytie)));
            tau = fmin(1.0, fmax(-1.0, tau));
        }
    }
    out_f[v * 4 + 0] = f1;
    out_f[v * 4 + 1] = rho;
    out_f[v * 4 + 2] = tau;
    out_f[v * 4 + 3] = static_cast<double>(mx);
    long long* oi = out_i + v * 8;
    oi[0] = tp;
    oi[1] = np_;
    oi[2] = nt;
    oi[3] = dis;
    oi[4] = xtie;
    oi[5] = ytie;
    oi[6] = ntie;
    oi[7] = nn;
}

// ---------------------------------------------------------------------------------------------
// cdist: 16 x 16 outputs per block, K staged through shared memory in chunks of 64 floats;
// every thread walks k = 0 .. D-1 in order with separate multiply and add (scipy's scalar loop).
constexpr int CT = 16, CK = 64;
__global__ void __launch_bounds__(CT * CT) cdist_kernel(const float* __restrict__ a, const float* __restrict__ b, int na,
                                                        int nb, int D, double* __restrict__ out) {
    __shared__ float sa[CT][CK + 1];
    __shared__ float sb[CT][CK + 1];
    const int tx = threadIdx.x % CT, ty = threadIdx.x / CT;
    const int i = blockIdx.y * CT + ty, j = blockIdx.x * CT + tx;
    double s = 0.0;
    for (int k0 = 0; k0 < D; k0 += CK) {
        for (int t = threadIdx.x; t < CT * CK; t += CT * CT) {
            const int r = t / CK, c = t % CK;
            const int gi = blockIdx.y * CT + r, gj = blockIdx.x * CT + r;
            sa[r][c] = (gi < na && k0 + c < D) ? a[static_cast<size_t>(gi) * D + k0 + c] : 0.f;
            sb[r][c] = (gj < nb && k0 + c < D) ? b[static_cast<size_t>(gj) * D + k0 + c] : 0.f;
        }
        __syncthreads();
        const int kc = min(CK, D - k0);
        for (int c = 0; c < kc; ++c) {
            const double d = __dsub_rn(static_cast<double>(sa[ty][c]), static_cast<double>(sb[tx][c]));
            s = __dadd_rn(s, __dmul_rn(d, d));
        }
        __syncthreads();
    }
    if (i < na && j < nb) out[static_cast<size_t>(i) * nb + j] = sqrt(s);
}

// out[k, :] = features[idx[k], :] * w[k]   (float32 multiply, as torch does for tensor * scalar)
__global__ void gather_scale_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx,
                                    const float* __restrict__ w, int U, int D, float* __restrict__ out) {
    const int k = blockIdx.x;
    const float wk = w[k];
    const float* src = feat + static_cast<size_t>(idx[k]) * D;
    float* dst = out + static_cast<size_t>(k) * D;
    if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        for (int c = threadIdx.x; c < D / 4; c += blockDim.x) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src) + c);
            reinterpret_cast<float4*>(dst)[c] = make_float4(__fmul_rn(v.x, wk), __fmul_rn(v.y, wk), __fmul_rn(v.z, wk),
                                                            __fmul_rn(v.w, wk));
        }
    } else {
        for (int c = threadIdx.x; c < D; c += blockDim.x) dst[c] = __fmul_rn(src[c], wk);
    }
}

// Exact DTW over a cost matrix [n, m] (float64).  acc uses fastdtw's recurrence
//   D[i, j] = min(D[i-1, j] + c, D[i, j-1] + c, D[i-1, j-1] + c), first minimum in that order on ties,
// with D = inf outside the matrix and D[-1, -1] = 0.  One block sweeps the anti-diagonals; the choice
// (0: from (i-1, j), 1: from (i, j-1), 2: from (i-1, j-1)) is recorded for the back-trace.
__global__ void __launch_bounds__(1024) dtw_kernel(const double* __restrict__ cost, int n, int m, double* __restrict__ acc,
                                                   uint8_t* __restrict__ choice, int32_t* __restrict__ path,
                                                   int32_t* __restrict__ path_len, double* __restrict__ total) {
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    for (int d = 0; d < n + m - 1; ++d) {
        const int i_lo = max(0, d - (m - 1)), i_hi = min(n - 1, d);
        for (int i = i_lo + threadIdx.x; i <= i_hi; i += blockDim.x) {
            const int j = d - i;
            const double c = cost[static_cast<size_t>(i) * m + j];
            const double up = i > 0 ? acc[static_cast<size_t>(i - 1) * m + j] : inf;
            const double left = j > 0 ? acc[static_cast<size_t>(i) * m + j - 1] : inf;
            const double diag = (i > 0 && j > 0) ? acc[static_cast<size_t>(i - 1) * m + j - 1] : ((i == 0 && j == 0) ? 0.0 : inf);
            const double s0 = __dadd_rn(up, c), s1 = __dadd_rn(left, c), s2 = __dadd_rn(diag, c);
            double best = s0;
            uint8_t ch = 0;
            if (s1 < best) { best = s1; ch = 1; }
            if (s2 < best) { best = s2; ch = 2; }
            acc[static_cast<size_t>(i) * m + j] = best;
            choice[static_cast<size_t>(i) * m + j] = ch;
        }
        __threadfence_block();
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // back-trace from (n-1, m-1); written reversed, then flipped in place
        int i = n - 1, j = m - 1, len = 0;
        while (i >= 0 && j >= 0) {
            path[2 * len] = i;
            path[2 * len + 1] = j;
            ++len;
            if (i == 0 && j == 0) break;
            const uint8_t ch = choice[static_cast<size_t>(i) * m + j];
            if (ch == 0) --i;
            else if (ch == 1) --j;
            else { --i; --j; }
        }
        for (int a = 0, b = len - 1; a < b; ++a, --b) {
            const int t0 = path[2 * a], t1 = path[2 * a + 1];
            path[2 * a] = path[2 * b];
            path[2 * a + 1] = path[2 * b + 1];
            path[2 * b] = t0;
            path[2 * b + 1] = t1;
        }
        *path_len = len;
        *total = acc[static_cast<size_t>(n - 1) * m + (m - 1)];
    }
}

}  // namespace

avs_status eval_metrics_device(const float* pred, const void* target, int tgt_f64, const int32_t* row_start,
                               const int32_t* lengths, int n_videos, double* out_f, long long* out_i,
                               cudaStream_t stream) {
    if (n_videos == 0) return AVS_OK;
    eval_metrics_kernel<<<n_videos, 256, 0, stream>>>(pred, target, tgt_f64, row_start, lengths, out_f, out_i);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

avs_status cdist_device(const float* a, const float* b, int na, int nb, int D, double* out, cudaStream_t stream) {
    if (na == 0 || nb == 0) return AVS_OK;
    dim3 grid((nb + CT - 1) / CT, (na + CT - 1) / CT);
    cdist_kernel<<<grid, CT * CT, 0, stream>>>(a, b, na, nb, D, out);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

avs_status gather_scale_device(const float* feat, const int32_t* idx, const float* w, int U, int D, float* out,
                               cudaStream_t stream) {
    if (U == 0 || D == 0) return AVS_OK;
    gather_scale_kernel<<<U, 256, 0, stream>>>(feat, idx, w, U, D, out);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

avs_status dtw_device(const double* cost, int n, int m, double* acc, uint8_t* choice, int32_t* path, int32_t* path_len,
                      double* total, cudaStream_t stream) {
    dtw_kernel<<<1, 1024, 0, stream>>>(cost, n, m, acc, choice, path, path_len, total);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

}  // namespace avs
