// Backward pass pieces for the training step of /root/reference/scripts/train_av_model.py:86-96
// (BASELINE.json configs[4]: batch-sharded data parallelism, gradients all-reduced over NCCL by the caller).
//
//   lstm_backward_kernel   BPTT through one (modality, direction) recurrence per cluster of 8 CTAs (CUDA-core
//                          fp32).  CTA r owns hidden units [32r, 32r+32): it turns dh of its units into gate
//                          gradients (recomputing the activations from the saved pre-activations), multiplies them
//                          with its register-resident 128 x 256 slice of W_hh to get a PARTIAL dh_{t-1} for all 256
//                          units, and reduce-scatters the partials through distributed shared memory.
//   transpose_kernel       [R, C] -> [C, R_pad] with tf32 round-to-nearest (operands of the tcgen05 kind::tf32
//                          GEMMs that compute dW = dY^T X and dX = dY W), optional LSTM gate-row un-permutation
//   colsum_kernel          bias gradients
//   shift_h_kernel         h_{t-1} for every frame of a recurrence (operand of dW_hh)
//
// The dense contractions of the backward pass reuse gemm_tc (C = A W^T) on transposed copies.
#include <cooperative_groups.h>

#include "common.cuh"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace avs {

namespace {

constexpr int HC = 256;
constexpr int CL = 8;
constexpr int UNITS = HC / CL;   // 32
constexpr int COLS = 4 * UNITS;  // 128
constexpr int XG_LD = 2 * 4 * HC;
constexpr int FUSED_LD = 4 * HC;

// Activations of the recomputation, on the SFU (ex2.approx + rcp.approx, ~1e-6 relative: far inside the stated
// gradient tolerances); the accurate expf / tanhf / division sequence was ~800 clk of every 3,000-clk step.
__device__ __forceinline__ float sigm(float x) {
    return __frcp_rn(1.0f + __expf(-fmaxf(x, -80.f)));
}
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = __expf(-2.0f * fminf(fmaxf(x, -40.f), 40.f));
    return __fdividef(1.0f - e, 1.0f + e);
}

template <int NB>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256, 1)
lstm_backward_kernel(const float* __restrict__ d_fused, const float4* __restrict__ save_pre,
                     const float* __restrict__ save_c, const float* __restrict__ whh, LstmBatch batch,
                     float* __restrict__ d_xg_v, float* __restrict__ d_xg_a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int r = static_cast<int>(cluster.block_rank());
    const int cid = blockIdx.x / CL;
    const int grp = cid >> 2;
    const int ld = cid & 3;
    const int dir = ld & 1;
    const int tid = threadIdx.x;

    __shared__ __align__(16) float dg[2][NB][COLS];       // gate gradients of this CTA's units (double-buffered by step)
    __shared__ __align__(16) float recv[2][CL][NB][UNITS];   // partial dh for my units, one slot per source CTA
    __shared__ __align__(8) uint64_t bar[2];              // "all partials of the step have landed in recv[i]"
    __shared__ int s_len[NB];
    __shared__ int s_row[NB];
    if (tid < NB) {
        s_len[tid] = batch.slot_len[grp * NB + tid];
        s_row[tid] = batch.slot_row_start[grp * NB + tid];
    }
    for (int i = tid; i < 2 * CL * NB * UNITS; i += blockDim.x) (&recv[0][0][0][0])[i] = 0.f;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }

    // thread k owns output unit k of the partial product: w_t[c] = W_hh[packed row r*128 + c][k]
    float w_t[COLS];
    {
        const float* src = whh + (static_cast<size_t>(ld) * 4 * HC + r * COLS) * HC + tid;
#pragma unroll
        for (int c = 0; c < COLS; ++c) w_t[c] = __ldg(src + static_cast<size_t>(c) * HC);
    }
    float* d_xg = ((ld >> 1) ? d_xg_a : d_xg_v) + dir * (4 * HC) + r * COLS;
    const int out_col = ld * HC + r * UNITS;
    const int maxlen = batch.group_maxlen[grp];

    const int jj = tid & 31;   // pointwise ownership: unit jj of video vb
    const int vb = tid >> 5;
    float dc_state = 0.f;
    __syncthreads();
    // Operands of the pointwise phase come from global memory and do not depend on the recurrence: they are
    // fetched one step ahead (registers), so their latency hides behind the previous step's partial products and
    // cluster barrier.  c_{s-1} of one step is c_s of the next: one new cell-state load per step.
    const int my_len = vb < NB ? s_len[vb] : 0;
    const size_t unit_off = static_cast<size_t>(ld) * HC + r * UNITS + jj;   // + row * 4 * HC
    const long long rstep = dir ? 1 : -1;            // frame consumed one forward step EARLIER
    long long row = 0;                               // frame consumed at the forward step being undone
    float dh_g = 0.f, c_cur = 0.f, c_nxt = 0.f;
    float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);
    if (my_len > 0) {
        row = static_cast<long long>(s_row[vb]) + (dir ? 0 : my_len - 1);
        dh_g = __ldg(d_fused + row * FUSED_LD + out_col + jj);
        pre = __ldg(save_pre + row * 4 * HC + unit_off);
        c_cur = __ldg(save_c + row * 4 * HC + unit_off);
        if (my_len > 1) c_nxt = __ldg(save_c + (row + rstep) * 4 * HC + unit_off);
    }
    cluster.sync();

    for (int u = 0; u < maxlen; ++u) {
        const int cur = u & 1, nxt = cur ^ 1;
        // ---- pointwise: dh -> gate gradients for (video vb, unit jj)
        if (vb < NB) {
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (u < my_len) {
                // prefetch the next step's operands (forward step s - 1, and the cell state of step s - 2)
                float dh_n = 0.f, c_new = 0.f;
                float4 pre_n = make_float4(0.f, 0.f, 0.f, 0.f);
                if (u + 1 < my_len) {
                    const long long rn = row + rstep;
                    dh_n = __ldg(d_fused + rn * FUSED_LD + out_col + jj);
                    pre_n = __ldg(save_pre + rn * 4 * HC + unit_off);
                    if (u + 2 < my_len) c_new = __ldg(save_c + (rn + rstep) * 4 * HC + unit_off);
                }
                float dh = dh_g;
                if (u > 0) {
                    mbar_wait_cluster(&bar[cur], ((u - 1) >> 1) & 1);   // the 8 partials of step u - 1 have landed
#pragma unroll
                    for (int src = 0; src < CL; ++src) dh += recv[cur][src][vb][jj];
                }
                const float c_t = c_cur;
                const float c_prev = (u + 1 < my_len) ? c_nxt : 0.f;   // forward step 0 starts from c = 0
                const float gi = sigm(pre.x), gf = sigm(pre.y), gg = tanh_fast(pre.z), go = sigm(pre.w);
                const float tc = tanh_fast(c_t);
                const float d_o = dh * tc;
                const float dc = dc_state + dh * go * (1.f - tc * tc);
                g4.x = dc * gg * gi * (1.f - gi);
                g4.y = dc * c_prev * gf * (1.f - gf);
                g4.z = dc * gi * (1.f - gg * gg);
                g4.w = d_o * go * (1.f - go);
                dc_state = dc * gf;
                *reinterpret_cast<float4*>(d_xg + row * XG_LD + 4 * jj) = g4;
                row += rstep;
                dh_g = dh_n;
                pre = pre_n;
                c_cur = c_nxt;
                c_nxt = c_new;
            }
            *reinterpret_cast<float4*>(&dg[cur][vb][4 * jj]) = g4;
        }
        // arm the barrier that collects this step's partials: 8 sources x 128 B per video that has an earlier step
        if (tid == 0) {
            int nact = 0;
#pragma unroll
            for (int b = 0; b < NB; ++b) nact += (u + 1 < s_len[b]) ? 1 : 0;
            if (nact > 0) mbar_expect_tx(&bar[nxt], CL * nact * UNITS * 4);
        }
        __syncthreads();   // dg[cur] complete; every warp is past its reads of recv[cur] and of dg[nxt] (step u - 1)
        // ---- partial dh_{t-1}[k] = sum_c W_hh[c][k] dg[c] over my 128 gate rows, sent to the owner of unit k:
        // warp w produces the 32 units of CTA w; four lanes' values travel as one 16-byte st.async whose
        // complete_tx counts on the owner's barrier (no cluster-wide barrier in the loop: a cluster.sync per step
        // cost ~1 us, and 4-byte remote stores ~2 us, of a 3.8 us step)
        const int dst_cta = tid >> 5, lane = tid & 31;
        const uint32_t remote_bar = mapa(smem_u32(&bar[nxt]), dst_cta);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (u + 1 < s_len[b]) {   // block-uniform: the video has an earlier step that needs dh
                const float4* gp = reinterpret_cast<const float4*>(&dg[cur][b][0]);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
#pragma unroll
                for (int c = 0; c < COLS / 4; c += 2) {   // eight independent chains: the phase is FMA-latency bound
                    const float4 g0 = gp[c], g1 = gp[c + 1];
                    a0 = fmaf(w_t[4 * c + 0], g0.x, a0);
                    a1 = fmaf(w_t[4 * c + 1], g0.y, a1);
                    a2 = fmaf(w_t[4 * c + 2], g0.z, a2);
                    a3 = fmaf(w_t[4 * c + 3], g0.w, a3);
                    a4 = fmaf(w_t[4 * c + 4], g1.x, a4);
                    a5 = fmaf(w_t[4 * c + 5], g1.y, a5);
                    a6 = fmaf(w_t[4 * c + 6], g1.z, a6);
                    a7 = fmaf(w_t[4 * c + 7], g1.w, a7);
                }
                const float v = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
                uint4 m;
                m.x = __float_as_uint(v);
                m.y = __float_as_uint(__shfl_down_sync(0xffffffffu, v, 1));
                m.z = __float_as_uint(__shfl_down_sync(0xffffffffu, v, 2));
                m.w = __float_as_uint(__shfl_down_sync(0xffffffffu, v, 3));
                if ((lane & 3) == 0)
                    st_async_v4(mapa(smem_u32(&recv[nxt][r][b][lane]), dst_cta), m, remote_bar);
            }
        }
    }
    cluster.sync();   // nobody exits while a peer's st.async could still target its shared memory
}

template <int NB>
avs_status launch_bwd(const float* d_fused, const void* save_pre, const float* save_c, const float* whh,
                      const LstmBatch& batch, float* d_xg_v, float* d_xg_a, cudaStream_t stream) {
    lstm_backward_kernel<NB><<<batch.n_groups * 4 * CL, 256, 0, stream>>>(
        d_fused, static_cast<const float4*>(save_pre), save_c, whh, batch, d_xg_v, d_xg_a);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

// dst[map(c)][r] = tf32_rn(src[r][c]) for r < R, c < C; 32 x 32 tiles through shared memory.
// perm: 0 none, 1 LSTM gate un-permutation (packed row dirblock*1024 + cta*128 + jj*4 + gate -> dirblock*1024 + gate*256 + cta*32 + jj)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, int64_t ld_src, int R, int C,
                                                        float* __restrict__ dst, int64_t ld_dst, int perm, int round) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int rr = r0 + i, cc = c0 + tx;
        tile[i][tx] = (rr < R && cc < C) ? src[static_cast<int64_t>(rr) * ld_src + cc] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int cc = c0 + i, rr = r0 + tx;
        if (cc < C && rr < R) {
            int oc = cc;
            if (perm == 1) {
                const int blk = cc >> 10, p = cc & 1023;
                oc = (blk << 10) + (p & 3) * HC + (p >> 7) * 32 + ((p >> 2) & 31);
            }
            const float v = tile[tx][i];
            dst[static_cast<int64_t>(oc) * ld_dst + rr] = round ? to_tf32_rn(v) : v;
        }
    }
}

// out[map(c)] = sum_r src[r][c]   (one block per 32 columns, 8 row lanes, fp32 tree)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ src, int64_t ld_src, int R, int C,
                                                     float* __restrict__ out, int perm) {
    __shared__ float part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int cc = blockIdx.x * 32 + tx;
    float acc = 0.f;
    if (cc < C)
        for (int rr = ty; rr < R; rr += 8) acc += src[static_cast<int64_t>(rr) * ld_src + cc];
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && cc < C) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += part[i][tx];
        int oc = cc;
        if (perm == 1) {
            const int blk = cc >> 10, p = cc & 1023;
            oc = (blk << 10) + (p & 3) * HC + (p >> 7) * 32 + ((p >> 2) & 31);
        }
        out[oc] = s;
    }
}

// Two-pass form for tall matrices (deterministic, no atomics): grid.y row slices write partial sums [S][C], a second
// launch adds the S partials in a fixed order.  (One block per 32 columns alone takes ~36 us for 2,560 x 2,048.)
__global__ void __launch_bounds__(256) colsum_part_kernel(const float* __restrict__ src, int64_t ld_src, int R, int C,
                                                          float* __restrict__ part_out) {
    __shared__ float part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int cc = blockIdx.x * 32 + tx;
    const int chunk = (R + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * chunk, r1 = min(R, r0 + chunk);
    float acc = 0.f;
    if (cc < C)
        for (int rr = r0 + ty; rr < r1; rr += 8) acc += src[static_cast<int64_t>(rr) * ld_src + cc];
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && cc < C) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += part[i][tx];
        part_out[static_cast<int64_t>(blockIdx.y) * C + cc] = s;
    }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ part, int S, int C,
                                                           float* __restrict__ out, int perm) {
    const int cc = blockIdx.x * 256 + threadIdx.x;
    if (cc >= C) return;
    float s = 0.f;
    for (int i = 0; i < S; ++i) s += part[static_cast<int64_t>(i) * C + cc];
    int oc = cc;
    if (perm == 1) {
        const int blk = cc >> 10, p = cc & 1023;
        oc = (blk << 10) + (p & 3) * HC + (p >> 7) * 32 + ((p >> 2) & 31);
    }
    out[oc] = s;
}

// hprev[row][ld*256 + j] = h of the forward step before the one that consumed `row` (0 at a sequence start):
// forward directions (ld even) read fused[row - 1], reverse directions (ld odd) fused[row + 1].
__global__ void shift_h_kernel(const float* __restrict__ fused, const int32_t* __restrict__ row_start,
                               const int32_t* __restrict__ lengths, int n_videos, float* __restrict__ hprev) {
    const int v = blockIdx.y;
    const int len = lengths[v];
    const int64_t base = row_start[v];
    for (int t = blockIdx.x; t < len; t += gridDim.x) {
        for (int c = threadIdx.x; c < FUSED_LD; c += blockDim.x) {
            const int ld = c >> 8;
            const int tp = (ld & 1) ? t + 1 : t - 1;
            hprev[(base + t) * FUSED_LD + c] = (tp >= 0 && tp < len) ? fused[(base + tp) * FUSED_LD + c] : 0.f;
        }
    }
}

}  // namespace

avs_status lstm_backward(const float* d_fused, const void* save_pre, const float* save_c, const float* whh_packed,
                         const LstmBatch& batch, float* d_xg_v, float* d_xg_a, cudaStream_t stream) {
    if (batch.n_groups == 0) return AVS_OK;
    switch (batch.nb) {
        case 1: return launch_bwd<1>(d_fused, save_pre, save_c, whh_packed, batch, d_xg_v, d_xg_a, stream);
        case 2: return launch_bwd<2>(d_fused, save_pre, save_c, whh_packed, batch, d_xg_v, d_xg_a, stream);
        case 4: return launch_bwd<4>(d_fused, save_pre, save_c, whh_packed, batch, d_xg_v, d_xg_a, stream);
        case 8: return launch_bwd<8>(d_fused, save_pre, save_c, whh_packed, batch, d_xg_v, d_xg_a, stream);
        default: set_error("lstm_backward: unsupported videos-per-cluster %d", batch.nb); return AVS_ERR_INVALID;
    }
}

avs_status transpose_f32(const float* src, int64_t ld_src, int R, int C, float* dst, int64_t ld_dst, int perm, int round,
                         cudaStream_t stream) {
    if (R == 0 || C == 0) return AVS_OK;
    dim3 grid((C + 31) / 32, (R + 31) / 32);
    transpose_kernel<<<grid, 256, 0, stream>>>(src, ld_src, R, C, dst, ld_dst, perm, round);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

avs_status colsum_f32(const float* src, int64_t ld_src, int R, int C, float* out, int perm, cudaStream_t stream) {
    if (C == 0) return AVS_OK;
    if (R < 512) {
        colsum_kernel<<<(C + 31) / 32, 256, 0, stream>>>(src, ld_src, R, C, out, perm);
        AVS_LAUNCH_CHECK();
        return AVS_OK;
    }
    const int S = 16;
    float* part = nullptr;
    AVS_CUDA(cudaMallocAsync(&part, static_cast<size_t>(S) * C * 4, stream));
    colsum_part_kernel<<<dim3((C + 31) / 32, S), 256, 0, stream>>>(src, ld_src, R, C, part);
    colsum_final_kernel<<<(C + 255) / 256, 256, 0, stream>>>(part, S, C, out, perm);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(part, stream);
    if (e != cudaSuccess) {
        set_error("colsum launch failed: %s", cudaGetErrorString(e));
        return AVS_ERR_CUDA;
    }
    count_launch();
    count_launch();
    return AVS_OK;
}

avs_status shift_h(const float* fused, const int32_t* row_start, const int32_t* lengths, int n_videos, int max_len,
                   float* hprev, cudaStream_t stream) {
    if (n_videos == 0 || max_len == 0) return AVS_OK;
    dim3 grid(std::min(max_len, 256), n_videos);
    shift_h_kernel<<<grid, 256, 0, stream>>>(fused, row_start, lengths, n_videos, hprev);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

}  // namespace avs
