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

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

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

    __shared__ __align__(16) float dg[NB][COLS];          // gate gradients of this CTA's units, this step
    __shared__ float recv[2][CL][NB][UNITS];              // partial dh for my units, one slot per source CTA
    __shared__ int s_len[NB];
    __shared__ int s_row[NB];
    if (tid < NB) {
        s_len[tid] = batch.slot_len[grp * NB + tid];
        s_row[tid] = batch.slot_row_start[grp * NB + tid];
    }
    for (int i = tid; i < 2 * CL * NB * UNITS; i += blockDim.x) (&recv[0][0][0][0])[i] = 0.f;

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
    cluster.sync();

    for (int u = 0; u < maxlen; ++u) {
        const int cur = u & 1, nxt = cur ^ 1;
        // ---- pointwise: dh -> gate gradients for (video vb, unit jj)
        if (vb < NB) {
            const int len = s_len[vb];
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (u < len) {
                const int s = len - 1 - u;                       // forward step being undone
                const int t = dir ? (len - 1 - s) : s;           // frame consumed at that step
                const size_t row = static_cast<size_t>(s_row[vb]) + t;
                float dh = __ldg(d_fused + row * FUSED_LD + out_col + jj);
                if (u > 0) {
#pragma unroll
                    for (int src = 0; src < CL; ++src) dh += recv[cur][src][vb][jj];
                }
                const size_t o = (row * 4 + ld) * HC + r * UNITS + jj;
                const float4 pre = __ldg(save_pre + o);
                const float c_t = __ldg(save_c + o);
                float c_prev = 0.f;
                if (s > 0) {
                    const size_t row_p = dir ? row + 1 : row - 1;   // frame consumed at forward step s - 1
                    c_prev = __ldg(save_c + (row_p * 4 + ld) * HC + r * UNITS + jj);
                }
                const float gi = sigm(pre.x), gf = sigm(pre.y), gg = tanhf(pre.z), go = sigm(pre.w);
                const float tc = tanhf(c_t);
                const float d_o = dh * tc;
                const float dc = dc_state + dh * go * (1.f - tc * tc);
                g4.x = dc * gg * gi * (1.f - gi);
                g4.y = dc * c_prev * gf * (1.f - gf);
                g4.z = dc * gi * (1.f - gg * gg);
                g4.w = d_o * go * (1.f - go);
                dc_state = dc * gf;
                *reinterpret_cast<float4*>(d_xg + row * XG_LD + 4 * jj) = g4;
            }
            *reinterpret_cast<float4*>(&dg[vb][4 * jj]) = g4;
        }
        __syncthreads();
        // ---- partial dh_{t-1}[k] = sum_c W_hh[c][k] dg[c] over my 128 gate rows; scatter to the owner of unit k
        const int dst_cta = tid >> 5, dst_jj = tid & 31;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (u + 1 < s_len[b]) {   // block-uniform: the video has an earlier step that needs dh
                const float4* gp = reinterpret_cast<const float4*>(&dg[b][0]);
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int c = 0; c < COLS / 4; c += 2) {
                    const float4 g0 = gp[c], g1 = gp[c + 1];
                    a0 = fmaf(w_t[4 * c + 0], g0.x, a0);
                    a0 = fmaf(w_t[4 * c + 1], g0.y, a0);
                    a0 = fmaf(w_t[4 * c + 2], g0.z, a0);
                    a0 = fmaf(w_t[4 * c + 3], g0.w, a0);
                    a1 = fmaf(w_t[4 * c + 4], g1.x, a1);
                    a1 = fmaf(w_t[4 * c + 5], g1.y, a1);
                    a1 = fmaf(w_t[4 * c + 6], g1.z, a1);
                    a1 = fmaf(w_t[4 * c + 7], g1.w, a1);
                }
                *cluster.map_shared_rank(&recv[nxt][r][b][dst_jj], dst_cta) = a0 + a1;
            }
        }
        cluster.sync();   // partials visible cluster-wide; also orders the reuse of recv[cur] and dg
    }
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
    colsum_kernel<<<(C + 31) / 32, 256, 0, stream>>>(src, ld_src, R, C, out, perm);
    AVS_LAUNCH_CHECK();
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
