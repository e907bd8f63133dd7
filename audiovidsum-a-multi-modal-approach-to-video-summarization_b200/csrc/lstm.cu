// K2b -- the recurrent half of both nn.LSTM modules of AVBiLSTMModel
// (/root/reference/models/av_model.py:18-23, 39-40; PyTorch LSTM semantics: gate order i,f,g,o,
// zero initial state, reverse direction walks each video from its own last frame).
//
// One thread-block CLUSTER of 8 CTAs runs one (modality, direction) recurrence for a group of up
// to NB videos (CUDA-core fp32 version: the AVS_PREC_FP32_SIMT path; lstm_tc.cu is the tensor-core
// one).  CTA r owns hidden units [32r, 32r+32) = 128 gate columns; its 128 x 256 slice of
// W_hh lives in REGISTERS for the whole kernel (128 floats per thread, 256 threads), the hidden
// state of all NB videos lives in shared memory and is re-broadcast to the 8 CTAs through
// distributed shared memory after every step.  The input projections (x W_ih^T + b_ih + b_hh)
// were computed for all frames at once by the tcgen05 GEMM and arrive as `xg`.
//
// This kernel is latency bound (T dependent steps); throughput comes from running the four
// recurrences of every video group concurrently across the chip.
#include <cooperative_groups.h>

#include "common.cuh"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace avs {

namespace {

constexpr int HC = 256;        // hidden units per direction
constexpr int CL = 8;          // CTAs per cluster
constexpr int UNITS = HC / CL; // hidden units per CTA (32)
constexpr int COLS = 4 * UNITS;// gate columns per CTA (128)
constexpr int XG_LD = 2 * 4 * HC;  // 2048: both directions of one modality
constexpr int FUSED_LD = 4 * HC;   // 1024

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int NB>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256, 1)
lstm_recurrence_kernel(const float* __restrict__ xg_v, const float* __restrict__ xg_a,
                       const float* __restrict__ whh, LstmBatch batch, float* __restrict__ fused, int round_tf32,
                       void* __restrict__ fused_lowp, int lowp_dtype) {
    cg::cluster_group cluster = cg::this_cluster();
    const int r = static_cast<int>(cluster.block_rank());
    const int cid = blockIdx.x / CL;
    const int grp = cid >> 2;
    const int ld = cid & 3;       // 0 visual fwd, 1 visual reverse, 2 audio fwd, 3 audio reverse
    const int dir = ld & 1;
    const int tid = threadIdx.x;
    const int c = tid & (COLS - 1);   // gate column inside this CTA's slice
    const int kh = tid >> 7;          // which half of the K = 256 reduction this thread owns

    __shared__ __align__(16) float h_buf[2][NB][HC];
    __shared__ float gates[NB][COLS];  // K-half partial sums first, then the finished gate values
    __shared__ int s_len[NB];
    __shared__ int s_row[NB];

    if (tid < NB) {
        s_len[tid] = batch.slot_len[grp * NB + tid];
        s_row[tid] = batch.slot_row_start[grp * NB + tid];
    }
    for (int i = tid; i < 2 * NB * HC; i += blockDim.x) (&h_buf[0][0][0])[i] = 0.f;

    // W_hh slice -> registers: packed row (ld, r*128 + c), K range [kh*128, kh*128+128)
    float w[128];
    {
        const float4* src = reinterpret_cast<const float4*>(
            whh + (static_cast<size_t>(ld) * 4 * HC + r * COLS + c) * HC + kh * 128);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float4 v = __ldg(src + i);
            w[4 * i + 0] = v.x;
            w[4 * i + 1] = v.y;
            w[4 * i + 2] = v.z;
            w[4 * i + 3] = v.w;
        }
    }
    const float* xg = ((ld >> 1) ? xg_a : xg_v) + dir * (4 * HC) + r * COLS + c;
    const int out_col = ld * HC + r * UNITS;  // + jj
    const int maxlen = batch.group_maxlen[grp];

    // pointwise ownership: hidden unit jj of videos vb, vb + 8, ...
    const int jj = tid & 31;
    const int vb = tid >> 5;
    constexpr int PW = (NB + 7) / 8;
    float c_state[PW];
#pragma unroll
    for (int i = 0; i < PW; ++i) c_state[i] = 0.f;

    __syncthreads();
    cluster.sync();  // every CTA's h_buf is zeroed before anyone writes into it remotely

    for (int s = 0; s < maxlen; ++s) {
        const int cur = s & 1, nxt = cur ^ 1;
        // gate pre-activations of this step (latency hidden behind the FMA phase)
        float xv[NB];
        if (kh == 0) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int len = s_len[b];
                if (s < len) {
                    const int t = dir ? (len - 1 - s) : s;
                    xv[b] = __ldg(xg + static_cast<size_t>(s_row[b] + t) * XG_LD);
                } else {
                    xv[b] = 0.f;
                }
            }
        }
        // h_prev . W_hh^T for this thread's column and K half
        float acc[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[b] = 0.f;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (s < s_len[b]) {  // block-uniform
                const float4* hp = reinterpret_cast<const float4*>(&h_buf[cur][b][kh * 128]);
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float4 h0 = hp[i];
                    const float4 h1 = hp[i + 1];
                    a0 = fmaf(w[4 * i + 0], h0.x, a0);
                    a0 = fmaf(w[4 * i + 1], h0.y, a0);
                    a0 = fmaf(w[4 * i + 2], h0.z, a0);
                    a0 = fmaf(w[4 * i + 3], h0.w, a0);
                    a1 = fmaf(w[4 * i + 4], h1.x, a1);
                    a1 = fmaf(w[4 * i + 5], h1.y, a1);
                    a1 = fmaf(w[4 * i + 6], h1.z, a1);
                    a1 = fmaf(w[4 * i + 7], h1.w, a1);
                }
                acc[b] = a0 + a1;
            }
        }
        if (kh == 1) {
#pragma unroll
            for (int b = 0; b < NB; ++b) gates[b][c] = acc[b];
        }
        __syncthreads();
        if (kh == 0) {
#pragma unroll
            for (int b = 0; b < NB; ++b) gates[b][c] = acc[b] + gates[b][c] + xv[b];
        }
        __syncthreads();
        // pointwise cell update + broadcast of the new hidden slice to all 8 CTAs
#pragma unroll
        for (int i = 0; i < PW; ++i) {
            const int b = vb + 8 * i;
            if (b < NB) {
                const int len = s_len[b];
                if (s < len) {
                    const float gi = sigmoid_acc(gates[b][4 * jj + 0]);  // packed column = 4*jj + gate
                    const float gf = sigmoid_acc(gates[b][4 * jj + 1]);
                    const float gg = tanhf(gates[b][4 * jj + 2]);
                    const float go = sigmoid_acc(gates[b][4 * jj + 3]);
                    const float cn = fmaf(gf, c_state[i], gi * gg);
                    c_state[i] = cn;
                    const float h = go * tanhf(cn);
                    const int t = dir ? (len - 1 - s) : s;
                    const size_t o = static_cast<size_t>(s_row[b] + t) * FUSED_LD + out_col + jj;
                    fused[o] = round_tf32 ? to_tf32_rn(h) : h;
                    if (fused_lowp != nullptr) {
                        if (lowp_dtype == DT_F16)
                            reinterpret_cast<__half*>(fused_lowp)[o] = __float2half_rn(h);
                        else
                            reinterpret_cast<__nv_bfloat16*>(fused_lowp)[o] = __float2bfloat16_rn(h);
                    }
                    float* local = &h_buf[nxt][b][r * UNITS + jj];
#pragma unroll
                    for (int rr = 0; rr < CL; ++rr) *cluster.map_shared_rank(local, rr) = h;
                }
            }
        }
        cluster.sync();  // new hidden state visible cluster-wide; also orders reuse of h_buf[cur]
    }
}

template <int NB>
avs_status launch_nb(const float* xg_v, const float* xg_a, const float* whh, const LstmBatch& batch, float* fused,
                     int round_tf32, void* fused_lowp, int lowp_dtype, cudaStream_t stream) {
    const int blocks = batch.n_groups * 4 * CL;
    lstm_recurrence_kernel<NB><<<blocks, 256, 0, stream>>>(xg_v, xg_a, whh, batch, fused, round_tf32, fused_lowp,
                                                          lowp_dtype);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

}  // namespace

avs_status lstm_recurrence(const float* xg_v, const float* xg_a, const float* whh_packed, const LstmBatch& batch,
                           float* fused, int round_tf32, void* fused_lowp, int lowp_dtype, cudaStream_t stream) {
    if (batch.n_groups == 0) return AVS_OK;
    switch (batch.nb) {
        case 1: return launch_nb<1>(xg_v, xg_a, whh_packed, batch, fused, round_tf32, fused_lowp, lowp_dtype, stream);
        case 2: return launch_nb<2>(xg_v, xg_a, whh_packed, batch, fused, round_tf32, fused_lowp, lowp_dtype, stream);
        case 4: return launch_nb<4>(xg_v, xg_a, whh_packed, batch, fused, round_tf32, fused_lowp, lowp_dtype, stream);
        case 8: return launch_nb<8>(xg_v, xg_a, whh_packed, batch, fused, round_tf32, fused_lowp, lowp_dtype, stream);
        case 16: return launch_nb<16>(xg_v, xg_a, whh_packed, batch, fused, round_tf32, fused_lowp, lowp_dtype, stream);
        default: set_error("lstm: unsupported videos-per-cluster %d", batch.nb); return AVS_ERR_INVALID;
    }
}

}  // namespace avs
