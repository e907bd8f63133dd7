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
#include <cstdlib>

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

// ------------------------------------------------------------------------------------------------ BPTT on tcgen05
// The same decomposition as the CUDA-core kernel above -- CTA r owns hidden units [32r, 32r+32), turns dh of its
// units into the 128 gate gradients of those units, and contributes, for ALL 256 units, the partial product over its
// own 128 gate rows of W_hh; the partials are reduce-scattered to the units' owners through DSMEM -- with the
// partial product on the tensor core:
//   partial[256 units, videos] = W_hh[packed rows 128r .. 128r+127, :]^T  *  dgates[128, videos]
//     A: W_hh^T slice, tf32 (RN), RESIDENT IN TENSOR MEMORY for the whole kernel: two M = 128 tiles x 128 columns
//        (lane = hidden unit, one 32-bit column per gate row)
//     B: the step's gate gradients, tf32 (RN), shared memory, K-major no-swizzle [k / 4][video][k % 4]: the thread
//        that owns (unit jj, video v) writes its four gate gradients (k = 4 jj + gate) as ONE 16-byte chunk
//     D: tensor memory, 2 tiles x 16 columns, double-buffered by step parity (a slow warp of this CTA may still be
//        reading step u's partials when the MMAs of step u + 1 start)
//   2 x 16 tcgen05.mma (kind::tf32, M = 128, N = 16, K = 8) per step instead of 128 FMAs per thread and video.
// Eight warps.  Warp w reads, after the MMAs, the 32 units that CTA w owns (tile w / 4, TMEM lane quarter w % 4) and
// sends them there with st.async, counted in bytes on the owner's mbarrier -- no CTA-wide or cluster-wide barrier in
// the loop; the owner adds the eight partials in a fixed order (deterministic).  The exchange is bound by the
// MESSAGE RATE an SM can receive (~2.5-3.5 clk per st.async whatever its size, as in the forward recurrence:
// 256 eight-byte messages took ~900 clk, 128 sixteen-byte ones ~450), so neighbouring lanes first merge their
// values into 16-byte messages with shuffles.  The first NBV warps are also the pointwise warps (unit = lane, video =
// warp), and the elected thread of warp 0 issues the MMAs right after a named barrier among them (a separate
// issuing warp behind an mbarrier cost ~330 clk per step from "operand staged" to "issuer awake").
// Per step: [partials landed] -> sum + 4 multiplies (everything that does not depend on dh -- the recomputed
// activations and their derivative factors -- is evaluated BEFORE the wait) -> B operand -> 32 MMAs -> tcgen05.ld ->
// st.async -> DSMEM.  tf32 operands are what every other GEMM of the backward pass uses (dW, dX).
constexpr int BW_N = 16;                              // MMA N: video columns (NBV of them real, the rest stay zero)
constexpr int BW_LBO = BW_N * 16;                     // bytes between core matrices adjacent in K (next 4 gate rows)
constexpr int BW_B_BYTES = (COLS / 4) * BW_LBO;       // 8 KB
constexpr int BW_TMEM_COLS = 512;                     // 256 (W_hh^T) + 2 x 2 x 16 (partials) -> next power of two
constexpr int BW_THREADS = 8 * 32;
constexpr int BW_SMEM_EXCLUSIVE = 120 * 1024;         // more than half an SM: one CTA per SM (each needs all of TMEM)

// This lane's NBV partials (its unit, every video) -> the owner's recv[source][unit][video] row.  Lanes merge their
// values into 16-byte messages: 4 units (NBV = 1) or 2 units (NBV = 2) per message.
template <int NBV>
__device__ __forceinline__ void send_partials(uint32_t dst, const uint32_t (&v)[NBV], uint32_t bar, int lane) {
    if constexpr (NBV == 1) {
        const uint32_t b = __shfl_down_sync(0xffffffffu, v[0], 1);
        const uint32_t c = __shfl_down_sync(0xffffffffu, v[0], 2);
        const uint32_t d = __shfl_down_sync(0xffffffffu, v[0], 3);
        if ((lane & 3) == 0) st_async_v4(dst, make_uint4(v[0], b, c, d), bar);
    } else if constexpr (NBV == 2) {
        const uint32_t c = __shfl_down_sync(0xffffffffu, v[0], 1);
        const uint32_t d = __shfl_down_sync(0xffffffffu, v[1], 1);
        if ((lane & 1) == 0) st_async_v4(dst, make_uint4(v[0], v[1], c, d), bar);
    } else {
#pragma unroll
        for (int i = 0; i < NBV; i += 4) st_async_v4(dst + i * 4, make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]), bar);
    }
}

// Optional phase trace (AVS_BPTT_TRACE=1, debugging aid): cluster 0 / CTA 0 accumulates clock64 deltas of the per-step
// chain; read back with avs_debug_bptt_trace().
__device__ unsigned long long g_bptt_trace[10];
__device__ __forceinline__ long long bw_clk() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    return t;
}

template <int NBV, bool TRACE>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(BW_THREADS, 1)
lstm_backward_tc_kernel(const float* __restrict__ d_fused, const float4* __restrict__ save_pre,
                        const float* __restrict__ save_c, const float* __restrict__ whh, LstmBatch batch,
                        float* __restrict__ d_xg_v, float* __restrict__ d_xg_a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int r = static_cast<int>(cluster.block_rank());
    const int cid = blockIdx.x / CL;
    const int grp = cid >> 2;
    const int ld = cid & 3;
    const int dir = ld & 1;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    extern __shared__ __align__(1024) uint8_t bw_smem[];
    uint8_t* b_sm = bw_smem;                                                   // B operand (gate gradients of the step)
    float* recv = reinterpret_cast<float*>(bw_smem + BW_B_BYTES);              // [2][CL][UNITS][NBV] partials for my units
    constexpr int RECV_BUF = CL * UNITS * NBV;                                 // floats per buffer
    uint64_t* bars = reinterpret_cast<uint64_t*>(recv + 2 * RECV_BUF);         // recv[2] | mma
    uint64_t* bar_recv = bars;
    uint64_t* bar_mma = bars + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
    int* s_len = reinterpret_cast<int*>(tmem_slot + 2);
    int* s_row = s_len + NBV;
    const bool tracing = TRACE && tid == 0 && blockIdx.x == 0;
    long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tr_sent = 0;

    if (tid < NBV) {
        s_len[tid] = batch.slot_len[grp * NBV + tid];
        s_row[tid] = batch.slot_row_start[grp * NBV + tid];
    }
    for (int i = tid; i < (BW_B_BYTES + 2 * RECV_BUF * 4) / 16; i += BW_THREADS)
        reinterpret_cast<uint4*>(bw_smem)[i] = make_uint4(0, 0, 0, 0);
    const int maxlen = batch.group_maxlen[grp];
    constexpr uint32_t RECV_TX = CL * UNITS * NBV * 4;    // 8 sources x 32 units x NBV videos, fp32
    if (tid == 0) {
        mbar_init(&bar_recv[0], 1);
        mbar_init(&bar_recv[1], 1);
        mbar_init(bar_mma, 1);
        fence_mbar_init();
        if (maxlen > 1) mbar_expect_tx(&bar_recv[1], RECV_TX);   // the partials of step 0
    }
    if (warp == 7) {
        tmem_alloc(tmem_slot, BW_TMEM_COLS);
        tmem_relinquish();
    }
    fence_proxy_async();   // the zeroed B operand (its padding columns stay zero) is visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_w = tmem_base;          // columns [0, 256): W_hh^T, tile t at column 128 t
    const uint32_t tmem_d = tmem_base + 256;    // partials: buffer b, tile t at column 32 b + 16 t
    const int t = warp >> 2, q = warp & 3;      // the accumulator rows this warp can read: tile t, lane quarter q
    {
        // W_hh^T slice -> tensor memory: lane = hidden unit m = 128 t + 32 q + lane, column k = gate row 128 r + k of the
        // packed W_hh; coalesced over the lanes (consecutive units)
        const float* src = whh + (static_cast<size_t>(ld) * 4 * HC + r * COLS) * HC + t * 128 + q * 32 + lane;
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) pk[i] = __float_as_uint(to_tf32_rn(__ldg(src + static_cast<size_t>(cb * 32 + i) * HC)));
            tmem_st_32x32(tmem_w + (static_cast<uint32_t>(q * 32) << 16) + t * 128 + cb * 32, pk);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    cluster.sync();   // every CTA's barriers and buffers exist before any remote access
    tc_fence_after();

    const int owner = warp;                          // the CTA that owns the 32 units this warp reads (4 t + q)
    const uint32_t taddr = tmem_d + (static_cast<uint32_t>(q * 32) << 16) + t * 16;
    const uint32_t remote_recv = mapa(smem_u32(recv + (r * UNITS + lane) * NBV), owner);
    const uint32_t remote_bar = mapa(smem_u32(bar_recv), owner);
    const uint32_t idesc = umma_idesc(UMMA_FMT_TF32, 128, BW_N);
    const uint64_t d00 = umma_desc_noswz_kmajor(smem_u32(b_sm), BW_LBO, 128);
    const uint32_t d_hi = static_cast<uint32_t>(d00 >> 32), d_lo = static_cast<uint32_t>(d00);
    constexpr uint32_t D_K = (2 * BW_LBO) >> 4;      // K = 8 per MMA = two core matrices
    // ---- pointwise role (warps 0 .. NBV-1): unit jj = lane of video vb = warp
    const bool pw = warp < NBV;
    const int jj = lane, vb = warp;
    const int my_len = pw ? s_len[vb] : 0;
    float* d_xg = ((ld >> 1) ? d_xg_a : d_xg_v) + dir * (4 * HC) + r * COLS;
    const int out_col = ld * HC + r * UNITS;
    const size_t unit_off = static_cast<size_t>(ld) * HC + r * UNITS + jj;   // + row * 4 * HC
    const long long rstep = dir ? 1 : -1;            // frame consumed one forward step EARLIER
    long long row = 0;
    float dc_state = 0.f, dh_g = 0.f, c_cur = 0.f, c_nxt = 0.f;
    float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);
    if (my_len > 0) {
        row = static_cast<long long>(s_row[vb]) + (dir ? 0 : my_len - 1);
        dh_g = __ldg(d_fused + row * FUSED_LD + out_col + jj);
        pre = __ldg(save_pre + row * 4 * HC + unit_off);
        c_cur = __ldg(save_c + row * 4 * HC + unit_off);
        if (my_len > 1) c_nxt = __ldg(save_c + (row + rstep) * 4 * HC + unit_off);
    }
    float4* const b_mine = reinterpret_cast<float4*>(b_sm + jj * BW_LBO + vb * 16);
    const float* recv_mine = recv + jj * NBV + vb;   // + (buffer * CL + source) * UNITS * NBV

    for (int u = 0; u < maxlen; ++u) {
        const int cur = u & 1, nxt = cur ^ 1;
        if (pw) {
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool on = u < my_len;
            // operands of the next step (registers; their latency hides behind this step's exchange)
            float dh_n = 0.f, c_new = 0.f;
            float4 pre_n = make_float4(0.f, 0.f, 0.f, 0.f);
            if (u + 1 < my_len) {
                const long long rn = row + rstep;
                dh_n = __ldg(d_fused + rn * FUSED_LD + out_col + jj);
                pre_n = __ldg(save_pre + rn * 4 * HC + unit_off);
                if (u + 2 < my_len) c_new = __ldg(save_c + (rn + rstep) * 4 * HC + unit_off);
            }
            // everything that does not depend on dh: activations and the derivative factors
            const float c_prev = (u + 1 < my_len) ? c_nxt : 0.f;   // forward step 0 starts from c = 0
            const float gi = sigm(pre.x), gf = sigm(pre.y), gg = tanh_fast(pre.z), go = sigm(pre.w);
            const float tc = tanh_fast(c_cur);
            const float k_o = tc * go * (1.f - go);          // d_o pre-activation = dh * k_o
            const float k_c = go * (1.f - tc * tc);          // dc += dh * k_c
            const float k_i = gg * gi * (1.f - gi);
            const float k_f = c_prev * gf * (1.f - gf);
            const float k_g = gi * (1.f - gg * gg);
            float dh = dh_g;
            long long t0 = 0;
            if (tracing) {
                t0 = bw_clk();
                if (u > 0) tr_acc[7] += t0 - tr_sent;        // partials sent -> dh-independent math of the next step done
            }
            if (u > 0) {
                // CTA-scope acquire, as in the forward recurrence: the payload is THIS SM's shared memory, written by
                // the peers' st.async before their complete_tx (acquire.cluster adds a CCTL.IVALL per step).
                mbar_wait(&bar_recv[cur], ((u - 1) >> 1) & 1);   // the 8 partials of step u - 1 have landed
                if (tracing) {
                    const long long tw = bw_clk();
                    tr_acc[6] += tw - tr_sent;               // partials sent -> all 8 partials landed (exchange)
                    t0 = tw;
                }
                const float* rp = recv_mine + cur * RECV_BUF;
                float acc = 0.f;
#pragma unroll
                for (int src = 0; src < CL; ++src) acc += rp[src * UNITS * NBV];
                dh += acc;
            }
            if (on) {
                const float dc = fmaf(dh, k_c, dc_state);
                g4 = make_float4(dc * k_i, dc * k_f, dc * k_g, dh * k_o);
                dc_state = dc * gf;
            }
            *b_mine = make_float4(to_tf32_rn(g4.x), to_tf32_rn(g4.y), to_tf32_rn(g4.z), to_tf32_rn(g4.w));
            if (u + 1 < maxlen) {                            // the first forward step needs no dh
                if constexpr (NBV > 1) named_bar_sync(1, 32 * NBV);   // every pointwise warp has staged its chunks
                else __syncwarp();
                long long tA = 0;
                if (tracing) {
                    tA = bw_clk();
                    tr_acc[0] += tA - t0;                    // partials landed -> B operand staged, barrier passed
                }
                if (warp == 0 && elect_one()) {
                    fence_proxy_async();                     // generic-proxy stores -> visible to the tensor core's reads
                    tc_fence_after();
                    const uint32_t d = tmem_d + cur * 32;
#pragma unroll
                    for (int tt = 0; tt < 2; ++tt)
#pragma unroll
                        for (int k = 0; k < 16; ++k)
                            umma_tf32_ts_lohi(d + tt * 16, tmem_w + tt * 128 + k * 8, d_lo + k * D_K, d_hi, idesc, k != 0);
                    tc_commit(bar_mma);
                    // arm the barrier that will collect the NEXT step's partials for my units (its previous phase, the
                    // partials of step u - 1, was consumed above; no peer can send step u + 1 before it has received
                    // this step's partial from this CTA, which is sent after these MMAs)
                    if (u + 2 < maxlen) mbar_expect_tx(&bar_recv[cur], RECV_TX);
                }
                __syncwarp();
                if (tracing) tr_acc[1] += bw_clk() - tA;     // 32 MMAs + commit issued
            }
            if (on) {
                *reinterpret_cast<float4*>(d_xg + row * XG_LD + 4 * jj) = g4;
                row += rstep;
            }
            dh_g = dh_n;
            pre = pre_n;
            c_cur = c_nxt;
            c_nxt = c_new;
        }
        if (u + 1 < maxlen) {
            long long tB = 0;
            if (tracing) tB = bw_clk();
            mbar_wait(bar_mma, u & 1);
            tc_fence_after();
            long long tC = 0, tD = 0;
            if (tracing) {
                tC = bw_clk();
                tr_acc[2] += tC - tB;                        // commit issued -> epilogue awake (MMA latency)
            }
            uint32_t v[NBV];
            tmem_ld_32xN<NBV>(taddr + cur * 32, v);
            tmem_ld_wait();
            if (tracing) {
                tD = bw_clk();
                tr_acc[3] += tD - tC;                        // tcgen05.ld
            }
            send_partials<NBV>(remote_recv + nxt * RECV_BUF * 4, v, remote_bar + nxt * 8, lane);
            tc_fence_before();
            if (tracing) {
                tr_sent = bw_clk();
                tr_acc[4] += tr_sent - tD;                   // shuffles + st.async issue
            }
        }
    }
    if (tracing) {
        for (int i = 0; i < 8; ++i) g_bptt_trace[i] = tr_acc[i];
        g_bptt_trace[8] = maxlen;
    }
    tc_fence_before();
    __syncthreads();
    cluster.sync();   // nobody exits while a peer's st.async could still target its shared memory
    if (warp == 7) tmem_dealloc(tmem_base, BW_TMEM_COLS);
}

template <int NBV>
avs_status launch_bwd_tc(const float* d_fused, const void* save_pre, const float* save_c, const float* whh,
                         const LstmBatch& batch, float* d_xg_v, float* d_xg_a, cudaStream_t stream) {
    static const bool trace = getenv("AVS_BPTT_TRACE") != nullptr;
    auto kern = trace ? lstm_backward_tc_kernel<NBV, true> : lstm_backward_tc_kernel<NBV, false>;
    static PerDeviceOnce configured;
    const int dev = current_device();
    if (configured.needed(dev)) {
        AVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM_EXCLUSIVE));
        configured.mark(dev);
    }
    kern<<<batch.n_groups * 4 * CL, BW_THREADS, BW_SMEM_EXCLUSIVE, stream>>>(
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
    static const bool simt = getenv("AVS_BPTT_SIMT") != nullptr;   // the CUDA-core fp32 kernel (exact-product debug aid)
    if (!simt) {
        switch (batch.nb) {
            case 1: return launch_bwd_tc<1>(d_fused, save_pre, save_c, whh_packed, batch, d_xg_v, d_xg_a, stream);
            case 2: return launch_bwd_tc<2>(d_fused, save_pre, save_c, whh_packed, batch, d_xg_v, d_xg_a, stream);
            case 4: return launch_bwd_tc<4>(d_fused, save_pre, save_c, whh_packed, batch, d_xg_v, d_xg_a, stream);
            case 8: return launch_bwd_tc<8>(d_fused, save_pre, save_c, whh_packed, batch, d_xg_v, d_xg_a, stream);
            default: set_error("lstm_backward: unsupported videos-per-cluster %d", batch.nb); return AVS_ERR_INVALID;
        }
    }
    switch (batch.nb) {
        case 1: return launch_bwd<1>(d_fused, save_pre, save_c, whh_packed, batch, d_xg_v, d_xg_a, stream);
        case 2: return launch_bwd<2>(d_fused, save_pre, save_c, whh_packed, batch, d_xg_v, d_xg_a, stream);
        case 4: return launch_bwd<4>(d_fused, save_pre, save_c, whh_packed, batch, d_xg_v, d_xg_a, stream);
        case 8: return launch_bwd<8>(d_fused, save_pre, save_c, whh_packed, batch, d_xg_v, d_xg_a, stream);
        default: set_error("lstm_backward: unsupported videos-per-cluster %d", batch.nb); return AVS_ERR_INVALID;
    }
}

// debugging aid: the phase trace of the last traced BPTT launch (see g_bptt_trace)
avs_status bptt_trace_read(unsigned long long* out10) {
    AVS_CUDA(cudaDeviceSynchronize());
    AVS_CUDA(cudaMemcpyFromSymbol(out10, g_bptt_trace, 10 * sizeof(unsigned long long)));
    return AVS_OK;
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
