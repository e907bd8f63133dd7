// K2b (tensor-core version) -- the recurrent half of both nn.LSTM modules of AVBiLSTMModel
// (/root/reference/models/av_model.py:18-23, 39-40) on tcgen05.
//
// One thread-block cluster of 8 CTAs runs one (modality, direction) recurrence for a group of up
// to NB videos.  CTA r owns hidden units [32r, 32r+32) = 128 gate columns (global packing p = 4*jj + gate); inside
// the CTA's accumulator the four gates (i,f,g,o) of a unit sit 8 TMEM lanes apart (lane 32 q + 8 gate + u).
//
// Per time step and CTA:
//   gates^T[128 cols, NB videos] = W_hh_slice[128, 256] * h_prev[NB, 256]^T      (16 x tcgen05.mma, fp16 in, fp32 acc)
//     - W_hh slice: fp16, resident in TENSOR MEMORY for the whole kernel (A operand from TMEM:
//       128 lanes x 128 columns), so a step never re-streams the 64 KB slice through shared memory
//     - h_prev: fp16, shared memory (no-swizzle K-major), rewritten every step by all 8 CTAs
//     - accumulator: tensor memory, NB columns
//   epilogue warps (4 per "part" of NB/4 videos): two tcgen05.ld of the 16-lane shapes (.16x128b / .16x256b), which
//     hand thread t the accumulator rows t/4, t/4 + 8 (+ 16, + 24 for the second load) of column t%4: the W_hh rows
//     are placed in TMEM so that those four lanes are the gates i, f, g, o of ONE hidden unit, i.e. every thread owns
//     all four gates of one (hidden unit, video) straight out of the load (round 1 read one lane per thread and
//     transposed 4x4 with two shuffle rounds, ~65 clk of every step's critical path) -> + x W_ih^T (precomputed by
//     the GEMM, fp32, one float4) -> cell update on the SFU with shared denominators (5 ex2 + 2 rcp per cell)
//     -> h staged as fp16 in the destination layout, then pushed into every peer CTA's next-step buffer per
//     warp, right after a __syncwarp: lane l sends the 16-byte chunk of video l/8 to CTA l%8 with st.async,
//     whose mbarrier complete_tx (release at cluster scope) counts the bytes on the receiver's barrier;
//     four lanes also write the chunks to the fused output (16-byte stores).
//     No CTA-wide or cluster-wide barrier, no bulk-copy engine, no fences at cluster scope.
//   Measured (tools/lstm_scaling.py, B200): 0.71 us per step for an isolated chain, 0.76 us with two chains
//   per CTA, 0.86 us when two CTAs share every SM (config 2: 50 videos -> 28 clusters on 148 SMs).
//
// fp16 operands have the same 11-bit significand as tf32 and |h| < 1, so the recurrent matmul
// carries tf32-level rounding (measured contribution to the final scores: < 3e-5 relative);
// cell state, gate math and accumulation stay fp32.  Latency bound: ~T dependent steps.
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace avs {

namespace {

constexpr int HC = 256;
constexpr int CL = 8;
constexpr int UNITS = HC / CL;      // 32 hidden units per CTA
constexpr int COLS = 4 * UNITS;     // 128 gate columns per CTA
constexpr int XG_LD = 2 * 4 * HC;   // 2048
constexpr int FUSED_LD = 4 * HC;    // 1024
// Epilogue warps come in "parts" of 4 (one warp per TMEM lane quarter); every part owns NB/4 video slots.
// PARTS = 4: all NB slots are real videos, one CTA per SM.  PARTS = 2 (NB = 16 only): 8 real videos per
// cluster (the other 8 MMA columns are padding), 9 warps and 256 TMEM columns per CTA, so TWO CTAs of
// different clusters share an SM and one cluster's DSMEM exchange overlaps the other's MMA + cell math.
// CHAINS = 2 (with PARTS = 2): the two parts are INDEPENDENT recurrences (4 videos each) with their own h
// buffers, accumulator, barriers and MMA-issuing warp; they share the TMEM-resident W_hh slice.  A step of
// one chain is exchange -> MMA -> cell math -> send; with two chains per CTA and two CTAs per SM four such
// chains interleave on every SM, so the DSMEM-bandwidth-bound exchange of one hides behind the others' work.
constexpr int W_TMEM_COLS = HC / 2;        // W_hh slice as packed fp16 pairs: 128 columns

template <int NB, int CHAINS = 1>
struct Smem {
    // h operand: fp16, UMMA K-major NO-swizzle ("interleaved") layout [k/8][video][k%8]:
    // core matrix = 8 videos x 16 B; LBO (next 8 k) = NB*16 B, SBO (next 8 videos) = 128 B.
    // The 32 hidden units a CTA produces are therefore ONE contiguous NB*64-byte block.
    static constexpr int H_LBO = NB * 16;
    static constexpr int H_BYTES = (HC / 8) * H_LBO;         // NB * 512
    static constexpr int SLICE_BYTES = (UNITS / 8) * H_LBO;  // NB * 64: this CTA's share of h
    static constexpr int OFF_H = 0;                           // per chain: two h buffers ...
    static constexpr int OFF_STAGE16 = OFF_H + 2 * H_BYTES;   // ... and the double-buffered staged slice
    static constexpr int CHAIN_BYTES = 2 * H_BYTES + 2 * SLICE_BYTES;
    static constexpr int OFF_META = CHAINS * CHAIN_BYTES;     // len[NB], row[NB]
    static constexpr int OFF_BAR = OFF_META + 2 * NB * 4;     // per chain bar_h[2], bar_mma; then the tmem slot
    static constexpr int TOTAL = 1024 + OFF_BAR + (3 * CHAINS + 1) * 8;
    static constexpr int TMEM_COLS = (W_TMEM_COLS + CHAINS * NB) <= 256 ? 256 : 512;
};

// The four gate pre-activations of this thread's (hidden unit, video) pairs.  TMEM lane 32 q + 8 g + u holds gate g of
// hidden unit 8 q + u; the .16x128b / .16x256b shapes give thread t (u = t / 4, c = t % 4) lanes u and u + 8 of a
// 16-lane window -- gates (i, f) from the window at lane 32 q, gates (g, o) from the window at 32 q + 16 -- for
// column c (NV = 4), columns 2c, 2c + 1 (NV = 8) or 2c, 2c + 1, 8 + 2c, 9 + 2c (NV = 16).
template <int NV>
__device__ __forceinline__ void tmem_ld_gates(uint32_t taddr, float (&gi)[NV / 4], float (&gf)[NV / 4], float (&gg)[NV / 4],
                                              float (&go)[NV / 4]) {
    const uint32_t hi = taddr + (16u << 16);
    if constexpr (NV == 4) {
        uint32_t a0, a1, b0, b1;
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(taddr) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(hi) : "memory");
        tmem_ld_wait();
        gi[0] = __uint_as_float(a0); gf[0] = __uint_as_float(a1); gg[0] = __uint_as_float(b0); go[0] = __uint_as_float(b1);
    } else if constexpr (NV == 8) {
        uint32_t a[4], b[4];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(hi) : "memory");
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            gi[k] = __uint_as_float(a[k]); gf[k] = __uint_as_float(a[2 + k]);
            gg[k] = __uint_as_float(b[k]); go[k] = __uint_as_float(b[2 + k]);
        }
    } else {
        static_assert(NV == 16, "video slots per part: 4, 8 or 16");
        uint32_t a[8], b[8];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7])
                     : "r"(hi) : "memory");
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // k = 2 * (column block) + (column in the pair)
            const int o = (k >> 1) * 4 + (k & 1);
            gi[k] = __uint_as_float(a[o]); gf[k] = __uint_as_float(a[o + 2]);
            gg[k] = __uint_as_float(b[o]); go[k] = __uint_as_float(b[o + 2]);
        }
    }
}
// video slot (inside the part) of this thread's k-th pair, c = lane % 4
template <int NV>
__device__ __forceinline__ int owned_video(int c, int k) {
    if constexpr (NV == 4) return c;
    else if constexpr (NV == 8) return 2 * c + k;
    else return 2 * c + (k & 1) + 8 * (k >> 1);
}

// SFU primitives of the cell update: ex2.approx / rcp.approx are ~1-2 ulp (absolute error ~1e-7 on the gate
// activations, far inside the fp16 operand rounding of this mode).
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Optional phase trace (AVS_LSTM_TRACE=1, debugging aid): cluster 0 / CTA 0 accumulates clock64 deltas of the
// per-step dependency chain; read back with avs_debug_lstm_trace().
__device__ unsigned long long g_lstm_trace[8];
__device__ __forceinline__ long long clk64() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    return t;
}

// The input projection of one (frame, hidden unit): the four gate pre-activations, fp32 (float4) or fp16 (uint2).
// Inference keeps them in fp16 -- half of the largest HBM item of the step (2 x 2048 values per frame written by the
// input-projection GEMMs and read back here); the rounding (2^-12 relative on a pre-activation) is the same size as
// that of the fp16 h operand of the recurrent product.  The training forward keeps fp32.
template <bool XG16> struct XgVec { using type = float4; };
template <> struct XgVec<true> { using type = uint2; };
__device__ __forceinline__ float4 xg_f4(const float4& v) { return v; }
__device__ __forceinline__ float4 xg_f4(const uint2& v) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void xg_zero(float4& v) { v = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void xg_zero(uint2& v) { v = make_uint2(0u, 0u); }

template <int NB, int PARTS, int CHAINS, bool TRACE, bool XG16>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(PARTS * 128 + CHAINS * 32, PARTS == 2 ? 2 : 1)
lstm_tc_kernel(const void* __restrict__ xg_v, const void* __restrict__ xg_a, const float* __restrict__ whh,
               LstmBatch batch, int op_dtype, void* __restrict__ fused_out, int out_dtype, int round_tf32,
               float4* __restrict__ save_pre, float* __restrict__ save_c, int grp_off) {
    static_assert(CHAINS == 1 || CHAINS == PARTS, "a chain is either the whole CTA or one part");
    using S = Smem<NB, CHAINS>;
    constexpr int NV = NB / 4;             // video slots per part
    constexpr int SLOTS = NV * PARTS;      // real video slots of this cluster (batch.nb)
    constexpr int CHAIN_SLOTS = SLOTS / CHAINS;    // real video slots of one chain
    constexpr int EPI_WARPS = 4 * PARTS;
    constexpr int THREADS = EPI_WARPS * 32 + CHAINS * 32;   // + one MMA-issuing warp per chain
    cg::cluster_group cluster = cg::this_cluster();
    const int r = static_cast<int>(cluster.block_rank());
    const int cid = blockIdx.x / CL;
    const int grp = (cid >> 2) + grp_off;   // grp_off: this launch covers groups [grp_off, grp_off + gridDim.x / 32)
    const int ld = cid & 3;
    const int dir = ld & 1;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    // which chain this warp works for: epilogue warps by part, MMA warps by index
    const int chain = CHAINS == 1 ? 0 : (warp < EPI_WARPS ? (warp >> 2) : (warp - EPI_WARPS));
    uint8_t* h_sm = sm + chain * S::CHAIN_BYTES + S::OFF_H;
    uint8_t* stage16 = sm + chain * S::CHAIN_BYTES + S::OFF_STAGE16;   // this CTA's h slice in destination layout
    int* s_len = reinterpret_cast<int*>(sm + S::OFF_META);
    int* s_row = s_len + NB;
    uint64_t* bar_all = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);
    uint64_t* bar_h = bar_all + 3 * chain;  // [2]
    uint64_t* bar_mma = bar_h + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_all + 3 * CHAINS);
    __shared__ volatile long long tr_ts[2];   // [0] MMA commit issued, [1] bulk copies issued (TRACE only)
    const bool tracing = TRACE && blockIdx.x == 0;
    long long tr_acc[7] = {0, 0, 0, 0, 0, 0, 0};

    // ---- one-time setup ---------------------------------------------------------------------
    if (tid < NB) {
        s_len[tid] = tid < SLOTS ? batch.slot_len[grp * SLOTS + tid] : 0;
        s_row[tid] = tid < SLOTS ? batch.slot_row_start[grp * SLOTS + tid] : 0;
    }
    for (int i = tid; i < CHAINS * S::CHAIN_BYTES / 16; i += THREADS)
        reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);   // h buffers + stages of every chain
    if (tid == 0) {
        for (int i = 0; i < 3 * CHAINS; ++i) mbar_init(bar_all + i, 1);
        fence_mbar_init();
    }
    if (warp == EPI_WARPS) {
        tmem_alloc(tmem_slot, S::TMEM_COLS);
        tmem_relinquish();
    }
    fence_proxy_async();  // zeroed h visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_w = tmem_base;                 // columns [0, 128): W_hh slice, fp16 pairs
    const uint32_t tmem_d = tmem_base + W_TMEM_COLS + chain * NB;   // this chain's gate accumulator (NB columns)

    if (warp < EPI_WARPS) {
        // W_hh slice -> tensor memory, resident for the whole kernel.  A-operand layout of
        // kind::f16 with M = 128: lane = row (gate column), 32-bit column c holds k = 2c, 2c+1.
        const int q = warp & 3;                            // lane quarter
        // TMEM lane 32 q + 8 g + u  <-  packed row 4 (8 q + u) + g  (gate g of hidden unit 8 q + u): the four gates
        // of a unit sit 8 lanes apart, where the 16-lane tcgen05.ld shapes deliver them to one thread
        const int prow = q * 32 + ((lane & 7) << 2) + (lane >> 3);
#pragma unroll 1
        for (int cpart = warp >> 2; cpart < 4; cpart += PARTS) {   // 32-column (64 k) part of the row
            const float4* src = reinterpret_cast<const float4*>(
                whh + (static_cast<size_t>(ld) * 4 * HC + r * COLS + prow) * HC + cpart * 64);
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4 v = __ldg(src + i);
                pk[2 * i] = pack_lowp2(v.x, v.y, op_dtype);
                pk[2 * i + 1] = pack_lowp2(v.z, v.w, op_dtype);
            }
            tmem_st_32x32(tmem_w + (static_cast<uint32_t>(q * 32) << 16) + cpart * 32, pk);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    cluster.sync();       // every CTA's barriers and buffers exist before any remote access
    tc_fence_after();
    // slots are sorted by length (longest first), so a chain runs as long as its first slot
    const int maxlen = CHAINS == 1 ? batch.group_maxlen[grp] : s_len[chain * CHAIN_SLOTS];

    if (warp >= EPI_WARPS) {
        // ------------------------------------------------------------------ MMA issuer (one thread per chain)
        if (elect_one()) {
            const uint32_t idesc = umma_idesc(op_dtype == DT_BF16 ? UMMA_FMT_BF16 : UMMA_FMT_F16, COLS, NB);
            // Everything this thread executes between "h has landed" and the 16th MMA is on the step's critical
            // path, so that stretch is kept to the MMAs and their operand moves.  The h operand descriptor of
            // (buffer b, K block k) differs from that of (0, 0) only in its start-address field (low word,
            // address >> 4); the step loop is unrolled by two so that b is a compile-time constant and all 32 low
            // words are loop-invariant.  (The first version rebuilt the 16 descriptors after every wait, ~100
            // scalar instructions: 384 clk from "h landed" to "commit issued", now 265; three other arrangements
            // of the operand moves measure within 1 % of this one, so what is left is the tensor pipe's own
            // dispatch rate for 128x16x16 products with A in tensor memory.)
            const uint64_t d00 = umma_desc_noswz_kmajor(smem_u32(h_sm), S::H_LBO, 128);
            const uint32_t d_hi = static_cast<uint32_t>(d00 >> 32), d_lo00 = static_cast<uint32_t>(d00);
            constexpr uint32_t D_BUF = S::H_BYTES >> 4, D_K = (2 * S::H_LBO) >> 4;
            const bool tr = tracing && chain == 0;
            auto step = [&](int s, auto buf) {
                constexpr int b = decltype(buf)::value;
                // arm the barrier that will collect h_{s+1}: 8 peers x 64 B per real video slot, one local arrival
                if (s + 1 < maxlen) mbar_expect_tx(bar_h + (b ^ 1), CL * CHAIN_SLOTS * 64);
                if (s > 0) {
                    // CTA-scope acquire: the payload is shared memory written by the peers' st.async, whose
                    // complete_tx is performed after the data has landed; the reader is the tensor core behind the
                    // proxy fence below.  (acquire.cluster adds a CCTL.IVALL -- an L1 invalidation that protects
                    // nothing here -- to every step.)
                    mbar_wait(bar_h + b, b ? ((s >> 1) & 1) : (((s >> 1) + 1) & 1));
                    fence_proxy_async();   // peers' st.async (generic proxy) -> visible to the tensor core's reads
                }
                tc_fence_after();
                long long tA = 0;
                if (tr) {
                    tA = clk64();
                    if (s > 0) tr_acc[6] += tA - tr_ts[1];   // copies issued -> all 8 slices of h landed
                }
#pragma unroll
                for (int k = 0; k < 16; ++k)   // K = 256 = 16 x 16; A from TMEM (8 columns per K step)
                    umma_f16_ts_lohi(tmem_d, tmem_w + k * 8, d_lo00 + b * D_BUF + k * D_K, d_hi, idesc, k != 0);
                tc_commit(bar_mma);
                if (tr) {
                    const long long tB = clk64();
                    tr_ts[0] = tB;
                    tr_acc[0] += tB - tA;                     // h landed -> 16 MMAs + commit issued
                }
            };
            for (int s = 0; s < maxlen; s += 2) {
                step(s, std::integral_constant<int, 0>{});
                if (s + 1 < maxlen) step(s + 1, std::integral_constant<int, 1>{});
            }
            if (tr) {
                g_lstm_trace[0] = tr_acc[0];
                g_lstm_trace[6] = tr_acc[6];
                g_lstm_trace[7] = maxlen;
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue warps
        // Thread (u = lane / 4, c = lane % 4) of quarter q owns hidden unit jj = 8 q + u and NP = NV / 4 videos
        // (owned_video): the cell update runs once per (unit, video) with no redundant lanes, the gates come
        // straight from the two 16-lane TMEM loads and the input projection arrives as one float4 per video.
        const int q = warp & 3;              // TMEM lane quarter
        const int part = warp >> 2;          // which quarter of the videos
        const int cq = lane & 3;             // column selector of the 16-lane load shapes
        const int jj = q * 8 + (lane >> 2);  // hidden unit inside the CTA's slice
        const int v0 = part * NV;                        // first video slot of this part (index into s_len / s_row)
        const int lv0 = CHAINS == 1 ? v0 : 0;            // the same, relative to this chain's buffers / accumulator
        using XgT = typename XgVec<XG16>::type;   // four gates of one hidden unit: 16 bytes (fp32) or 8 (fp16)
        const XgT* xg4 = reinterpret_cast<const XgT*>((ld >> 1) ? xg_a : xg_v) + (dir * (4 * HC) + r * COLS + 4 * jj) / 4;
        constexpr int XG_LD4 = XG_LD / 4;
        const int out_col = ld * HC + r * UNITS;
        const uint32_t taddr = tmem_d + (static_cast<uint32_t>(q * 32) << 16) + lv0;
        // this thread's slot in the staged slice: [jj/8][video][jj%8] 16-bit values
        uint16_t* stage_mine = reinterpret_cast<uint16_t*>(stage16 + (jj >> 3) * S::H_LBO) + (jj & 7);
        float* const fcol = reinterpret_cast<float*>(fused_out) + out_col + jj;
        uint16_t* const fcol_h = reinterpret_cast<uint16_t*>(fused_out) + out_col + jj;
        const bool op_bf16 = op_dtype == DT_BF16;
        const bool lowp_out = out_dtype == op_dtype;   // fused output straight from the staged operand slice
        const int rstep = dir ? -1 : 1;
        constexpr int NP = NV / 4;             // (unit, video) pairs owned by this thread
        constexpr float LOG2E = 1.4426950408889634f;
        // h exchange, warp-local: this warp produces the 16-byte chunks (8 hidden units of chunk q) of videos
        // v0 .. v0+NV-1.  After a __syncwarp lane l sends chunk (video v0 + 4k + (l >> 3)) to peer CTA (l & 7)
        // with one st.async: 4 videos x 8 peers = 32 lanes.  No CTA-wide barrier, no bulk-copy engine.
        const uint32_t peer = batch.lane_map ? (lane >> 2) : (lane & 7), cvid = batch.lane_map ? (lane & 3) : (lane >> 3);
        const uint32_t stage_rd = smem_u32(stage16) + q * S::H_LBO + (lv0 + cvid) * 16;     // + k*64, + slot
        const uint32_t remote_h = mapa(smem_u32(h_sm) + (r * 4 + q) * S::H_LBO + (lv0 + cvid) * 16, peer);
        const uint32_t remote_bar = mapa(smem_u32(bar_h), peer);

        float c_state[NP];
        XgT xv0[NP], xv1[NP];                  // xv0: this step's input projection, xv1: next step's
        int len_r[NP], row_r[NP], vid_r[NP];   // row_r: global row of the frame consumed at step s
        XgT zero4;
        xg_zero(zero4);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            c_state[k] = 0.f;
            vid_r[k] = v0 + owned_video<NV>(cq, k);
            const int len = s_len[vid_r[k]];
            len_r[k] = len;
            row_r[k] = s_row[vid_r[k]] + (dir ? (len > 0 ? len - 1 : 0) : 0);
            xv0[k] = len > 0 ? __ldg(xg4 + static_cast<size_t>(row_r[k]) * XG_LD4) : zero4;
            xv1[k] = len > 1 ? __ldg(xg4 + static_cast<size_t>(row_r[k] + rstep) * XG_LD4) : zero4;
        }

        for (int s = 0; s < maxlen; ++s) {
            // two-step-deep register prefetch of the input projections (DRAM latency >> one step)
            XgT xv2[NP];
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                xv2[k] = zero4;
                if (s + 2 < len_r[k]) xv2[k] = __ldg(xg4 + static_cast<size_t>(row_r[k] + 2 * rstep) * XG_LD4);
            }
            mbar_wait(bar_mma, s & 1);
            tc_fence_after();
            long long tC = 0, tD = 0;
            if (tracing && tid == 0) {
                tC = clk64();
                tr_acc[1] += tC - tr_ts[0];                   // commit issued -> epilogue awake (MMA latency)
            }
            float gi[NP], gf[NP], gg[NP], go[NP];
            tmem_ld_gates<NV>(taddr, gi, gf, gg, go);
            if (tracing && tid == 0) {
                tD = clk64();
                tr_acc[2] += tD - tC;                         // tcgen05.ld
            }
            float h_out[NP];
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const float4 xin = xg_f4(xv0[k]);
                const float t_i = gi[k] + xin.x;
                const float t_f = gf[k] + xin.y;
                const float t_g = gg[k] + xin.z;
                const float t_o = go[k] + xin.w;
                // ---- LSTM cell on the SFU with shared denominators (5 ex2 + 2 rcp per cell):
                //   sigmoid(x) = 1 / (1 + e^-x),  tanh(x) = (1 - e^-2x) / (1 + e^-2x)
                //   c' = sig(f) c + sig(i) tanh(g) = [c (1+ei)(1+eg) + (1-eg)(1+ef)] / [(1+ei)(1+eg)(1+ef)]
                //   h  = sig(o) tanh(c')           = (1-ec) / [(1+eo)(1+ec)]
                // arguments clamped from below so that the products stay far inside fp32 (sigmoid(-25) = 1.4e-11).
                // No upper clamp: beyond +25 (+12.5) the exponential is < 2^-24 and 1 + e rounds to 1 either way,
                // down to the flushed zero -- same bits, one dependent instruction less per gate.
#ifdef AVS_LSTM_TANH_APPROX
                // EXPERIMENT (not the shipped build): one MUFU.TANH per activation -- sigmoid(x) =
                // 0.5 tanh(x / 2) + 0.5 -- i.e. a dependent chain of two SFU operations per cell instead of four, at
                // tanh.approx's ~2^-11 relative error (the formulation below is ~1e-7)
                const float s_i = fmaf(tanh_approx(0.5f * t_i), 0.5f, 0.5f);
                const float s_f = fmaf(tanh_approx(0.5f * t_f), 0.5f, 0.5f);
                const float s_o = fmaf(tanh_approx(0.5f * t_o), 0.5f, 0.5f);
                const float cn = fmaf(s_f, c_state[k], s_i * tanh_approx(t_g));
                const float h = s_o * tanh_approx(cn);
#else
                const float ei = ex2_approx(fmaxf(t_i, -25.f) * -LOG2E);
                const float ef = ex2_approx(fmaxf(t_f, -25.f) * -LOG2E);
                const float eg = ex2_approx(fmaxf(t_g, -12.5f) * (-2.f * LOG2E));
                const float eo = ex2_approx(fmaxf(t_o, -25.f) * -LOG2E);
                const float A = (1.f + ei) * (1.f + eg);
                const float opf = 1.f + ef;
                const float cn = fmaf(c_state[k], A, (1.f - eg) * opf) * rcp_approx(A * opf);
                const float ec = ex2_approx(fmaxf(cn, -12.5f) * (-2.f * LOG2E));
                const float h = (1.f - ec) * rcp_approx((1.f + eo) * (1.f + ec));
#endif
                const bool on = s < len_r[k];
                c_state[k] = on ? cn : c_state[k];
                h_out[k] = h;
                if (save_pre != nullptr && on) {   // training: gate pre-activations and the new cell state
                    const size_t o = (static_cast<size_t>(row_r[k]) * 4 + ld) * HC + r * UNITS + jj;
                    save_pre[o] = make_float4(t_i, t_f, t_g, t_o);
                    save_c[o] = cn;
                }
                if (on)   // |h| < 1: no saturation needed
                    stage_mine[(s & 1) * (S::SLICE_BYTES / 2) + (vid_r[k] - v0 + lv0) * 8] =
                        op_bf16 ? __bfloat16_as_ushort(__float2bfloat16_rn(h)) : __half_as_ushort(__float2half_rn(h));
            }
            long long tE = 0, tF = 0;
            if (tracing && tid == 0) {
                tE = clk64();
                tr_acc[3] += tE - tD;                         // transpose + cell math + stage write
            }
            tc_fence_before();     // this warp's TMEM reads of the step are complete
            __syncwarp();          // the warp's chunks are staged
            if (tracing && tid == 0) {
                tF = clk64();
                tr_acc[4] += tF - tE;                         // warp sync
            }
            uint4 hv[NP];
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(hv[k].x), "=r"(hv[k].y), "=r"(hv[k].z), "=r"(hv[k].w)
                             : "r"(stage_rd + (s & 1) * S::SLICE_BYTES + k * 64));
                if (s + 1 < maxlen) {
                    const uint32_t nb = (s + 1) & 1;
                    st_async_v4(remote_h + nb * S::H_BYTES + k * 64, hv[k], remote_bar + nb * 8);
                }
            }
            if (tracing && tid == 0) {
                const long long tG = clk64();
                tr_ts[1] = tG;
                tr_acc[5] += tG - tF;                         // ld.shared + st.async issue
            }
            // h -> global fused output (off the critical path).  16-bit output in the operand format: the chunk
            // this lane just read is exactly 16 contiguous bytes of a fused row; the lanes with peer == 0 store it.
            if (lowp_out && peer == 0) {
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    const int v = v0 + 4 * k + cvid;
                    const int len = s_len[v];
                    if (s < len) {
                        const int row = s_row[v] + (dir ? len - 1 - s : s);
                        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(fused_out) +
                                                  static_cast<size_t>(row) * FUSED_LD + out_col + q * 8) = hv[k];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const bool on = s < len_r[k];
                if (!lowp_out && on) {
                    if (out_dtype != DT_F32)
                        fcol_h[static_cast<size_t>(row_r[k]) * FUSED_LD] = to_lowp_bits(h_out[k], out_dtype);
                    else
                        fcol[static_cast<size_t>(row_r[k]) * FUSED_LD] = round_tf32 ? to_tf32_rn(h_out[k]) : h_out[k];
                }
                row_r[k] += on ? rstep : 0;
                xv0[k] = xv1[k];
                xv1[k] = xv2[k];
            }
        }
        if (tracing && tid == 0)
            for (int i = 1; i <= 5; ++i) g_lstm_trace[i] = tr_acc[i];
    }
    tc_fence_before();
    __syncthreads();
    cluster.sync();  // nobody exits while a peer could still touch its shared memory
    if (warp == EPI_WARPS) tmem_dealloc(tmem_base, S::TMEM_COLS);
}

__global__ void stagger_kernel(unsigned int ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < ns);
}

// Streams / events for the split launch below (one set per device; the library serialises calls per handle).
struct SplitCtx {
    cudaStream_t aux = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool ok = false;
};
SplitCtx& split_ctx() {
    static SplitCtx per_device[kMaxDevices];
    SplitCtx& c = per_device[current_device()];
    if (!c.ok && c.aux == nullptr) {
        if (cudaStreamCreateWithFlags(&c.aux, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming) == cudaSuccess)
            c.ok = true;
    }
    return c;
}

template <int NB, int PARTS, int CHAINS, bool XG16>
avs_status launch_tc(const void* xg_v, const void* xg_a, const float* whh, const LstmBatch& batch_in, int op_dtype,
                     void* fused, int out_dtype, int round_tf32, float4* save_pre, float* save_c, cudaStream_t stream) {
    static const bool trace = getenv("AVS_LSTM_TRACE") != nullptr;
    static const bool no_split = getenv("AVS_LSTM_NO_SPLIT") != nullptr;
    LstmBatch batch = batch_in;
    batch.lane_map = 1;   // adjacent lanes address adjacent 16-byte chunks of ONE peer (measured: an isolated chain
                          // takes 0.67 instead of 0.71 us per step; no difference once chains share an SM)
    if (const char* e = getenv("AVS_LSTM_LANEMAP")) batch.lane_map = atoi(e);
    auto kern = trace ? lstm_tc_kernel<NB, PARTS, CHAINS, NB == 16, XG16> : lstm_tc_kernel<NB, PARTS, CHAINS, false, XG16>;
    constexpr int SMEM = Smem<NB, CHAINS>::TOTAL;
    constexpr int SMEM_EXCLUSIVE = 200 * 1024;   // more than half an SM: no second CTA of either launch fits beside it
    static PerDeviceOnce configured;
    const int dev = current_device();
    if (configured.needed(dev)) {
        AVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      PARTS == 2 ? SMEM_EXCLUSIVE : SMEM));
        configured.mark(dev);
    }
    const int threads = PARTS * 128 + CHAINS * 32;
    // The kernel lasts as long as the LONGEST group's chain of dependent steps, and a step is ~13 % slower on
    // an SM that two CTAs share (tools/lstm_scaling.py: 0.76 vs 0.86 us).  When the batch needs shared SMs
    // anyway, the longest groups (groups are sorted by length) are launched on their own, with a shared-memory
    // request that keeps their CTAs alone on their SMs; the other groups, which have fewer steps to run, share
    // the remaining SMs two CTAs each.  Small batches (<= 16 clusters) run entirely on exclusive SMs.
    // A = number of leading (longest) groups that get exclusive SMs.  Clusters are placed inside one GPC (16-20
    // SMs on B200): a GPC holds 2 exclusive clusters or 4 shared ones, so with 8 GPCs 2A + (n - A) <= 8 must hold
    // for n groups; measured on B200, 16 exclusive clusters do NOT fit at once (one waits for a second wave), 12
    // do, and n = 7 with A = 1 does -- hence the table (one GPC of slack whenever A >= 2).
    SplitCtx& sc = split_ctx();
    int n_excl = 0;
    if (PARTS == 2 && !no_split && batch.excl >= 0) {
        n_excl = lstm_exclusive_groups(batch.n_groups);
        if (n_excl < batch.n_groups && !sc.ok) n_excl = 0;
    }
    if (n_excl == 0 || n_excl == batch.n_groups) {
        kern<<<batch.n_groups * 4 * CL, threads, n_excl ? SMEM_EXCLUSIVE : SMEM, stream>>>(
            xg_v, xg_a, whh, batch, op_dtype, fused, out_dtype, round_tf32, save_pre, save_c, 0);
        AVS_LAUNCH_CHECK();
        return AVS_OK;
    }
    AVS_CUDA(cudaEventRecord(sc.fork, stream));
    kern<<<n_excl * 4 * CL, threads, SMEM_EXCLUSIVE, stream>>>(xg_v, xg_a, whh, batch, op_dtype, fused, out_dtype,
                                                               round_tf32, save_pre, save_c, 0);
    AVS_LAUNCH_CHECK();
    AVS_CUDA(cudaStreamWaitEvent(sc.aux, sc.fork, 0));
    // the exclusive CTAs above must be PLACED first (they need empty SMs): both launches become eligible at the same
    // moment, and which one the hardware picks up first depends on how the streams map to its queues (measured: after
    // other streams had been used the recurrence stage took 0.50 - 1.86 ms instead of 0.45 in one run out of two);
    // the shorter groups start ~3 us later, which they have to spare
    stagger_kernel<<<1, 1, 0, sc.aux>>>(3000);
    AVS_LAUNCH_CHECK();
    kern<<<(batch.n_groups - n_excl) * 4 * CL, threads, SMEM, sc.aux>>>(xg_v, xg_a, whh, batch, op_dtype, fused,
                                                                        out_dtype, round_tf32, save_pre, save_c, n_excl);
    AVS_LAUNCH_CHECK();
    AVS_CUDA(cudaEventRecord(sc.join, sc.aux));
    AVS_CUDA(cudaStreamWaitEvent(stream, sc.join, 0));
    return AVS_OK;
}

}  // namespace

avs_status launch_stagger(cudaStream_t stream, unsigned int ns) {
    stagger_kernel<<<1, 1, 0, stream>>>(ns);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

int lstm_exclusive_groups(int n_groups) {
    static const int kExclusive[8] = {0, 1, 2, 3, 3, 2, 1, 1};
    return (n_groups >= 0 && n_groups <= 7) ? kExclusive[n_groups] : 0;
}

avs_status lstm_recurrence_tc_groups(const void* xg_v, const void* xg_a, int xg_dtype, const float* whh_packed,
                                     const LstmBatch& batch_in, int g_lo, int g_hi, int exclusive, int op_dtype, void* fused,
                                     int out_dtype, cudaStream_t stream) {
    AVS_CHECK(xg_dtype == DT_F32 || xg_dtype == DT_F16, AVS_ERR_INVALID, "lstm_tc: input projections must be fp32 or fp16");
    AVS_CHECK(batch_in.nb == 8 && g_lo >= 0 && g_lo < g_hi && g_hi <= batch_in.n_groups, AVS_ERR_INVALID,
              "lstm_recurrence_tc_groups: bad group range [%d, %d) of %d (8-slot variant only)", g_lo, g_hi, batch_in.n_groups);
    AVS_CHECK(op_dtype == DT_F16 || op_dtype == DT_BF16, AVS_ERR_INVALID, "lstm_tc: operand dtype must be fp16 or bf16");
    LstmBatch batch = batch_in;
    batch.lane_map = 1;
    auto kern = xg_dtype == DT_F16 ? lstm_tc_kernel<16, 2, 2, false, true> : lstm_tc_kernel<16, 2, 2, false, false>;
    constexpr int SMEM = Smem<16, 2>::TOTAL;
    constexpr int SMEM_EXCLUSIVE = 200 * 1024;
    static PerDeviceOnce configured[2];
    const int dev = current_device();
    if (configured[xg_dtype == DT_F16].needed(dev)) {
        AVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_EXCLUSIVE));
        configured[xg_dtype == DT_F16].mark(dev);
    }
    // exclusive: 0 = the kernel's own shared memory (two CTAs per SM), 1 = an SM-exclusive request, >= 1024 = that many
    // bytes (a request between the two keeps CTAs of the same class apart while smaller ones still fit beside them)
    const int smem_req = exclusive >= 1024 ? std::min(std::max(exclusive, SMEM), SMEM_EXCLUSIVE) : (exclusive ? SMEM_EXCLUSIVE : SMEM);
    kern<<<(g_hi - g_lo) * 4 * CL, 2 * 128 + 2 * 32, smem_req, stream>>>(
        xg_v, xg_a, whh_packed, batch, op_dtype, fused, out_dtype, 0, nullptr, nullptr, g_lo);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

// debugging aid: the phase trace of the last traced launch (8 values, see g_lstm_trace)
avs_status lstm_trace_read(unsigned long long* out8) {
    AVS_CUDA(cudaDeviceSynchronize());
    AVS_CUDA(cudaMemcpyFromSymbol(out8, g_lstm_trace, 8 * sizeof(unsigned long long)));
    return AVS_OK;
}

template <bool XG16>
static avs_status recurrence_tc_dispatch(const void* xg_v, const void* xg_a, const float* whh_packed, const LstmBatch& batch,
                                         int op_dtype, void* fused, int out_dtype, int round_tf32, cudaStream_t stream,
                                         void* save_pre, float* save_c) {
    float4* const pre = static_cast<float4*>(save_pre);
    switch (batch.nb) {   // video slots per cluster
        case 8:
            if (getenv("AVS_LSTM_ONE_CHAIN") != nullptr)   // experiment: 8 videos in ONE chain per CTA
                return launch_tc<16, 2, 1, XG16>(xg_v, xg_a, whh_packed, batch, op_dtype, fused, out_dtype, round_tf32, pre,
                                                 save_c, stream);
            return launch_tc<16, 2, 2, XG16>(xg_v, xg_a, whh_packed, batch, op_dtype, fused, out_dtype, round_tf32, pre,
                                             save_c, stream);
        case 16: return launch_tc<16, 4, 1, XG16>(xg_v, xg_a, whh_packed, batch, op_dtype, fused, out_dtype, round_tf32, pre,
                                                  save_c, stream);
        case 32: return launch_tc<32, 4, 1, XG16>(xg_v, xg_a, whh_packed, batch, op_dtype, fused, out_dtype, round_tf32, pre,
                                                  save_c, stream);
        case 64: return launch_tc<64, 4, 1, XG16>(xg_v, xg_a, whh_packed, batch, op_dtype, fused, out_dtype, round_tf32, pre,
                                                  save_c, stream);
        default: set_error("lstm_tc: unsupported videos-per-cluster %d", batch.nb); return AVS_ERR_INVALID;
    }
}

avs_status lstm_recurrence_tc(const void* xg_v, const void* xg_a, int xg_dtype, const float* whh_packed,
                              const LstmBatch& batch, int op_dtype, void* fused, int out_dtype, int round_tf32,
                              cudaStream_t stream, void* save_pre, float* save_c) {
    if (batch.n_groups == 0) return AVS_OK;
    AVS_CHECK(op_dtype == DT_F16 || op_dtype == DT_BF16, AVS_ERR_INVALID, "lstm_tc: operand dtype must be fp16 or bf16");
    AVS_CHECK(xg_dtype == DT_F32 || xg_dtype == DT_F16, AVS_ERR_INVALID, "lstm_tc: input projections must be fp32 or fp16");
    if (xg_dtype == DT_F16)
        return recurrence_tc_dispatch<true>(xg_v, xg_a, whh_packed, batch, op_dtype, fused, out_dtype, round_tf32, stream,
                                            save_pre, save_c);
    return recurrence_tc_dispatch<false>(xg_v, xg_a, whh_packed, batch, op_dtype, fused, out_dtype, round_tf32, stream,
                                         save_pre, save_c);
}

}  // namespace avs
