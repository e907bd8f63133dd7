// K4 -- temporal multi-head self-attention core on tcgen05 / TMEM / TMA (flash-style, never
// materialises the [T, T] score matrix that /root/reference/models/attention.py:21-22 builds).
//
// Arithmetic of attention.py:17-23 (== nn.MultiheadAttention's core, av_model.py:26,44 on the
// transposed tensor): contiguous dh = 256 wide heads, softmax(Q K^T / sqrt(dh)) V over the frames
// of ONE video, keys >= len masked (variable-length batches, packed rows).
//
// One CTA per (128-query block, head, video):
//   warp 0      TMA producer: K_j / V_j blocks of 64 keys (32 KB each, 3 stages each)
//   warp 1      tcgen05.mma issuer:  S_j = Q K_j^T   (TS: Q from TMEM, K_j from smem, 128 x 64 x 256, fp16 in / fp32 acc)
//                                    O  += P_j V_j   (TS: P_j from TMEM, V_j MN-major from smem, 128 x 256 x 64)
//   Long sequences (QT = true, max length >= 1024): Q lives in TENSOR MEMORY (128 lanes x 128 columns of 16-bit
//   pairs, written once by the softmax warps).  With Q in shared memory every 128 x 64 x 16 product reads 4 KB of Q +
//   2 KB of K per 32 tensor-pipe clocks = 192 B/clk, more than the 128 B/clk shared memory delivers, so the QK^T half
//   of the kernel runs at 2/3 of the tensor rate (round 1 / r02b ncu: 59 % tensor-pipe active at T = 8192); from TMEM
//   the product reads 2 KB of K per 32 clk (1.92 -> 1.76 ms for 8 x T = 8192).  Short sequences (QT = false) keep Q in
//   shared memory, loaded by TMA next to K_0: their time is the prologue, and the per-thread global loads of the
//   TMEM variant lengthen it (config 2, T <= 700: 0.095 -> 0.110 ms).
//   warps 2..9  softmax: TWO warps per TMEM lane quarter, each takes 32 of the 64 keys of a block (and half of O's
//               columns when O is rescaled / written): tcgen05.ld S_j, block max exchanged between the two warps of a
//               row through shared memory + a 64-thread named barrier, running max with LAZY rescaling (O is only
//               rescaled when the max grows by > 2^8), exp2 on the SFU, P_j -> fp16 -> tcgen05.st over S_j's
//               columns; final O / l -> global.  (Round 1 had four softmax warps, one per SM sub-partition: 8,192
//               ex2 + ~4 FMA-pipe instructions per key and row from ONE warp per scheduler could not keep up with the
//               1,024 tensor-pipe clocks of a key block -- ncu: 59 % tensor-pipe active at T = 8192.)
// TMEM columns: O [0,256), S_0 / P_0 [256,320), S_1 / P_1 [320,384), Q [384,512).  S is double buffered so the
// tensor pipe computes S_{j+1} while the softmax warps work on S_j.
// fp16 operands carry the same 11-bit significand as tf32; |q.k| stays far below the fp16 range
// because the inputs of the projection are LSTM outputs in (-1, 1).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace avs {

namespace {

constexpr int DH = 256;
constexpr int BM = 128;   // queries per CTA
constexpr int BN = 64;    // keys per block
constexpr int KV_SUB = BN * 128;       // one dh-block of a K / V block: 64 rows x 128 B
constexpr int KV_BYTES = 4 * KV_SUB;   // 32 KB
constexpr int Q_SUB = BM * 128;        // one dh-block (64 elements) of Q: 128 rows x 128 B   (QT = false only)
constexpr int Q_BYTES = 4 * Q_SUB;     // 64 KB
template <bool QT>
struct Lay {
    static constexpr int KV_STAGES = QT ? 3 : 2;
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = QT ? 0 : Q_BYTES;
    static constexpr int OFF_V = OFF_K + KV_STAGES * KV_BYTES;
    static constexpr int OFF_BAR = OFF_V + KV_STAGES * KV_BYTES;
    static constexpr int N_BARS = 1 + 4 * KV_STAGES + 2 + 2 + 1 + 1;  // q, k_full/empty, v_full/empty, s[2], p[2], o, done
    static constexpr int OFF_XCH = OFF_BAR + N_BARS * 8 + 16;  // softmax exchange: block max [2][2][128], row sum [2][128]
    static constexpr int SMEM_TOTAL = 1024 + OFF_XCH + (2 * 2 * BM + 2 * BM) * 4;
};
constexpr int SOFTMAX_WARPS = 8;
constexpr int ATT_THREADS = 64 + SOFTMAX_WARPS * 32;
constexpr uint32_t TM_O = 0, TM_S = 256, TM_Q = 384;   // TMEM column offsets
constexpr float LAZY_THRESHOLD = 8.0f;         // rescale O only when the scaled max grows by more than this

// shared-memory matrix descriptor, MN-major operand with 128-byte swizzle (V block as B of P*V):
// 64 contiguous N elements (128 B) x 8 K rows per swizzle atom; LBO = stride between 64-element
// N blocks, SBO = stride between 8-row K groups.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool QT>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const uint16_t* __restrict__ qkv, int64_t total_rows,
                    SeqDesc seqs, int E, int in_dtype, void* __restrict__ ctx, int64_t ld_ctx, int out_dtype,
                    int round_tf32, float scale_log2) {
    const int seq = blockIdx.z, head = blockIdx.y;
    const int len = seqs.len[seq];
    const int q0 = blockIdx.x * BM;
    if (q0 >= len) return;
    const int base = seqs.base[seq];
    const int nblk = (len + BN - 1) / BN;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    constexpr int KV_STAGES = Lay<QT>::KV_STAGES, OFF_K = Lay<QT>::OFF_K, OFF_V = Lay<QT>::OFF_V, OFF_BAR = Lay<QT>::OFF_BAR,
                  OFF_XCH = Lay<QT>::OFF_XCH, N_BARS = Lay<QT>::N_BARS, OFF_Q = Lay<QT>::OFF_Q;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint64_t* bar_q = bars;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = k_full + KV_STAGES;
    uint64_t* v_full = k_empty + KV_STAGES;
    uint64_t* v_empty = v_full + KV_STAGES;
    uint64_t* bar_s = v_empty + KV_STAGES;
    uint64_t* bar_p = bar_s + 2;
    uint64_t* bar_o = bar_p + 2;
    uint64_t* bar_done = bar_o + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_qkv);
        mbar_init(bar_q, QT ? SOFTMAX_WARPS * 32 : 1);
        for (int i = 0; i < KV_STAGES; ++i) {
            mbar_init(k_full + i, 1);
            mbar_init(k_empty + i, 1);
            mbar_init(v_full + i, 1);
            mbar_init(v_empty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_s + i, 1);
            mbar_init(bar_p + i, SOFTMAX_WARPS * 32);
        }
        mbar_init(bar_o, 1);
        mbar_init(bar_done, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            const int kcol = E + head * DH, vcol = 2 * E + head * DH;
            if (!QT) {
                mbar_expect_tx(bar_q, Q_BYTES);
                for (int d = 0; d < 4; ++d)
                    for (int hf = 0; hf < 2; ++hf)
                        tma_load_2d(sm + OFF_Q + d * Q_SUB + hf * (64 * 128), &tm_qkv, bar_q, head * DH + d * 64, base + q0 + hf * 64);
            }
            for (int j = 0; j < nblk; ++j) {
                const int st = j % KV_STAGES;
                const uint32_t ph = (j / KV_STAGES) & 1;
                mbar_wait(k_empty + st, ph ^ 1);
                mbar_expect_tx(k_full + st, KV_BYTES);
                for (int d = 0; d < 4; ++d)
                    tma_load_2d(sm + OFF_K + st * KV_BYTES + d * KV_SUB, &tm_qkv, k_full + st, kcol + d * 64, base + j * BN);
                mbar_wait(v_empty + st, ph ^ 1);
                mbar_expect_tx(v_full + st, KV_BYTES);
                for (int d = 0; d < 4; ++d)
                    tma_load_2d(sm + OFF_V + st * KV_BYTES + d * KV_SUB, &tm_qkv, v_full + st, vcol + d * 64, base + j * BN);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            const uint32_t fmt = in_dtype == DT_BF16 ? UMMA_FMT_BF16 : UMMA_FMT_F16;
            const uint32_t idesc_s = umma_idesc(fmt, BM, BN);
            const uint32_t idesc_o = umma_idesc(fmt, BM, DH) | (1u << 16);  // B (= V) is MN-major
            auto issue_s = [&](int j) {
                const int st = j & 1, ks = j % KV_STAGES;
                mbar_wait(k_full + ks, (j / KV_STAGES) & 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sm + OFF_K + ks * KV_BYTES);
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) {
                    const uint64_t bd = umma_desc_sw128_kmajor(k_addr + (kk >> 2) * KV_SUB + (kk & 3) * 32);
                    if (QT) {   // A = Q from tensor memory: 8 columns (16 x 16-bit pairs) per K step
                        umma_f16_ts(tmem + TM_S + st * BN, tmem + TM_Q + kk * 8, bd, idesc_s, kk != 0);
                    } else {
                        const uint64_t ad = umma_desc_sw128_kmajor(smem_u32(sm + OFF_Q) + (kk >> 2) * Q_SUB + (kk & 3) * 32);
                        umma_f16_ss(tmem + TM_S + st * BN, ad, bd, idesc_s, kk != 0);
                    }
                }
                tc_commit(k_empty + ks);   // K stage reusable once these MMAs retire
                tc_commit(bar_s + st);     // S_j ready for the softmax warps
            };
            mbar_wait(bar_q, 0);      // Q stored in tensor memory by the softmax warps / landed in shared memory
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < nblk; ++j) {
                const int st = j & 1, vs = j % KV_STAGES;
                if (j + 1 < nblk) issue_s(j + 1);   // overlaps the softmax of block j
                mbar_wait(bar_p + st, (j >> 1) & 1);  // P_j stored (and O rescaled if it had to be)
                mbar_wait(v_full + vs, (j / KV_STAGES) & 1);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(sm + OFF_V + vs * KV_BYTES);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {     // 16 keys per MMA
                    const uint64_t bd = umma_desc_sw128_mnmajor(v_addr + kk * 2048, KV_SUB, 1024);
                    umma_f16_ts(tmem + TM_O, tmem + TM_S + st * BN + kk * 8, bd, idesc_o, (j | kk) != 0);
                }
                tc_commit(v_empty + vs);
                tc_commit(bar_o);          // phase j: O includes blocks 0..j
            }
            tc_commit(bar_done);           // every product has retired: O is final
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ softmax / epilogue warps
        const int q = warp & 3;                                   // TMEM lane quarter
        const int half = (warp - 2) >> 2;                         // which 32 of a block's 64 keys / which half of O
        const int row = q0 + q * 32 + lane;                       // query index inside the video
        const int r128 = q * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const bool p_bf16 = in_dtype == DT_BF16;
        float* xmax = reinterpret_cast<float*>(sm + OFF_XCH);     // [2][2][BM]
        float* xsum = xmax + 2 * 2 * BM;                          // [2][BM]
        // ---- Q -> tensor memory (A operand of kind::f16 with M = 128: lane = query row, 32-bit column c = elements
        // 2c, 2c + 1): this thread's row, its half of the head dimension (128 elements = 64 columns), straight from
        // global memory.  Rows past the end of the buffer read as zero (they are never stored).
        if (QT) {
            const int64_t grow_q = static_cast<int64_t>(base) + row;
            const uint4* src = reinterpret_cast<const uint4*>(qkv + grow_q * (3 * static_cast<int64_t>(E)) + head * DH + half * (DH / 2));
            const bool in_buf = grow_q < total_rows;
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                uint32_t qv[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 v = in_buf ? __ldg(src + part * 8 + i) : make_uint4(0, 0, 0, 0);
                    qv[4 * i] = v.x; qv[4 * i + 1] = v.y; qv[4 * i + 2] = v.z; qv[4 * i + 3] = v.w;
                }
                tmem_st_32x32(tmem + lane_addr + TM_Q + half * 64 + part * 32, qv);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_q);
        }
        float m_used = -INFINITY;   // the max the stored exponentials are relative to (scaled units); identical in both warps of a row
        float l = 0.f;              // this warp's share of the row sum
        for (int j = 0; j < nblk; ++j) {
            const int st = j & 1;
            mbar_wait(bar_s + st, (j >> 1) & 1);
            tc_fence_after();
            uint32_t sv[32];
            tmem_ld_32x32(tmem + lane_addr + TM_S + st * BN + half * 32, sv);
            tmem_ld_wait();
            const int nk = min(BN, len - j * BN) - half * 32;   // valid keys among this warp's 32 columns
            float bmax = -INFINITY;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float a = (c < nk) ? __uint_as_float(sv[c]) * scale_log2 : -INFINITY;
                sv[c] = __float_as_uint(a);
                bmax = fmaxf(bmax, a);
            }
            // block max of the row = max over both warps' halves.  Double-buffered by block parity: the barrier of
            // block j + 1 orders every read of buffer (j & 1) before its next writes in block j + 2.  The barrier also
            // orders this warp's tcgen05.ld of S_j (complete: wait::ld above) before the OTHER warp's store of P_j,
            // which overwrites S_j's columns [16, 32).
            xmax[(st * 2 + half) * BM + r128] = bmax;
            named_bar_sync(1 + q, 64);
            bmax = fmaxf(bmax, xmax[(st * 2 + (half ^ 1)) * BM + r128]);
            const bool grow = bmax > m_used + LAZY_THRESHOLD;   // also true on the first block (m_used = -inf)
            if (__any_sync(0xffffffffu, grow)) {
                const float m_new = grow ? bmax : m_used;
                const float alpha = grow ? ex2f(m_used - m_new) : 1.0f;   // exp2(-inf) = 0 on the first block
                l *= alpha;
                m_used = m_new;
                if (j > 0) {
                    // P_{j-1} V_{j-1} has landed in O.  A parity wait only separates ADJACENT phases of bar_o; it is
                    // safe here because S_j (which this warp has just read) was issued after P_{j-2} V_{j-2} and
                    // the tensor pipe retires in order: bar_o has completed at least phase j - 2, and cannot
                    // complete phase j before this warp arrives on bar_p[j].
                    mbar_wait(bar_o, (j - 1) & 1);
                    tc_fence_after();
#pragma unroll 1
                    for (int c = half * (DH / 2); c < (half + 1) * (DH / 2); c += 32) {
                        uint32_t o[32];
                        tmem_ld_32x32(tmem + lane_addr + TM_O + c, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int t = 0; t < 32; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
                        tmem_st_32x32(tmem + lane_addr + TM_O + c, o);
                    }
                    tmem_st_wait();
                }
            }
            uint32_t pk[16];
            float lsum = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                const float p0 = ex2f(__uint_as_float(sv[c]) - m_used), p1 = ex2f(__uint_as_float(sv[c + 1]) - m_used);
                // the row sum uses the SAME rounded values the tensor core will multiply with
                if (p_bf16) {
                    __nv_bfloat162 h01 = __floats2bfloat162_rn(p0, p1);
                    lsum += __low2float(h01) + __high2float(h01);
                    pk[c >> 1] = *reinterpret_cast<uint32_t*>(&h01);
                } else {
                    __half2 h01 = __floats2half2_rn(p0, p1);
                    lsum += __low2float(h01) + __high2float(h01);
                    pk[c >> 1] = *reinterpret_cast<uint32_t*>(&h01);
                }
            }
            l += lsum;
            // P_j (16-bit pairs) over S_j's first 32 columns: this warp's keys [32 half, 32 half + 32) -> columns [16 half, + 16)
            tmem_st_32x16(tmem + lane_addr + TM_S + st * BN + half * 16, pk);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_p + st);
        }
        // ---- epilogue: O / l -> ctx
        // bar_o completes one phase per key block and a parity wait can only tell ADJACENT phases apart: a warp
        // that runs ahead of the slowest warp finishes its last block while P_{n-2} V_{n-2} may not even have been
        // issued (the MMA thread still waits for the slow warp's P_{n-2}), and a wait on the last phase's parity
        // would pass at once on the parity of phase n-3.  The end of the last product therefore has its own barrier.
        xsum[half * BM + r128] = l;
        named_bar_sync(1 + q, 64);
        l += xsum[(half ^ 1) * BM + r128];
        mbar_wait(bar_done, 0);
        tc_fence_after();
        const float inv_l = 1.0f / l;
        const bool row_ok = row < len;
        float* dst = reinterpret_cast<float*>(ctx) + static_cast<int64_t>(base + row) * ld_ctx + head * DH;
        uint16_t* dst_h = reinterpret_cast<uint16_t*>(ctx) + static_cast<int64_t>(base + row) * ld_ctx + head * DH;
#pragma unroll 1
        for (int c = half * (DH / 2); c < (half + 1) * (DH / 2); c += 32) {
            uint32_t o[32];
            tmem_ld_32x32(tmem + lane_addr + TM_O + c, o);
            tmem_ld_wait();
            if (row_ok && out_dtype != DT_F32) {
#pragma unroll
                for (int t = 0; t < 32; t += 8) {
                    uint32_t pk[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        pk[u] = pack_lowp2(__uint_as_float(o[t + 2 * u]) * inv_l, __uint_as_float(o[t + 2 * u + 1]) * inv_l,
                                           out_dtype);
                    *reinterpret_cast<uint4*>(dst_h + c + t) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            } else if (row_ok) {
#pragma unroll
                for (int t = 0; t < 32; t += 4) {
                    float4 v = make_float4(__uint_as_float(o[t]) * inv_l, __uint_as_float(o[t + 1]) * inv_l,
                                           __uint_as_float(o[t + 2]) * inv_l, __uint_as_float(o[t + 3]) * inv_l);
                    if (round_tf32) v = make_float4(to_tf32_rn(v.x), to_tf32_rn(v.y), to_tf32_rn(v.z), to_tf32_rn(v.w));
                    *reinterpret_cast<float4*>(dst + c + t) = v;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// qkv_h: fp16 [rows, 3E] (q | k | v).  Sequences must be contiguous rows (stride 1) and E / H == 256.
avs_status attention_tc(const void* qkv_h, int in_dtype, int64_t rows, int E, int H, const SeqDesc& seqs, void* ctx,
                        int64_t ld_ctx, int out_dtype, int round_tf32, cudaStream_t stream) {
    if (seqs.n_seqs == 0 || seqs.max_len == 0) return AVS_OK;
    AVS_CHECK(in_dtype == DT_F16 || in_dtype == DT_BF16, AVS_ERR_INVALID, "attention_tc: q|k|v must be fp16 or bf16");
    AVS_CHECK(H > 0 && E == H * DH, AVS_ERR_UNSUPPORTED, "attention_tc: head dim must be 256 (E=%d, heads=%d)", E, H);
    AVS_CHECK(seqs.n_seqs <= 65535, AVS_ERR_UNSUPPORTED, "attention_tc: too many sequences in one launch");
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        AVS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        AVS_CHECK(qres == cudaDriverEntryPointSuccess && p, AVS_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        encode = reinterpret_cast<EncodeTiledFn>(p);
    }
    CUtensorMap tm;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(3 * E), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(3 * E) * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm, in_dtype == DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(qkv_h), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AVS_CHECK(r == CUDA_SUCCESS, AVS_ERR_CUDA, "cuTensorMapEncodeTiled(qkv) failed with CUresult %d", static_cast<int>(r));
    static PerDeviceOnce configured;
    const int dev = current_device();
    if (configured.needed(dev)) {
        AVS_CUDA(cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<true>::SMEM_TOTAL));
        AVS_CUDA(cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<false>::SMEM_TOTAL));
        configured.mark(dev);
    }
    dim3 grid((seqs.max_len + BM - 1) / BM, H, seqs.n_seqs);
    const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(DH));
    // Q in tensor memory for long sequences (AVS_ATTN_Q: "tmem" / "smem" forces one variant -- tests run both)
    const char* force = getenv("AVS_ATTN_Q");
    const bool qt = force ? force[0] == 't' : seqs.max_len >= 1024;
    if (qt)
        attention_tc_kernel<true><<<grid, ATT_THREADS, Lay<true>::SMEM_TOTAL, stream>>>(
            tm, static_cast<const uint16_t*>(qkv_h), rows, seqs, E, in_dtype, ctx, ld_ctx, out_dtype, round_tf32, scale_log2);
    else
        attention_tc_kernel<false><<<grid, ATT_THREADS, Lay<false>::SMEM_TOTAL, stream>>>(
            tm, static_cast<const uint16_t*>(qkv_h), rows, seqs, E, in_dtype, ctx, ld_ctx, out_dtype, round_tf32, scale_log2);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

}  // namespace avs
