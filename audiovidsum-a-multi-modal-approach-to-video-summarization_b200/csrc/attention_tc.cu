// K4 -- temporal multi-head self-attention core on tcgen05 / TMEM / TMA (flash-style, never
// materialises the [T, T] score matrix that /root/reference/models/attention.py:21-22 builds).
//
// Arithmetic of attention.py:17-23 (== nn.MultiheadAttention's core, av_model.py:26,44 on the
// transposed tensor): contiguous dh = 256 wide heads, softmax(Q K^T / sqrt(dh)) V over the frames
// of ONE video, keys >= len masked (variable-length batches, packed rows).
//
// One CTA per (128-query block, head, video):
//   warp 0      TMA producer: Q once (64 KB), then K_j / V_j blocks of 64 keys (32 KB each, 2 stages)
//   warp 1      tcgen05.mma issuer:  S_j = Q K_j^T   (SS, 128 x 64 x 256, fp16 in / fp32 acc, TMEM)
//                                    O  += P_j V_j   (TS: P_j from TMEM, V_j MN-major from smem, 128 x 256 x 64)
//   warps 2..5  softmax: one query row per thread; tcgen05.ld S_j, running max with LAZY rescaling
//               (O is only rescaled when the max grows by > 2^8), exp2 on the SFU, P_j -> fp16 ->
//               tcgen05.st over S_j's columns; final O / l -> global.
// TMEM columns: O [0,256), S_0 / P_0 [256,320), S_1 / P_1 [320,384).  S is double buffered so the
// tensor pipe computes S_{j+1} while the softmax warps work on S_j.
// fp16 operands carry the same 11-bit significand as tf32; |q.k| stays far below the fp16 range
// because the inputs of the projection are LSTM outputs in (-1, 1).
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace avs {

namespace {

constexpr int DH = 256;
constexpr int BM = 128;   // queries per CTA
constexpr int BN = 64;    // keys per block
constexpr int Q_SUB = BM * 128;        // one dh-block (64 elements) of Q: 128 rows x 128 B
constexpr int KV_SUB = BN * 128;       // one dh-block of a K / V block: 64 rows x 128 B
constexpr int Q_BYTES = 4 * Q_SUB;     // 64 KB
constexpr int KV_BYTES = 4 * KV_SUB;   // 32 KB
constexpr int OFF_Q = 0;
constexpr int OFF_K = Q_BYTES;
constexpr int OFF_V = OFF_K + 2 * KV_BYTES;
constexpr int OFF_BAR = OFF_V + 2 * KV_BYTES;
constexpr int N_BARS = 1 + 4 + 4 + 2 + 2 + 1 + 1;  // q, k_full/empty[2], v_full/empty[2], s[2], p[2], o, done
constexpr int SMEM_TOTAL = 1024 + OFF_BAR + N_BARS * 8 + 16;
constexpr int ATT_THREADS = 192;
constexpr uint32_t TM_O = 0, TM_S = 256;       // TMEM column offsets
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

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, SeqDesc seqs, int E, int in_dtype,
                    void* __restrict__ ctx, int64_t ld_ctx, int out_dtype, int round_tf32, float scale_log2) {
    const int seq = blockIdx.z, head = blockIdx.y;
    const int len = seqs.len[seq];
    const int q0 = blockIdx.x * BM;
    if (q0 >= len) return;
    const int base = seqs.base[seq];
    const int nblk = (len + BN - 1) / BN;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint64_t* bar_q = bars;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = bars + 3;
    uint64_t* v_full = bars + 5;
    uint64_t* v_empty = bars + 7;
    uint64_t* bar_s = bars + 9;
    uint64_t* bar_p = bars + 11;
    uint64_t* bar_o = bars + 13;
    uint64_t* bar_done = bars + 14;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_qkv);
        mbar_init(bar_q, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(k_full + i, 1);
            mbar_init(k_empty + i, 1);
            mbar_init(v_full + i, 1);
            mbar_init(v_empty + i, 1);
            mbar_init(bar_s + i, 1);
            mbar_init(bar_p + i, 128);
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
            const int qcol = head * DH, kcol = E + head * DH, vcol = 2 * E + head * DH;
            mbar_expect_tx(bar_q, Q_BYTES);
            for (int d = 0; d < 4; ++d)
                for (int hf = 0; hf < 2; ++hf)
                    tma_load_2d(sm + OFF_Q + d * Q_SUB + hf * (64 * 128), &tm_qkv, bar_q, qcol + d * 64, base + q0 + hf * 64);
            for (int j = 0; j < nblk; ++j) {
                const int st = j & 1;
                const uint32_t ph = (j >> 1) & 1;
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
            const uint32_t q_addr = smem_u32(sm + OFF_Q);
            auto issue_s = [&](int j) {
                const int st = j & 1;
                mbar_wait(k_full + st, (j >> 1) & 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sm + OFF_K + st * KV_BYTES);
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) {
                    const uint64_t ad = umma_desc_sw128_kmajor(q_addr + (kk >> 2) * Q_SUB + (kk & 3) * 32);
                    const uint64_t bd = umma_desc_sw128_kmajor(k_addr + (kk >> 2) * KV_SUB + (kk & 3) * 32);
                    umma_f16_ss(tmem + TM_S + st * BN, ad, bd, idesc_s, kk != 0);
                }
                tc_commit(k_empty + st);   // K stage reusable once these MMAs retire
                tc_commit(bar_s + st);     // S_j ready for the softmax warps
            };
            mbar_wait(bar_q, 0);
            issue_s(0);
            for (int j = 0; j < nblk; ++j) {
                const int st = j & 1;
                if (j + 1 < nblk) issue_s(j + 1);   // overlaps the softmax of block j
                mbar_wait(bar_p + st, (j >> 1) & 1);  // P_j stored (and O rescaled if it had to be)
                mbar_wait(v_full + st, (j >> 1) & 1);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(sm + OFF_V + st * KV_BYTES);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {     // 16 keys per MMA
                    const uint64_t bd = umma_desc_sw128_mnmajor(v_addr + kk * 2048, KV_SUB, 1024);
                    umma_f16_ts(tmem + TM_O, tmem + TM_S + st * BN + kk * 8, bd, idesc_o, (j | kk) != 0);
                }
                tc_commit(v_empty + st);
                tc_commit(bar_o);          // phase j: O includes blocks 0..j
            }
            tc_commit(bar_done);           // every product has retired: O is final
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ softmax / epilogue warps
        const int q = warp & 3;
        const int row = q0 + q * 32 + lane;                       // query index inside the video
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const bool p_bf16 = in_dtype == DT_BF16;
        float m_used = -INFINITY;   // the max the stored exponentials are relative to (scaled units)
        float l = 0.f;
        for (int j = 0; j < nblk; ++j) {
            const int st = j & 1;
            mbar_wait(bar_s + st, (j >> 1) & 1);
            tc_fence_after();
            uint32_t s0[32], s1[32];
            tmem_ld_32x32(tmem + lane_addr + TM_S + st * BN, s0);
            tmem_ld_32x32(tmem + lane_addr + TM_S + st * BN + 32, s1);
            tmem_ld_wait();
            const int nk = min(BN, len - j * BN);   // valid keys in this block
            float bmax = -INFINITY;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float a = (c < nk) ? __uint_as_float(s0[c]) * scale_log2 : -INFINITY;
                const float b = (c + 32 < nk) ? __uint_as_float(s1[c]) * scale_log2 : -INFINITY;
                s0[c] = __float_as_uint(a);
                s1[c] = __float_as_uint(b);
                bmax = fmaxf(bmax, fmaxf(a, b));
            }
            const bool grow = bmax > m_used + LAZY_THRESHOLD;   // also true on the first block (m_used = -inf)
            if (__any_sync(0xffffffffu, grow)) {
                const float m_new = grow ? bmax : m_used;
                const float alpha = grow ? ex2f(m_used - m_new) : 1.0f;   // exp2(-inf) = 0 on the first block
                l *= alpha;
                m_used = m_new;
                if (j > 0) {
                    mbar_wait(bar_o, (j - 1) & 1);   // P_{j-1} V_{j-1} has landed in O
                    tc_fence_after();
#pragma unroll 1
                    for (int c = 0; c < DH; c += 32) {
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
            uint32_t pk[32];
            float lsum = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
                const float p0 = ex2f(__uint_as_float(s0[c]) - m_used), p1 = ex2f(__uint_as_float(s0[c + 1]) - m_used);
                const float p2 = ex2f(__uint_as_float(s1[c]) - m_used), p3 = ex2f(__uint_as_float(s1[c + 1]) - m_used);
                // the row sum uses the SAME rounded values the tensor core will multiply with
                if (p_bf16) {
                    __nv_bfloat162 h01 = __floats2bfloat162_rn(p0, p1), h23 = __floats2bfloat162_rn(p2, p3);
                    lsum += (__low2float(h01) + __high2float(h01)) + (__low2float(h23) + __high2float(h23));
                    pk[c >> 1] = *reinterpret_cast<uint32_t*>(&h01);
                    pk[16 + (c >> 1)] = *reinterpret_cast<uint32_t*>(&h23);
                } else {
                    __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
                    lsum += (__low2float(h01) + __high2float(h01)) + (__low2float(h23) + __high2float(h23));
                    pk[c >> 1] = *reinterpret_cast<uint32_t*>(&h01);
                    pk[16 + (c >> 1)] = *reinterpret_cast<uint32_t*>(&h23);
                }
            }
            l += lsum;
            tmem_st_32x32(tmem + lane_addr + TM_S + st * BN, pk);   // P_j (fp16 pairs) over S_j's first 32 columns
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_p + st);
        }
        // ---- epilogue: O / l -> ctx
        // bar_o completes one phase per key block and a parity wait can only tell ADJACENT phases apart: a warp
        // that runs ahead of the slowest warp finishes its last block while P_{n-2} V_{n-2} may not even have been
        // issued (the MMA thread still waits for the slow warp's P_{n-2}), and a wait on the last phase's parity
        // would pass at once on the parity of phase n-3.  The end of the last product therefore has its own barrier.
        mbar_wait(bar_done, 0);
        tc_fence_after();
        const float inv_l = 1.0f / l;
        const bool row_ok = row < len;
        float* dst = reinterpret_cast<float*>(ctx) + static_cast<int64_t>(base + row) * ld_ctx + head * DH;
        uint16_t* dst_h = reinterpret_cast<uint16_t*>(ctx) + static_cast<int64_t>(base + row) * ld_ctx + head * DH;
#pragma unroll 1
        for (int c = 0; c < DH; c += 32) {
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
        AVS_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        configured.mark(dev);
    }
    dim3 grid((seqs.max_len + BM - 1) / BM, H, seqs.n_seqs);
    const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(DH));
    attention_tc_kernel<<<grid, ATT_THREADS, SMEM_TOTAL, stream>>>(tm, seqs, E, in_dtype, ctx, ld_ctx, out_dtype,
                                                                   round_tf32, scale_log2);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

}  // namespace avs
