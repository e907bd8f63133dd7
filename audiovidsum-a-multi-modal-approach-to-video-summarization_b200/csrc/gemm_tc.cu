// K1/K2a/K3/K5/K6 -- the dense contractions of AVBiLSTMModel.forward
// (nn.Linear call sites /root/reference/models/av_model.py:10-15, 26, 29-31 and the
// LSTM input projections of av_model.py:18-23) as ONE tcgen05 + TMA + TMEM GEMM:
//
//     C[M, N] = epilogue(A[M, K] * W[N, K]^T)          (both operands K-major)
//
// Persistent kernel, one CTA per SM, static round-robin over 128 x BN output tiles:
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B-swizzled tiles, 6-8 stage mbarrier ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (kind::tf32 or kind::f16)
//   warps 2..9  : epilogue (two warps per TMEM lane quarter, half of the tile's columns each; measured: with four
//                 warps the fp32-output GEMMs were epilogue bound, tools/gemm_trace.py) -- tcgen05.ld the fp32 accumulator (double buffered in TMEM, so it
//                 overlaps the next tile's main loop), fuse bias / ReLU / tf32 rounding /
//                 fp16-bf16 cast, stage the warp's 32 rows in 128B-swizzled shared memory and write
//                 them with TMA stores (cp.async.bulk.tensor ... bulk_group: full-line coalesced
//                 writes, rows >= M clipped by the tensor map); or the whole frame-score head
//                 (64-wide ReLU, dot with scorer.2.weight, sigmoid), one float per row.
//
// Tile 128 x BN x 128 B of K per stage (32 tf32 or 64 half elements), 4 MMAs (K = 32 B) per stage.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace avs {

namespace {

constexpr int BM = 128;
constexpr int A_STAGE_BYTES = BM * 128;
constexpr int EPI_WARPS = 8;                      // two per TMEM lane quarter: each takes half of the tile's columns
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;

template <int BN, int STAGES>
struct SmemLayout {
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + BN * 128;
    static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
    // one epilogue warp's staging: its 32 rows of the tile in fp32; the 256-wide tile has room for half of
    // that (4 boxes of 32 rows x 128 B) and cycles through it twice for fp32 outputs
    static constexpr int OUT_WARP_BYTES = 32 * BN * (BN == 256 ? 2 : 4);
    static constexpr int OUT_BYTES = 4 * OUT_WARP_BYTES;
    static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 8;
    static constexpr int TOTAL = 1024 /*alignment slack*/ + TILE_BYTES + OUT_BYTES + BAR_BYTES;
};

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// Optional pipeline trace (AVS_GEMM_TRACE=1, debugging aid): block 0 accumulates clock64 totals --
// [0] MMA thread busy span, [1] MMA waiting for operands (full), [2] MMA waiting for a drained accumulator,
// [3] producer waiting for a free stage, [4] epilogue warp waiting for an accumulator, [5] epilogue span, [6] tiles.
__device__ int g_gemm_trace_on = 0;
__device__ unsigned long long g_gemm_trace[8];
__device__ __forceinline__ long long gclk() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    return t;
}

// TMA store shared::cta -> global (bulk async group completion).
__device__ __forceinline__ void tma_store_2d(const void* desc, uint32_t smem_src, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(desc)),
                 "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most n of this thread's most recent bulk groups still READ their shared-memory source
__device__ __forceinline__ void bulk_wait_group_read(int n) {
    switch (n) {
        case 0: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
        case 3: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
        default: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
    }
}
// 32 consecutive bias values for columns [nb, nb + 32) as independent vector loads (every lane reads the same
// addresses: one broadcast transaction each).  N % 16 == 0, so validity is decided per 16 columns.
__device__ __forceinline__ void load_bias32(const float* __restrict__ bias, int nb, int N, bool vec_ok, float (&bv)[32]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (bias != nullptr && nb + 16 * h < N) {
            if (vec_ok) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(bias + nb + 16 * h) + j);
                    bv[16 * h + 4 * j] = v.x;
                    bv[16 * h + 4 * j + 1] = v.y;
                    bv[16 * h + 4 * j + 2] = v.z;
                    bv[16 * h + 4 * j + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) bv[16 * h + j] = __ldg(bias + nb + 16 * h + j);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) bv[16 * h + j] = 0.f;
        }
    }
}
// arrive on an mbarrier given by its shared::cluster address (own CTA or the pair's leader)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// The epilogue warps (4 warps, one per TMEM lane quarter) of both GEMM kernels.  Tiles t = tile0, tile0 + tile_step,
// ...; this CTA owns rows [m0, m0 + 128) of the tile with m0 = (t / tiles_n) * tile_rows + m_off.  tmem_empty_addr:
// shared::cluster address of the two "accumulator drained" barriers (in the leader CTA for the 2-CTA kernel).
template <int BN, class L>
__device__ __forceinline__ void gemm_epilogue(uint8_t* tiles, uint32_t tmem_base, uint64_t* tmem_full,
                                              uint32_t tmem_empty_addr, const CUtensorMap& tmC, int M, int N, int tiles_n,
                                              int num_tiles, int tile0, int tile_step, int tile_rows, int m_off,
                                              const GemmEpilogue& epi, int warp, int lane) {
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int cg = (warp - 2) >> 2;   // column group: which half of the tile's columns this warp handles
        const uint32_t lane_taddr = static_cast<uint32_t>(q * 32) << 16;
        if (epi.scores != nullptr) {
            // fused frame-score head: one float per row, straight from registers (64 columns: column group 0 only)
            const bool bias_vec = (reinterpret_cast<uintptr_t>(epi.bias) & 15) == 0;
            const bool w2_vec = (reinterpret_cast<uintptr_t>(epi.score_w2) & 15) == 0;
            uint32_t lt = 0;
            for (int t = tile0; t < num_tiles; t += tile_step, ++lt) {
                const int m0 = (t / tiles_n) * tile_rows + m_off, n0 = (t % tiles_n) * BN;
                const uint32_t acc = lt & 1;
                mbar_wait(tmem_full + acc, (lt >> 1) & 1);
                tc_fence_after();
                if (cg != 0) {
                    mbar_arrive_cluster(tmem_empty_addr + acc * 8);
                    continue;
                }
                const int row = m0 + q * 32 + lane;
                float score_acc = 0.f;
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    uint32_t r[32];
                    float bv[32], wv[32];
                    const int nb = n0 + c;
                    tmem_ld_32x32(tmem_base + acc * BN + lane_taddr + c, r);
                    load_bias32(epi.bias, nb, N, bias_vec, bv);
                    load_bias32(epi.score_w2, nb, N, w2_vec, wv);   // zero beyond N: those columns drop out
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = fmaf(__uint_as_float(r[j]), epi.acc_scale, bv[j]);
                        if (epi.relu) x = fmaxf(x, 0.f);
                        score_acc = fmaf(x, wv[j], score_acc);
                    }
                }
                tc_fence_before();
                mbar_arrive_cluster(tmem_empty_addr + acc * 8);
                if (row < M) epi.scores[row] = sigmoidf_acc(score_acc + __ldg(epi.score_b2));
            }
        } else {
            // stage 32 rows x 128 B boxes (32 fp32 or 64 half columns) in swizzled shared memory, TMA-store each
            const bool wide = epi.out_dtype != DT_F32;          // 16-bit output: 64 columns per 128-byte box
            const int box_cols = wide ? 64 : 32;
            const int n_boxes = BN / box_cols;
            const int bx_mid = (n_boxes + 1) / 2;               // column group 0: boxes [0, mid), group 1: [mid, n_boxes)
            const int bx_lo = cg ? bx_mid : 0, bx_hi = cg ? n_boxes : bx_mid;
            constexpr int SLOTS = L::OUT_WARP_BYTES / 2 / 4096;   // staging boxes per warp (cycled when a tile has more)
            static_assert(SLOTS >= 1, "staging area too small");
            uint32_t issued = 0;                                // boxes handed to the TMA engine so far
            const uint32_t out_base = smem_u32(tiles + L::TILE_BYTES) + (q * 2 + cg) * (L::OUT_WARP_BYTES / 2);
            const uint32_t my_row = out_base + lane * 128;
            const uint32_t sw = static_cast<uint32_t>(lane & 7);
            const bool bias_vec = (reinterpret_cast<uintptr_t>(epi.bias) & 15) == 0;
            const bool tr = g_gemm_trace_on && blockIdx.x == 0 && warp == 2 && lane == 0;   // (warp 2: q = 2, cg = 0)
            long long tr_wait = 0, tr_t0 = tr ? gclk() : 0;
            // range watch of fp16 outputs (epi.sat_flag): the cast saturates at 65504 instead of overflowing, which is
            // silent; the largest magnitude this thread produced is tracked in a register and reported ONCE, after the
            // last tile, with a plain store into the (mapped, pinned) flag word.  (NaN does not count: fmaxf drops it,
            // and a NaN feature shows up as a NaN score.)
            float sat_big = 0.f;
            uint32_t lt = 0;
            for (int t = tile0; t < num_tiles; t += tile_step, ++lt) {
                const int m0 = (t / tiles_n) * tile_rows + m_off, n0 = (t % tiles_n) * BN;
                const uint32_t acc = lt & 1;
                const long long w0 = tr ? gclk() : 0;
                mbar_wait(tmem_full + acc, (lt >> 1) & 1);
                if (tr) tr_wait += gclk() - w0;
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * BN + lane_taddr;
#pragma unroll 1
                for (int bx = bx_lo; bx < bx_hi; ++bx) {
                    const int nb = n0 + bx * box_cols;
                    // the TMA store that read this staging slot SLOTS boxes ago must be done reading
                    const uint32_t slot = issued % SLOTS;
                    if (issued >= SLOTS) {
                        if (lane == 0) bulk_wait_group_read(SLOTS - 1);
                        __syncwarp();
                    }
                    ++issued;
                    const uint32_t dst = my_row + slot * 4096;
                    if (!wide) {
                        uint32_t r[32];
                        float bv[32];
                        tmem_ld_32x32(taddr + bx * 32, r);
                        load_bias32(epi.bias, nb, N, bias_vec, bv);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float x = fmaf(__uint_as_float(r[j]), epi.acc_scale, bv[j]);
                            if (epi.relu) x = fmaxf(x, 0.f);
                            if (epi.round_tf32) x = to_tf32_rn(x);
                            r[j] = __float_as_uint(x);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            st_shared_v4(dst + ((static_cast<uint32_t>(j) ^ sw) << 4), r[4 * j], r[4 * j + 1],
                                         r[4 * j + 2], r[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            uint32_t r[32];
                            float bv[32];
                            tmem_ld_32x32(taddr + bx * 64 + hf * 32, r);
                            load_bias32(epi.bias, nb + hf * 32, N, bias_vec, bv);
                            tmem_ld_wait();
                            uint32_t pk[16];
                            float big = 0.f;   // largest magnitude of this thread's values (range watch, see below)
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                float x0 = fmaf(__uint_as_float(r[j]), epi.acc_scale, bv[j]), x1 = fmaf(__uint_as_float(r[j + 1]), epi.acc_scale, bv[j + 1]);
                                if (epi.relu) {
                                    x0 = fmaxf(x0, 0.f);
                                    x1 = fmaxf(x1, 0.f);
                                }
                                big = fmaxf(big, fmaxf(fabsf(x0), fabsf(x1)));
                                pk[j >> 1] = pack_lowp2(x0, x1, epi.out_dtype);
                            }
                            sat_big = fmaxf(sat_big, big);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                st_shared_v4(dst + ((static_cast<uint32_t>(hf * 4 + j) ^ sw) << 4), pk[4 * j],
                                             pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                        }
                    }
                    fence_proxy_async();   // generic-proxy smem writes -> visible to the TMA engine
                    __syncwarp();
                    if (lane == 0) {
                        if (nb < N) tma_store_2d(&tmC, out_base + slot * 4096, nb, m0 + q * 32);
                        bulk_commit_group();   // (possibly empty) keeps the group count per tile fixed
                    }
                }
                // all TMEM reads of this accumulator are complete (tmem_ld_wait above): hand it back
                tc_fence_before();
                mbar_arrive_cluster(tmem_empty_addr + acc * 8);
            }
            if (lane == 0) bulk_wait_group_read(0);   // shared memory must outlive the last stores' reads
            __syncwarp();
            // (rows beyond M and columns beyond N are zero-filled by TMA and carry the bias only: they cannot trip it)
            if (epi.sat_flag != nullptr && epi.out_dtype == DT_F16 && !(sat_big < 65504.f))
                *reinterpret_cast<volatile unsigned int*>(epi.sat_flag) = 1u;
            if (tr) {
                g_gemm_trace[4] += tr_wait;
                g_gemm_trace[5] += gclk() - tr_t0;
                g_gemm_trace[6] += lt;
            }
        }
}

// Persistent: grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, + gridDim.x, ...
// (n fastest, so CTAs that run concurrently share the same A row block in L2).  The TMA and MMA
// pipelines run straight through tile boundaries; the accumulator is double buffered in TMEM
// (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i + 1.
template <int BN, int STAGES, bool TF32>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, int M, int N, int k_blocks, int bk_elems, uint32_t idesc,
               int tiles_n, int num_tiles, GemmEpilogue epi) {
    using L = SmemLayout<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* tiles = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + L::TILE_BYTES + L::OUT_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (epi.scores == nullptr) tma_prefetch_desc(&tmC);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full + a, 1);
            mbar_init(tmem_empty + a, EPI_WARPS * 32);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 2 * BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one()) {
            const bool tr = g_gemm_trace_on && blockIdx.x == 0;
            long long tr_wait = 0;
            uint32_t it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    const long long w0 = tr ? gclk() : 0;
                    mbar_wait(empty + s, ph ^ 1);
                    if (tr) tr_wait += gclk() - w0;
                    mbar_expect_tx(full + s, L::STAGE_BYTES);
                    uint8_t* a_dst = tiles + s * L::STAGE_BYTES;
                    tma_load_2d(a_dst, &tmA, full + s, kb * bk_elems, m0);
                    tma_load_2d(a_dst + A_STAGE_BYTES, &tmB, full + s, kb * bk_elems, n0);
                }
            }
            if (tr) g_gemm_trace[3] += tr_wait;
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            const bool tr = g_gemm_trace_on && blockIdx.x == 0;
            long long tr_full = 0, tr_acc = 0, tr_t0 = tr ? gclk() : 0;
            uint32_t it = 0, lt = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
                const uint32_t acc = lt & 1;
                long long w0 = tr ? gclk() : 0;
                mbar_wait(tmem_empty + acc, ((lt >> 1) & 1) ^ 1);   // epilogue drained this accumulator
                if (tr) tr_acc += gclk() - w0;
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    w0 = tr ? gclk() : 0;
                    mbar_wait(full + s, ph);
                    if (tr) tr_full += gclk() - w0;
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + s * L::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // 4 x (K = 32 bytes) per 128-byte swizzled row
                        const uint64_t ad = umma_desc_sw128_kmajor(a_addr + k * 32);
                        const uint64_t bd = umma_desc_sw128_kmajor(b_addr + k * 32);
                        if (TF32)
                            umma_tf32_ss(tmem_acc, ad, bd, idesc, (kb | k) != 0);
                        else
                            umma_f16_ss(tmem_acc, ad, bd, idesc, (kb | k) != 0);
                    }
                    tc_commit(empty + s);  // smem slot reusable once these MMAs retire
                }
                tc_commit(tmem_full + acc);  // accumulator complete
            }
            if (tr) {
                g_gemm_trace[0] += gclk() - tr_t0;
                g_gemm_trace[1] += tr_full;
                g_gemm_trace[2] += tr_acc;
            }
        }
        __syncwarp();
    } else {
        gemm_epilogue<BN, L>(tiles, tmem_base, tmem_full, smem_u32(tmem_empty), tmC, M, N, tiles_n, num_tiles,
                             static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x), BM, 0, epi, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

// ---------------------------------------------------------------- 2-CTA (cta_group::2) variant
// A CTA pair (cluster of 2, same TPC) computes one 256 x BN output tile: CTA r owns rows [m0 + 128 r, + 128) of A and
// of the accumulator (its own TMEM) and loads HALF of the B tile (BN / 2 weight rows); the leader's
// tcgen05.mma.cta_group::2 (M = 256) reads both halves.  Per CTA and k-block the operands are 16 KB of A + BN/4 KB of
// B instead of 16 + BN/2: a third less L2 -> SM traffic per FLOP, which is what bounds these GEMMs.
//   full[s]   (leader's): both CTAs' TMA loads complete_tx on it; the leader's producer arms it for both
//   empty[s]  (each CTA's own): tcgen05.commit.cta_group::2 with multicast to both CTAs
//   tmem_full[a]  (each CTA's own): multicast commit;  tmem_empty[a] (leader's): 2 x 128 epilogue threads arrive
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load into this CTA's shared memory, bytes counted on the barrier at shared::cluster address bar_cluster_addr
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const void* desc, uint32_t bar_cluster_addr, int32_t c0,
                                                 int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :
        : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// arrive (count 1) on the barrier at this offset in every CTA of cta_mask when the issued MMAs retire
__device__ __forceinline__ void tc_commit_2cta(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma_ss_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    if (TF32)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}\n"
            :
            : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n"
            :
            : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
// mbarrier wait with acquire at cluster scope (the arrivals come from the peer CTA's epilogue threads)
__device__ __forceinline__ void mbar_wait_cluster_acq(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0, ok = 0;
    const uint32_t addr = smem_u32(bar);
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred P;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > AVS_SPIN_LIMIT) {
#ifdef AVS_DEBUG_WAITS
            printf("avsum_b200: gemm accumulator wait timed out (block %d parity %u)\n", blockIdx.x, parity);
#endif
            __trap();
        }
    }
}

template <int BN, int STAGES>
struct SmemLayout2 {
    static constexpr int B_HALF_BYTES = (BN / 2) * 128;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_HALF_BYTES;
    static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int OUT_WARP_BYTES = 32 * BN * (BN == 256 ? 2 : 4);
    static constexpr int OUT_BYTES = 4 * OUT_WARP_BYTES;
    static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 8;
    static constexpr int TOTAL = 1024 /*alignment slack*/ + TILE_BYTES + OUT_BYTES + BAR_BYTES;
};

template <int BN, int STAGES, bool TF32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, int M, int N, int k_blocks, int bk_elems, uint32_t idesc,
                int tiles_n, int num_tiles, GemmEpilogue epi) {
    using L = SmemLayout2<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* tiles = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + L::TILE_BYTES + L::OUT_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = static_cast<int>(blockIdx.x >> 1);
    const int n_clusters = static_cast<int>(gridDim.x >> 1);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full + a, 1);
            mbar_init(tmem_empty + a, 2 * EPI_WARPS * 32);   // the epilogue threads of both CTAs
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc_2cta(tmem_slot, 2 * BN);
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // the peer's barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (elect_one()) {
            const uint32_t full_leader = mapa_u32(smem_u32(full), 0);
            const bool tr = g_gemm_trace_on && blockIdx.x == 0;
            long long tr_wait = 0;
            uint32_t it = 0;
            for (int t = cluster_id; t < num_tiles; t += n_clusters) {
                const int m0 = (t / tiles_n) * (2 * BM) + static_cast<int>(rank) * BM;
                const int n0 = (t % tiles_n) * BN + static_cast<int>(rank) * (BN / 2);
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    const long long w0 = tr ? gclk() : 0;
                    mbar_wait(empty + s, ph ^ 1);
                    if (tr) tr_wait += gclk() - w0;
                    if (rank == 0) mbar_expect_tx(full + s, 2 * L::STAGE_BYTES);
                    uint8_t* a_dst = tiles + s * L::STAGE_BYTES;
                    tma_load_2d_2cta(a_dst, &tmA, full_leader + s * 8, kb * bk_elems, m0);
                    tma_load_2d_2cta(a_dst + A_STAGE_BYTES, &tmB, full_leader + s * 8, kb * bk_elems, n0);
                }
            }
            if (tr) g_gemm_trace[3] += tr_wait;
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0 && elect_one()) {
            const bool tr = g_gemm_trace_on && blockIdx.x == 0;
            long long tr_full = 0, tr_acc = 0, tr_t0 = tr ? gclk() : 0;
            uint32_t it = 0, lt = 0;
            for (int t = cluster_id; t < num_tiles; t += n_clusters, ++lt) {
                const uint32_t acc = lt & 1;
                long long w0 = tr ? gclk() : 0;
                mbar_wait_cluster_acq(tmem_empty + acc, ((lt >> 1) & 1) ^ 1);   // both epilogues drained it
                if (tr) tr_acc += gclk() - w0;
                tc_fence_after();
                const uint32_t tmem_acc = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    w0 = tr ? gclk() : 0;
                    mbar_wait(full + s, ph);
                    if (tr) tr_full += gclk() - w0;
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + s * L::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + A_STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t ad = umma_desc_sw128_kmajor(a_addr + k * 32);
                        const uint64_t bd = umma_desc_sw128_kmajor(b_addr + k * 32);
                        umma_ss_2cta<TF32>(tmem_acc, ad, bd, idesc, (kb | k) != 0);
                    }
                    tc_commit_2cta(empty + s, 3);   // the slot is reusable in both CTAs once these MMAs retire
                }
                tc_commit_2cta(tmem_full + acc, 3);   // accumulator complete (both CTAs' epilogues)
            }
            if (tr) {
                g_gemm_trace[0] += gclk() - tr_t0;
                g_gemm_trace[1] += tr_full;
                g_gemm_trace[2] += tr_acc;
            }
        }
        __syncwarp();
    } else {
        gemm_epilogue<BN, L>(tiles, tmem_base, tmem_full, mapa_u32(smem_u32(tmem_empty), 0), tmC, M, N, tiles_n, num_tiles,
                             cluster_id, n_clusters, 2 * BM, static_cast<int>(rank) * BM, epi, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();      // nobody leaves while the pair's MMAs / commits may still touch its memory
    if (warp == 1) tmem_dealloc_2cta(tmem_base, 2 * BN);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D row-major tensor [rows, K] with leading dimension ld (elements): box = 128 bytes of a row x box_rows,
// 128-byte swizzle.  Used for the K-major operands (TMA loads) and for the output C (TMA stores).
avs_status make_tmap(CUtensorMap* tm, const void* ptr, int dtype, int64_t rows, int K, int64_t ld, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    AVS_CHECK(fn != nullptr, AVS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const int esz = dtype_size(dtype);
    AVS_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, AVS_ERR_INVALID, "GEMM operand not 16-byte aligned");
    AVS_CHECK((ld * esz) % 16 == 0, AVS_ERR_INVALID, "GEMM operand row pitch %lld B not a multiple of 16",
              static_cast<long long>(ld * esz));
    CUtensorMapDataType dt = dtype == DT_F32   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                             : dtype == DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                               : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld * esz)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, dt, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AVS_CHECK(r == CUDA_SUCCESS, AVS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return AVS_OK;
}

}  // namespace

static avs_status gemm_tc2(const void* A, int64_t lda, const void* W, int64_t ldw, int in_dtype, int64_t M, int N,
                           int K, const GemmEpilogue& epi, cudaStream_t stream) {
    constexpr int BN = 256;
    const int esz = dtype_size(in_dtype);
    const int bk_elems = 128 / esz;
    const int k_blocks = (K + bk_elems - 1) / bk_elems;
    const bool tf32 = in_dtype == DT_F32;
    const uint32_t fmt = tf32 ? UMMA_FMT_TF32 : (in_dtype == DT_F16 ? UMMA_FMT_F16 : UMMA_FMT_BF16);
    CUtensorMap tmA, tmB, tmC;
    AVS_TRY(make_tmap(&tmA, A, in_dtype, M, K, lda, BM));
    AVS_TRY(make_tmap(&tmB, W, in_dtype, N, K, ldw, BN / 2));
    AVS_TRY(make_tmap(&tmC, epi.C, epi.out_dtype, M, N, epi.ldc, 32));
    const int tiles_n = N / BN;
    const int64_t tiles_total = ((M + 2 * BM - 1) / (2 * BM)) * tiles_n;
    AVS_CHECK(tiles_total < (1ll << 31), AVS_ERR_UNSUPPORTED, "gemm: too many tiles");
    const int num_tiles = static_cast<int>(tiles_total);
    const int num_pairs = device_sm_count() / 2;
    AVS_CHECK(num_pairs > 0, AVS_ERR_CUDA, "gemm: could not read the SM count");
    int pairs_avail = num_pairs;
    if (epi.max_ctas > 0) pairs_avail = std::max(1, std::min(num_pairs, epi.max_ctas / 2));
    const int grid = 2 * (num_tiles < pairs_avail ? num_tiles : pairs_avail);
    const uint32_t idesc = umma_idesc(fmt, 2 * BM, BN);
#define AVS_GEMM2_LAUNCH(ST_, TF_)                                                                            \
    do {                                                                                                     \
        using L = SmemLayout2<BN, ST_>;                                                                      \
        auto kern = gemm_tc2_kernel<BN, ST_, TF_>;                                                           \
        static PerDeviceOnce configured;                                                                     \
        const int dev_ = current_device();                                                                   \
        if (configured.needed(dev_)) {                                                                       \
            AVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));     \
            configured.mark(dev_);                                                                           \
        }                                                                                                    \
        kern<<<grid, GEMM_THREADS, L::TOTAL, stream>>>(tmA, tmB, tmC, static_cast<int>(M), N, k_blocks,      \
                                                       bk_elems, idesc, tiles_n, num_tiles, epi);            \
    } while (0)
    if (tf32) AVS_GEMM2_LAUNCH(5, true);
    else AVS_GEMM2_LAUNCH(5, false);
#undef AVS_GEMM2_LAUNCH
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

// debugging aid: read and clear the pipeline trace (AVS_GEMM_TRACE=1)
avs_status gemm_trace_read(unsigned long long* out8) {
    AVS_CUDA(cudaDeviceSynchronize());
    AVS_CUDA(cudaMemcpyFromSymbol(out8, g_gemm_trace, 8 * sizeof(unsigned long long)));
    unsigned long long zero[8] = {};
    AVS_CUDA(cudaMemcpyToSymbol(g_gemm_trace, zero, sizeof(zero)));
    return AVS_OK;
}

avs_status gemm_tc(const void* A, int64_t lda, const void* W, int64_t ldw, int in_dtype, int64_t M, int N, int K,
                   const GemmEpilogue& epi, cudaStream_t stream) {
    if (M == 0) return AVS_OK;
    static const bool want_trace = getenv("AVS_GEMM_TRACE") != nullptr;
    static PerDeviceOnce trace_set;
    if (want_trace && trace_set.needed(current_device())) {
        const int on = 1;
        AVS_CUDA(cudaMemcpyToSymbol(g_gemm_trace_on, &on, sizeof(on)));
        trace_set.mark(current_device());
    }
    AVS_CHECK(M > 0 && N > 0 && K > 0, AVS_ERR_INVALID, "gemm: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
    AVS_CHECK(M < (1ll << 31), AVS_ERR_UNSUPPORTED, "gemm: M too large");
    // the bias / score epilogues read 16 columns at a time; a plain store only needs a 16-byte row pitch
    // (columns >= N are computed from zero-filled operand rows and clipped by the TMA store)
    AVS_CHECK(N % 16 == 0 || (epi.bias == nullptr && epi.scores == nullptr && (N * dtype_size(epi.out_dtype)) % 16 == 0),
              AVS_ERR_UNSUPPORTED, "gemm: N=%d must be a multiple of 16 (or of 4 without a bias, fp32 output)", N);
    if (epi.scores != nullptr)
        AVS_CHECK(N == 64 && epi.score_w2 && epi.score_b2, AVS_ERR_INVALID, "score epilogue needs N == 64");
    else
        AVS_CHECK(epi.C != nullptr && (epi.ldc * dtype_size(epi.out_dtype)) % 16 == 0 &&
                      (reinterpret_cast<uintptr_t>(epi.C) & 15) == 0,
                  AVS_ERR_INVALID, "gemm: output must be 16-byte aligned with a row pitch that is a multiple of 16 B");
    const int esz = dtype_size(in_dtype);
    const int bk_elems = 128 / esz;
    const int k_blocks = (K + bk_elems - 1) / bk_elems;
    const bool tf32 = in_dtype == DT_F32;
    const uint32_t fmt = tf32 ? UMMA_FMT_TF32 : (in_dtype == DT_F16 ? UMMA_FMT_F16 : UMMA_FMT_BF16);
    // 128 x 256 tiles move 25 % fewer operand bytes per FLOP through L2 (the bound of these GEMMs); used for
    // 16-bit operands when the tile count still balances over the SMs
    const int64_t m_tiles = (M + BM - 1) / BM;
    const bool wide_ok = N % 256 == 0 && epi.scores == nullptr && !(in_dtype == DT_F32) && m_tiles * (N / 256) >= 4 * 148;
    const int BN = wide_ok ? 256 : ((N % 128 == 0) ? 128 : 64);
    // CTA pairs (256 x 256 tiles, cta_group::2) when the pair tiles still spread over the 74 SM pairs
    static const bool no_pairs = getenv("AVS_GEMM_1CTA") != nullptr;
    static const int pair_min = getenv("AVS_GEMM_PAIR_MIN") ? atoi(getenv("AVS_GEMM_PAIR_MIN")) : 2 * 74;   // test hook
    const int64_t pair_tiles = ((M + 2 * BM - 1) / (2 * BM)) * (N / 256);
    if (!no_pairs && N % 256 == 0 && epi.scores == nullptr && pair_tiles >= (epi.prefer_pairs ? 1 : pair_min))
        return gemm_tc2(A, lda, W, ldw, in_dtype, M, N, K, epi, stream);

    CUtensorMap tmA, tmB, tmC;
    AVS_TRY(make_tmap(&tmA, A, in_dtype, M, K, lda, BM));
    AVS_TRY(make_tmap(&tmB, W, in_dtype, N, K, ldw, BN));
    if (epi.scores == nullptr) AVS_TRY(make_tmap(&tmC, epi.C, epi.out_dtype, M, N, epi.ldc, 32));
    else tmC = tmA;   // unused by the score epilogue
    const int tiles_n = (N + BN - 1) / BN;
    const int64_t tiles_total = static_cast<int64_t>((M + BM - 1) / BM) * tiles_n;
    AVS_CHECK(tiles_total < (1ll << 31), AVS_ERR_UNSUPPORTED, "gemm: too many tiles");
    const int num_tiles = static_cast<int>(tiles_total);
    const int num_sms = device_sm_count();
    AVS_CHECK(num_sms > 0, AVS_ERR_CUDA, "gemm: could not read the SM count");
    int sms_avail = num_sms;
    if (epi.max_ctas > 0) sms_avail = std::max(1, std::min(num_sms, epi.max_ctas));
    const int grid = num_tiles < sms_avail ? num_tiles : sms_avail;   // persistent: one CTA per SM
    const uint32_t idesc = umma_idesc(fmt, BM, BN);

#define AVS_GEMM_LAUNCH(BN_, ST_, TF_)                                                                       \
    do {                                                                                                     \
        using L = SmemLayout<BN_, ST_>;                                                                      \
        auto kern = gemm_tc_kernel<BN_, ST_, TF_>;                                                           \
        static PerDeviceOnce configured;                                                                     \
        const int dev_ = current_device();                                                                   \
        if (configured.needed(dev_)) {                                                                       \
            AVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));     \
            configured.mark(dev_);                                                                           \
        }                                                                                                    \
        kern<<<grid, GEMM_THREADS, L::TOTAL, stream>>>(tmA, tmB, tmC, static_cast<int>(M), N, k_blocks,      \
                                                       bk_elems, idesc, tiles_n, num_tiles, epi);            \
    } while (0)

    if (BN == 256) {
        AVS_GEMM_LAUNCH(256, 3, false);
    } else if (BN == 128) {
        if (tf32) AVS_GEMM_LAUNCH(128, 4, true);
        else AVS_GEMM_LAUNCH(128, 4, false);
    } else {
        if (tf32) AVS_GEMM_LAUNCH(64, 6, true);
        else AVS_GEMM_LAUNCH(64, 6, false);
    }
#undef AVS_GEMM_LAUNCH
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

}  // namespace avs
