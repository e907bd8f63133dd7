// K7 / K8 -- summary generation: shot pooling over change points, 0/1 knapsack at the length
// budget, keyshot bitmap; plus the batched overlap-F1 of
// /root/reference/evaluation/metrics.py:1-9 (== utils/shot_metrics.py:4-16).
//
// The reference has NO pooling / knapsack code; the algorithm is specified
// by oracle/av_oracle.py (shot_pool, knapsack, generate_summary) in integer arithmetic and these
// kernels are bit-exact against it:
//   q_i        = clamp(rint(score_i * 2^24), 0, 2^24)                (the only fp step)
//   seg_sum_s  = sum_i q_i * |[pos_i, pos_{i+1}) ^ [a_s, b_s]|        (int64, order independent)
//   seg_mean_s = floor((2 seg_sum_s + nf_s) / (2 nf_s)),  nf_s = b_s - a_s + 1
//   dp_s[w]    = max(dp_{s-1}[w], dp_{s-1}[w - nf_s] + seg_mean_s), item kept only on strict gain
//   back-trace from (S-1, capacity), capacity = floor(n_frames * num / den).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace avs {

namespace {

constexpr int SCORE_FRAC_BITS = 24;

__device__ __forceinline__ long long quantize_score(float s) {
    if (!(s == s)) return 0;  // NaN
    const float x = s * 16777216.0f;  // exact power-of-two scaling
    if (x <= 0.f) return 0;
    if (x >= 16777216.0f) return 1ll << SCORE_FRAC_BITS;
    return __float2ll_rn(x);  // round half to even == numpy.rint
}

// ---------------------------------------------------------------------------- K7
// One block per video; thread i owns sampled frame i (stride loop).  HBM-bound:
// 8 B read per sampled frame + 8 B per shot, 8 B written per shot.
// acc: per-shot accumulators of this video (shared or global memory), indexed by the shot number in the video
__device__ __forceinline__ void pool_video(const float* __restrict__ scores, const int32_t* __restrict__ positions,
                                           int row0, int T, int nf, const int2* __restrict__ cps, int S,
                                           unsigned long long* acc) {
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const long long q = quantize_score(scores[row0 + i]);
        const int lo = positions[row0 + i];
        const int hi = (i + 1 < T) ? positions[row0 + i + 1] : nf;
        if (hi <= lo || q == 0) continue;
        // first shot whose (inclusive) end >= lo; shots are sorted and disjoint (validated on host)
        int a = 0, z = S;
        while (a < z) {
            const int mid = (a + z) >> 1;
            if (cps[mid].y < lo) a = mid + 1; else z = mid;
        }
        for (int s = a; s < S; ++s) {
            const int2 seg = cps[s];
            if (seg.x >= hi) break;
            const int ov = min(hi, seg.y + 1) - max(lo, seg.x);
            if (ov > 0) atomicAdd(acc + s, static_cast<unsigned long long>(q * ov));
        }
    }
}

__global__ void __launch_bounds__(256) shot_pool_kernel(const float* __restrict__ scores,
                                                        const int32_t* __restrict__ positions, SummaryBatch b,
                                                        unsigned long long* __restrict__ seg_sum) {
    const int v = blockIdx.x;
    const int s0 = b.cps_start[v], S = b.cps_start[v + 1] - s0;
    pool_video(scores, positions, b.row_start[v], b.lengths[v], b.n_frames[v],
               reinterpret_cast<const int2*>(b.cps) + s0, S, seg_sum + s0);
}

// ---------------------------------------------------------------------------- K8
// One block per video.  Everything the item loop touches lives in shared memory when it fits: the two DP rows
// (2 * (cap + 1) int64), the items (weight, pooled value -- computed once, in parallel, instead of a global load
// and a 64-bit division per item on the critical path) and the per-item "kept" bits (one ballot word per warp).
// Oversized videos fall back to a global (L2-resident) workspace for the DP rows / keep bits.  A single thread
// walks the keep bits backwards.
constexpr int KNAP_THREADS = 1024;

__global__ void __launch_bounds__(KNAP_THREADS) knapsack_kernel(SummaryBatch b,
                                                                const unsigned long long* __restrict__ seg_sum,
                                                                long long* __restrict__ seg_mean_out,
                                                                uint8_t* __restrict__ picks,
                                                                uint8_t* __restrict__ summary,
                                                                uint32_t* __restrict__ keep_bits,
                                                                long long* __restrict__ dp_ws, int smem_rows_cap,
                                                                int smem_items_cap, int smem_keep_words,
                                                                const float* __restrict__ scores,
                                                                const int32_t* __restrict__ positions) {
    extern __shared__ long long dp_smem[];
    const int v = blockIdx.x;
    const int tid = threadIdx.x;
    const int nf = b.n_frames[v];
    const int s0 = b.cps_start[v], S = b.cps_start[v + 1] - s0;
    const int2* cps = reinterpret_cast<const int2*>(b.cps) + s0;
    const long long cap_ll = (static_cast<long long>(nf) * b.prop_num) / b.prop_den;
    const int cap = static_cast<int>(cap_ll < 0 ? 0 : cap_ll);
    const int words = (cap + 32) >> 5;  // ceil((cap + 1) / 32)

    long long* item_val = dp_smem + 2 * smem_rows_cap;                            // [smem_items_cap]
    int* item_wt = reinterpret_cast<int*>(item_val + smem_items_cap);             // [smem_items_cap]
    int* item_beg = item_wt + smem_items_cap;                                     // [smem_items_cap] first frame; < 0: not picked
    uint32_t* keep_sm = reinterpret_cast<uint32_t*>(item_beg + smem_items_cap);
    const bool items_in_smem = S <= smem_items_cap;
    const bool keep_in_smem = static_cast<long long>(S) * words <= smem_keep_words;
    uint32_t* keep = keep_in_smem ? keep_sm : keep_bits + b.keep_start[v];

    long long* cur;
    long long* nxt;
    if (cap + 1 <= smem_rows_cap) {
        cur = dp_smem;
        nxt = dp_smem + smem_rows_cap;
    } else {
        cur = dp_ws + b.dp_start[v];
        nxt = cur + (cap + 1);
    }
    for (int w = tid; w <= cap; w += KNAP_THREADS) cur[w] = 0;
    // fused K7 (scores != nullptr; the host only asks for it when the items fit in shared memory): pool this
    // video's frame scores into per-shot sums held in item_val, no global accumulators / memset / extra launch
    const bool fused_pool = scores != nullptr;
    if (fused_pool) {
        for (int s = tid; s < S; s += KNAP_THREADS) item_val[s] = 0;
        __syncthreads();
        pool_video(scores, positions, b.row_start[v], b.lengths[v], nf, cps, S,
                   reinterpret_cast<unsigned long long*>(item_val));
        __syncthreads();
    }
    // pooled shot values: mean rounded half up in fixed point (oracle: shot_pool)
    for (int s = tid; s < S; s += KNAP_THREADS) {
        const int2 seg = cps[s];
        const int wt = seg.y - seg.x + 1;
        const unsigned long long sum = fused_pool ? static_cast<unsigned long long>(item_val[s]) : seg_sum[s0 + s];
        const long long val = wt > 0 ? static_cast<long long>((2ull * sum + wt) / (2ull * wt)) : 0ll;
        if (seg_mean_out != nullptr) seg_mean_out[s0 + s] = val;
        if (items_in_smem) {
            item_val[s] = val;
            item_wt[s] = wt;
            item_beg[s] = seg.x;
        }
    }
    __syncthreads();

    for (int s = 0; s < S; ++s) {
        int wt;
        long long val;
        if (items_in_smem) {
            wt = item_wt[s];
            val = item_val[s];
        } else {
            const int2 seg = cps[s];
            wt = seg.y - seg.x + 1;
            const unsigned long long sum = seg_sum[s0 + s];
            val = wt > 0 ? static_cast<long long>((2ull * sum + wt) / (2ull * wt)) : 0ll;
        }
        for (int wbase = 0; wbase <= cap; wbase += KNAP_THREADS) {  // warp-uniform trip count
            const int w = wbase + tid;
            bool better = false;
            if (w <= cap) {
                const long long old = cur[w];
                long long cand = old;
                if (wt > 0 && w >= wt) {
                    cand = cur[w - wt] + val;
                    better = cand > old;
                } else if (wt == 0 && val > 0) {
                    cand = old + val;
                    better = true;
                }
                nxt[w] = better ? cand : old;
            }
            const uint32_t bits = __ballot_sync(0xffffffffu, better);
            if ((tid & 31) == 0 && (w >> 5) < words) keep[static_cast<size_t>(s) * words + (w >> 5)] = bits;
        }
        __syncthreads();
        long long* t = cur; cur = nxt; nxt = t;
    }
    __threadfence_block();
    __syncthreads();
    if (tid == 0) {
        int w = cap;
        for (int s = S - 1; s >= 0; --s) {
            const uint32_t word = keep[static_cast<size_t>(s) * words + (w >> 5)];
            const int take = (word >> (w & 31)) & 1;
            picks[s0 + s] = static_cast<uint8_t>(take);
            if (take) w -= items_in_smem ? item_wt[s] : (cps[s].y - cps[s].x + 1);
            if (items_in_smem && !take) item_wt[s] = -1;      // the bitmap pass below reads the selection from smem
        }
    }
    if (summary != nullptr) {
        __syncthreads();
        uint8_t* out = summary + b.summary_start[v];
        // zero fill with 16-byte stores where the alignment allows (summary_start is arbitrary), bytes at the edges
        {
            const uintptr_t addr = reinterpret_cast<uintptr_t>(out);
            const int head = min(nf, static_cast<int>((16 - (addr & 15)) & 15));
            const int body = (nf - head) / 16;
            for (int f = tid; f < head; f += KNAP_THREADS) out[f] = 0;
            uint4* o4 = reinterpret_cast<uint4*>(out + head);
            for (int i = tid; i < body; i += KNAP_THREADS) o4[i] = make_uint4(0, 0, 0, 0);
            for (int f = head + body * 16 + tid; f < nf; f += KNAP_THREADS) out[f] = 0;
        }
        __syncthreads();
        for (int s = 0; s < S; ++s) {
            int a, z;
            if (items_in_smem) {
                if (item_wt[s] < 0) continue;
                a = max(item_beg[s], 0);
                z = min(item_beg[s] + item_wt[s], nf);
            } else {
                if (!picks[s0 + s]) continue;
                a = max(cps[s].x, 0);
                z = min(cps[s].y + 1, nf);
            }
            for (int f = a + tid; f < z; f += KNAP_THREADS) out[f] = 1;
        }
    }
}

// ---------------------------------------------------------------------------- K7 + K8, fast path
// The common case (TVSum / SumMe-sized videos: capacity < 4096 frames, everything fits in shared memory):
//   * the shots are loaded into shared memory once; pooling binary-searches them there,
//   * every thread keeps its DP cells in REGISTERS (cell w = tid + c * 1024); per item it publishes them to one of
//     two shared buffers and reads dp[w - wt] from the other -- one load + one store per cell and item instead of
//     two loads + one store, in int32 when the total value fits (S <= 127 shots of <= 2^24 each),
//   * the keyshot bitmap is produced 16 bytes per thread with one binary search per block.
// Bit-exact against the same oracle as the general kernel above (which remains for oversized videos).
constexpr int KNAP_CPT = 4;   // DP cells per thread (fast path)
// Long videos (BASELINE configs[3]: T = 8192 -> 122,880 frames, capacity 18,432, ~750 shots): the same scheme with
// 512 threads x up to 40 register-resident cells.  One DP row no longer fits twice in shared memory (2 x 147 KB), so
// the row is published into a SINGLE buffer (two barriers per item instead of one), the keep bits (S x words x 4 B =
// 1.7 MB) go to the global workspace, and the back-trace stages them back through the (then free) row buffer in
// chunks of ~60 items, so that the single-thread walk reads shared memory instead of 750 dependent L2 round trips.
// Measured on B200, 8 videos x T = 8192: 5.46 ms (general kernel: DP rows in L2) -> see DESIGN.md.
constexpr int KNAP_BIG_THREADS = 512;
constexpr int KNAP_BIG_CPT = 40;

template <typename V, int CPT, int NT, bool BIG>
__global__ void __launch_bounds__(NT) knapsack_fast_kernel(SummaryBatch b, const float* __restrict__ scores,
                                                           const int32_t* __restrict__ positions,
                                                           long long* __restrict__ seg_mean_out,
                                                           uint8_t* __restrict__ picks,
                                                           uint8_t* __restrict__ summary, int rows_cap,
                                                           int items_cap, uint32_t* __restrict__ keep_bits) {
    extern __shared__ long long fsm[];
    const int v = blockIdx.x;
    const int tid = threadIdx.x;
    const int nf = b.n_frames[v];
    const int s0 = b.cps_start[v], S = b.cps_start[v + 1] - s0;
    const int2* cps = reinterpret_cast<const int2*>(b.cps) + s0;
    const long long cap_ll = (static_cast<long long>(nf) * b.prop_num) / b.prop_den;
    const int cap = static_cast<int>(cap_ll < 0 ? 0 : cap_ll);
    const int words = (cap + 32) >> 5;

    long long* item_val = fsm;                                           // [items_cap]  sums, then pooled values
    V* buf_a = reinterpret_cast<V*>(item_val + items_cap);               // [rows_cap]
    V* buf_b = BIG ? buf_a : buf_a + rows_cap + (rows_cap & 1);          // [rows_cap]  (kept 8-byte aligned)
    int* item_beg = reinterpret_cast<int*>(buf_b + rows_cap + (rows_cap & 1));   // [items_cap] first frame of the shot
    int* item_len = item_beg + items_cap;                                // [items_cap] frames; negated once not picked
    uint32_t* keep = BIG ? keep_bits + b.keep_start[v]                   // [S][kwords]: global workspace (BIG) ...
                         : reinterpret_cast<uint32_t*>(item_len + items_cap);   // ... or shared memory [S][words]
    // BIG: every thread owns cpt_pad cells (a multiple of 4); row buffer and keep rows are padded accordingly, so the
    // item loop needs no per-cell bounds checks (cells beyond the capacity hold values nobody reads)
    const int cpt_pad = BIG ? (((cap + NT) / NT + 3) & ~3) : CPT;
    const int kwords = BIG ? cpt_pad * (NT / 32) : words;

    for (int s = tid; s < S; s += NT) {
        const int2 seg = cps[s];
        item_beg[s] = seg.x;
        item_len[s] = seg.y - seg.x + 1;
        item_val[s] = 0;
    }
    __syncthreads();
    // ---- K7: pool the frame scores of this video into per-shot sums (int64 shared-memory atomics)
    {
        const int row0 = b.row_start[v], T = b.lengths[v];
        unsigned long long* acc = reinterpret_cast<unsigned long long*>(item_val);
        for (int i = tid; i < T; i += NT) {
            const long long q = quantize_score(scores[row0 + i]);
            const int lo = positions[row0 + i];
            const int hi = (i + 1 < T) ? positions[row0 + i + 1] : nf;
            if (hi <= lo || q == 0) continue;
            int a = 0, z = S;   // first shot whose inclusive end >= lo
            while (a < z) {
                const int mid = (a + z) >> 1;
                if (item_beg[mid] + item_len[mid] - 1 < lo) a = mid + 1; else z = mid;
            }
            for (int s = a; s < S; ++s) {
                const int sb = item_beg[s];
                if (sb >= hi) break;
                const int ov = min(hi, sb + item_len[s]) - max(lo, sb);
                if (ov > 0) atomicAdd(acc + s, static_cast<unsigned long long>(q * ov));
            }
        }
    }
    __syncthreads();
    for (int s = tid; s < S; s += NT) {
        const int wt = item_len[s];
        const unsigned long long sum = static_cast<unsigned long long>(item_val[s]);
        const long long val = wt > 0 ? static_cast<long long>((2ull * sum + wt) / (2ull * wt)) : 0ll;
        if (seg_mean_out != nullptr) seg_mean_out[s0 + s] = val;
        item_val[s] = val;
    }
    // ---- K8: DP with register-resident cells
    V mine[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        mine[c] = 0;
        const int w = tid + c * NT;
        if (BIG ? (c < cpt_pad) : (w <= cap)) buf_a[w] = 0;
    }
    __syncthreads();
    if constexpr (BIG) {
        // ---- long videos: branch-free cells.  Per cell: clamp the source index, one 64-bit shared load, add, 64-bit
        // compare, two selects, a ballot and (lane 0) one keep word -- ~13 instructions; the first version, with bounds
        // checks and divergent branches per cell, executed 43 and was instruction bound (ncu: 1,566 warp
        // instructions per item and warp).
        uint32_t* keep_w = keep + (tid >> 5);          // this warp's word of cell block c is keep_w[c * (NT / 32)]
        const bool lane0 = (tid & 31) == 0;
        for (int s = 0; s < S; ++s) {
            const int wt = item_len[s];
            const V val = static_cast<V>(item_val[s]);
            uint32_t* krow = keep_w + static_cast<size_t>(s) * kwords;
            const int base = tid - wt;
            const bool bump = wt == 0 && val > 0;      // defensive: validated shots have wt >= 1
#pragma unroll
            for (int c0 = 0; c0 < CPT; c0 += 4) {
                if (c0 < cpt_pad) {   // uniform
                    V prev[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) prev[k] = buf_a[max(base + (c0 + k) * NT, 0)];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int c = c0 + k;
                        const V cand = (bump ? mine[c] : prev[k]) + val;
                        const bool better = bump || (wt > 0 && base + c * NT >= 0 && cand > mine[c]);
                        mine[c] = better ? cand : mine[c];
                        const uint32_t bits = __ballot_sync(0xffffffffu, better);
                        if (lane0) krow[c * (NT / 32)] = bits;
                    }
                }
            }
            __syncthreads();   // everybody has read row s - 1
#pragma unroll
            for (int c0 = 0; c0 < CPT; c0 += 4) {
                if (c0 < cpt_pad) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) buf_a[tid + (c0 + k) * NT] = mine[c0 + k];
                }
            }
            __syncthreads();   // row s published
        }
    }
    for (int s = 0; s < (BIG ? 0 : S); ++s) {
        const int wt = item_len[s];
        const V val = static_cast<V>(item_val[s]);
        const V* rd = (s & 1) ? buf_b : buf_a;
        V* wr = (s & 1) ? buf_a : buf_b;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int w = tid + c * NT;
            if (c * NT <= cap) {   // warp-uniform
                bool better = false;
                if (w <= cap) {
                    const V old = mine[c];
                    V cand = old;
                    if (wt > 0 && w >= wt) {
                        cand = rd[w - wt] + val;
                        better = cand > old;
                    } else if (wt == 0 && val > 0) {
                        cand = old + val;
                        better = true;
                    }
                    mine[c] = better ? cand : old;
                    wr[w] = mine[c];
                }
                const uint32_t bits = __ballot_sync(0xffffffffu, better);
                if ((tid & 31) == 0 && (w >> 5) < words) keep[s * words + (w >> 5)] = bits;
            }
        }
        __syncthreads();
    }
    if (!BIG) {
        if (tid == 0) {
            int w = cap;
            for (int s = S - 1; s >= 0; --s) {
                const int take = (keep[s * words + (w >> 5)] >> (w & 31)) & 1;
                picks[s0 + s] = static_cast<uint8_t>(take);
                if (take) w -= item_len[s];
                else item_len[s] = -item_len[s];          // the bitmap pass reads the selection from the sign
            }
        }
    } else {
        // back-trace in chunks of items whose keep rows fit in the row buffer (the DP values are no longer needed)
        __threadfence_block();
        uint32_t* stage = reinterpret_cast<uint32_t*>(buf_a);
        const int chunk = max(1, static_cast<int>((static_cast<size_t>(rows_cap) * sizeof(V)) / (static_cast<size_t>(kwords) * 4)));
        __shared__ int w_cursor;
        if (tid == 0) w_cursor = cap;
        for (int hi = S; hi > 0; hi -= chunk) {
            const int lo = max(0, hi - chunk);
            __syncthreads();   // stage free (and, first time, every keep word written)
            const size_t n_words = static_cast<size_t>(hi - lo) * kwords;
            const uint32_t* src = keep + static_cast<size_t>(lo) * kwords;
#pragma unroll 8
            for (size_t i = tid; i < n_words; i += NT) stage[i] = src[i];
            __syncthreads();
            if (tid == 0) {
                int w = w_cursor;
                for (int s = hi - 1; s >= lo; --s) {
                    const int take = (stage[static_cast<size_t>(s - lo) * kwords + (w >> 5)] >> (w & 31)) & 1;
                    picks[s0 + s] = static_cast<uint8_t>(take);
                    if (take) w -= item_len[s];
                    else item_len[s] = -item_len[s];
                }
                w_cursor = w;
            }
        }
    }
    if (summary == nullptr) return;
    __syncthreads();
    // ---- keyshot bitmap: frame f is set iff it lies inside a picked shot
    uint8_t* out = summary + b.summary_start[v];
    auto shot_of = [&](int f) {   // last shot with beg <= f, or -1
        int a = 0, z = S;
        while (a < z) {
            const int mid = (a + z) >> 1;
            if (item_beg[mid] <= f) a = mid + 1; else z = mid;
        }
        return a - 1;
    };
    auto bit_at = [&](int f, int& s) -> uint32_t {   // s: cursor, advanced monotonically
        while (s + 1 < S && item_beg[s + 1] <= f) ++s;
        if (s < 0) return 0u;
        const int len = item_len[s];
        return (len > 0 && f < item_beg[s] + len) ? 1u : 0u;
    };
    const uintptr_t addr = reinterpret_cast<uintptr_t>(out);
    const int head = min(nf, static_cast<int>((16 - (addr & 15)) & 15));
    const int body = (nf - head) / 16;
    for (int f = tid; f < head; f += NT) {
        int s = shot_of(f);
        out[f] = static_cast<uint8_t>(bit_at(f, s));
    }
    uint4* o4 = reinterpret_cast<uint4*>(out + head);
    for (int i = tid; i < body; i += NT) {
        const int f0 = head + 16 * i;
        int s = shot_of(f0);
        uint32_t wv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) x |= bit_at(f0 + 4 * k + j, s) << (8 * j);
            wv[k] = x;
        }
        o4[i] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
    for (int f = head + body * 16 + tid; f < nf; f += NT) {
        int s = shot_of(f);
        out[f] = static_cast<uint8_t>(bit_at(f, s));
    }
}

// ---------------------------------------------------------------------------- K7 + K8, long videos on a cluster
// The long-video kernel above runs one video on ONE SM and is bound by instruction issue there (36 cells per thread
// and item); a configs[3] batch has 8 videos for 148 SMs.  Here a cluster of KC CTAs shares a video: CTA k owns the
// capacity cells [k * slice, (k + 1) * slice) (register-resident, published into a double-buffered shared row), and
// cell w reads dp[w - wt] from the row of whichever CTA owns it -- its own (plain shared loads; warp-uniform test)
// or the left neighbour's.  No cluster-wide barrier in the item loop (barrier.cluster -- or an mbarrier equivalent --
// cost ~1.3 us per item, 745 items per video): after its item every CTA PUSHES the last `hmax` cells of its new row
// (hmax >= the longest shot) into the right neighbour's halo buffer with 16-byte st.async whose byte count completes
// the neighbour's "halo landed" mbarrier -- data and synchronisation in one DSMEM hop, as in the recurrence's h
// exchange -- and the right neighbour returns a credit (remote mbarrier arrive) once it has finished reading a halo
// buffer, before which the left one must not overwrite it.  Rows and halos are double-buffered by item parity.
// Every CTA pools the shots itself (cheap, and it avoids a broadcast); CTA 0 walks the keep bits back and writes
// picks / bitmap.
constexpr int KC_THREADS = 512;
constexpr int KC_CELLS = 24576;       // capacity cells a cluster covers: KC CTAs x 512 threads x (48 / KC) cells
// (8 videos x T = 8192 on B200: 4 CTAs per video 1.20 ms, 8 CTAs per video 0.99 ms)

__device__ __forceinline__ long long ld_dsmem_s64(uint32_t cluster_addr) {
    long long v;
    asm volatile("ld.shared::cluster.b64 %0, [%1];" : "=l"(v) : "r"(cluster_addr) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

template <int KC>
__global__ void __cluster_dims__(KC, 1, 1) __launch_bounds__(KC_THREADS)
knapsack_cluster_kernel(SummaryBatch b, const float* __restrict__ scores, const int32_t* __restrict__ positions,
                        long long* __restrict__ seg_mean_out, uint8_t* __restrict__ picks, uint8_t* __restrict__ summary,
                        int slice, int items_cap, uint32_t* __restrict__ keep_bits, int hmax) {
    extern __shared__ long long csm[];
    constexpr int NT = KC_THREADS;
    constexpr int KC_CPT = KC_CELLS / (KC * KC_THREADS);   // cells per thread: 12 (KC = 4) or 6 (KC = 8)
    const int v = blockIdx.x / KC;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tid = threadIdx.x;
    const int nf = b.n_frames[v];
    const int s0 = b.cps_start[v], S = b.cps_start[v + 1] - s0;
    const int2* cps = reinterpret_cast<const int2*>(b.cps) + s0;
    const long long cap_ll = (static_cast<long long>(nf) * b.prop_num) / b.prop_den;
    const int cap = static_cast<int>(cap_ll < 0 ? 0 : cap_ll);
    // keep rows as the planner lays them out for long videos (knapsack_keep_words): padded to whole 2,048-cell groups
    const int kwords = (cap + 4 * KNAP_BIG_THREADS) / (4 * KNAP_BIG_THREADS) * (4 * KNAP_BIG_THREADS) / 32;
    const int cpt = slice / NT;                       // cells per thread actually used (slice is a multiple of NT)
    const int lo = static_cast<int>(rank) * slice;    // first cell of this CTA

    long long* item_val = csm;                                        // [items_cap]
    long long* row0 = item_val + items_cap;                           // [2][slice] double-buffered row slice
    long long* halo0 = row0 + 2 * static_cast<size_t>(slice);         // [2][hmax] the left neighbour's last hmax cells
    int* item_beg = reinterpret_cast<int*>(halo0 + 2 * static_cast<size_t>(hmax));
    int* item_len = item_beg + items_cap;
    // keep bits of the last KEEP_BATCH items, flushed to the global workspace in one coalesced burst: a global store
    // per item would be outstanding at every cluster barrier, whose release semantics then wait for it (measured:
    // 2.3 us per item with the stores in the loop)
    constexpr int KEEP_BATCH = 16;
    const int wpc = slice / 32;                                       // keep words of this CTA's slice per item
    uint32_t* keep_sm = reinterpret_cast<uint32_t*>(item_len + items_cap);   // [KEEP_BATCH][wpc]
    uint32_t* keep = keep_bits + b.keep_start[v];

    for (int s = tid; s < S; s += NT) {
        const int2 seg = cps[s];
        item_beg[s] = seg.x;
        item_len[s] = seg.y - seg.x + 1;
        item_val[s] = 0;
    }
    __syncthreads();
    {   // K7 (every CTA of the cluster computes the same per-shot sums)
        const int row0v = b.row_start[v], T = b.lengths[v];
        unsigned long long* acc = reinterpret_cast<unsigned long long*>(item_val);
        for (int i = tid; i < T; i += NT) {
            const long long q = quantize_score(scores[row0v + i]);
            const int plo = positions[row0v + i];
            const int phi = (i + 1 < T) ? positions[row0v + i + 1] : nf;
            if (phi <= plo || q == 0) continue;
            int a = 0, z = S;
            while (a < z) {
                const int mid = (a + z) >> 1;
                if (item_beg[mid] + item_len[mid] - 1 < plo) a = mid + 1; else z = mid;
            }
            for (int s = a; s < S; ++s) {
                const int sb = item_beg[s];
                if (sb >= phi) break;
                const int ov = min(phi, sb + item_len[s]) - max(plo, sb);
                if (ov > 0) atomicAdd(acc + s, static_cast<unsigned long long>(q * ov));
            }
        }
    }
    __syncthreads();
    for (int s = tid; s < S; s += NT) {
        const int wt = item_len[s];
        const unsigned long long sum = static_cast<unsigned long long>(item_val[s]);
        const long long val = wt > 0 ? static_cast<long long>((2ull * sum + wt) / (2ull * wt)) : 0ll;
        if (rank == 0 && seg_mean_out != nullptr) seg_mean_out[s0 + s] = val;
        item_val[s] = val;
    }
    long long mine[KC_CPT];
#pragma unroll
    for (int c = 0; c < KC_CPT; ++c) {
        mine[c] = 0;
        if (c < cpt) {
            row0[tid + c * NT] = 0;
            row0[slice + tid + c * NT] = 0;
        }
    }
    for (int i = tid; i < 2 * hmax; i += NT) halo0[i] = 0;   // item 0 reads the left neighbour's initial row: zeros
    __shared__ __align__(8) uint64_t data_bar[2];     // halo[p] landed (byte-counted st.async from the left neighbour)
    __shared__ __align__(8) uint64_t credit_bar[2];   // the right neighbour has finished reading its halo[p]
    if (tid == 0) {
        mbar_init(&data_bar[0], 1);
        mbar_init(&data_bar[1], 1);
        mbar_init(&credit_bar[0], 1);
        mbar_init(&credit_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    cluster_barrier();     // every CTA's rows, halos and barriers exist before any remote access
    const bool has_left = rank > 0, has_right = rank + 1 < KC;
    const uint32_t right_halo = has_right ? mapa(smem_u32(halo0), rank + 1) : 0;
    const uint32_t right_data_bar = has_right ? mapa(smem_u32(&data_bar[0]), rank + 1) : 0;
    const uint32_t left_credit_bar = has_left ? mapa(smem_u32(&credit_bar[0]), rank - 1) : 0;
    const bool lane0 = (tid & 31) == 0;
    const int my_words = max(0, min(wpc, kwords - (lo >> 5)));     // words of this slice that exist in the keep rows
    for (int s = 0; s < S; ++s) {
        const int wt = item_len[s];
        const long long val = item_val[s];
        const long long* rd = row0 + static_cast<size_t>(s & 1) * slice;
        long long* wr = row0 + static_cast<size_t>((s & 1) ^ 1) * slice;
        const long long* halo = halo0 + static_cast<size_t>(s & 1) * hmax;   // left neighbour's row after item s - 1
        uint32_t* krow = keep_sm + (s % KEEP_BATCH) * wpc + (tid >> 5);
        const bool bump = wt == 0 && val > 0;
        if (has_left) {
            // arm the barrier that collects the halo for item s + 1, then wait for this item's halo
            if (tid == 0 && s + 1 < S) mbar_expect_tx(&data_bar[(s + 1) & 1], static_cast<uint32_t>(hmax) * 8);
            if (s > 0) mbar_wait(&data_bar[s & 1], ((s - 1) >> 1) & 1);
        }
#pragma unroll
        for (int c = 0; c < KC_CPT; ++c) {
            if (c < cpt) {   // uniform
                const int wl = tid + c * NT;            // local cell
                const int src = lo + wl - wt;           // global source cell
                long long prev = 0;
                if (src >= lo) prev = rd[src - lo];                  // own slice (all lanes, except in the first wt cells)
                else if (src >= 0) prev = halo[hmax + (src - lo)];   // the left neighbour's tail (wt <= hmax)
                // (loading all cells' sources first, then comparing, measured SLOWER here: 1.74 vs 1.20 ms)
                const long long cand = (bump ? mine[c] : prev) + val;
                const bool better = bump || (wt > 0 && src >= 0 && cand > mine[c]);
                mine[c] = better ? cand : mine[c];
                wr[wl] = mine[c];
                const uint32_t bits = __ballot_sync(0xffffffffu, better);
                if (lane0) krow[c * (NT / 32)] = bits;
            }
        }
        __syncthreads();   // the new row is complete in shared memory; this CTA's reads of halo[s & 1] are over
        if (s + 1 < S) {
            if (has_left && tid == 0) mbar_arrive_remote(left_credit_bar + (s & 1) * 8);   // credit: halo[s & 1] may be overwritten
            if (has_right && tid < hmax / 2) {
                // the right neighbour read its halo[(s + 1) & 1] during item s - 1: wait for that credit, then push
                if (s > 0) mbar_wait_cluster(&credit_bar[(s + 1) & 1], ((s - 1) >> 1) & 1);
                const uint4 v4 = *reinterpret_cast<const uint4*>(wr + slice - hmax + 2 * tid);
                st_async_v4(right_halo + (static_cast<uint32_t>((s + 1) & 1) * hmax + 2 * tid) * 8, v4,
                            right_data_bar + ((s + 1) & 1) * 8);
            }
        }
        if ((s % KEEP_BATCH) == KEEP_BATCH - 1 || s == S - 1) {
            // flush the batch (a video shorter than the batch's longest has narrower keep rows than KC slices: only
            // my_words of them exist; cells past them are never read)
            const int first = s - (s % KEEP_BATCH), n_it = s - first + 1;
            for (int i = tid; i < n_it * my_words; i += NT) {
                const int it = i / my_words, wd = i - it * my_words;
                keep[static_cast<size_t>(first + it) * kwords + (lo >> 5) + wd] = keep_sm[it * wpc + wd];
            }
            __syncthreads();   // keep_sm is rewritten by the next item
        }
    }
    __threadfence();
    cluster_barrier();       // every CTA's last keep-bit flush is visible; last access to the peers' shared memory is over
    if (rank != 0) return;
    // ---- back-trace on CTA 0, keep rows staged through the (now free) row buffers in chunks
    __threadfence();
    uint32_t* stage = reinterpret_cast<uint32_t*>(row0);
    const int chunk = max(1, static_cast<int>((2 * static_cast<size_t>(slice) * 8) / (static_cast<size_t>(kwords) * 4)));
    __shared__ int w_cursor;
    if (tid == 0) w_cursor = cap;
    for (int hi = S; hi > 0; hi -= chunk) {
        const int lo_s = max(0, hi - chunk);
        __syncthreads();
        const size_t n_words = static_cast<size_t>(hi - lo_s) * kwords;
        const uint32_t* src = keep + static_cast<size_t>(lo_s) * kwords;
#pragma unroll 8
        for (size_t i = tid; i < n_words; i += NT) stage[i] = __ldcg(src + i);   // written by other SMs: read through L2
        __syncthreads();
        if (tid == 0) {
            int w = w_cursor;
            for (int s = hi - 1; s >= lo_s; --s) {
                const int take = (stage[static_cast<size_t>(s - lo_s) * kwords + (w >> 5)] >> (w & 31)) & 1;
                picks[s0 + s] = static_cast<uint8_t>(take);
                if (take) w -= item_len[s];
                else item_len[s] = -item_len[s];
            }
            w_cursor = w;
        }
    }
    if (summary == nullptr) return;
    __syncthreads();
    uint8_t* out = summary + b.summary_start[v];
    auto shot_of = [&](int f) {
        int a = 0, z = S;
        while (a < z) {
            const int mid = (a + z) >> 1;
            if (item_beg[mid] <= f) a = mid + 1; else z = mid;
        }
        return a - 1;
    };
    auto bit_at = [&](int f, int& s) -> uint32_t {
        while (s + 1 < S && item_beg[s + 1] <= f) ++s;
        if (s < 0) return 0u;
        const int len = item_len[s];
        return (len > 0 && f < item_beg[s] + len) ? 1u : 0u;
    };
    const uintptr_t addr = reinterpret_cast<uintptr_t>(out);
    const int head = min(nf, static_cast<int>((16 - (addr & 15)) & 15));
    const int body = (nf - head) / 16;
    for (int f = tid; f < head; f += NT) {
        int s = shot_of(f);
        out[f] = static_cast<uint8_t>(bit_at(f, s));
    }
    uint4* o4 = reinterpret_cast<uint4*>(out + head);
    for (int i = tid; i < body; i += NT) {
        const int f0 = head + 16 * i;
        int s = shot_of(f0);
        uint32_t wv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) x |= bit_at(f0 + 4 * k + j, s) << (8 * j);
            wv[k] = x;
        }
        o4[i] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
    for (int f = head + body * 16 + tid; f < nf; f += NT) {
        int s = shot_of(f);
        out[f] = static_cast<uint8_t>(bit_at(f, s));
    }
}

// ------------------------------------------------------------------- overlap F1
__global__ void __launch_bounds__(256) temporal_f1_kernel(const int2* __restrict__ pred,
                                                          const int32_t* __restrict__ pred_start,
                                                          const int2* __restrict__ gt,
                                                          const int32_t* __restrict__ gt_start,
                                                          double* __restrict__ f1) {
    const int v = blockIdx.x;
    const int p0 = pred_start[v], P = pred_start[v + 1] - p0;
    const int g0 = gt_start[v], G = gt_start[v + 1] - g0;
    long long ov = 0, plen = 0, glen = 0;
    for (long long idx = threadIdx.x; idx < static_cast<long long>(P) * G; idx += blockDim.x) {
        const int2 p = pred[p0 + idx / G], g = gt[g0 + idx % G];
        const int o = min(p.y, g.y) - max(p.x, g.x);
        if (o > 0) ov += o;
    }
    for (int i = threadIdx.x; i < P; i += blockDim.x) plen += pred[p0 + i].y - pred[p0 + i].x;
    for (int i = threadIdx.x; i < G; i += blockDim.x) glen += gt[g0 + i].y - gt[g0 + i].x;
    __shared__ long long red[3][256];
    red[0][threadIdx.x] = ov; red[1][threadIdx.x] = plen; red[2][threadIdx.x] = glen;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st)
            for (int k = 0; k < 3; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // evaluation/metrics.py:7-9 in IEEE double, no contraction
        const double precision = __ddiv_rn(static_cast<double>(red[0][0]), static_cast<double>(red[1][0]));
        const double recall = __ddiv_rn(static_cast<double>(red[0][0]), static_cast<double>(red[2][0]));
        const double num = __dmul_rn(2.0, __dmul_rn(precision, recall));
        const double den = __dadd_rn(__dadd_rn(precision, recall), 1e-8);
        f1[v] = __ddiv_rn(num, den);
    }
}

}  // namespace

avs_status shot_pool(const float* scores, const int32_t* positions, const SummaryBatch& b, unsigned long long* seg_sum,
                     cudaStream_t stream) {
    if (b.n == 0) return AVS_OK;
    shot_pool_kernel<<<b.n, 256, 0, stream>>>(scores, positions, b, seg_sum);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

// keep-bit words per item that the planner must reserve for a video of this capacity: the long-video kernel pads
// its keep rows to whole groups of 4 cells per thread
int64_t knapsack_keep_words(long long cap) {
    const long long plain = (cap + 32) >> 5;
    const long long padded = (cap + 4 * KNAP_BIG_THREADS) / (4 * KNAP_BIG_THREADS) * (4 * KNAP_BIG_THREADS) / 32;
    return std::max(plain, padded);
}

bool knapsack_can_fuse_pool(const SummaryBatch& b) {
    // the fused pooling needs every video's items in shared memory (see the budget below)
    const size_t limit = 200 * 1024;
    size_t smem = 2ull * (static_cast<size_t>(b.max_cap) + 1) * sizeof(long long);
    if (smem > limit) smem = 0;
    return smem + static_cast<size_t>(b.max_S) * 16 + 8 <= limit;
}

avs_status knapsack_select(const SummaryBatch& b, const unsigned long long* seg_sum, long long* seg_mean,
                           uint8_t* picks, uint8_t* summary, uint32_t* keep_bits, long long* dp_ws,
                           cudaStream_t stream, const float* scores, const int32_t* positions) {
    if (b.n == 0) return AVS_OK;
    const size_t limit = 200 * 1024;
    // ---- fast path: fused pooling, register-resident DP cells, everything in shared memory
    if (scores != nullptr && b.max_cap + 1 <= KNAP_CPT * KNAP_THREADS) {
        const bool narrow = b.max_S <= 127;                       // total value < 2^31: int32 DP
        const int rows = b.max_cap + 1, items = std::max(b.max_S, 1);
        const size_t vsz = narrow ? 4 : 8;
        const size_t words_all = static_cast<size_t>(b.max_S) * ((static_cast<size_t>(b.max_cap) + 32) >> 5);
        const size_t need = static_cast<size_t>(items) * 8 + 2 * (static_cast<size_t>(rows) + 1) * vsz +
                            static_cast<size_t>(items) * 8 + words_all * 4 + 64;
        if (need <= limit) {
            static PerDeviceOnce cfg32, cfg64;
            const int dev = current_device();
            if (narrow) {
                if (cfg32.needed(dev) && need > 48 * 1024) {
                    AVS_CUDA(cudaFuncSetAttribute(knapsack_fast_kernel<int, KNAP_CPT, KNAP_THREADS, false>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(limit)));
                    cfg32.mark(dev);
                }
                knapsack_fast_kernel<int, KNAP_CPT, KNAP_THREADS, false><<<b.n, KNAP_THREADS, need, stream>>>(
                    b, scores, positions, seg_mean, picks, summary, rows, items, nullptr);
            } else {
                if (cfg64.needed(dev) && need > 48 * 1024) {
                    AVS_CUDA(cudaFuncSetAttribute(knapsack_fast_kernel<long long, KNAP_CPT, KNAP_THREADS, false>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(limit)));
                    cfg64.mark(dev);
                }
                knapsack_fast_kernel<long long, KNAP_CPT, KNAP_THREADS, false><<<b.n, KNAP_THREADS, need, stream>>>(
                    b, scores, positions, seg_mean, picks, summary, rows, items, nullptr);
            }
            AVS_LAUNCH_CHECK();
            return AVS_OK;
        }
    }
    // ---- long videos, few of them: a cluster of 8 (or 4) CTAs per video (capacity range split, halo pushed to the
    // right neighbour)
    static const bool no_cluster = getenv("AVS_KNAPSACK_NO_CLUSTER") != nullptr;
    const int kc = (b.n * 8 <= device_sm_count()) ? 8 : ((b.n * 4 <= device_sm_count()) ? 4 : 0);
    if (scores != nullptr && keep_bits != nullptr && !no_cluster && kc > 0 && b.max_cap + 1 <= KC_CELLS) {
        const int slice = ((b.max_cap + kc) / kc + KC_THREADS - 1) / KC_THREADS * KC_THREADS;   // ceil((cap+1)/kc) -> x512
        const int items = (std::max(b.max_S, 1) + 1) / 2 * 2;   // even: the rows behind the item values stay 16-byte aligned
        const int hmax = std::max(2, (b.max_wt + 1) / 2 * 2);   // halo cells: >= the longest shot, even (16-byte messages)
        const size_t need = static_cast<size_t>(items) * 8 + 2 * static_cast<size_t>(slice) * 8 + 2 * static_cast<size_t>(hmax) * 8 +
                            static_cast<size_t>(items) * 8 + 16 * static_cast<size_t>(slice / 32) * 4 + 64;
        // the keep rows are written with the long-video stride (knapsack_keep_words): all slices must fit in it
        const int kwords = (b.max_cap + 4 * KNAP_BIG_THREADS) / (4 * KNAP_BIG_THREADS) * (4 * KNAP_BIG_THREADS) / 32;
        if (need <= limit && kc * slice <= kwords * 32 && hmax <= slice && hmax / 2 <= KC_THREADS) {
            static PerDeviceOnce cfg;
            const int dev = current_device();
            if (cfg.needed(dev)) {
                AVS_CUDA(cudaFuncSetAttribute(knapsack_cluster_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(limit)));
                AVS_CUDA(cudaFuncSetAttribute(knapsack_cluster_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(limit)));
                cfg.mark(dev);
            }
            if (kc == 8)
                knapsack_cluster_kernel<8><<<b.n * 8, KC_THREADS, need, stream>>>(b, scores, positions, seg_mean, picks,
                                                                                  summary, slice, items, keep_bits, hmax);
            else
                knapsack_cluster_kernel<4><<<b.n * 4, KC_THREADS, need, stream>>>(b, scores, positions, seg_mean, picks,
                                                                                  summary, slice, items, keep_bits, hmax);
            AVS_LAUNCH_CHECK();
            return AVS_OK;
        }
    }
    // ---- long videos: register-resident DP cells, ONE row buffer in shared memory, keep bits in the workspace
    if (scores != nullptr && b.max_cap + 1 <= KNAP_BIG_CPT * KNAP_BIG_THREADS && keep_bits != nullptr) {
        // row buffer padded to whole groups of 4 cells per thread (see the kernel); the keep rows in the workspace
        // are padded the same way by the host planner (knapsack_keep_words)
        const int rows = (b.max_cap + 4 * KNAP_BIG_THREADS) / (4 * KNAP_BIG_THREADS) * (4 * KNAP_BIG_THREADS);
        const int items = std::max(b.max_S, 1);
        const size_t need = static_cast<size_t>(items) * 8 + (static_cast<size_t>(rows) + 1) * 8 +
                            static_cast<size_t>(items) * 8 + 64;
        if (need <= limit) {
            static PerDeviceOnce cfg;
            const int dev = current_device();
            auto kern = knapsack_fast_kernel<long long, KNAP_BIG_CPT, KNAP_BIG_THREADS, true>;
            if (cfg.needed(dev)) {
                AVS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(limit)));
                cfg.mark(dev);
            }
            kern<<<b.n, KNAP_BIG_THREADS, need, stream>>>(b, scores, positions, seg_mean, picks, summary, rows, items,
                                                          keep_bits);
            AVS_LAUNCH_CHECK();
            return AVS_OK;
        }
    }
    // ---- general path.  Shared-memory budget (<= 200 KiB): DP rows first, then the items, then the keep bits --
    // each only if it fits
    int rows_cap = b.max_cap + 1;
    size_t smem = 2ull * rows_cap * sizeof(long long);
    if (smem > limit) { rows_cap = 0; smem = 0; }
    int items_cap = b.max_S;
    const size_t items_bytes = static_cast<size_t>(items_cap) * 16 + 8;
    if (smem + items_bytes > limit) items_cap = 0;
    else smem += items_bytes;
    const size_t words_max = static_cast<size_t>(b.max_S) * ((static_cast<size_t>(b.max_cap) + 32) >> 5);
    int keep_words = 0;
    if (smem + words_max * 4 <= limit && words_max < (1u << 30)) {
        keep_words = static_cast<int>(words_max);
        smem += words_max * 4;
    }
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        AVS_CUDA(cudaFuncSetAttribute(knapsack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(limit)));
        configured = limit;
    }
    knapsack_kernel<<<b.n, KNAP_THREADS, smem, stream>>>(b, seg_sum, seg_mean, picks, summary, keep_bits, dp_ws,
                                                        rows_cap, items_cap, keep_words,
                                                        items_cap > 0 ? scores : nullptr, positions);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

avs_status temporal_f1_device(const int32_t* pred, const int32_t* pred_start, const int32_t* gt,
                              const int32_t* gt_start, int n, double* f1_dev, cudaStream_t stream) {
    if (n == 0) return AVS_OK;
    temporal_f1_kernel<<<n, 256, 0, stream>>>(reinterpret_cast<const int2*>(pred), pred_start,
                                              reinterpret_cast<const int2*>(gt), gt_start, f1_dev);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

}  // namespace avs
