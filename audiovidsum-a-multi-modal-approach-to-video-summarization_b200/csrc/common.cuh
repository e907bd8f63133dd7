// Internal declarations shared by the translation units of libavsum_b200.so.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include "../../include/avsum_b200.h"

namespace avs {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local char g_err[512];
void count_launch(int n = 1);

#define AVS_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::avs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return _e == cudaErrorMemoryAllocation ? AVS_ERR_OOM : AVS_ERR_CUDA;                    \
        }                                                                                           \
    } while (0)

#define AVS_CHECK(cond, code, ...)            \
    do {                                      \
        if (!(cond)) {                        \
            ::avs::set_error(__VA_ARGS__);    \
            return (code);                    \
        }                                     \
    } while (0)

#define AVS_TRY(expr)                     \
    do {                                  \
        avs_status _s = (expr);           \
        if (_s != AVS_OK) return _s;      \
    } while (0)

#define AVS_LAUNCH_CHECK()                \
    do {                                  \
        ::avs::count_launch();            \
        AVS_CUDA(cudaGetLastError());     \
    } while (0)

// ---- per-device one-time setup ------------------------------------------------
// Function attributes (cudaFuncSetAttribute), __device__ symbols, streams, events and SM counts belong to ONE
// device's context, and the API takes a device per handle: everything cached "once" is cached per device.
constexpr int kMaxDevices = 64;
inline int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < kMaxDevices) ? d : 0;
}
struct PerDeviceOnce {
    std::atomic<bool> done[kMaxDevices] = {};
    // true while the current device still needs the setup; call mark() after it succeeded (a concurrent second
    // setup is harmless: the calls it guards are idempotent)
    bool needed(int dev) const { return !done[dev].load(std::memory_order_acquire); }
    void mark(int dev) { done[dev].store(true, std::memory_order_release); }
};
int device_sm_count();   // multiprocessors of the current device (cached per device)

// ---- element types of GEMM operands / outputs -------------------------------
enum DType : int { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2 };
static inline int dtype_size(int dt) { return dt == DT_F32 ? 4 : 2; }

#ifdef __CUDACC__
// fp32 -> 16-bit container (round to nearest; fp16 saturates instead of overflowing to inf)
// (fminf / fmaxf return the non-NaN operand, so NaN is passed through explicitly)
__device__ __forceinline__ float sat_f16(float x) { return x != x ? x : fminf(fmaxf(x, -65504.f), 65504.f); }
__device__ __forceinline__ uint16_t to_lowp_bits(float x, int dt) {
    if (dt == DT_F16) return __half_as_ushort(__float2half_rn(sat_f16(x)));
    return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}
__device__ __forceinline__ uint32_t pack_lowp2(float a, float b, int dt) {
    if (dt == DT_F16) {
        __half2 h = __floats2half2_rn(sat_f16(a), sat_f16(b));
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
#endif

// ---- GEMM: C[M,N] = act(A[M,K] * W[N,K]^T + bias) ---------------------------
struct GemmEpilogue {
    const float* bias = nullptr;  // [N] fp32 or null
    void* C = nullptr;            // output, row-major, leading dimension ldc (elements)
    int64_t ldc = 0;
    int out_dtype = DT_F32;
    int relu = 0;
    int round_tf32 = 0;           // fp32 outputs rounded (RN) to tf32 for a following tf32 GEMM
    float acc_scale = 1.0f;       // epilogue computes acc * acc_scale + bias (see AVS_PREC_TF32 in api.cu)
    // fused frame-score head (N must be 64): scores[m] = sigmoid(relu(acc + bias) . w2 + b2)
    const float* score_w2 = nullptr;
    const float* score_b2 = nullptr;   // device pointer to 1 float
    float* scores = nullptr;
    // fp16 outputs only: *sat_flag = 1 when a value reached the fp16 range limit (|x| >= 65504) or was not finite
    // (the cast saturates silently).  Points into mapped pinned host memory: the host reads it after a stream sync.
    unsigned int* sat_flag = nullptr;
    // host-side launch hint: at most this many CTAs of the persistent grid (0 = one per SM).  Used when part of the
    // GPU is known to be occupied by concurrently running recurrence clusters: CTAs of a statically scheduled
    // persistent kernel that cannot be placed at once would delay their share of the tiles.
    int max_ctas = 0;
    // host-side launch hint: take the CTA-pair kernel (256 x 256 tiles) even when the pair tiles do not fill the GPU --
    // for GEMMs over a group's rows that run beside other work, where the 128 x 128 tiles' operand traffic (125 B/clk/SM
    // against the 42.6 the L2 delivers) costs more than the unbalanced last round of wide tiles
    int prefer_pairs = 0;
};

// tcgen05 / TMA path.  in_dtype: DT_F32 (kind::tf32), DT_F16 or DT_BF16 (kind::f16).
avs_status gemm_tc(const void* A, int64_t lda, const void* W, int64_t ldw, int in_dtype, int64_t M, int N, int K,
                   const GemmEpilogue& epi, cudaStream_t stream);
avs_status gemm_trace_read(unsigned long long* out8);   // debugging aid (AVS_GEMM_TRACE=1)
// CUDA-core fp32 path (debug / exact-order aid).
avs_status gemm_simt(const float* A, int64_t lda, const float* W, int64_t ldw, int64_t M, int N, int K,
                     const GemmEpilogue& epi, cudaStream_t stream);

// ---- elementwise helpers -----------------------------------------------------
// dst = tf32_rn(src) (fp32 container) or bf16/f16 cast; n elements.
avs_status convert_f32(const float* src, void* dst, int64_t n, int dst_dtype, int round_tf32, cudaStream_t stream);
// *count_dev += number of fp16 elements that are saturated (|x| = 65504), infinite or NaN (AVS_CHECK_RANGE=1)
avs_status count_saturated_f16(const void* x, int64_t n, unsigned int* count_dev, cudaStream_t stream);

// ---- LSTM recurrence -----------------------------------------------------------
struct LstmBatch {            // device arrays, one entry per slot (n_groups * NB slots)
    const int32_t* slot_row_start;
    const int32_t* slot_len;      // 0 for empty slots
    const int32_t* group_maxlen;  // [n_groups]
    int n_groups;
    int nb;                       // videos per cluster (1, 2, 4, 8, 16)
    int excl = 0;                 // tensor-core kernel, 8-slot variant: 0 = launch policy decides which groups get
                                  // exclusive SMs, -1 = none (the call shares the GPU with other video groups)
    int lane_map = 0;             // h exchange: 0 = lane -> (peer = lane % 8, video = lane / 8), 1 = (lane / 4, lane % 4)
};
// xg_v, xg_a: [rows, 2048] gate pre-activations (biases included) in the packed column order
//   col = dir * 1024 + cta * 128 + jj * 4 + gate   (hidden unit j = cta * 32 + jj)
// whh: [4][1024][256] fp32 in the same packed row order;  fused: [rows, 1024].
avs_status lstm_recurrence(const float* xg_v, const float* xg_a, const float* whh_packed, const LstmBatch& batch,
                           float* fused, int round_tf32, void* fused_lowp, int lowp_dtype, cudaStream_t stream);

// tcgen05 version (16-bit operands: op_dtype = DT_F16 or DT_BF16; fp32 accumulate / state); batch.nb must be
// 8 (two CTAs per SM), 16, 32 or 64.  fused output: fp32 (optionally tf32-rounded), fp16 or bf16 (out_dtype).
// save_pre (float4 [rows, 4, 256]: i, f, g, o pre-activations) / save_c (float [rows, 4, 256]: new cell state) are
// written when non-null (training forward; consumed by lstm_backward).
// xg_dtype: DT_F32 or DT_F16 (the element type of xg_v / xg_a; the inference path keeps them in fp16).
avs_status lstm_recurrence_tc(const void* xg_v, const void* xg_a, int xg_dtype, const float* whh_packed,
                              const LstmBatch& batch, int op_dtype, void* fused, int out_dtype, int round_tf32,
                              cudaStream_t stream, void* save_pre = nullptr, float* save_c = nullptr);
// The 8-slot variant (batch.nb == 8) for groups [g_lo, g_hi) only, on `stream`, with (exclusive != 0) or without an
// SM-exclusive shared-memory request: the pipelined forward launches the groups one by one as their input
// projections become available.  lstm_exclusive_groups: how many leading groups the one-launch policy would give
// exclusive SMs for this group count.
avs_status lstm_recurrence_tc_groups(const void* xg_v, const void* xg_a, int xg_dtype, const float* whh_packed,
                                     const LstmBatch& batch, int g_lo, int g_hi, int exclusive, int op_dtype, void* fused,
                                     int out_dtype, cudaStream_t stream);
int lstm_exclusive_groups(int n_groups);
// Holds `stream` for ~ns nanoseconds (one sleeping thread).  Orders the PLACEMENT of kernels that become eligible at the
// same moment on different streams: SM-exclusive recurrence CTAs need EMPTY SMs, so they must be placed before the
// shared CTAs of the other groups spread over every SM (otherwise the longest chain waits for whole groups to finish).
avs_status launch_stagger(cudaStream_t stream, unsigned int ns);

// BPTT through the four recurrences (CUDA-core fp32, cluster of 8 CTAs, DSMEM reduce-scatter of dh):
// d_fused [rows, 1024] -> d_xg_v / d_xg_a [rows, 2048] (gate gradients in the packed column order of xg).
avs_status lstm_backward(const float* d_fused, const void* save_pre, const float* save_c, const float* whh_packed,
                         const LstmBatch& batch, float* d_xg_v, float* d_xg_a, cudaStream_t stream);

avs_status lstm_trace_read(unsigned long long* out8);   // AVS_LSTM_TRACE=1 debugging aid
avs_status bptt_trace_read(unsigned long long* out10);  // AVS_BPTT_TRACE=1 debugging aid

// ---- attention core -------------------------------------------------------------
struct SeqDesc {  // device arrays [n_seqs]
    const int32_t* base;
    const int32_t* stride;
    const int32_t* len;
    int n_seqs;
    int max_len;
};
// CUDA-core reference implementation (fp32), any sequence layout.
avs_status attention_simt(const float* qkv, int64_t ld_qkv, int E, int H, const SeqDesc& seqs, void* ctx,
                          int64_t ld_ctx, int out_dtype, int round_tf32, cudaStream_t stream);

// tcgen05 flash-style kernel: fp16 / bf16 (in_dtype) qkv [rows, 3E], contiguous sequences (stride 1), head dim 256.
avs_status attention_tc(const void* qkv_h, int in_dtype, int64_t rows, int E, int H, const SeqDesc& seqs, void* ctx,
                        int64_t ld_ctx, int out_dtype, int round_tf32, cudaStream_t stream);

// ---- summary generation -----------------------------------------------------------
struct SummaryBatch {   // device arrays, [n] unless noted
    const int32_t* row_start;
    const int32_t* lengths;
    const int32_t* n_frames;
    const int32_t* cps;          // [sum S, 2]
    const int32_t* cps_start;    // [n + 1]
    const int64_t* summary_start;// [n + 1] or null
    const int64_t* keep_start;   // [n + 1] word offsets into keep bits workspace
    const int64_t* dp_start;     // [n + 1] element offsets into global dp workspace (2 rows per video)
    int n;
    int prop_num, prop_den;
    int max_cap;                 // max capacity over the batch
    int max_S;                   // max number of shots of one video
    int max_wt = 0;              // longest shot (frames) of the batch
};
avs_status shot_pool(const float* scores, const int32_t* positions, const SummaryBatch& b, unsigned long long* seg_sum,
                     cudaStream_t stream);
// scores / positions non-null (and knapsack_can_fuse_pool(b)): the kernel pools the frame scores itself (fused K7)
// and seg_sum is not read.
bool knapsack_can_fuse_pool(const SummaryBatch& b);
int64_t knapsack_keep_words(long long cap);   // keep-bit words per item the workspace must hold for this capacity
avs_status knapsack_select(const SummaryBatch& b, const unsigned long long* seg_sum, long long* seg_mean,
                           uint8_t* picks, uint8_t* summary, uint32_t* keep_bits, long long* dp_ws,
                           cudaStream_t stream, const float* scores = nullptr, const int32_t* positions = nullptr);
avs_status temporal_f1_device(const int32_t* pred, const int32_t* pred_start, const int32_t* gt,
                              const int32_t* gt_start, int n, double* f1_dev, cudaStream_t stream);

// ---- training helpers (train.cu) -------------------------------------------------------------
// dst[map(c)][r] = src[r][c] (optionally tf32-rounded); perm 1 = LSTM gate-row un-permutation per 1024-column block
avs_status transpose_f32(const float* src, int64_t ld_src, int R, int C, float* dst, int64_t ld_dst, int perm, int round,
                         cudaStream_t stream);
avs_status colsum_f32(const float* src, int64_t ld_src, int R, int C, float* out, int perm, cudaStream_t stream);
avs_status shift_h(const float* fused, const int32_t* row_start, const int32_t* lengths, int n_videos, int max_len,
                   float* hprev, cudaStream_t stream);

// ---- evaluation metrics / fusion helpers (metrics.cu) -------------------------------------------
avs_status eval_metrics_device(const float* pred, const void* target, int tgt_f64, const int32_t* row_start,
                               const int32_t* lengths, int n_videos, double* out_f, long long* out_i,
                               cudaStream_t stream);
avs_status cdist_device(const float* a, const float* b, int na, int nb, int D, double* out, cudaStream_t stream);
avs_status gather_scale_device(const float* feat, const int32_t* idx, const float* w, int U, int D, float* out,
                               cudaStream_t stream);
avs_status dtw_device(const double* cost, int n, int m, double* acc, uint8_t* choice, int32_t* path, int32_t* path_len,
                      double* total, cudaStream_t stream);

}  // namespace avs
