/*
 * avsum_b200 -- C ABI of the B200-native AudioVidSum hot path (libavsum_b200.so).
 *
 * The reference (Research-Implementation/AudioVidSum) is pure Python over torch.nn and
 * defines NO plugin / operator / FFI interface; its boundary is the
 * Python surface of models/av_model.py.  These entry points are what a binding for that
 * surface calls; each one cites the reference code it replaces.  INTEGRATION.md shows the
 * ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every function returns an avs_status (0 = OK) and never throws;
 *     avs_last_error() returns a thread-local message for the last failure;
 *   - all tensors are dense row-major; "rows" are frames, videos are described by
 *     (row_start[b], length[b]) so both packed (sum T rows) and padded [B, Tmax] batches work;
 *   - `space` says where the caller's data buffers live: AVS_DEVICE pointers are used in
 *     place on `stream`; AVS_HOST buffers (pinned recommended) are copied H2D / D2H inside the
 *     call and the call returns after the results have landed (stream synchronised);
 *   - small int32 descriptor arrays (row_start, lengths, cps offsets ...) are ALWAYS host memory;
 *   - the model handle owns packed device weights and a grow-only device workspace; calls on
 *     one handle must not overlap (one handle per stream / thread).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     AVS_ERR_CUDA.
 */
#ifndef AVSUM_B200_H
#define AVSUM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int avs_status;
enum {
    AVS_OK = 0,
    AVS_ERR_INVALID = 1,     /* bad argument (message says which)                     */
    AVS_ERR_UNSUPPORTED = 2, /* shape outside what the sm_100a kernels are built for   */
    AVS_ERR_CUDA = 3,        /* CUDA runtime / driver error, or no sm_100 device       */
    AVS_ERR_OOM = 4
};

enum { AVS_HOST = 0, AVS_DEVICE = 1 };

/* Which axis nn.MultiheadAttention attends over.
 * AVS_ATTN_LITERAL  : exactly models/av_model.py:44 -- the module is built without batch_first
 *                     (av_model.py:26) yet fed [B, T, E], so frame t attends over the B videos
 *                     of the batch (weights == 1 when B == 1).  Requires equal lengths.
 * AVS_ATTN_TEMPORAL : frame self-attention inside each video with variable-length masking
 *                     (the arithmetic of models/attention.py:15-25 on in_proj/out_proj weights).
 * AVS_ATTN_LITERAL_B1: every video is its own B = 1 call of the reference, i.e. the loop of
 *                     scripts/evaluate.py:12-18 batched into one launch chain (lengths may differ). */
enum { AVS_ATTN_LITERAL = 0, AVS_ATTN_TEMPORAL = 1, AVS_ATTN_LITERAL_B1 = 2 };

/* Arithmetic of the dense contractions.
 * AVS_PREC_TF32 : fp32 storage, tcgen05 kind::tf32 with round-to-nearest operands, fp32 accumulate
 *                 (attention core: fp16 operands = the same 11-bit significand); LSTM recurrence,
 *                 softmax and all pointwise math in fp32.  Target: <= 1e-3 relative on scores.
 * AVS_PREC_BF16 : bf16 operands/activations, fp32 accumulate (stated looser tolerance).
 * AVS_PREC_FP32_SIMT : CUDA-core fp32 contractions (slow; exact-order debugging aid). */
enum { AVS_PREC_TF32 = 0, AVS_PREC_BF16 = 1, AVS_PREC_FP32_SIMT = 2 };

/* Element type of the caller's visual / audio feature buffers (avs_model_set_feature_format).
 * AVS_FEAT_F32 (default): float32, what the reference's data/dataset.py:19-32 loads.
 * AVS_FEAT_F16: IEEE half -- an opt-in 16-bit feature cache (half the PCIe / HBM bytes per frame).  The fc layers
 *   then run kind::f16 on the stored values against fp16 weights (11-bit significands on both sides, exactly what
 *   the default mode's tf32 operands carry); needs AVS_PREC_TF32 and feature dims that are multiples of 8. */
enum { AVS_FEAT_F32 = 0, AVS_FEAT_F16 = 1 };

/* fp32 parameters in the reference's state_dict layout (av_model.py:10-31).
 * lstm_*[i]: i = 0 visual fwd, 1 visual reverse, 2 audio fwd, 3 audio reverse.
 * Pointers may be host or device memory (copied with cudaMemcpyDefault). */
typedef struct avs_weights {
    int32_t visual_dim, audio_dim, hidden_dim, num_heads;
    const float *visual_fc_w, *visual_fc_b; /* [H, Dv], [H]                       av_model.py:10-12 */
    const float *audio_fc_w, *audio_fc_b;   /* [H, Da], [H]                       av_model.py:13-15 */
    const float* lstm_w_ih[4];              /* [4*H/2, H]   gate order i,f,g,o    av_model.py:18-23 */
    const float* lstm_w_hh[4];              /* [4*H/2, H/2]                                         */
    const float* lstm_b_ih[4];              /* [4*H/2]                                              */
    const float* lstm_b_hh[4];              /* [4*H/2]                                              */
    const float *attn_in_w, *attn_in_b;     /* [3E, E], [3E]  E = 2H              av_model.py:26    */
    const float *attn_out_w, *attn_out_b;   /* [E, E], [E]                                          */
    const float *scorer0_w, *scorer0_b;     /* [64, E], [64]                      av_model.py:29-31 */
    const float *scorer2_w, *scorer2_b;     /* [1, 64], [1]                                         */
} avs_weights;

typedef struct avs_model avs_model;

const char* avs_last_error(void);
int avs_version(void);
/* 1 if a usable sm_100 device is visible, else 0 (never fails). */
int avs_device_ok(void);

/* Replaces AVBiLSTMModel.__init__ + .cuda() (av_model.py:7-31; scripts/evaluate.py:13):
 * packs the weights for the kernels (gate interleave for the LSTM slices, tf32 rounding,
 * bf16 copies) on `device`.  The packing runs on a stream of the handle: the weight tensors must be complete (the
 * stream that wrote them synchronised) when this is called; it returns after the packing has finished.  Handles on
 * several devices may coexist in one process (kernel attributes, SM counts and helper streams are kept per device). */
avs_status avs_model_create(const avs_weights* w, int device, avs_model** out);
/* Re-pack after the caller changed parameters (load_state_dict / optimizer step).  Synchronous: waits for all work
 * queued on the device (cudaDeviceSynchronize), packs, returns when the handle is up to date. */
avs_status avs_model_update(avs_model* m, const avs_weights* w);
/* The same without a host synchronisation: the packing kernels are queued on `cuda_stream`, so the source tensors
 * must not change before the stream gets there (true for parameters updated by work on the same stream -- the
 * training loop of scripts/train_av_model.py:94-96).  lstm_only != 0 re-packs only the four recurrences' tensors,
 * the only ones a training step reads through the handle (avs_bilstm_pair_train / _bwd). */
avs_status avs_model_update_async(avs_model* m, const avs_weights* w, int lstm_only, void* cuda_stream);
/* How the `visual` / `audio` arguments of avs_forward, avs_forward_summarize and avs_forward_summarize_async are
 * read from now on: AVS_FEAT_F32 (the declared `const float*`) or AVS_FEAT_F16 (the same pointers address IEEE half
 * values, row pitch visual_dim / audio_dim halves).  The reference has no such switch (its features are float32
 * .npy files, data/dataset.py:25-28); this is the 16-bit host feature cache of data.dataset.packed_batches.
 * No asynchronous step may be in flight. */
avs_status avs_model_set_feature_format(avs_model* m, int format);
void avs_model_destroy(avs_model* m);

/* Replaces AVBiLSTMModel.forward (av_model.py:33-46) for a batch of n_videos videos.
 *   visual [total_rows, Dv], audio [total_rows, Da], scores [total_rows]  (fp32, `space`)
 *   row_start[b], lengths[b] : host int32; rows outside every video (padding) never influence
 *   a video's result and their scores are unspecified.  The caller applies the final
 *   .squeeze() (a view). */
avs_status avs_forward(avs_model* m, const float* visual, const float* audio, int64_t total_rows,
                       int32_t n_videos, const int32_t* row_start, const int32_t* lengths, int attn_axis,
                       int precision, float* scores, int space, void* cuda_stream);

/* Summary generation for a batch (shot pooling over change points + 0/1 knapsack at the
 * length budget + keyshot bitmap).  NOT IN THE REFERENCE: specified by
 * oracle/av_oracle.py (integer arithmetic, bit-exact).
 *   scores     fp32 [total_rows] (`space`), same row indexing as avs_forward
 *   positions  int32 [total_rows] (`space`): original frame index of every sampled frame
 *   n_frames   host int32 [n]; cps host int32 [sum S, 2] inclusive; cps_start host int32 [n+1]
 *   capacity_b = floor(n_frames_b * prop_num / prop_den)
 * outputs (`space`): picks uint8 [sum S]; seg_mean int64 [sum S] (2^-24 fixed point, may be NULL);
 *   summary uint8 [sum n_frames] addressed by host int64 summary_start[n+1] (may be NULL). */
avs_status avs_summarize(avs_model* m, const float* scores, const int32_t* positions, int32_t n_videos,
                         const int32_t* row_start, const int32_t* lengths, const int32_t* n_frames,
                         const int32_t* cps, const int32_t* cps_start, int32_t prop_num, int32_t prop_den,
                         uint8_t* picks, int64_t* seg_mean, uint8_t* summary, const int64_t* summary_start,
                         int space, void* cuda_stream);

/* avs_forward followed by avs_summarize in one call (the "scored + summarised" step of BASELINE.json): scores stay
 * on the device between the two halves; in host space the call returns after ONE synchronisation with scores,
 * picks, seg_mean (may be NULL) and summary (may be NULL) in the caller's host buffers. */
avs_status avs_forward_summarize(avs_model* m, const float* visual, const float* audio, const int32_t* positions,
                                 int64_t total_rows, int32_t n_videos, const int32_t* row_start,
                                 const int32_t* lengths, int attn_axis, int precision, const int32_t* n_frames,
                                 const int32_t* cps, const int32_t* cps_start, int32_t prop_num, int32_t prop_den,
                                 float* scores, uint8_t* picks, int64_t* seg_mean, uint8_t* summary,
                                 const int64_t* summary_start, int space, void* cuda_stream);

/* Streaming form of avs_forward_summarize for HOST-space (pinned) buffers -- the scripts/evaluate.py:12-18 loop over
 * many batches: the call enqueues the whole step on `cuda_stream` and returns WITHOUT synchronising, so the
 * features of batch i+1 cross PCIe while batch i is still being computed.  `slot` (0 or 1) selects one of two
 * device staging areas; a slot may be reused only after avs_slot_wait(m, slot) has returned, which is also when
 * the slot's output buffers are valid.  The input buffers must stay untouched until then. */
avs_status avs_forward_summarize_async(avs_model* m, const float* visual, const float* audio,
                                       const int32_t* positions, int64_t total_rows, int32_t n_videos,
                                       const int32_t* row_start, const int32_t* lengths, int attn_axis,
                                       int precision, const int32_t* n_frames, const int32_t* cps,
                                       const int32_t* cps_start, int32_t prop_num, int32_t prop_den, float* scores,
                                       uint8_t* picks, int64_t* seg_mean, uint8_t* summary,
                                       const int64_t* summary_start, int slot, void* cuda_stream);
/* fp16 range watch.  The default precision keeps the activations behind the fc layers in fp16, whose cast SATURATES at
 * 65504 instead of overflowing -- silent for un-normalised features of huge magnitude (the reference computes in fp32).
 * The fc GEMM epilogues therefore watch the magnitudes they write (dense batches: every row owned by a video) and set
 * a flag; HOST-space calls (avs_forward, avs_forward_summarize, avs_slot_wait) fail with AVS_ERR_UNSUPPORTED when it is
 * set.  DEVICE-space calls return before the work has run: their callers synchronise and ask here.  *saturated = 1
 * when an activation was clamped since the last report; reading clears the flag. */
avs_status avs_model_range_status(avs_model* m, int32_t* saturated);

/* Block until the asynchronous step that used `slot` has completed (no-op for an idle slot). */
avs_status avs_slot_wait(avs_model* m, int slot);

/* ---- building blocks (device pointers only), exported so each kernel can be parity-tested
 * against the reference sub-module it replaces and reused by
 * models/attention.py's drop-in. ---- */

/* nn.Linear (+ReLU): C[M, N] = act(A[M, K] * W[N, K]^T + bias[N]); fp32 in/out.
 * `precision` as above; K % 4 == 0, N % 16 == 0 for the tensor-core paths. */
avs_status avs_linear(const float* A, const float* W, const float* bias, int64_t M, int32_t N, int32_t K,
                      int relu, int precision, float* C, void* cuda_stream);

/* Both nn.LSTM modules of the model (av_model.py:39-40) on packed rows:
 * v_emb, a_emb [total_rows, H] -> fused [total_rows, 2H] = [v_fwd | v_bwd | a_fwd | a_bwd]. */
avs_status avs_bilstm_pair(avs_model* m, const float* v_emb, const float* a_emb, int64_t total_rows,
                           int32_t n_videos, const int32_t* row_start, const int32_t* lengths, int precision,
                           float* fused, void* cuda_stream);

/* Scaled-dot-product multi-head attention core on a packed qkv buffer [rows, 3E] (q | k | v),
 * heads = contiguous dh-wide column slices (attention.py:17-23).  Sequence s consists of rows
 * seq_base[s] + i * seq_stride[s], i < seq_len[s] (host int32 arrays).  ctx [rows, E] fp32. */
avs_status avs_attention(const float* qkv, int64_t rows, int32_t E, int32_t num_heads, int32_t n_seqs,
                         const int32_t* seq_base, const int32_t* seq_stride, const int32_t* seq_len,
                         int precision, float* ctx, void* cuda_stream);

/* Overlap-F1 between predicted and ground-truth shot lists for a batch (evaluation/metrics.py:1-9,
 * utils/shot_metrics.py:4-16): shots int32 [*, 2] half-open (start, end), offsets host int32 [n+1].
 * f1 double [n] written to host memory. */
avs_status avs_temporal_f1(const int32_t* pred, const int32_t* pred_start, const int32_t* gt,
                           const int32_t* gt_start, int32_t n_videos, double* f1_host, void* cuda_stream);

/* ---- training step building blocks (scripts/train_av_model.py:86-96: forward in train mode, loss.backward()).
 * The reference trains through torch autograd; these are the backward kernels an autograd.Function around the
 * forward building blocks calls (device pointers, fp32).  Optimiser and loss stay the caller's (torch.optim.AdamW,
 * F.mse_loss -- train_av_model.py:68,91); gradients are all-reduced by the caller (NCCL) in data-parallel runs. */

/* avs_bilstm_pair (tensor-core mode) that also saves what BPTT needs:
 * save_pre float [rows, 4, 256, 4] = gate pre-activations (i, f, g, o) per (frame, recurrence, hidden unit),
 * save_c   float [rows, 4, 256]    = cell state after the step; recurrence order as lstm_* in avs_weights. */
avs_status avs_bilstm_pair_train(avs_model* m, const float* v_emb, const float* a_emb, int64_t total_rows,
                                 int32_t n_videos, const int32_t* row_start, const int32_t* lengths, float* fused,
                                 float* save_pre, float* save_c, void* cuda_stream);

/* Backward of avs_bilstm_pair: d_fused [rows, 1024] -> d_v_emb, d_a_emb [rows, 512] and, per recurrence i (order of
 * avs_weights.lstm_*), dW_ih[i] [1024, 512], dW_hh[i] [1024, 256], db[i] [1024] (= grad of b_ih = grad of b_hh) in the
 * reference's gate-row order.  `fused`, `v_emb`, `a_emb` are the forward's output / inputs. */
avs_status avs_bilstm_pair_bwd(avs_model* m, const float* d_fused, const float* save_pre, const float* save_c,
                               const float* fused, const float* v_emb, const float* a_emb, int64_t total_rows,
                               int32_t n_videos, const int32_t* row_start, const int32_t* lengths, float* d_v_emb,
                               float* d_a_emb, float* const* dW_ih, float* const* dW_hh, float* const* db,
                               void* cuda_stream);

/* Backward of nn.Linear y = x W^T + b:  dX[M, K] = dY W,  dW[N, K] = dY^T X,  db[N] = column sums of dY; any of
 * dX / dW / db may be NULL.  tcgen05 kind::tf32 GEMMs on round-to-nearest operands; N % 4 == 0 and K % 4 == 0. */
avs_status avs_linear_bwd(const float* dY, const float* X, const float* W, int64_t M, int32_t N, int32_t K, float* dX,
                          float* dW, float* db, void* cuda_stream);

/* The metric block of scripts/evaluate.py:25-36 for a batch of videos (the caller right after the forward):
 *   binary_pred = pred > np.mean(pred), binary_target = target > np.mean(target)   (np.mean's pairwise
 *   summation restated exactly), tp / precision / recall / F1 with the 1e-8 guard, scipy.stats.spearmanr
 *   (Pearson correlation of average ranks) and scipy.stats.kendalltau (tau-b) per video.
 * pred fp32 [rows]; target fp32 or fp64 (target_is_f64) [rows]; same (row_start, lengths) indexing as avs_forward.
 * metrics double [n, 4] = f1, spearman, kendall, np.mean(pred);  counts int64 [n, 8] = tp, sum(binary_pred),
 * sum(binary_target), discordant pairs, x ties, y ties, joint ties, n.  `space` applies to pred / target /
 * metrics / counts.  F1 and tau are bit-exact against numpy / scipy; rho to ~1e-15 (exact integer moments). */
avs_status avs_eval_metrics(const float* pred, const void* target, int target_is_f64, int32_t n_videos,
                            const int32_t* row_start, const int32_t* lengths, double* metrics, int64_t* counts,
                            int space, void* cuda_stream);

/* features/fusion.py:7-12 compute_dtw: scipy cdist(a, b, "euclidean") -- a [na, D], b [nb, D] fp32, out [na, nb]
 * float64, summed in scipy's order (bit-exact). */
avs_status avs_cdist(const float* a, const float* b, int32_t na, int32_t nb, int32_t D, double* out, int space,
                     void* cuda_stream);

/* features/fusion.py:21-32 interpolate_features: out[k, :] = features[idx[k], :] * weights[k]  (fp32 multiply).
 * idx / weights: host arrays [U] (np.unique of the path's first column and counts / counts.sum()). */
avs_status avs_interpolate(const float* features, int64_t n_rows, int32_t D, const int32_t* idx, const float* weights,
                           int32_t U, float* out, int space, void* cuda_stream);

/* Exact dynamic-time-warping path through a cost matrix [n, m] (float64, host or device; results to host):
 * fastdtw's published recurrence D[i,j] = c[i,j] + min(D[i-1,j], D[i,j-1], D[i-1,j-1]) with its tie order.
 * This is what features/fusion.py:15-18 evidently intends; the reference's own call raises TypeError.
 * path int32 [(n + m - 1), 2] (first *path_len rows valid), total = accumulated cost. */
avs_status avs_dtw_path(const double* cost, int32_t n, int32_t m, int32_t* path, int32_t* path_len, double* total,
                        void* cuda_stream);

/* Counters: number of kernel launches issued by this library since load (gpu_launches in bench.py). */
int64_t avs_launch_count(void);

/* Per-stage device timing with CUDA events on the launching stream (used by bench.py's roofline).
 * avs_profile(1) enables, (0) disables, (2) resets the accumulators and enables.  Events are
 * resolved lazily by avs_profile_read, which fills ms[avs_profile_stages()] and calls[...]. */
void avs_profile(int enable);
int avs_profile_stages(void);
const char* avs_profile_stage_name(int stage);
void avs_profile_read(double* ms, int64_t* calls);

/* Debugging aid (no reference counterpart): phase trace of the tensor-core LSTM recurrence, filled when the
 * environment variable AVS_LSTM_TRACE is set.  out8[0..6] = summed clock64 deltas of the per-step chain
 * (h landed -> MMAs issued -> epilogue awake -> tcgen05.ld -> cell math -> fence+barrier -> copies issued ->
 * next h landed), out8[7] = number of steps. */
avs_status avs_debug_lstm_trace(uint64_t* out8);
/* The same for the tensor-core BPTT kernel of the training step (AVS_BPTT_TRACE): out10[0..4] = partials landed ->
 * B operand staged -> MMAs issued -> epilogue awake -> tcgen05.ld -> st.async issued, [6] = partials sent -> next
 * partials landed, [7] = partials sent -> next step's dh-independent math done, out10[8] = number of steps. */
avs_status avs_debug_bptt_trace(uint64_t* out10);

/* Page-locked host buffers for the host-space entry points (the features a loader packs).  write_combined != 0:
 * write-combined pages -- the CPU should only write them; device reads across PCIe then do not snoop CPU caches. */
avs_status avs_host_alloc(void** out, size_t bytes, int write_combined);
avs_status avs_host_free(void* p);

/* Debugging aid: clock64 totals of the tensor-core GEMM pipeline (block 0) since the last read, recorded when the
 * environment variable AVS_GEMM_TRACE is set: out8[0] MMA thread span, [1] MMA waiting for operands, [2] MMA waiting
 * for a drained accumulator, [3] producer waiting for a free stage, [4] epilogue warp waiting for an accumulator,
 * [5] epilogue span, [6] tiles. */
avs_status avs_debug_gemm_trace(uint64_t* out8);

/* Debugging aid: timeline of the last pipelined host-space avs_forward_summarize call, recorded when the
 * environment variable AVS_E2E_TRACE is 1.  out20[0] = number of video groups G; out20[1..5] = host clock (ms after
 * entry) at entry / copies queued / groups queued / tail queued / synchronised; in ms after the first device
 * timestamp: out20[6..6+G) = features of group g landed, out20[12..12+G) = forward of group g finished,
 * out20[18] = pooling + knapsack finished, out20[19] = last D2H copy finished. */
avs_status avs_debug_e2e_trace(double* out20);

/* Host logic only (no GPU work; callable on a machine without a device): the recurrence plan avs_forward derives from
 * a batch's descriptors -- videos sorted by length (stable) and cut into groups of one recurrence cluster set each.
 * group_of[n_videos] = group of every video (-1: empty video); info[0] = number of groups, info[1] = slots per group,
 * info[2] = 1 when the groups tile the rows in order (rows laid out longest video first: what the per-group
 * schedule of DESIGN.md section 4 needs; otherwise the forward runs one launch per stage over all rows),
 * info[3] = 1 when the groups end at different times (the shortest group's longest video <= 0.9 x the longest
 * group's); group_rows[2 g], [2 g + 1] = row range of group g when info[2] (at most max_groups pairs; may be NULL).
 * Used by the CPU tests of the callers' row layout (evaluate, summarize_videos, packed_batches, DeviceDataset). */
avs_status avs_debug_plan(int32_t n_videos, const int32_t* row_start, const int32_t* lengths, int64_t total_rows,
                          int32_t* group_of, int32_t* info, int64_t* group_rows, int32_t max_groups);

#ifdef __cplusplus
}
#endif
#endif /* AVSUM_B200_H */
