// C ABI of libavsum_b200.so (declared in include/avsum_b200.h): model handle, weight packing,
// workspace management and the orchestration of AVBiLSTMModel.forward
// (/root/reference/models/av_model.py:33-46) as a short chain of sm_100a kernels.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <numeric>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace avs {

thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int device_sm_count() {
    static std::atomic<int> cached[kMaxDevices] = {};
    const int dev = current_device();
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// ---- optional per-stage device timing (CUDA events on the launching stream) -------------
// Enabled by avs_profile(1); events are resolved lazily in avs_profile_read so the hot path
// never synchronises because of profiling.
enum Stage : int { ST_CONVERT = 0, ST_FC, ST_IH_PROJ, ST_LSTM, ST_QKV_PROJ, ST_ATTENTION, ST_OUT_PROJ, ST_SCORER,
                   ST_POOL, ST_KNAPSACK, ST_FRONT, ST_FRONT_LSTM, ST_TAIL_EXPOSED, ST_COUNT };
static const char* const kStageNames[ST_COUNT] = {"convert_tf32", "fc_gemm", "lstm_input_gemm", "lstm_recurrence",
                                                  "attn_in_proj_gemm", "attention_core", "attn_out_proj_gemm",
                                                  "score_head_gemm", "shot_pool", "knapsack_select",
                                                  "frontend_gemms", "frontend_lstm_pipelined",
                                                  "tail_behind_longest_recurrence"};
struct Profiler {
    bool enabled = false;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    // span != 0: the records of one span (same id) time concurrent pieces of ONE stage that start together; the stage
    // lasts until the last of them ends (maximum, counted once)
    struct Rec { int stage; cudaEvent_t a, b; long long span = 0; };
    long long next_span = 1;
    std::vector<Rec> pending;
    double ms[ST_COUNT] = {};
    long long calls[ST_COUNT] = {};
    cudaEvent_t get() {
        if (used == pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
        }
        return pool[used++];
    }
    void resolve() {
        long long cur_span = 0;
        int span_stage = 0;
        float span_ms = 0.f;
        auto close_span = [&]() {
            if (cur_span) {
                ms[span_stage] += span_ms;
                calls[span_stage] += 1;
            }
            cur_span = 0;
            span_ms = 0.f;
        };
        for (auto& r : pending) {
            cudaEventSynchronize(r.b);
            float t = 0.f;
            if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) continue;
            if (r.span != cur_span) close_span();
            if (r.span) {   // records of a span are pushed back to back
                cur_span = r.span;
                span_stage = r.stage;
                span_ms = std::max(span_ms, t);
            } else {
                ms[r.stage] += t;
                calls[r.stage] += 1;
            }
        }
        close_span();
        pending.clear();
        used = 0;
    }
};
static Profiler g_prof;
struct StageTimer {
    int stage;
    cudaStream_t st;
    cudaEvent_t a = nullptr;
    StageTimer(int stage_, cudaStream_t st_) : stage(stage_), st(st_) {
        if (g_prof.enabled) {
            a = g_prof.get();
            cudaEventRecord(a, st);
        }
    }
    ~StageTimer() {
        if (a) {
            cudaEvent_t b = g_prof.get();
            cudaEventRecord(b, st);
            g_prof.pending.push_back({stage, a, b, 0});
        }
    }
};

// ---- optional timeline of a pipelined host-space call (AVS_E2E_TRACE=1; read with avs_debug_e2e_trace) ----------
// CUDA events with timing next to the (timing-disabled) dependency events of forward_entry / avs_forward_summarize,
// plus host clock samples; a debugging aid that explains where an end-to-end step goes.
struct E2ETrace {
    int on = -1;
    cudaEvent_t t0 = nullptr, chunk[6] = {}, grp[6] = {}, knap = nullptr, end = nullptr;
    int n_groups = 0;
    double host[6] = {};   // seconds: [0] entry, [1] copies queued, [2] groups queued, [3] tail queued, [4] synchronised
    bool enabled() {
        if (on < 0) {
            const char* e = getenv("AVS_E2E_TRACE");
            on = (e && e[0] == '1') ? 1 : 0;
            if (on) {
                cudaEventCreate(&t0);
                cudaEventCreate(&knap);
                cudaEventCreate(&end);
                for (int i = 0; i < 6; ++i) { cudaEventCreate(&chunk[i]); cudaEventCreate(&grp[i]); }
            }
        }
        return on == 1;
    }
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + ts.tv_nsec * 1e-9;
    }
};
static E2ETrace g_e2e;

namespace {

constexpr int HC = 256;  // the kernels are built for hidden_dim = 512 (SURVEY.md 8a defaults)
constexpr int H = 512;
constexpr int E = 1024;
constexpr int G4 = 4 * HC;  // 1024 gate rows per direction

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// feature buffers are fp32 or (avs_model_set_feature_format) fp16: pointer arithmetic in bytes
inline const float* feat_at(const float* base, int64_t row, int dim, size_t fsz) {
    return reinterpret_cast<const float*>(reinterpret_cast<const char*>(base) + static_cast<size_t>(row) * dim * fsz);
}
inline float* feat_at(float* base, int64_t row, int dim, size_t fsz) {
    return reinterpret_cast<float*>(reinterpret_cast<char*>(base) + static_cast<size_t>(row) * dim * fsz);
}

struct Arena {
    char* base = nullptr;
    size_t cap = 0;
    size_t off = 0;
    avs_status reserve(size_t bytes) {
        if (bytes <= cap) return AVS_OK;
        if (base) AVS_CUDA(cudaFree(base));  // synchronises with outstanding work
        base = nullptr;
        cap = 0;
        const size_t want = align_up(bytes + bytes / 8, 1 << 20);
        AVS_CUDA(cudaMalloc(&base, want));
        cap = want;
        return AVS_OK;
    }
    void reset() { off = 0; }
    template <typename T>
    T* take(size_t n) {
        off = align_up(off, 256);
        T* p = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        return p;
    }
    void release() {
        if (base) cudaFree(base);
        base = nullptr;
        cap = off = 0;
    }
};

// dst[p, :] = src[perm(p), :] for the LSTM gate interleave:
//   packed row p = cta*128 + jj*4 + gate   <-   original row gate*256 + cta*32 + jj
// (the four gates of a hidden unit are adjacent, so they land in adjacent accumulator lanes)
__global__ void permute_gate_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int cols,
                                         int round_tf32) {
    const int p = blockIdx.x;
    const int cta = p >> 7, gate = p & 3, jj = (p >> 2) & 31;
    const int o = gate * HC + cta * 32 + jj;
    for (int k = threadIdx.x; k < cols; k += blockDim.x) {
        const float v = src[static_cast<size_t>(o) * cols + k];
        dst[static_cast<size_t>(p) * cols + k] = round_tf32 ? to_tf32_rn(v) : v;
    }
}
// Small host arrays (launch plans, sequence descriptors) travel to the device INSIDE kernel parameters instead of
// through cudaMemcpyAsync: a pageable-memory copy would queue on the H2D copy engine behind the multi-megabyte
// feature transfers of the next video group and stall the pipeline.
struct UploadPayload { uint32_t w[896]; };   // 3,584 bytes, inside the 4 KB kernel-parameter limit
__global__ void upload_kernel(UploadPayload p, uint32_t* __restrict__ dst, int n_words) {
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = p.w[i];
}

// Row packing of padded / sparse batches: rows [src_start[v], + len[v]) of src -> rows [dst_start[v], + len[v]) of dst
// for every video v (blockIdx.y), ROWS rows per block, 16-byte (features) or 4-byte (scores) elements.
// desc = src_start[n] | dst_start[n] | len[n].
template <typename T>
__global__ void copy_video_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, const int32_t* __restrict__ desc,
                                       int n, int row_elems, int rows_per_block) {
    const int v = blockIdx.y;
    const int len = desc[2 * n + v];
    const int r0 = blockIdx.x * rows_per_block;
    if (r0 >= len) return;
    const int rows = min(rows_per_block, len - r0);
    const T* s = src + (static_cast<size_t>(desc[v]) + r0) * row_elems;
    T* d = dst + (static_cast<size_t>(desc[n + v]) + r0) * row_elems;
    const size_t total = static_cast<size_t>(rows) * row_elems;
    for (size_t i = threadIdx.x; i < total; i += blockDim.x) d[i] = s[i];
}

__global__ void permute_gate_bias_kernel(const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                         float* __restrict__ dst) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= G4) return;
    const int cta = p >> 7, gate = p & 3, jj = (p >> 2) & 31;
    const int o = gate * HC + cta * 32 + jj;
    dst[p] = b_ih[o] + b_hh[o];
}

}  // namespace
}  // namespace avs

using namespace avs;

struct avs_model {
    int device = 0;
    int Dv = 0, Da = 0, heads = 4;
    int feat_f16 = 0;   // avs_model_set_feature_format: the caller's visual / audio buffers hold IEEE fp16, not fp32
    // fp16 range watch: one word of MAPPED pinned host memory that the fc GEMM epilogues set when an activation
    // reaches the fp16 limit (the cast saturates silently).  Host-space calls read it after their final stream
    // synchronisation (no copy); device-space callers ask avs_model_range_status after synchronising themselves.
    unsigned int* range_flag = nullptr;       // host view
    unsigned int* range_flag_dev = nullptr;   // device view of the same word
    // packed parameters (one slab); *_x = exact fp32, *_t = tf32-rounded copies
    char* slab = nullptr;
    float *fc_v_w_x, *fc_v_w_t, *fc_a_w_x, *fc_a_w_t, *fc_v_b, *fc_a_b;
    float *ih_v_x, *ih_v_t, *ih_a_x, *ih_a_t;  // [2048, 512] packed gate order, both directions
    float *ih_v_b, *ih_a_b;                    // [2048]
    float* whh;                                // [4][1024][256] packed gate order, exact fp32
    float *in_w_x, *in_w_t, *in_b, *out_w_x, *out_w_t, *out_b;
    float *sc0_w_x, *sc0_w_t, *sc0_b, *sc2_w, *sc2_b;
    // 16-bit copies of the GEMM weights, [0] = fp16 (AVS_PREC_TF32 mode, everything behind the fc layers),
    // [1] = bf16 (AVS_PREC_BF16 mode); same packed row order as the fp32 tensors
    uint16_t *fc_v_w_l[2], *fc_a_w_l[2], *ih_v_l[2], *ih_a_l[2], *in_w_l[2], *out_w_l[2], *sc0_w_l[2];
    Arena ws;       // activations
    Arena staging;  // raw weights during packing
    Arena pack_ws;  // padded / sparse batches: the valid rows packed densely (features in, scores out)
    // device copies of host-space inputs when a call is pipelined by video group, and the pooling / knapsack
    // workspace; two slots so that avs_forward_summarize_async can stream batch i+1 while batch i finishes
    static constexpr int SLOTS = 2;
    Arena host_in[SLOTS];
    // pooling / knapsack workspaces: [0], [1] belong to the asynchronous slots, [2], [3] alternate between the
    // other calls.  `done` marks the end of an arena's last user, so the descriptor uploads of the next user can
    // run on the side stream while the caller's stream is still busy with the forward.
    struct SumArena {
        Arena a;
        cudaEvent_t done = nullptr;
        bool used = false;
    };
    SumArena sum_ws[SLOTS + 2];
    int sum_flip = 0;
    cudaStream_t sum_stream = nullptr;  // side stream of the summary descriptor uploads
    cudaEvent_t ev_sum_out = nullptr;
    cudaEvent_t ev_slot[SLOTS] = {};   // recorded at the end of an asynchronous step
    bool slot_busy[SLOTS] = {};
    Arena ws_grp[5];                 // activations of groups 1..5 (group 0 uses ws): the groups run concurrently
    cudaStream_t grp_stream[5] = {}; // compute streams of groups 1..5 (group 0 runs on the caller's stream)
    cudaEvent_t ev_grp[5] = {};
    // host-space calls: H2D copies run on their own stream, chunked, so that the row-parallel
    // front of the pipeline (tf32 rounding, fc and LSTM-input GEMMs) overlaps the PCIe transfer
    static constexpr int MAX_CHUNKS = 6;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_start = nullptr;
    cudaEvent_t ev_chunk[MAX_CHUNKS] = {};
    // the audio branch (audio_fc -> audio LSTM input projection) of group g runs beside the visual branch
    cudaStream_t branch_stream[MAX_CHUNKS] = {};
    cudaEvent_t ev_branch_in[MAX_CHUNKS] = {}, ev_branch_out[MAX_CHUNKS] = {};
    // pipelined front: the recurrence of video group k starts (on its own high-priority stream) as soon as the input
    // projections of its rows exist, while the GEMMs of the following groups still run
    static constexpr int PIPE_SEGS = 8;
    cudaStream_t pipe_stream[PIPE_SEGS] = {};
    cudaEvent_t ev_pipe_front[PIPE_SEGS] = {}, ev_pipe_done[PIPE_SEGS] = {}, ev_pipe_rec[PIPE_SEGS] = {};
};

namespace {

avs_status check_dims(const avs_weights* w) {
    AVS_CHECK(w != nullptr, AVS_ERR_INVALID, "weights pointer is null");
    AVS_CHECK(w->hidden_dim == H, AVS_ERR_UNSUPPORTED,
              "hidden_dim=%d: the sm_100a kernels are built for hidden_dim=512 (reference default)", w->hidden_dim);
    AVS_CHECK(w->visual_dim > 0 && w->audio_dim > 0, AVS_ERR_INVALID, "feature dims must be positive");
    AVS_CHECK(w->visual_dim % 4 == 0 && w->audio_dim % 4 == 0, AVS_ERR_UNSUPPORTED,
              "visual_dim/audio_dim must be multiples of 4 (16-byte TMA row pitch); got %d/%d", w->visual_dim,
              w->audio_dim);
    AVS_CHECK(w->num_heads > 0 && E % w->num_heads == 0 && (E / w->num_heads) % 32 == 0 &&
                  E / w->num_heads <= 256,
              AVS_ERR_UNSUPPORTED, "num_heads=%d unsupported (head dim must be a multiple of 32, <= 256)",
              w->num_heads);
    const void* ptrs[] = {w->visual_fc_w, w->visual_fc_b, w->audio_fc_w, w->audio_fc_b, w->attn_in_w, w->attn_in_b,
                          w->attn_out_w,  w->attn_out_b,  w->scorer0_w,  w->scorer0_b,  w->scorer2_w, w->scorer2_b};
    for (const void* p : ptrs) AVS_CHECK(p != nullptr, AVS_ERR_INVALID, "a weight pointer is null");
    for (int i = 0; i < 4; ++i)
        AVS_CHECK(w->lstm_w_ih[i] && w->lstm_w_hh[i] && w->lstm_b_ih[i] && w->lstm_b_hh[i], AVS_ERR_INVALID,
                  "an LSTM weight pointer is null");
    return AVS_OK;
}

// lstm_only: just the four recurrences' tensors (what a training step changes and reads through the handle);
// sync: wait for the packing kernels (model creation) -- otherwise the caller's stream orders them
avs_status pack_weights(avs_model* m, const avs_weights* w, cudaStream_t st, bool lstm_only, bool sync) {
    const int Dv = m->Dv, Da = m->Da;
    // raw copies
    const size_t raw_elems = static_cast<size_t>(H) * (Dv + Da) + 2 * H + 4 * (static_cast<size_t>(G4) * H + G4 * HC + 2 * G4) +
                             3ull * E * E + 3 * E + static_cast<size_t>(E) * E + E + 64ull * E + 64 + 64 + 1;
    AVS_TRY(m->staging.reserve(raw_elems * sizeof(float) + 64 * 256));
    m->staging.reset();
    auto up = [&](const float* src, size_t n, float** dst) -> avs_status {
        *dst = m->staging.take<float>(n);
        AVS_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyDefault, st));
        return AVS_OK;
    };
    float *r_fcv = nullptr, *r_fca = nullptr, *r_ih[4], *r_hh[4], *r_bih[4], *r_bhh[4], *r_inw = nullptr,
          *r_outw = nullptr, *r_sc0 = nullptr;
    if (!lstm_only) {
        AVS_TRY(up(w->visual_fc_w, static_cast<size_t>(H) * Dv, &r_fcv));
        AVS_TRY(up(w->audio_fc_w, static_cast<size_t>(H) * Da, &r_fca));
    }
    for (int i = 0; i < 4; ++i) {
        AVS_TRY(up(w->lstm_w_ih[i], static_cast<size_t>(G4) * H, &r_ih[i]));
        AVS_TRY(up(w->lstm_w_hh[i], static_cast<size_t>(G4) * HC, &r_hh[i]));
        AVS_TRY(up(w->lstm_b_ih[i], G4, &r_bih[i]));
        AVS_TRY(up(w->lstm_b_hh[i], G4, &r_bhh[i]));
    }
    if (!lstm_only) {
        AVS_TRY(up(w->attn_in_w, 3ull * E * E, &r_inw));
        AVS_TRY(up(w->attn_out_w, static_cast<size_t>(E) * E, &r_outw));
        AVS_TRY(up(w->scorer0_w, 64ull * E, &r_sc0));
        // direct (unpacked) small tensors
        AVS_CUDA(cudaMemcpyAsync(m->fc_v_b, w->visual_fc_b, H * sizeof(float), cudaMemcpyDefault, st));
        AVS_CUDA(cudaMemcpyAsync(m->fc_a_b, w->audio_fc_b, H * sizeof(float), cudaMemcpyDefault, st));
        AVS_CUDA(cudaMemcpyAsync(m->in_b, w->attn_in_b, 3 * E * sizeof(float), cudaMemcpyDefault, st));
        AVS_CUDA(cudaMemcpyAsync(m->out_b, w->attn_out_b, E * sizeof(float), cudaMemcpyDefault, st));
        AVS_CUDA(cudaMemcpyAsync(m->sc0_b, w->scorer0_b, 64 * sizeof(float), cudaMemcpyDefault, st));
        AVS_CUDA(cudaMemcpyAsync(m->sc2_w, w->scorer2_w, 64 * sizeof(float), cudaMemcpyDefault, st));
        AVS_CUDA(cudaMemcpyAsync(m->sc2_b, w->scorer2_b, sizeof(float), cudaMemcpyDefault, st));
        // exact + tf32 copies of the GEMM weights
        auto both = [&](const float* raw, size_t n, float* exact, float* tf) -> avs_status {
            AVS_CUDA(cudaMemcpyAsync(exact, raw, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
            return convert_f32(raw, tf, static_cast<int64_t>(n), DT_F32, 1, st);
        };
        AVS_TRY(both(r_fcv, static_cast<size_t>(H) * Dv, m->fc_v_w_x, m->fc_v_w_t));
        AVS_TRY(both(r_fca, static_cast<size_t>(H) * Da, m->fc_a_w_x, m->fc_a_w_t));
        AVS_TRY(both(r_inw, 3ull * E * E, m->in_w_x, m->in_w_t));
        AVS_TRY(both(r_outw, static_cast<size_t>(E) * E, m->out_w_x, m->out_w_t));
        AVS_TRY(both(r_sc0, 64ull * E, m->sc0_w_x, m->sc0_w_t));
        for (int l = 0; l < 2; ++l) {
            const int dt = l ? DT_BF16 : DT_F16;
            AVS_TRY(convert_f32(r_fcv, m->fc_v_w_l[l], static_cast<int64_t>(H) * Dv, dt, 0, st));
            AVS_TRY(convert_f32(r_fca, m->fc_a_w_l[l], static_cast<int64_t>(H) * Da, dt, 0, st));
            AVS_TRY(convert_f32(r_inw, m->in_w_l[l], 3ll * E * E, dt, 0, st));
            AVS_TRY(convert_f32(r_outw, m->out_w_l[l], static_cast<int64_t>(E) * E, dt, 0, st));
            AVS_TRY(convert_f32(r_sc0, m->sc0_w_l[l], 64ll * E, dt, 0, st));
        }
    }
    // LSTM: gate-interleaved row order so that cluster CTA r owns 128 contiguous gate columns
    for (int i = 0; i < 4; ++i) {
        const int mod = i >> 1, dir = i & 1;
        float* ih_x = (mod ? m->ih_a_x : m->ih_v_x) + static_cast<size_t>(dir) * G4 * H;
        float* ih_t = (mod ? m->ih_a_t : m->ih_v_t) + static_cast<size_t>(dir) * G4 * H;
        float* bias = (mod ? m->ih_a_b : m->ih_v_b) + dir * G4;
        permute_gate_rows_kernel<<<G4, 128, 0, st>>>(r_ih[i], ih_x, H, 0);
        AVS_LAUNCH_CHECK();
        permute_gate_rows_kernel<<<G4, 128, 0, st>>>(r_ih[i], ih_t, H, 1);
        AVS_LAUNCH_CHECK();
        permute_gate_rows_kernel<<<G4, 128, 0, st>>>(r_hh[i], m->whh + static_cast<size_t>(i) * G4 * HC, HC, 0);
        AVS_LAUNCH_CHECK();
        permute_gate_bias_kernel<<<(G4 + 255) / 256, 256, 0, st>>>(r_bih[i], r_bhh[i], bias);
        AVS_LAUNCH_CHECK();
    }
    for (int l = 0; l < 2; ++l) {
        const int dt = l ? DT_BF16 : DT_F16;
        AVS_TRY(convert_f32(m->ih_v_x, m->ih_v_l[l], 2ll * G4 * H, dt, 0, st));
        AVS_TRY(convert_f32(m->ih_a_x, m->ih_a_l[l], 2ll * G4 * H, dt, 0, st));
    }
    if (sync) AVS_CUDA(cudaStreamSynchronize(st));
    return AVS_OK;
}

avs_status upload_small(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st) {
    AVS_CHECK(bytes % 4 == 0, AVS_ERR_INVALID, "upload_small: size must be a multiple of 4");
    const uint32_t* src = static_cast<const uint32_t*>(src_host);
    uint32_t* dst = static_cast<uint32_t*>(dst_dev);
    size_t words = bytes / 4;
    while (words > 0) {
        UploadPayload p;
        const int n = static_cast<int>(std::min<size_t>(words, 896));
        std::memcpy(p.w, src, static_cast<size_t>(n) * 4);
        upload_kernel<<<1, 256, 0, st>>>(p, dst, n);
        AVS_LAUNCH_CHECK();
        src += n;
        dst += n;
        words -= n;
    }
    return AVS_OK;
}

struct Guard {  // restore the caller's current device
    int prev = -1;
    explicit Guard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~Guard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

avs_status validate_videos(int64_t total_rows, int32_t n, const int32_t* row_start, const int32_t* lengths,
                           int* max_len) {
    AVS_CHECK(total_rows >= 0 && total_rows < (1ll << 31), AVS_ERR_INVALID, "total_rows out of range");
    AVS_CHECK(n >= 0, AVS_ERR_INVALID, "n_videos negative");
    AVS_CHECK(n == 0 || (row_start && lengths), AVS_ERR_INVALID, "row_start / lengths null");
    int mx = 0;
    for (int b = 0; b < n; ++b) {
        AVS_CHECK(lengths[b] >= 0 && row_start[b] >= 0 &&
                      static_cast<int64_t>(row_start[b]) + lengths[b] <= total_rows,
                  AVS_ERR_INVALID, "video %d: rows [%d, %d) outside [0, %lld)", b, row_start[b],
                  row_start[b] + lengths[b], static_cast<long long>(total_rows));
        mx = std::max(mx, lengths[b]);
    }
    *max_len = mx;
    return AVS_OK;
}

// Length-sorted grouping of videos into LSTM clusters.
struct LstmPlan {
    std::vector<int32_t> host;  // [slot_row_start | slot_len | group_maxlen]
    std::vector<int32_t> video; // [slot] -> index of the video in the caller's descriptors, -1 = empty slot
    int nb = 1, n_groups = 0;
};
LstmPlan plan_lstm(int32_t n, const int32_t* row_start, const int32_t* lengths, bool tensor_core) {
    std::vector<int> order;
    for (int b = 0; b < n; ++b)
        if (lengths[b] > 0) order.push_back(b);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lengths[a] > lengths[b]; });
    LstmPlan p;
    const int B = static_cast<int>(order.size());
    if (B == 0) return p;
    // one wave = 16 clusters of 8 CTAs (4 recurrences x 4 groups) with one CTA per SM; the tensor-core kernel
    // also has an 8-slot variant whose CTAs are small enough for two per SM (32 clusters per wave), so that
    // one cluster's DSMEM exchange overlaps another's MMA + cell math on the same SM
    if (tensor_core) {
        p.nb = 64;
        for (int nb : {8, 16, 32, 64}) {
            if (((B + nb - 1) / nb) * 4 <= (nb == 8 ? 32 : 16)) { p.nb = nb; break; }
        }
    } else {
        p.nb = 16;
        for (int nb : {1, 2, 4, 8, 16}) {
            if (((B + nb - 1) / nb) * 4 <= 16) { p.nb = nb; break; }
        }
    }
    // The 8-slot kernel runs two independent chains of 4 videos per CTA; an isolated chain is faster (0.67 vs 0.76
    // us per step, tools/lstm_scaling.py).  A small batch therefore takes 4 videos per cluster -- the second chain of
    // every CTA stays empty -- as long as all clusters still get SMs of their own (16 videos -> 16 clusters).
    // (8 videos x T = 8192, B200: 2 per cluster 4.95 ms, 4 per cluster 5.09 ms, 1 per cluster -- two CTAs per SM -- 5.18 ms)
    int per_cluster = (tensor_core && p.nb == 8 && B <= 16) ? (B <= 8 ? 2 : 4) : p.nb;
    if (tensor_core && p.nb == 8) {   // tuning aid: videos per cluster for small batches (1, 2, 4 or 8)
        if (const char* e = getenv("AVS_LSTM_PER_CLUSTER")) {
            const int want = atoi(e);
            if ((want == 1 || want == 2 || want == 4 || want == 8) && ((B + want - 1) / want) * 4 <= 32) per_cluster = want;
        }
    }
    p.n_groups = (B + per_cluster - 1) / per_cluster;
    const int slots = p.n_groups * p.nb;
    p.host.assign(2 * slots + p.n_groups, 0);
    p.video.assign(slots, -1);
    // Uneven split: the extra videos go to the SHORTEST groups.  The longest group is the critical chain and the last
    // to finish; the row-parallel work behind its recurrence (pipelined tail) is what the step still has to wait for,
    // so it gets the fewest rows.  (AVS_PLAN_REM_FIRST=1: extras to the longest groups, the round-1 split; A/B aid.)
    static const bool rem_first = getenv("AVS_PLAN_REM_FIRST") != nullptr;
    const int base = B / p.n_groups, rem = B % p.n_groups;
    int idx = 0;
    for (int g = 0; g < p.n_groups; ++g) {
        const int cnt = base + ((rem_first ? g < rem : g >= p.n_groups - rem) ? 1 : 0);
        p.host[2 * slots + g] = lengths[order[idx]];
        for (int i = 0; i < cnt; ++i, ++idx) {
            p.host[g * p.nb + i] = row_start[order[idx]];
            p.host[slots + g * p.nb + i] = lengths[order[idx]];
            p.video[g * p.nb + i] = order[idx];
        }
    }
    return p;
}

// Row range [lo, hi) of every recurrence group when the groups tile the batch's rows in order (batches ordered longest
// video first, the order packed_batches produces): group g then owns a contiguous block of rows and row-parallel work
// can be issued per group.  False when a group's rows are not contiguous or the groups do not follow each other.
bool group_row_ranges(const LstmPlan& plan, int64_t R, std::vector<int64_t>& glo, std::vector<int64_t>& ghi) {
    const int G = plan.n_groups, nbp = plan.nb, slots = G * nbp;
    glo.assign(G, 0);
    ghi.assign(G, 0);
    for (int gq = 0; gq < G; ++gq) {
        int64_t lo = INT64_MAX, hi = 0, sum = 0;
        for (int i = 0; i < nbp; ++i) {
            const int64_t len = plan.host[slots + gq * nbp + i], rs0 = plan.host[gq * nbp + i];
            if (len > 0) {
                lo = std::min(lo, rs0);
                hi = std::max(hi, rs0 + len);
                sum += len;
            }
        }
        if (!(sum > 0 && sum == hi - lo && lo == (gq ? ghi[gq - 1] : 0))) return false;
        glo[gq] = lo;
        ghi[gq] = hi;
    }
    return G > 0 && ghi[G - 1] == R;
}

// One weight matrix in every operand format the kernels consume.
struct GemmW {
    const float* x;              // exact fp32 (CUDA-core path)
    const float* t;              // tf32-rounded fp32
    uint16_t* const* l;          // [0] fp16, [1] bf16
};

// C = epilogue(A * W[off : off + N*ldw]^T).  The operand format follows the dtype of A:
// fp32 -> kind::tf32 (or the CUDA-core path in AVS_PREC_FP32_SIMT), fp16 / bf16 -> kind::f16.
avs_status run_gemm(int precision, const void* A, int a_dtype, int64_t lda, const GemmW& w, int64_t w_off, int64_t ldw,
                    int64_t M, int N, int K, GemmEpilogue epi, cudaStream_t st) {
    if (precision == AVS_PREC_FP32_SIMT) {
        AVS_CHECK(a_dtype == DT_F32, AVS_ERR_INVALID, "fp32_simt GEMM needs fp32 operands");
        epi.round_tf32 = 0;
        return gemm_simt(static_cast<const float*>(A), lda, w.x + w_off, ldw, M, N, K, epi, st);
    }
    if (a_dtype == DT_F32) return gemm_tc(A, lda, w.t + w_off, ldw, DT_F32, M, N, K, epi, st);
    return gemm_tc(A, lda, w.l[a_dtype == DT_BF16 ? 1 : 0] + w_off, ldw, a_dtype, M, N, K, epi, st);
}

bool precision_ok(int precision) {
    return precision == AVS_PREC_TF32 || precision == AVS_PREC_BF16 || precision == AVS_PREC_FP32_SIMT;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

const char* avs_last_error(void) { return g_err; }
int avs_version(void) { return 100; }
int64_t avs_launch_count(void) { return g_launches.load(); }

void avs_profile(int enable) {
    g_prof.resolve();
    g_prof.enabled = enable != 0;
    if (enable == 2) {  // reset accumulators
        for (int i = 0; i < ST_COUNT; ++i) { g_prof.ms[i] = 0; g_prof.calls[i] = 0; }
        g_prof.enabled = true;
    }
}
int avs_profile_stages(void) { return ST_COUNT; }
const char* avs_profile_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }
void avs_profile_read(double* ms, int64_t* calls) {
    g_prof.resolve();
    for (int i = 0; i < ST_COUNT; ++i) {
        if (ms) ms[i] = g_prof.ms[i];
        if (calls) calls[i] = g_prof.calls[i];
    }
}

/* Debugging aid (not part of the reference surface): with AVS_LSTM_TRACE=1 in the environment the NB=16
 * recurrence kernel accumulates clock64 deltas of its per-step dependency chain on cluster 0 / CTA 0:
 * [0] h landed -> MMAs issued, [1] -> epilogue awake, [2] tcgen05.ld, [3] cell math + staging,
 * [4] fence + barrier, [5] bulk-copy issue, [6] copies issued -> next h landed, [7] steps. */
avs_status avs_debug_lstm_trace(uint64_t* out8) {
    AVS_CHECK(out8 != nullptr, AVS_ERR_INVALID, "null pointer");
    return lstm_trace_read(reinterpret_cast<unsigned long long*>(out8));
}

/* Host logic only (no GPU work, callable without a device): the recurrence plan avs_forward derives from a batch's
 * descriptors.  group_of[n_videos] = recurrence group of every video (-1: empty video); info[0] = groups, info[1] =
 * slots per group, info[2] = 1 when the groups tile the rows in order (what the per-group schedule needs: rows laid out
 * longest video first), info[3] = 1 when the groups end at different times (shortest group's longest video <= 0.9 x
 * the longest group's); group_rows[2 g], [2 g + 1] = row range of group g when info[2] (max_groups entries at most). */
avs_status avs_debug_plan(int32_t n_videos, const int32_t* row_start, const int32_t* lengths, int64_t total_rows,
                          int32_t* group_of, int32_t* info, int64_t* group_rows, int32_t max_groups) {
    AVS_CHECK(info != nullptr && (n_videos == 0 || group_of != nullptr), AVS_ERR_INVALID, "null pointer");
    int max_len = 0;
    AVS_TRY(validate_videos(total_rows, n_videos, row_start, lengths, &max_len));
    const LstmPlan plan = plan_lstm(n_videos, row_start, lengths, true);
    for (int b = 0; b < n_videos; ++b) group_of[b] = -1;
    for (int g = 0; g < plan.n_groups; ++g)
        for (int i = 0; i < plan.nb; ++i)
            if (plan.video[g * plan.nb + i] >= 0) group_of[plan.video[g * plan.nb + i]] = g;
    std::vector<int64_t> glo, ghi;
    const bool ordered = plan.n_groups > 0 && group_row_ranges(plan, total_rows, glo, ghi);
    const int slots = plan.n_groups * plan.nb;
    info[0] = plan.n_groups;
    info[1] = plan.nb;
    info[2] = ordered ? 1 : 0;
    info[3] = (plan.n_groups > 0 && plan.host[2 * slots + plan.n_groups - 1] * 10ll <= plan.host[2 * slots] * 9ll) ? 1 : 0;
    if (ordered && group_rows)
        for (int g = 0; g < plan.n_groups && g < max_groups; ++g) {
            group_rows[2 * g] = glo[g];
            group_rows[2 * g + 1] = ghi[g];
        }
    return AVS_OK;
}

/* Debugging aid: with AVS_BPTT_TRACE=1 the tensor-core BPTT kernel accumulates clock64 deltas of its per-step chain on
 * cluster 0 / CTA 0: [0] partials landed -> B operand staged, [1] 32 MMAs + commit issued, [2] -> epilogue awake,
 * [3] tcgen05.ld, [4] shuffles + st.async issue, [6] partials sent -> next partials landed,
 * [7] partials sent -> dh-independent math of the next step done, [8] steps. */
avs_status avs_debug_bptt_trace(uint64_t* out10) {
    AVS_CHECK(out10 != nullptr, AVS_ERR_INVALID, "null pointer");
    return bptt_trace_read(reinterpret_cast<unsigned long long*>(out10));
}

/* Debugging aid: clock64 totals of the tensor-core GEMM pipeline on block 0 since the last read (AVS_GEMM_TRACE=1):
 * [0] MMA thread span, [1] MMA waiting for operands, [2] MMA waiting for a drained accumulator, [3] producer waiting
 * for a free stage, [4] epilogue warp waiting for an accumulator, [5] epilogue span, [6] tiles. */
avs_status avs_debug_gemm_trace(uint64_t* out8) {
    AVS_CHECK(out8 != nullptr, AVS_ERR_INVALID, "null pointer");
    return gemm_trace_read(reinterpret_cast<unsigned long long*>(out8));
}

/* Debugging aid: timeline of the last pipelined host-space avs_forward_summarize call (AVS_E2E_TRACE=1).
 * out[0] = number of video groups G; out[1..5] = host clock at entry / copies queued / groups queued / tail queued /
 * synchronised (ms after entry); then, in ms after the first device timestamp: out[6..6+G) = group g's features
 * landed, out[12..12+G) = group g's forward finished, out[18] = knapsack finished, out[19] = last D2H finished. */
avs_status avs_debug_e2e_trace(double* out20) {
    AVS_CHECK(out20 != nullptr, AVS_ERR_INVALID, "null pointer");
    AVS_CHECK(g_e2e.on == 1 && g_e2e.n_groups > 0, AVS_ERR_INVALID, "no trace recorded (AVS_E2E_TRACE=1, pipelined call)");
    for (int i = 0; i < 20; ++i) out20[i] = 0.0;
    out20[0] = g_e2e.n_groups;
    for (int i = 0; i < 5; ++i) out20[1 + i] = (g_e2e.host[i] - g_e2e.host[0]) * 1e3;
    float t = 0.f;
    for (int g = 0; g < g_e2e.n_groups; ++g) {
        if (cudaEventElapsedTime(&t, g_e2e.t0, g_e2e.chunk[g]) == cudaSuccess) out20[6 + g] = t;
        if (cudaEventElapsedTime(&t, g_e2e.t0, g_e2e.grp[g]) == cudaSuccess) out20[12 + g] = t;
    }
    if (cudaEventElapsedTime(&t, g_e2e.t0, g_e2e.knap) == cudaSuccess) out20[18] = t;
    if (cudaEventElapsedTime(&t, g_e2e.t0, g_e2e.end) == cudaSuccess) out20[19] = t;
    cudaGetLastError();
    return AVS_OK;
}

/* Page-locked host memory for the batches a caller feeds to the host-space entry points.  write_combined != 0 asks
 * for write-combined pages: the CPU should only WRITE them (reads are uncached and slow), and device reads across
 * PCIe do not snoop the CPU caches. */
avs_status avs_host_alloc(void** out, size_t bytes, int write_combined) {
    AVS_CHECK(out != nullptr, AVS_ERR_INVALID, "null pointer");
    *out = nullptr;
    if (bytes == 0) return AVS_OK;
    AVS_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
    return AVS_OK;
}
avs_status avs_host_free(void* p) {
    if (p != nullptr) AVS_CUDA(cudaFreeHost(p));
    return AVS_OK;
}

int avs_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return 0;
    }
    cudaDeviceProp p;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
    return p.major == 10 ? 1 : 0;
}

avs_status avs_model_create(const avs_weights* w, int device, avs_model** out) {
    AVS_CHECK(out != nullptr, AVS_ERR_INVALID, "out pointer is null");
    *out = nullptr;
    AVS_TRY(check_dims(w));
    int n = 0;
    AVS_CUDA(cudaGetDeviceCount(&n));
    AVS_CHECK(device >= 0 && device < n, AVS_ERR_CUDA, "CUDA device %d not available (%d visible)", device, n);
    cudaDeviceProp prop;
    AVS_CUDA(cudaGetDeviceProperties(&prop, device));
    AVS_CHECK(prop.major == 10, AVS_ERR_CUDA, "device %d is sm_%d%d; this library contains sm_100a code only",
              device, prop.major, prop.minor);
    Guard g(device);
    avs_model* m = new avs_model();
    m->device = device;
    m->Dv = w->visual_dim;
    m->Da = w->audio_dim;
    m->heads = w->num_heads;
    const size_t Dv = m->Dv, Da = m->Da;
    struct Item { void** p; size_t bytes; };
    std::vector<Item> items;
    auto f32 = [&](float** p, size_t n) { items.push_back({reinterpret_cast<void**>(p), n * sizeof(float)}); };
    auto l16 = [&](uint16_t** p, size_t n) { items.push_back({reinterpret_cast<void**>(p), n * 2}); };
    f32(&m->fc_v_w_x, H * Dv); f32(&m->fc_v_w_t, H * Dv); f32(&m->fc_a_w_x, H * Da); f32(&m->fc_a_w_t, H * Da);
    f32(&m->fc_v_b, H); f32(&m->fc_a_b, H);
    f32(&m->ih_v_x, 2ull * G4 * H); f32(&m->ih_v_t, 2ull * G4 * H);
    f32(&m->ih_a_x, 2ull * G4 * H); f32(&m->ih_a_t, 2ull * G4 * H);
    f32(&m->ih_v_b, 2 * G4); f32(&m->ih_a_b, 2 * G4); f32(&m->whh, 4ull * G4 * HC);
    f32(&m->in_w_x, 3ull * E * E); f32(&m->in_w_t, 3ull * E * E); f32(&m->in_b, 3 * E);
    f32(&m->out_w_x, 1ull * E * E); f32(&m->out_w_t, 1ull * E * E); f32(&m->out_b, E);
    f32(&m->sc0_w_x, 64ull * E); f32(&m->sc0_w_t, 64ull * E); f32(&m->sc0_b, 64); f32(&m->sc2_w, 64); f32(&m->sc2_b, 64);
    for (int l = 0; l < 2; ++l) {
        l16(&m->fc_v_w_l[l], H * Dv); l16(&m->fc_a_w_l[l], H * Da);
        l16(&m->ih_v_l[l], 2ull * G4 * H); l16(&m->ih_a_l[l], 2ull * G4 * H);
        l16(&m->in_w_l[l], 3ull * E * E); l16(&m->out_w_l[l], 1ull * E * E); l16(&m->sc0_w_l[l], 64ull * E);
    }
    size_t total = 0;
    for (auto& it : items) total += align_up(it.bytes, 256);
    cudaError_t e = cudaMalloc(&m->slab, total);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu) for packed weights failed: %s", total, cudaGetErrorString(e));
        delete m;
        return AVS_ERR_OOM;
    }
    size_t off = 0;
    for (auto& it : items) {
        *it.p = m->slab + off;
        off += align_up(it.bytes, 256);
    }
    if (cudaHostAlloc(reinterpret_cast<void**>(&m->range_flag), sizeof(unsigned int), cudaHostAllocMapped) == cudaSuccess)
        *m->range_flag = 0u;
    else {
        m->range_flag = nullptr;
    }
    if (m->range_flag == nullptr ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&m->range_flag_dev), m->range_flag, 0) != cudaSuccess) {
        if (m->range_flag) cudaFreeHost(m->range_flag);
        m->range_flag = m->range_flag_dev = nullptr;   // the watch is an aid: without the page the forward does not report
        cudaGetLastError();
    }
    // the packing kernels run on the handle's own stream: the caller guarantees that the parameters are complete
    // (include/avsum_b200.h), nothing else is ordered against the legacy default stream
    avs_status s = AVS_OK;
    if (cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("creating the copy stream failed");
        s = AVS_ERR_CUDA;
    }
    if (s == AVS_OK) s = pack_weights(m, w, m->copy_stream, false, true);
    if (s == AVS_OK) {
        cudaError_t ce = cudaEventCreateWithFlags(&m->ev_start, cudaEventDisableTiming);
        for (int i = 0; i < avs_model::MAX_CHUNKS && ce == cudaSuccess; ++i)
            ce = cudaEventCreateWithFlags(&m->ev_chunk[i], cudaEventDisableTiming);
        for (int i = 0; i < 5 && ce == cudaSuccess; ++i) {
            ce = cudaStreamCreateWithFlags(&m->grp_stream[i], cudaStreamNonBlocking);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&m->ev_grp[i], cudaEventDisableTiming);
        }
        for (int i = 0; i < avs_model::SLOTS && ce == cudaSuccess; ++i)
            ce = cudaEventCreateWithFlags(&m->ev_slot[i], cudaEventDisableTiming);
        for (int i = 0; i < avs_model::MAX_CHUNKS && ce == cudaSuccess; ++i) {
            ce = cudaStreamCreateWithFlags(&m->branch_stream[i], cudaStreamNonBlocking);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&m->ev_branch_in[i], cudaEventDisableTiming);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&m->ev_branch_out[i], cudaEventDisableTiming);
        }
        int prio_lo = 0, prio_hi = 0;
        if (ce == cudaSuccess) ce = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        for (int i = 0; i < avs_model::PIPE_SEGS && ce == cudaSuccess; ++i) {
            ce = cudaStreamCreateWithPriority(&m->pipe_stream[i], cudaStreamNonBlocking, prio_hi);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&m->ev_pipe_front[i], cudaEventDisableTiming);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&m->ev_pipe_done[i], cudaEventDisableTiming);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&m->ev_pipe_rec[i], cudaEventDisableTiming);
        }
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&m->sum_stream, cudaStreamNonBlocking);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&m->ev_sum_out, cudaEventDisableTiming);
        for (int i = 0; i < avs_model::SLOTS + 2 && ce == cudaSuccess; ++i)
            ce = cudaEventCreateWithFlags(&m->sum_ws[i].done, cudaEventDisableTiming);
        if (ce != cudaSuccess) {
            set_error("creating the copy stream / events failed: %s", cudaGetErrorString(ce));
            s = AVS_ERR_CUDA;
        }
    }
    if (s != AVS_OK) {
        avs_model_destroy(m);
        return s;
    }
    *out = m;
    return AVS_OK;
}

avs_status avs_model_update(avs_model* m, const avs_weights* w) {
    AVS_CHECK(m != nullptr, AVS_ERR_INVALID, "model handle is null");
    AVS_TRY(check_dims(w));
    AVS_CHECK(w->visual_dim == m->Dv && w->audio_dim == m->Da, AVS_ERR_INVALID,
              "avs_model_update: feature dims differ from the handle's");
    m->heads = w->num_heads;
    Guard g(m->device);
    // synchronous form: everything queued on this device (earlier asynchronous packs share the staging arena, the
    // writer of the parameters may be any stream) finishes first, then the pack runs on the handle's own stream
    AVS_CUDA(cudaDeviceSynchronize());
    return pack_weights(m, w, m->copy_stream, false, true);
}

avs_status avs_model_update_async(avs_model* m, const avs_weights* w, int lstm_only, void* cuda_stream) {
    AVS_CHECK(m != nullptr, AVS_ERR_INVALID, "model handle is null");
    AVS_TRY(check_dims(w));
    AVS_CHECK(w->visual_dim == m->Dv && w->audio_dim == m->Da, AVS_ERR_INVALID,
              "avs_model_update_async: feature dims differ from the handle's");
    m->heads = w->num_heads;
    Guard g(m->device);
    return pack_weights(m, w, static_cast<cudaStream_t>(cuda_stream), lstm_only != 0, false);
}

avs_status avs_model_set_feature_format(avs_model* m, int format) {
    AVS_CHECK(m != nullptr, AVS_ERR_INVALID, "model handle is null");
    AVS_CHECK(format == AVS_FEAT_F32 || format == AVS_FEAT_F16, AVS_ERR_INVALID, "bad feature format %d", format);
    AVS_CHECK(format == AVS_FEAT_F32 || (m->Dv % 8 == 0 && m->Da % 8 == 0), AVS_ERR_UNSUPPORTED,
              "fp16 features need visual_dim / audio_dim that are multiples of 8 (16-byte TMA row pitch); got %d / %d",
              m->Dv, m->Da);
    for (int i = 0; i < avs_model::SLOTS; ++i)
        AVS_CHECK(!m->slot_busy[i], AVS_ERR_INVALID, "slot %d is in flight: call avs_slot_wait first", i);
    m->feat_f16 = format == AVS_FEAT_F16;
    return AVS_OK;
}

void avs_model_destroy(avs_model* m) {
    if (!m) return;
    Guard g(m->device);
    cudaDeviceSynchronize();
    m->ws.release();
    m->staging.release();
    m->pack_ws.release();
    if (m->range_flag) cudaFreeHost(m->range_flag);
    for (int i = 0; i < avs_model::SLOTS; ++i) {
        m->host_in[i].release();
        if (m->ev_slot[i]) cudaEventDestroy(m->ev_slot[i]);
    }
    for (int i = 0; i < 5; ++i) {
        m->ws_grp[i].release();
        if (m->grp_stream[i]) cudaStreamDestroy(m->grp_stream[i]);
        if (m->ev_grp[i]) cudaEventDestroy(m->ev_grp[i]);
    }
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    if (m->sum_stream) cudaStreamDestroy(m->sum_stream);
    for (int i = 0; i < avs_model::PIPE_SEGS; ++i) {
        if (m->pipe_stream[i]) cudaStreamDestroy(m->pipe_stream[i]);
        if (m->ev_pipe_front[i]) cudaEventDestroy(m->ev_pipe_front[i]);
        if (m->ev_pipe_done[i]) cudaEventDestroy(m->ev_pipe_done[i]);
        if (m->ev_pipe_rec[i]) cudaEventDestroy(m->ev_pipe_rec[i]);
    }
    for (int i = 0; i < avs_model::MAX_CHUNKS; ++i) {
        if (m->branch_stream[i]) cudaStreamDestroy(m->branch_stream[i]);
        if (m->ev_branch_in[i]) cudaEventDestroy(m->ev_branch_in[i]);
        if (m->ev_branch_out[i]) cudaEventDestroy(m->ev_branch_out[i]);
    }
    if (m->ev_sum_out) cudaEventDestroy(m->ev_sum_out);
    for (int i = 0; i < avs_model::SLOTS + 2; ++i) {
        m->sum_ws[i].a.release();
        if (m->sum_ws[i].done) cudaEventDestroy(m->sum_ws[i].done);
    }
    if (m->ev_start) cudaEventDestroy(m->ev_start);
    for (cudaEvent_t e : m->ev_chunk)
        if (e) cudaEventDestroy(e);
    if (m->slab) cudaFree(m->slab);
    delete m;
}

// Host side of the fp16 range watch: call after a stream synchronisation that covers the forward(s) in question.
// Sticky until reported: with two batches in flight the report may come one call early.
static avs_status range_report(avs_model* m) {
    if (m == nullptr || m->range_flag == nullptr) return AVS_OK;
    volatile unsigned int* f = m->range_flag;
    if (*f == 0u) return AVS_OK;
    *f = 0u;
    set_error("fc activations reached the fp16 range limit (|x| >= 65504) in this or a concurrently running call and "
              "were clamped there: the scores are not trustworthy.  Normalise the features or use AVS_PREC_BF16 (fp32 "
              "exponent range)");
    return AVS_ERR_UNSUPPORTED;
}

static avs_status forward_impl(avs_model* m, const float* visual, const float* audio, int64_t total_rows,
                               int32_t n_videos, const int32_t* row_start, const int32_t* lengths, int attn_axis,
                               int precision, float* scores, int space, void* cuda_stream, Arena* arena = nullptr,
                               int lstm_excl = 0);

// scores_dev_out != nullptr (host space only): leave the scores on the device (pointer returned), skip the D2H copy
// and the final synchronisation -- the caller continues on the stream (avs_forward_summarize).
static avs_status forward_entry(avs_model* m, const float* visual, const float* audio, int64_t total_rows,
                                int32_t n_videos, const int32_t* row_start, const int32_t* lengths, int attn_axis,
                                int precision, float* scores, int space, void* cuda_stream, float** scores_dev_out,
                                const int32_t* positions_host, int32_t** positions_dev_out, int slot = 0,
                                bool async = false) {
    // Host-space calls on large packed batches are pipelined by VIDEO GROUP: all H2D copies are queued on the
    // copy stream up front, and the complete forward of group g (GEMMs, recurrences, attention, score head) runs
    // on the caller's stream as soon as its rows have landed -- i.e. while group g+1 is still crossing PCIe.
    // The recurrence of a group lasts as long as its longest video, so callers that order the batch longest
    // video first (data/dataset.py packed_batches does) hide most of the compute behind the transfer.
    const bool per_video = attn_axis == AVS_ATTN_TEMPORAL || attn_axis == AVS_ATTN_LITERAL_B1;
    const bool host = space == AVS_HOST;
    // Padded / sparse layouts (BASELINE configs[2]: [B, Tmax] rows with lengths[B]): every GEMM would process the rows
    // no video owns -- 47 % of the rows of the SumMe-shaped batch.  The valid rows are packed densely first (one
    // copy per video: device-to-device, or straight from the caller's host buffer, which also saves the PCIe bytes
    // of the padding), the forward runs on the packed rows, and the scores are copied back to each video's rows.
    // Rows no video owns are not written.
    if (m != nullptr && (host || space == AVS_DEVICE) && per_video && visual && audio && scores && !scores_dev_out &&
        row_start && lengths && n_videos >= 1 && total_rows >= 256 && total_rows < (1ll << 31) &&
        getenv("AVS_NO_ROW_PACKING") == nullptr) {
        int64_t covered = 0;
        bool ok = true;
        for (int b = 0; b < n_videos && ok; ++b) {
            ok = lengths[b] >= 0 && row_start[b] >= 0 && static_cast<int64_t>(row_start[b]) + lengths[b] <= total_rows;
            covered += ok ? lengths[b] : 0;
        }
        // device space gathers with 16-byte vectors: unaligned caller buffers (a raw C-ABI caller) take the unpacked path
        const size_t fsz0 = m->feat_f16 ? 2 : 4;
        const bool vec_ok = host || ((reinterpret_cast<uintptr_t>(visual) & 15) == 0 && (reinterpret_cast<uintptr_t>(audio) & 15) == 0 &&
                                     (m->Dv * fsz0) % 16 == 0 && (m->Da * fsz0) % 16 == 0);
        if (ok && vec_ok && covered > 0 && covered * 10 < total_rows * 9) {
            AVS_CHECK(precision_ok(precision), AVS_ERR_INVALID, "bad precision %d", precision);
            Guard gp(m->device);
            cudaStream_t sp = static_cast<cudaStream_t>(cuda_stream);
            const size_t C_ = static_cast<size_t>(covered), Dv_ = m->Dv, Da_ = m->Da, fsz = m->feat_f16 ? 2 : 4;
            AVS_TRY(m->pack_ws.reserve(C_ * (Dv_ + Da_) * fsz + C_ * 4 + 6 * static_cast<size_t>(n_videos) * 4 + 8 * 256));
            m->pack_ws.reset();
            float* pv = reinterpret_cast<float*>(m->pack_ws.take<char>(C_ * Dv_ * fsz));
            float* pa = reinterpret_cast<float*>(m->pack_ws.take<char>(C_ * Da_ * fsz));
            float* ps = m->pack_ws.take<float>(C_);
            int32_t* desc_dev = m->pack_ws.take<int32_t>(3 * static_cast<size_t>(n_videos));
            std::vector<int32_t> prs(n_videos), desc(3 * static_cast<size_t>(n_videos));
            int64_t at = 0;
            int max_len = 0;
            // packed longest video first (the order of the recurrence groups): every group then owns one contiguous
            // block of packed rows, which is what the pipelined tail of the forward needs
            std::vector<int> pack_order(n_videos);
            for (int b = 0; b < n_videos; ++b) pack_order[b] = b;
            std::stable_sort(pack_order.begin(), pack_order.end(), [&](int x, int y) { return lengths[x] > lengths[y]; });
            for (int b : pack_order) {
                prs[b] = static_cast<int32_t>(at);
                desc[b] = row_start[b];
                desc[n_videos + b] = prs[b];
                desc[2 * n_videos + b] = lengths[b];
                max_len = std::max(max_len, lengths[b]);
                at += lengths[b];
            }
            if (host) {   // straight from the caller's (pinned) buffer: one DMA copy per video, padding never crosses PCIe
                for (int b = 0; b < n_videos; ++b) {
                    const size_t n = static_cast<size_t>(lengths[b]);
                    if (!n) continue;
                    AVS_CUDA(cudaMemcpyAsync(feat_at(pv, prs[b], m->Dv, fsz), feat_at(visual, row_start[b], m->Dv, fsz),
                                             n * Dv_ * fsz, cudaMemcpyHostToDevice, sp));
                    AVS_CUDA(cudaMemcpyAsync(feat_at(pa, prs[b], m->Da, fsz), feat_at(audio, row_start[b], m->Da, fsz),
                                             n * Da_ * fsz, cudaMemcpyHostToDevice, sp));
                }
            } else {      // device space: ONE gather kernel per tensor (per-video copies cost more than the padding)
                AVS_TRY(upload_small(desc_dev, desc.data(), desc.size() * 4, sp));
                constexpr int ROWS = 8;
                const dim3 grid((max_len + ROWS - 1) / ROWS, n_videos);
                copy_video_rows_kernel<uint4><<<grid, 256, 0, sp>>>(reinterpret_cast<const uint4*>(visual),
                                                                    reinterpret_cast<uint4*>(pv), desc_dev, n_videos,
                                                                    static_cast<int>(Dv_ * fsz / 16), ROWS);
                AVS_LAUNCH_CHECK();
                copy_video_rows_kernel<uint4><<<grid, 256, 0, sp>>>(reinterpret_cast<const uint4*>(audio),
                                                                    reinterpret_cast<uint4*>(pa), desc_dev, n_videos,
                                                                    static_cast<int>(Da_ * fsz / 16), ROWS);
                AVS_LAUNCH_CHECK();
            }
            AVS_TRY(forward_entry(m, pv, pa, covered, n_videos, prs.data(), lengths, attn_axis, precision, ps, AVS_DEVICE,
                                  cuda_stream, nullptr, nullptr, nullptr));
            if (host) {
                for (int b = 0; b < n_videos; ++b)
                    if (lengths[b])
                        AVS_CUDA(cudaMemcpyAsync(scores + row_start[b], ps + prs[b], static_cast<size_t>(lengths[b]) * 4,
                                                 cudaMemcpyDeviceToHost, sp));
            } else {
                // scatter: packed scores -> each video's rows (descriptor roles swapped: src = packed, dst = caller's)
                std::vector<int32_t> back(desc);
                for (int b = 0; b < n_videos; ++b) std::swap(back[b], back[n_videos + b]);
                int32_t* back_dev = m->pack_ws.take<int32_t>(back.size());
                AVS_TRY(upload_small(back_dev, back.data(), back.size() * 4, sp));
                const dim3 grid((max_len + 255) / 256, n_videos);
                copy_video_rows_kernel<float><<<grid, 256, 0, sp>>>(ps, scores, back_dev, n_videos, 1, 256);
                AVS_LAUNCH_CHECK();
            }
            if (host) {
                AVS_CUDA(cudaStreamSynchronize(sp));
                AVS_TRY(range_report(m));
            }
            return AVS_OK;
        }
    }
    int n_groups = 1;
    if (m != nullptr && (host || space == AVS_DEVICE) && per_video && visual && audio && (scores || scores_dev_out) &&
        row_start && lengths && n_videos >= 2 && total_rows >= 4096 && total_rows < (1ll << 31)) {
        bool ordered = true;   // groups must be contiguous, disjoint row ranges
        for (int b = 0; b < n_videos && ordered; ++b) {
            ordered = lengths[b] >= 0 && row_start[b] >= 0 &&
                      static_cast<int64_t>(row_start[b]) + lengths[b] <= total_rows &&
                      (b == 0 || row_start[b] >= row_start[b - 1] + lengths[b - 1]);
        }
        if (ordered && host) {
            n_groups = (total_rows >= 16384 && n_videos >= 24) ? 6
                       : (total_rows >= 16384 && n_videos >= 12) ? 4 : ((total_rows >= 8192 && n_videos >= 6) ? 3 : 2);
            // a streamed step hides its tail behind the next batch's transfer, so it only needs enough groups to
            // start computing before the last byte has landed; fewer groups = fewer, longer copies (measured on
            // config 2: 1.87 ms per step with six groups, 1.83 with two; a bare copy takes 1.79)
            if (async) n_groups = std::min(n_groups, 2);
            if (const char* e = getenv("AVS_HOST_GROUPS")) {   // tuning aid
                const int want = atoi(e);
                if (want >= 1 && want <= n_groups && want != 5) n_groups = want;
            }
        }
        if (ordered && !host) {
            // device-resident inputs: nothing to hide a transfer behind, but the recurrence of the longest videos
            // is a chain of dependent steps that leaves the tensor pipe idle -- the other groups' GEMMs run beside it
            const char* e = getenv("AVS_DEV_GROUPS");
            const int want = e ? atoi(e) : 1;
            if (want >= 2 && want <= avs_model::MAX_CHUNKS && want != 5 && n_videos >= 8 * want) n_groups = want;
        }
    }
    if (n_groups == 1) {
        if (scores_dev_out == nullptr)
            return forward_impl(m, visual, audio, total_rows, n_videos, row_start, lengths, attn_axis, precision, scores,
                                space, cuda_stream);
        // small / unsplittable host batch, scores wanted on the device: stage the inputs, run in device space
        AVS_CHECK(m && visual && audio && total_rows >= 0, AVS_ERR_INVALID, "avs_forward_summarize: bad arguments");
        Guard g1(m->device);
        cudaStream_t s1 = static_cast<cudaStream_t>(cuda_stream);
        const size_t r1 = static_cast<size_t>(total_rows), fsz1 = m->feat_f16 ? 2 : 4;
        Arena& HI = m->host_in[slot];
        AVS_TRY(HI.reserve(r1 * (m->Dv + m->Da) * fsz1 + r1 * 8 + 5 * 256));
        HI.reset();
        float* v1 = reinterpret_cast<float*>(HI.take<char>(r1 * m->Dv * fsz1));
        float* a1 = reinterpret_cast<float*>(HI.take<char>(r1 * m->Da * fsz1));
        float* sc1 = HI.take<float>(r1);
        int32_t* p1 = HI.take<int32_t>(r1);
        AVS_CUDA(cudaMemcpyAsync(v1, visual, r1 * m->Dv * fsz1, cudaMemcpyHostToDevice, s1));
        AVS_CUDA(cudaMemcpyAsync(a1, audio, r1 * m->Da * fsz1, cudaMemcpyHostToDevice, s1));
        if (positions_host) AVS_CUDA(cudaMemcpyAsync(p1, positions_host, r1 * 4, cudaMemcpyHostToDevice, s1));
        *scores_dev_out = sc1;
        if (positions_dev_out) *positions_dev_out = p1;
        return forward_impl(m, v1, a1, total_rows, n_videos, row_start, lengths, attn_axis, precision, sc1, AVS_DEVICE,
                            cuda_stream);
    }

    AVS_CHECK(precision_ok(precision), AVS_ERR_INVALID, "bad precision %d", precision);
    Guard g(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int Dv = m->Dv, Da = m->Da;
    const size_t uR = static_cast<size_t>(total_rows), fsz = m->feat_f16 ? 2 : 4;
    const float* in_v = visual;
    const float* in_a = audio;
    float* sc_dev = scores;
    int32_t* pos_dev = nullptr;
    if (host) {
        Arena& HI = m->host_in[slot];
        AVS_TRY(HI.reserve(uR * (Dv + Da) * fsz + uR * 8 + 5 * 256));
        HI.reset();
        in_v = reinterpret_cast<float*>(HI.take<char>(uR * Dv * fsz));
        in_a = reinterpret_cast<float*>(HI.take<char>(uR * Da * fsz));
        sc_dev = HI.take<float>(uR);
        pos_dev = HI.take<int32_t>(uR);
    }
    // group boundaries at video boundaries.  Host space: shares shrink towards the end -- what remains after the
    // last byte has crossed PCIe is the last group's compute, so it should be the smallest (and, for a
    // longest-first batch, the one with the shortest recurrence).  Device space: the first group holds the longest
    // videos (the critical chain) and is kept small so that its recurrence starts early.
    static const int kShareHost[7][6] = {{0}, {100}, {60, 100}, {45, 80, 100}, {40, 70, 90, 100}, {0}, {32, 58, 77, 89, 96, 100}};
    static const int kShareDev[7][6] = {{0}, {100}, {30, 100}, {25, 60, 100}, {24, 48, 74, 100}, {0}, {24, 45, 62, 76, 88, 100}};
    int share[6];
    for (int i = 0; i < 6; ++i) share[i] = host ? kShareHost[n_groups][i] : kShareDev[n_groups][i];
    if (const char* e = getenv(host ? "AVS_HOST_SHARES" : "AVS_DEV_SHARES")) {   // tuning aid: "24,48,74,100"
        int v[6], k = 0;
        for (const char* p = e; *p && k < 6;) {
            v[k++] = atoi(p);
            while (*p && *p != ',') ++p;
            if (*p == ',') ++p;
        }
        if (k == n_groups && v[k - 1] == 100)
            for (int i = 0; i < k; ++i) share[i] = v[i];
    }
    int first[avs_model::MAX_CHUNKS + 1];
    first[0] = 0;
    for (int gi = 1; gi < n_groups; ++gi) {
        const int64_t want = total_rows * share[gi - 1] / 100;
        int b = first[gi - 1] + 1;
        const int b_max = n_videos - (n_groups - gi);
        while (b < b_max && row_start[b] < want) ++b;
        // a cluster of the recurrence kernel carries up to 8 videos: groups of 8k videos waste no cluster slots
        const int cnt = b - first[gi - 1];
        if (cnt >= 6) {
            const int snapped = first[gi - 1] + (cnt + 3) / 8 * 8;
            if (snapped > first[gi - 1] && snapped <= b_max) b = snapped;
        }
        first[gi] = b;
    }
    first[n_groups] = n_videos;
    const bool trace = host && g_e2e.enabled();
    if (trace) {
        g_e2e.n_groups = n_groups;
        cudaEventRecord(g_e2e.t0, st);
    }
    // Synchronous calls: the staging area is free once prior work on st is done.  Asynchronous calls own a staging
    // slot whose previous user has completed (avs_slot_wait), so the copies start at once -- while the previous
    // batch (other slot) is still being computed.
    if (!async) AVS_CUDA(cudaEventRecord(m->ev_start, st));
    int64_t lo[avs_model::MAX_CHUNKS], hi[avs_model::MAX_CHUNKS];
    for (int gi = 0; gi < n_groups; ++gi) {
        lo[gi] = row_start[first[gi]];
        hi[gi] = static_cast<int64_t>(row_start[first[gi + 1] - 1]) + lengths[first[gi + 1] - 1];
    }
    if (host) {
        if (!async) AVS_CUDA(cudaStreamWaitEvent(m->copy_stream, m->ev_start, 0));
        if (positions_host)   // summary input, a few KB: goes first so it never waits behind the features
            AVS_CUDA(cudaMemcpyAsync(pos_dev, positions_host, uR * 4, cudaMemcpyHostToDevice, m->copy_stream));
        for (int gi = 0; gi < n_groups; ++gi) {
            const size_t rows = static_cast<size_t>(hi[gi] - lo[gi]);
            if (rows) {
                AVS_CUDA(cudaMemcpyAsync(feat_at(const_cast<float*>(in_v), lo[gi], Dv, fsz), feat_at(visual, lo[gi], Dv, fsz),
                                         rows * Dv * fsz, cudaMemcpyHostToDevice, m->copy_stream));
                AVS_CUDA(cudaMemcpyAsync(feat_at(const_cast<float*>(in_a), lo[gi], Da, fsz), feat_at(audio, lo[gi], Da, fsz),
                                         rows * Da * fsz, cudaMemcpyHostToDevice, m->copy_stream));
            }
            AVS_CUDA(cudaEventRecord(m->ev_chunk[gi], m->copy_stream));
            if (trace) cudaEventRecord(g_e2e.chunk[gi], m->copy_stream);
        }
    }
    if (trace) g_e2e.host[1] = E2ETrace::now();
    // every group on its own stream and workspace: the recurrence of group g (a chain of max(T_g) dependent steps
    // that leaves the tensor pipe idle) keeps running while the other groups' GEMMs execute.  The groups share
    // the GPU, so their recurrences do not ask for exclusive SMs -- except the first group of a device-space call,
    // whose chain (the longest videos) is the critical path.
    std::vector<int32_t> rs;
    for (int gi = 0; gi < n_groups; ++gi) {
        const int nv = first[gi + 1] - first[gi];
        rs.assign(nv, 0);
        for (int b = 0; b < nv; ++b) rs[b] = row_start[first[gi] + b] - static_cast<int32_t>(lo[gi]);
        cudaStream_t gs = gi == 0 ? st : m->grp_stream[gi - 1];
        if (gi > 0 && !async) AVS_CUDA(cudaStreamWaitEvent(gs, m->ev_start, 0));   // not before earlier work on st is done
        if (host) AVS_CUDA(cudaStreamWaitEvent(gs, m->ev_chunk[gi], 0));
        AVS_TRY(forward_impl(m, feat_at(in_v, lo[gi], Dv, fsz), feat_at(in_a, lo[gi], Da, fsz), hi[gi] - lo[gi], nv, rs.data(),
                             lengths + first[gi], attn_axis, precision, sc_dev + lo[gi], AVS_DEVICE, gs,
                             gi == 0 ? nullptr : &m->ws_grp[gi - 1], (!host && gi == 0) ? 0 : -1));
        if (trace) cudaEventRecord(g_e2e.grp[gi], gs);
        if (gi > 0) {
            AVS_CUDA(cudaEventRecord(m->ev_grp[gi - 1], gs));
            AVS_CUDA(cudaStreamWaitEvent(st, m->ev_grp[gi - 1], 0));
        }
    }
    if (trace) g_e2e.host[2] = E2ETrace::now();
    if (!host) return AVS_OK;   // device space: the caller's stream now depends on every group
    if (scores_dev_out != nullptr) {   // the caller continues on the stream with the scores still on the device
        *scores_dev_out = sc_dev;
        if (positions_dev_out) *positions_dev_out = pos_dev;
        return AVS_OK;
    }
    AVS_CUDA(cudaMemcpyAsync(scores, sc_dev, uR * 4, cudaMemcpyDeviceToHost, st));
    AVS_CUDA(cudaStreamSynchronize(st));
    return range_report(m);
}

avs_status avs_forward(avs_model* m, const float* visual, const float* audio, int64_t total_rows, int32_t n_videos,
                       const int32_t* row_start, const int32_t* lengths, int attn_axis, int precision, float* scores,
                       int space, void* cuda_stream) {
    return forward_entry(m, visual, audio, total_rows, n_videos, row_start, lengths, attn_axis, precision, scores, space,
                         cuda_stream, nullptr, nullptr, nullptr);
}

static avs_status forward_impl(avs_model* m, const float* visual, const float* audio, int64_t total_rows,
                               int32_t n_videos, const int32_t* row_start, const int32_t* lengths, int attn_axis,
                               int precision, float* scores, int space, void* cuda_stream, Arena* arena, int lstm_excl) {
    AVS_CHECK(m != nullptr, AVS_ERR_INVALID, "model handle is null");
    Arena& WS = arena ? *arena : m->ws;
    AVS_CHECK(space == AVS_HOST || space == AVS_DEVICE, AVS_ERR_INVALID, "bad memory space %d", space);
    AVS_CHECK(precision_ok(precision), AVS_ERR_INVALID, "bad precision %d", precision);
    AVS_CHECK(attn_axis == AVS_ATTN_LITERAL || attn_axis == AVS_ATTN_TEMPORAL || attn_axis == AVS_ATTN_LITERAL_B1,
              AVS_ERR_INVALID, "bad attn_axis %d", attn_axis);
    int max_len = 0;
    AVS_TRY(validate_videos(total_rows, n_videos, row_start, lengths, &max_len));
    if (total_rows == 0 || n_videos == 0 || max_len == 0) return AVS_OK;
    AVS_CHECK(visual && audio && scores, AVS_ERR_INVALID, "visual / audio / scores pointer is null");
    Guard g(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int64_t R = total_rows;
    const int Dv = m->Dv, Da = m->Da;
    const bool simt = precision == AVS_PREC_FP32_SIMT;
    const bool bf16 = precision == AVS_PREC_BF16;
    // Operand formats.  AVS_PREC_TF32: the user's fp32 features are rounded (RN) to tf32 and feed kind::tf32
    // GEMMs (fp32 exponent range); every activation behind the fc layers is fp16 (same 11-bit significand,
    // saturating) and feeds kind::f16.  AVS_PREC_BF16: bf16 everywhere.  Accumulation, gate pre-activations,
    // cell state, softmax statistics and the score head stay fp32 in all modes.
    const int act = simt ? DT_F32 : (bf16 ? DT_BF16 : DT_F16);   // internal activations
    // fp16 feature format (avs_model_set_feature_format): the caller's buffers already hold 16-bit operands -- half
    // the PCIe / HBM bytes; they feed kind::f16 GEMMs against the fp16 weight copies directly (exact operands, no
    // truncation to compensate).  Only with the default precision (fp16 activations).
    const bool f16_in = m->feat_f16 != 0;
    AVS_CHECK(!f16_in || precision == AVS_PREC_TF32, AVS_ERR_UNSUPPORTED,
              "fp16 feature buffers need the default precision mode (AVS_PREC_TF32: fp16 operands behind the fc layers)");
    const size_t fsz = f16_in ? 2 : 4;
    const int in_dt = bf16 ? DT_BF16 : (f16_in ? DT_F16 : DT_F32);   // what the fc GEMMs read
    const size_t asz = dtype_size(act);
    // gate pre-activations (input projections of the four recurrences): fp16 on the tensor-core paths -- the largest
    // HBM item of the step (2 x 2048 values per frame, written by two GEMMs and read back by the recurrence), halved;
    // fp16 in the bf16 mode too (pre-activations need the significand, and beyond +-65504 every gate is saturated
    // anyway: the cast clamps).  AVS_XG_F32=1 keeps fp32 (A/B aid).
    static const bool xg_f32 = getenv("AVS_XG_F32") != nullptr;
    const int xg_dt = (simt || xg_f32) ? DT_F32 : DT_F16;
    const size_t xsz = dtype_size(xg_dt);

    bool literal_rows = attn_axis == AVS_ATTN_LITERAL_B1 || (attn_axis == AVS_ATTN_LITERAL && n_videos == 1);
    if (attn_axis == AVS_ATTN_LITERAL && n_videos > 1) {
        for (int b = 0; b < n_videos; ++b)
            AVS_CHECK(lengths[b] == lengths[0] && row_start[b] == row_start[0] + b * lengths[0], AVS_ERR_INVALID,
                      "AVS_ATTN_LITERAL mixes the videos of a batch (av_model.py:44) and needs a dense [B, T] "
                      "layout with equal lengths; video %d breaks it", b);
    }

    // fp16 range watch (see avs_model::range_flag): only when every row belongs to a video -- rows no video owns may
    // hold anything (the contract is that they never reach a valid frame), and the GEMM epilogue cannot tell them apart
    int64_t owned_rows = 0;
    for (int b = 0; b < n_videos; ++b) owned_rows += lengths[b];
    unsigned int* const sat_flag = (!simt && !bf16 && owned_rows == total_rows) ? m->range_flag_dev : nullptr;
    // ---- plan + workspace
    LstmPlan plan = plan_lstm(n_videos, row_start, lengths, !simt);
    const int n_seqs = literal_rows ? 0 : (attn_axis == AVS_ATTN_TEMPORAL ? n_videos : lengths[0]);
    const bool tc_attn = !simt && attn_axis == AVS_ATTN_TEMPORAL && E == m->heads * 256;
    const int qkv_dt = tc_attn ? act : DT_F32;
    const size_t uR = static_cast<size_t>(R);
    const size_t bytes = uR * (Dv + Da) * fsz + (bf16 ? uR * (Dv + Da) * 2 : 0) + 2 * uR * H * asz + 2 * uR * 2 * G4 * xsz +
                         3 * uR * E * asz + (literal_rows ? 0 : uR * 3 * E * dtype_size(qkv_dt)) + uR * 4 +
                         (plan.host.size() + 3 * static_cast<size_t>(n_seqs)) * 4 + 64 * 256;
    AVS_TRY(WS.reserve(bytes));
    WS.reset();
    float* in_v = reinterpret_cast<float*>(WS.take<char>(uR * Dv * fsz));
    float* in_a = reinterpret_cast<float*>(WS.take<char>(uR * Da * fsz));
    uint16_t* in_v16 = bf16 ? WS.take<uint16_t>(uR * Dv) : nullptr;
    uint16_t* in_a16 = bf16 ? WS.take<uint16_t>(uR * Da) : nullptr;
    char* v_emb = WS.take<char>(uR * H * asz);
    char* a_emb = WS.take<char>(uR * H * asz);
    char* xg_v = WS.take<char>(uR * 2 * G4 * xsz);
    char* xg_a = WS.take<char>(uR * 2 * G4 * xsz);
    char* fused = WS.take<char>(uR * E * asz);
    char* qkv = literal_rows ? nullptr : WS.take<char>(uR * 3 * E * dtype_size(qkv_dt));
    char* ctx = WS.take<char>(uR * E * asz);
    char* attn_out = WS.take<char>(uR * E * asz);
    float* scores_dev = space == AVS_DEVICE ? scores : WS.take<float>(uR);
    int32_t* plan_dev = WS.take<int32_t>(plan.host.size());
    int32_t* seq_dev = WS.take<int32_t>(3 * static_cast<size_t>(std::max(n_seqs, 1)));

    const GemmW w_fc_v{m->fc_v_w_x, m->fc_v_w_t, m->fc_v_w_l}, w_fc_a{m->fc_a_w_x, m->fc_a_w_t, m->fc_a_w_l};
    const GemmW w_ih_v{m->ih_v_x, m->ih_v_t, m->ih_v_l}, w_ih_a{m->ih_a_x, m->ih_a_t, m->ih_a_l};
    const GemmW w_in{m->in_w_x, m->in_w_t, m->in_w_l}, w_out{m->out_w_x, m->out_w_t, m->out_w_l};
    const GemmW w_sc0{m->sc0_w_x, m->sc0_w_t, m->sc0_w_l};

    // ---- front of the pipeline, row-parallel and therefore chunked: H2D (host space) -> operand conversion of the
    // user features -> K1 visual_fc / audio_fc (Linear + ReLU; Dropout is identity in eval, av_model.py:35-36)
    // -> K2a LSTM input projections for both directions (av_model.py:39-40).  In host space chunk c+1 is in
    // flight on the copy stream while chunk c is being computed.
    AVS_TRY(upload_small(plan_dev, plan.host.data(), plan.host.size() * 4, st));
    // ---- pipelined front (device-resident batches ordered longest video first, the order packed_batches produces):
    // the recurrence is a chain of max(T) dependent steps that leaves most of every SM idle, and it cannot start before
    // the input projections of its rows exist.  Instead of ALL front GEMMs followed by ALL recurrences, the rows are
    // processed group by group (a group = the videos of one recurrence cluster): as soon as the GEMMs of group k are
    // done its recurrence starts on its own high-priority stream, and the GEMMs of groups k+1.. run beside it on the
    // SMs the running recurrences leave free (their persistent grids are sized for those).  The longest videos come
    // first, so the critical chain starts after ~1/4 of the GEMM work instead of all of it.
    // MEASURED (B200, config 2, profiles/r02f_pipeline_ab.log): front + recurrences 0.749 ms pipelined vs 0.174 + 0.481 =
    // 0.655 ms one after the other -- the GEMMs lose the SMs the recurrence clusters hold (and wait for clusters to be
    // placed), the second-longest group becomes the critical chain and starts later than it does today.  Same
    // conclusion as round 1's video-group experiment: OFF by default, AVS_PIPELINE=1 turns it on.
    bool pipelined = false;
    if (!simt && space == AVS_DEVICE && arena == nullptr && lstm_excl == 0 && plan.nb == 8 && plan.n_groups >= 3 &&
        getenv("AVS_PIPELINE") != nullptr && !(reinterpret_cast<uintptr_t>(visual) & 15) &&
        !(reinterpret_cast<uintptr_t>(audio) & 15)) {
        const int G = plan.n_groups, nbp = plan.nb;
        std::vector<int64_t> glo, ghi;
        const bool ordered = group_row_ranges(plan, R, glo, ghi);
        if (ordered) {
            pipelined = true;
            const int n_seg = std::min(G, avs_model::PIPE_SEGS);   // groups 0 .. n_seg-2 alone, the rest together
            const int n_excl = lstm_exclusive_groups(G);
            const int sms = device_sm_count();
            const int slots_all = G * nbp;
            LstmBatch lb{plan_dev, plan_dev + slots_all, plan_dev + 2 * slots_all, G, nbp, 0};
            cudaStream_t sa = m->branch_stream[0];
            StageTimer tm(ST_FRONT_LSTM, st);
            int busy = 0;   // SMs held by the recurrences launched so far (exclusive: 32 per group, shared: 16)
            for (int k = 0; k < n_seg; ++k) {
                const int g_lo = k, g_hi = (k == n_seg - 1) ? G : k + 1;
                const int64_t r0 = glo[g_lo], Rc = ghi[g_hi - 1] - r0;
                GemmEpilogue e1;
                e1.relu = 1;
                e1.out_dtype = act;
                e1.sat_flag = sat_flag;
                e1.ldc = H;
                if (!bf16 && !f16_in) e1.acc_scale = 1.0f + 1.0f / 2048.0f;
                GemmEpilogue e2;
                e2.ldc = 2 * G4;
                e2.out_dtype = xg_dt;
                e1.max_ctas = e2.max_ctas = busy ? std::max(sms - busy, 16) : 0;
                const void* xv = feat_at(visual, r0, Dv, fsz);
                const void* xa = feat_at(audio, r0, Da, fsz);
                if (bf16) {
                    void* dv = in_v16 + r0 * Dv;
                    void* da = in_a16 + r0 * Da;
                    AVS_TRY(convert_f32(static_cast<const float*>(xv), dv, Rc * Dv, in_dt, 0, st));
                    AVS_TRY(convert_f32(static_cast<const float*>(xa), da, Rc * Da, in_dt, 0, st));
                    xv = dv;
                    xa = da;
                }
                AVS_CUDA(cudaEventRecord(m->ev_branch_in[0], st));
                AVS_CUDA(cudaStreamWaitEvent(sa, m->ev_branch_in[0], 0));
                e1.bias = m->fc_v_b;
                e1.C = v_emb + r0 * H * asz;
                AVS_TRY(run_gemm(precision, xv, in_dt, Dv, w_fc_v, 0, Dv, Rc, H, Dv, e1, st));
                e1.bias = m->fc_a_b;
                e1.C = a_emb + r0 * H * asz;
                AVS_TRY(run_gemm(precision, xa, in_dt, Da, w_fc_a, 0, Da, Rc, H, Da, e1, sa));
                e2.bias = m->ih_v_b;
                e2.C = xg_v + r0 * 2 * G4 * xsz;
                AVS_TRY(run_gemm(precision, v_emb + r0 * H * asz, act, H, w_ih_v, 0, H, Rc, 2 * G4, H, e2, st));
                e2.bias = m->ih_a_b;
                e2.C = xg_a + r0 * 2 * G4 * xsz;
                AVS_TRY(run_gemm(precision, a_emb + r0 * H * asz, act, H, w_ih_a, 0, H, Rc, 2 * G4, H, e2, sa));
                AVS_CUDA(cudaEventRecord(m->ev_branch_out[0], sa));
                AVS_CUDA(cudaStreamWaitEvent(st, m->ev_branch_out[0], 0));
                AVS_CUDA(cudaEventRecord(m->ev_pipe_front[k], st));
                AVS_CUDA(cudaStreamWaitEvent(m->pipe_stream[k], m->ev_pipe_front[k], 0));
                // ONE launch per segment (kernels on one stream would run one after the other); only single-group
                // segments can be exclusive
                const int excl = (g_hi - g_lo == 1 && g_lo < n_excl) ? 1 : 0;
                AVS_TRY(lstm_recurrence_tc_groups(xg_v, xg_a, xg_dt, m->whh, lb, g_lo, g_hi, excl, act, fused, act, m->pipe_stream[k]));
                busy += (g_hi - g_lo) * (excl ? 32 : 16);
                AVS_CUDA(cudaEventRecord(m->ev_pipe_done[k], m->pipe_stream[k]));
            }
            for (int k = 0; k < n_seg; ++k) AVS_CUDA(cudaStreamWaitEvent(st, m->ev_pipe_done[k], 0));
        }
    }
    const int n_chunks = pipelined ? 0 : ((space == AVS_HOST && R >= 2048) ? avs_model::MAX_CHUNKS : 1);
    const int64_t chunk_rows = ((R + std::max(n_chunks, 1) - 1) / std::max(n_chunks, 1) + 127) / 128 * 128;
    if (space == AVS_HOST) {
        AVS_CUDA(cudaEventRecord(m->ev_start, st));              // the workspace is free once prior work on st is done
        AVS_CUDA(cudaStreamWaitEvent(m->copy_stream, m->ev_start, 0));
        for (int c = 0; c < n_chunks; ++c) {
            const int64_t r0 = c * chunk_rows, r1 = std::min<int64_t>(R, r0 + chunk_rows);
            if (r0 >= r1) break;
            AVS_CUDA(cudaMemcpyAsync(feat_at(in_v, r0, Dv, fsz), feat_at(visual, r0, Dv, fsz),
                                     static_cast<size_t>(r1 - r0) * Dv * fsz, cudaMemcpyHostToDevice, m->copy_stream));
            AVS_CUDA(cudaMemcpyAsync(feat_at(in_a, r0, Da, fsz), feat_at(audio, r0, Da, fsz),
                                     static_cast<size_t>(r1 - r0) * Da * fsz, cudaMemcpyHostToDevice, m->copy_stream));
            AVS_CUDA(cudaEventRecord(m->ev_chunk[c], m->copy_stream));
        }
    }
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t r0 = c * chunk_rows, r1 = std::min<int64_t>(R, r0 + chunk_rows);
        if (r0 >= r1) break;
        const int64_t Rc = r1 - r0;
        const float* src_v = feat_at(space == AVS_HOST ? in_v : visual, r0, Dv, fsz);
        const float* src_a = feat_at(space == AVS_HOST ? in_a : audio, r0, Da, fsz);
        if (space == AVS_HOST) AVS_CUDA(cudaStreamWaitEvent(st, m->ev_chunk[c], 0));
        // the GEMMs read the features through TMA, which needs 16-byte aligned rows: a caller-owned device buffer
        // that is not (never the case for a torch tensor) is staged in the workspace first
        if (space == AVS_DEVICE && !simt && (reinterpret_cast<uintptr_t>(src_v) & 15)) {
            AVS_CUDA(cudaMemcpyAsync(feat_at(in_v, r0, Dv, fsz), src_v, static_cast<size_t>(Rc) * Dv * fsz, cudaMemcpyDeviceToDevice, st));
            src_v = feat_at(in_v, r0, Dv, fsz);
        }
        if (space == AVS_DEVICE && !simt && (reinterpret_cast<uintptr_t>(src_a) & 15)) {
            AVS_CUDA(cudaMemcpyAsync(feat_at(in_a, r0, Da, fsz), src_a, static_cast<size_t>(Rc) * Da * fsz, cudaMemcpyDeviceToDevice, st));
            src_a = feat_at(in_a, r0, Da, fsz);
        }
        const void* xv = src_v;
        const void* xa = src_a;
        if (bf16) {
            StageTimer tm(ST_CONVERT, st);
            void* dv = in_v16 + r0 * Dv;
            void* da = in_a16 + r0 * Da;
            AVS_TRY(convert_f32(src_v, dv, Rc * Dv, in_dt, 0, st));
            AVS_TRY(convert_f32(src_a, da, Rc * Da, in_dt, 0, st));
            xv = dv;
            xa = da;
        }
        // AVS_PREC_TF32: the fc GEMMs read the user's fp32 features as they are.  kind::tf32 ignores the low 13
        // mantissa bits (truncation towards zero): x -> x (1 - d), d uniform in [0, 2^-10).  Its variance equals
        // that of round-to-nearest (2^-10 / sqrt(12) == 2^-11 / sqrt(3)); only the mean, -2^-11, differs, and
        // that is a common factor of every product, removed exactly by scaling the accumulator with 1 + 2^-11
        // in the epilogue.  No separate rounding pass over the 4.6 KB/frame of features (weights are RN-rounded
        // once at pack time).
        // The visual and the audio branch are independent until the recurrences: the audio branch runs on a side
        // stream, so the tail of one GEMM (its last, partly filled round of tiles) overlaps the head of another.
        static const bool two_branches = getenv("AVS_ONE_BRANCH_STREAM") == nullptr;
        const int gidx = arena ? static_cast<int>(arena - m->ws_grp) + 1 : 0;
        cudaStream_t sa = (two_branches && n_chunks == 1) ? m->branch_stream[gidx] : st;
        GemmEpilogue e1;
        e1.relu = 1;
        e1.out_dtype = act;
        e1.sat_flag = sat_flag;
        e1.ldc = H;
        if (!simt && !bf16 && !f16_in) e1.acc_scale = 1.0f + 1.0f / 2048.0f;
        GemmEpilogue e2;
        e2.ldc = 2 * G4;
        e2.out_dtype = xg_dt;
        if (sa != st) {
            // concurrent kernels: per-kernel event pairs would count the overlap twice, so the four GEMMs are
            // timed as ONE stage ("frontend_gemms": fork -> join on st)
            StageTimer tm(ST_FRONT, st);
            AVS_CUDA(cudaEventRecord(m->ev_branch_in[gidx], st));
            AVS_CUDA(cudaStreamWaitEvent(sa, m->ev_branch_in[gidx], 0));
            e1.bias = m->fc_v_b;
            e1.C = v_emb + r0 * H * asz;
            AVS_TRY(run_gemm(precision, xv, in_dt, Dv, w_fc_v, 0, Dv, Rc, H, Dv, e1, st));
            e1.bias = m->fc_a_b;
            e1.C = a_emb + r0 * H * asz;
            AVS_TRY(run_gemm(precision, xa, in_dt, Da, w_fc_a, 0, Da, Rc, H, Da, e1, sa));
            e2.bias = m->ih_v_b;
            e2.C = xg_v + r0 * 2 * G4 * xsz;
            AVS_TRY(run_gemm(precision, v_emb + r0 * H * asz, act, H, w_ih_v, 0, H, Rc, 2 * G4, H, e2, st));
            e2.bias = m->ih_a_b;
            e2.C = xg_a + r0 * 2 * G4 * xsz;
            AVS_TRY(run_gemm(precision, a_emb + r0 * H * asz, act, H, w_ih_a, 0, H, Rc, 2 * G4, H, e2, sa));
            AVS_CUDA(cudaEventRecord(m->ev_branch_out[gidx], sa));
            AVS_CUDA(cudaStreamWaitEvent(st, m->ev_branch_out[gidx], 0));
        } else {
            {
                StageTimer tm(ST_FC, st);
                e1.bias = m->fc_v_b;
                e1.C = v_emb + r0 * H * asz;
                AVS_TRY(run_gemm(precision, xv, in_dt, Dv, w_fc_v, 0, Dv, Rc, H, Dv, e1, st));
                e1.bias = m->fc_a_b;
                e1.C = a_emb + r0 * H * asz;
                AVS_TRY(run_gemm(precision, xa, in_dt, Da, w_fc_a, 0, Da, Rc, H, Da, e1, st));
            }
            {
                StageTimer tm(ST_IH_PROJ, st);
                e2.bias = m->ih_v_b;
                e2.C = xg_v + r0 * 2 * G4 * xsz;
                AVS_TRY(run_gemm(precision, v_emb + r0 * H * asz, act, H, w_ih_v, 0, H, Rc, 2 * G4, H, e2, st));
                e2.bias = m->ih_a_b;
                e2.C = xg_a + r0 * 2 * G4 * xsz;
                AVS_TRY(run_gemm(precision, a_emb + r0 * H * asz, act, H, w_ih_a, 0, H, Rc, 2 * G4, H, e2, st));
            }
        }
    }

    // ---- optional range check (AVS_CHECK_RANGE=1): the default mode keeps the activations behind the fc layers in
    // fp16, whose casts SATURATE at 65504 instead of overflowing -- silent for un-normalised features with huge
    // magnitudes.  With the switch on, a saturated / non-finite embedding fails the call loudly (one extra pass over
    // 2 KB per frame and a synchronisation: a validation aid, not the default).
    const bool check_range = getenv("AVS_CHECK_RANGE") != nullptr;
    if (check_range && act == DT_F16) {
        unsigned int* cnt = reinterpret_cast<unsigned int*>(seq_dev);   // reuses the (not yet written) descriptor slot
        unsigned int host_cnt = 0;
        AVS_CUDA(cudaMemsetAsync(cnt, 0, 4, st));
        AVS_TRY(count_saturated_f16(v_emb, R * H, cnt, st));
        AVS_TRY(count_saturated_f16(a_emb, R * H, cnt, st));
        AVS_CUDA(cudaMemcpyAsync(&host_cnt, cnt, 4, cudaMemcpyDeviceToHost, st));
        AVS_CUDA(cudaStreamSynchronize(st));
        AVS_CHECK(host_cnt == 0, AVS_ERR_UNSUPPORTED,
                  "%u fc activations saturate fp16 (|x| >= 65504) or are not finite: normalise the features or use "
                  "AVS_PREC_BF16 (fp32 exponent range)", host_cnt);
    }

    // ---- pipelined tail (device-resident batches ordered longest video first; sequence-length-1 or temporal
    // attention -- both work per video): the recurrence of a group lasts max(T of the group) dependent steps, so the groups finish one after the other,
    // shortest first, and the SMs of a finished group stay idle until the longest group is done (config 2: 128 of 148
    // SMs hold recurrence CTAs, and the groups end between ~1/3 and all of the kernel's time).  Everything behind the
    // recurrence is parallel over rows or videos (value projection -- or q | k | v projection + attention core --,
    // out_proj, score head: av_model.py:44-46), so each group's recurrence is launched on its own stream, followed on the
    // same stream by the tail of that group's rows: the tails of the shorter groups run on the SMs their recurrences have left while the longest
    // chain is still going, and only the tail of the longest group (~1/4 of the rows) remains after it.  Row r of a
    // GEMM is computed identically whatever the row range of the launch, so the scores are bit-identical to the
    // one-launch schedule.  Per-stage profiling (avs_profile) then reports the recurrence stage as front-done -> the
    // LAST group's recurrence done, and "tail_behind_longest_recurrence" = the longest group's recurrence done -> all
    // streams joined; the individual tail GEMMs are only timed in the one-launch schedule (AVS_PIPE_TAIL=0).
    const char* const pipe_tail_env = getenv("AVS_PIPE_TAIL");   // read per call: the tests toggle it
    const int pipe_tail = pipe_tail_env ? atoi(pipe_tail_env) : 1;
    bool tail_done = false;
    const bool tail_temporal = attn_axis == AVS_ATTN_TEMPORAL && tc_attn;
    if (!pipelined && pipe_tail && !simt && (literal_rows || tail_temporal) && space == AVS_DEVICE && arena == nullptr && lstm_excl == 0 &&
        plan.nb == 8 && plan.n_groups >= 3 && plan.n_groups <= avs_model::PIPE_SEGS && owned_rows == total_rows) {
        std::vector<int64_t> glo, ghi;
        // Worth it when the groups finish at different times (the shortest group's longest video is <= 0.9 x the longest
        // group's) or when the tails are long (>= 16k rows: the tails of equally long groups then overlap each other's
        // GEMMs and attention).  Small batches of equally long videos end together and only pay for the stagger and the
        // smaller GEMMs (measured, 5 - 12 videos x 320 frames: +3 .. 7 %).  AVS_PIPE_TAIL=2 forces the schedule.
        const int slots_chk = plan.n_groups * plan.nb;
        const bool spread = plan.host[2 * slots_chk + plan.n_groups - 1] * 10ll <= plan.host[2 * slots_chk] * 9ll;
        if ((spread || R >= 16384 || pipe_tail >= 2) && group_row_ranges(plan, R, glo, ghi)) {
            tail_done = true;
            const int G = plan.n_groups, slots_all = G * plan.nb;
            const int n_excl = lstm_exclusive_groups(G);
            // groups 2.. each start behind one more short stagger kernel, so that the groups are placed in order of
            // length (measured: 0.685 vs 0.695 ms per config-2 forward; AVS_PIPE_STAGGER_ONCE=1: one stagger for all)
            const bool stagger_each = getenv("AVS_PIPE_STAGGER_ONCE") == nullptr;
            const int sms = device_sm_count();
            LstmBatch lb{plan_dev, plan_dev + slots_all, plan_dev + 2 * slots_all, G, plan.nb, 0};
            // temporal attention: one descriptor table in PLAN order (group by group), so that the videos of a group
            // are a contiguous range of it; the attention core of a group is one launch over that range
            std::vector<int> seq_off(G + 1, 0);
            if (tail_temporal) {
                int n_sq = 0;
                for (int i = 0; i < slots_all; ++i) n_sq += plan.host[slots_all + i] > 0;
                std::vector<int32_t> sd(3 * static_cast<size_t>(n_sq));
                int q = 0;
                for (int k = 0; k < G; ++k) {
                    seq_off[k] = q;
                    for (int i = 0; i < plan.nb; ++i) {
                        const int32_t len = plan.host[slots_all + k * plan.nb + i];
                        if (len <= 0) continue;
                        sd[q] = plan.host[k * plan.nb + i];
                        sd[n_sq + q] = 1;
                        sd[2 * n_sq + q] = len;
                        ++q;
                    }
                }
                seq_off[G] = q;
                AVS_CHECK(n_sq <= std::max(n_seqs, 1), AVS_ERR_INVALID, "pipelined tail: descriptor table too small");
                AVS_TRY(upload_small(seq_dev, sd.data(), sd.size() * 4, st));
            }
            // The longest group runs on the caller's stream itself: its recurrence follows the front without an event
            // hop and its SM-exclusive CTAs are placed first (exclusive CTAs need EMPTY SMs; were the 192 shared CTAs
            // of the other groups placed first, one on every SM, the longest chain would wait for the shortest group
            // to finish -- measured: +150..260 us in a third of the steps).  The other groups start a few microseconds
            // later, behind a stagger kernel; they have that much slack many times over.
            AVS_CUDA(cudaEventRecord(m->ev_pipe_front[0], st));
            AVS_CUDA(cudaStreamWaitEvent(m->pipe_stream[0], m->ev_pipe_front[0], 0));
            AVS_TRY(launch_stagger(m->pipe_stream[0], 4000));
            AVS_CUDA(cudaEventRecord(m->ev_pipe_front[1], m->pipe_stream[0]));
            avs_status rc = AVS_OK;
            int launched = 0;
            cudaEvent_t prof_a = nullptr, prof_g0 = nullptr;
            const long long prof_span = g_prof.enabled ? g_prof.next_span++ : 0;
            if (g_prof.enabled) {
                prof_a = g_prof.get();
                cudaEventRecord(prof_a, st);
            }
            // AVS_PIPE_TRACE=1: timeline of the schedule on stderr (synchronises: a debugging aid)
            static const bool ptrace = getenv("AVS_PIPE_TRACE") != nullptr;
            static cudaEvent_t tev[1 + 2 * avs_model::PIPE_SEGS] = {};
            if (ptrace) {
                if (!tev[0]) for (auto& e : tev) cudaEventCreate(&e);
                cudaEventRecord(tev[0], st);
            }
            // Phase 1: every group's recurrence on its own stream, placed in order of length.
            for (int k = 0; k < G && rc == AVS_OK; ++k) {
                cudaStream_t ps = k ? m->pipe_stream[k] : st;
                if (stagger_each && k >= 2) {
                    launch_stagger(m->pipe_stream[0], 1500);
                    cudaEventRecord(m->ev_pipe_front[k], m->pipe_stream[0]);
                }
                if (k && cudaStreamWaitEvent(ps, m->ev_pipe_front[(stagger_each && k >= 2) ? k : 1], 0) != cudaSuccess) { rc = AVS_ERR_CUDA; break; }
                ++launched;
                rc = lstm_recurrence_tc_groups(xg_v, xg_a, xg_dt, m->whh, lb, k, k + 1, k < n_excl ? 1 : 0, act, fused, act, ps);
                cudaEventRecord(m->ev_pipe_rec[k], ps);
                if (ptrace) cudaEventRecord(tev[1 + 2 * k], ps);
                if (prof_a) {
                    cudaEvent_t b = g_prof.get();
                    cudaEventRecord(b, ps);
                    g_prof.pending.push_back({ST_LSTM, prof_a, b, prof_span});
                    if (k == 0) prof_g0 = b;
                }
            }
            // Phase 2: the tails, one per SEGMENT of groups.  A segment is one group, except that groups expected to
            // finish about when the longest one does share ITS tail: a step of a group on shared SMs takes ~1.16x a
            // step on exclusive SMs (0.78 vs 0.67 us), so a group whose longest video has >= 0.86x the steps of the
            // longest group's ends no earlier -- its own tail would run beside the longest group's, both on half a GPU
            // with a partly filled last round of tiles each (measured, config 2: 69 + 57 us side by side).
            static const bool no_merge = getenv("AVS_PIPE_NO_MERGE") != nullptr;
            int seg_end0 = 1;   // groups [0, seg_end0) share the first tail
            if (!no_merge && n_excl >= 1)
                while (seg_end0 < G && seg_end0 >= n_excl &&
                       plan.host[2 * slots_all + seg_end0] * 100ll >= plan.host[2 * slots_all] * 86ll) ++seg_end0;
            std::vector<std::pair<int, int>> segs;   // [first group, end group) of every tail
            for (int k = 0; k < launched; k = segs.back().second) {
                segs.push_back({k, std::min(launched, k == 0 ? seg_end0 : k + 1)});
            }
            for (size_t si = 0; si < segs.size() && rc == AVS_OK; ++si) {
                const int k = segs[si].first, k_hi = segs[si].second;
                cudaStream_t ps = k ? m->pipe_stream[k] : st;
                for (int j = k + 1; j < k_hi; ++j) cudaStreamWaitEvent(ps, m->ev_pipe_rec[j], 0);
                const int64_t r0 = glo[k], Rc = ghi[k_hi - 1] - r0;
                // SMs the longer groups' recurrences still hold when this tail runs (the groups behind it have
                // finished): 32 for an exclusive group, up to 32 for a shared one (its 32 CTAs are spread)
                const int free_sms = std::max(sms - 32 * k, 20);
                GemmEpilogue e3;
                e3.bias = m->in_b + 2 * E;
                e3.C = ctx + r0 * E * asz;
                e3.ldc = E;
                e3.out_dtype = act;
                e3.max_ctas = k ? free_sms : 0;
                e3.prefer_pairs = 1;
                if (tail_temporal) {   // q | k | v projection of the segment's rows, then the attention core of its videos
                    GemmEpilogue eq = e3;
                    eq.bias = m->in_b;
                    eq.C = qkv + r0 * 3 * E * dtype_size(qkv_dt);
                    eq.ldc = 3 * E;
                    eq.out_dtype = qkv_dt;
                    rc = run_gemm(precision, fused + r0 * E * asz, act, E, w_in, 0, E, Rc, 3 * E, E, eq, ps);
                    const int n_sq = seq_off[G];
                    SeqDesc sq{seq_dev + seq_off[k], seq_dev + n_sq + seq_off[k], seq_dev + 2 * n_sq + seq_off[k],
                               seq_off[k_hi] - seq_off[k], plan.host[2 * slots_all + k]};
                    if (rc == AVS_OK) rc = attention_tc(qkv, act, R, E, m->heads, sq, ctx, E, act, 0, ps);
                } else {
                    rc = run_gemm(precision, fused + r0 * E * asz, act, E, w_in, 2ll * E * E, E, Rc, E, E, e3, ps);
                }
                GemmEpilogue e5 = e3;
                e5.bias = m->out_b;
                e5.C = attn_out + r0 * E * asz;
                if (rc == AVS_OK) rc = run_gemm(precision, ctx + r0 * E * asz, act, E, w_out, 0, E, Rc, E, E, e5, ps);
                GemmEpilogue e6;
                e6.bias = m->sc0_b;
                e6.relu = 1;
                e6.score_w2 = m->sc2_w;
                e6.score_b2 = m->sc2_b;
                e6.scores = scores_dev + r0;
                e6.max_ctas = k ? free_sms : 0;
                if (rc == AVS_OK) rc = run_gemm(precision, attn_out + r0 * E * asz, act, E, w_sc0, 0, E, Rc, 64, E, e6, ps);
                if (ptrace)
                    for (int j = k; j < k_hi; ++j) cudaEventRecord(tev[2 + 2 * j], ps);
            }
            for (int k = 1; k < launched; ++k) cudaEventRecord(m->ev_pipe_done[k], m->pipe_stream[k]);
            // join every forked stream, also after an error (pipe_stream[0] only ran the stagger kernel, which the
            // other streams waited for)
            for (int k = 1; k < launched; ++k) cudaStreamWaitEvent(st, m->ev_pipe_done[k], 0);
            if (launched < 2) cudaStreamWaitEvent(st, m->ev_pipe_front[1], 0);
            if (prof_g0) {
                cudaEvent_t e = g_prof.get();
                cudaEventRecord(e, st);
                g_prof.pending.push_back({ST_TAIL_EXPOSED, prof_g0, e, 0});
            }
            if (rc != AVS_OK) return rc;
            if (ptrace) {
                cudaDeviceSynchronize();
                fprintf(stderr, "[pipe tail] group: rows, recurrence done / tail done (us after the front)\n");
                for (int k = 0; k < G; ++k) {
                    float a = 0.f, b = 0.f;
                    cudaEventElapsedTime(&a, tev[0], tev[1 + 2 * k]);
                    cudaEventElapsedTime(&b, tev[0], tev[2 + 2 * k]);
                    fprintf(stderr, "  g%d: %6lld rows, maxT %4d  %7.1f / %7.1f\n", k, static_cast<long long>(ghi[k] - glo[k]),
                            plan.host[2 * slots_all + k], a * 1e3f, b * 1e3f);
                }
            }
        }
    }
    if (tail_done) return AVS_OK;

    // ---- K2b: recurrences; writes [v_fwd | v_bwd | a_fwd | a_bwd] = torch.cat of av_model.py:43
    if (!pipelined) {
        const int slots = plan.n_groups * plan.nb;
        LstmBatch lb{plan_dev, plan_dev + slots, plan_dev + 2 * slots, plan.n_groups, plan.nb, lstm_excl};
        int64_t covered = 0;
        for (int b = 0; b < n_videos; ++b) covered += lengths[b];
        if (covered < R)  // padded layout: rows no video owns must stay finite (0 * NaN would poison P*V)
            AVS_CUDA(cudaMemsetAsync(fused, 0, uR * E * asz, st));
        StageTimer tm(ST_LSTM, st);
        if (simt)
            AVS_TRY(lstm_recurrence(reinterpret_cast<const float*>(xg_v), reinterpret_cast<const float*>(xg_a), m->whh, lb,
                                    reinterpret_cast<float*>(fused), 0, nullptr, 0, st));
        else AVS_TRY(lstm_recurrence_tc(xg_v, xg_a, xg_dt, m->whh, lb, act, fused, act, 0, st));
    }

    // ---- K3/K4: nn.MultiheadAttention  av_model.py:44
    if (literal_rows) {
        // sequence length 1: softmax weight == 1, context == value projection
        GemmEpilogue e3;
        e3.bias = m->in_b + 2 * E;
        e3.C = ctx;
        e3.ldc = E;
        e3.out_dtype = act;
        StageTimer tm(ST_QKV_PROJ, st);
        AVS_TRY(run_gemm(precision, fused, act, E, w_in, 2ll * E * E, E, R, E, E, e3, st));
    } else {
        GemmEpilogue e3;
        e3.bias = m->in_b;
        e3.C = qkv;
        e3.ldc = 3 * E;
        e3.out_dtype = qkv_dt;   // q | k | v straight to 16 bit for the tcgen05 attention core
        {
            StageTimer tm(ST_QKV_PROJ, st);
            AVS_TRY(run_gemm(precision, fused, act, E, w_in, 0, E, R, 3 * E, E, e3, st));
        }
        std::vector<int32_t> sd(3 * static_cast<size_t>(n_seqs));
        int seq_max = 0;
        if (attn_axis == AVS_ATTN_TEMPORAL) {
            for (int b = 0; b < n_videos; ++b) {
                sd[b] = row_start[b];
                sd[n_seqs + b] = 1;
                sd[2 * n_seqs + b] = lengths[b];
            }
            seq_max = max_len;
        } else {
            for (int t = 0; t < n_seqs; ++t) {
                sd[t] = row_start[0] + t;
                sd[n_seqs + t] = lengths[0];
                sd[2 * n_seqs + t] = n_videos;
            }
            seq_max = n_videos;
        }
        AVS_TRY(upload_small(seq_dev, sd.data(), sd.size() * 4, st));
        SeqDesc seqs{seq_dev, seq_dev + n_seqs, seq_dev + 2 * n_seqs, n_seqs, seq_max};
        StageTimer tm(ST_ATTENTION, st);
        if (tc_attn) AVS_TRY(attention_tc(qkv, act, R, E, m->heads, seqs, ctx, E, act, 0, st));
        else AVS_TRY(attention_simt(reinterpret_cast<const float*>(qkv), 3 * E, E, m->heads, seqs, ctx, E, act, 0, st));
    }

    // ---- K5: out_proj
    GemmEpilogue e5;
    e5.bias = m->out_b;
    e5.C = attn_out;
    e5.ldc = E;
    e5.out_dtype = act;
    {
        StageTimer tm(ST_OUT_PROJ, st);
        AVS_TRY(run_gemm(precision, ctx, act, E, w_out, 0, E, R, E, E, e5, st));
    }

    // ---- K6: scorer (Linear 1024->64 + ReLU + Linear 64->1 + Sigmoid fused)  av_model.py:29-31,46
    GemmEpilogue e6;
    e6.bias = m->sc0_b;
    e6.relu = 1;
    e6.score_w2 = m->sc2_w;
    e6.score_b2 = m->sc2_b;
    e6.scores = scores_dev;
    {
        StageTimer tm(ST_SCORER, st);
        AVS_TRY(run_gemm(precision, attn_out, act, E, w_sc0, 0, E, R, 64, E, e6, st));
    }

    if (space == AVS_HOST) {
        AVS_CUDA(cudaMemcpyAsync(scores, scores_dev, uR * 4, cudaMemcpyDeviceToHost, st));
        AVS_CUDA(cudaStreamSynchronize(st));
        AVS_TRY(range_report(m));
    }
    return AVS_OK;
}

// Pooling + knapsack in three steps.  summarize_prepare (host only) validates the change points and lays the
// workspace out in the slot's arena; summarize_upload sends the descriptors as kernel parameters (a pageable-memory
// copy would queue on the H2D copy engine behind the feature transfers of a pipelined call) -- it does not depend
// on the scores, so the fused call runs it on a side stream, off the critical path; summarize_run launches the
// kernels and, in host space, the D2H copies.
struct SummaryPlan {
    SummaryBatch sb;
    unsigned long long* seg_sum = nullptr;
    long long* seg_mean_dev = nullptr;
    uint32_t* keep = nullptr;
    long long* dp_ws = nullptr;
    float* sc_stage = nullptr;       // device copies of host-space scores / positions (avs_summarize in host space)
    int32_t* pos_stage = nullptr;
    uint8_t* picks_dev = nullptr;    // device staging of host-space outputs
    uint8_t* summary_dev = nullptr;
    int total_S = 0;
    int64_t rows = 0, sum_bytes = 0;
    bool fuse_pool = false;
    // host copies of the descriptors until summarize_upload has sent them
    std::vector<int32_t> desc;
    std::vector<int64_t> off64;
    int32_t* desc_dev = nullptr;
    int64_t* off_dev = nullptr;
};

static avs_status summarize_prepare(avs_model* m, Arena& A, int32_t n_videos, const int32_t* row_start,
                                    const int32_t* lengths, const int32_t* n_frames, const int32_t* cps,
                                    const int32_t* cps_start, int32_t prop_num, int32_t prop_den, bool want_summary,
                                    const int64_t* summary_start, int in_space, int space, SummaryPlan* P) {
    // in_space: where scores / positions live; space: where picks / seg_mean / summary go
    AVS_CHECK(m != nullptr, AVS_ERR_INVALID, "model handle is null");
    AVS_CHECK(space == AVS_HOST || space == AVS_DEVICE, AVS_ERR_INVALID, "bad memory space %d", space);
    AVS_CHECK(n_videos > 0, AVS_ERR_INVALID, "n_videos must be positive");
    AVS_CHECK(row_start && lengths && n_frames && cps_start, AVS_ERR_INVALID, "a required pointer is null");
    AVS_CHECK(prop_den > 0 && prop_num >= 0, AVS_ERR_INVALID, "bad proportion %d/%d", prop_num, prop_den);
    AVS_CHECK(!want_summary || summary_start != nullptr, AVS_ERR_INVALID, "summary_start is null");
    const int n = n_videos;
    AVS_CHECK(cps_start[0] == 0, AVS_ERR_INVALID, "cps_start[0] must be 0");
    const int total_S = cps_start[n];
    AVS_CHECK(total_S == 0 || cps != nullptr, AVS_ERR_INVALID, "cps is null");
    int64_t rows = 0;
    std::vector<int64_t> off(2 * (static_cast<size_t>(n) + 1), 0);  // keep_start | dp_start
    int max_cap = 0;
    for (int v = 0; v < n; ++v) {
        AVS_CHECK(lengths[v] >= 0 && row_start[v] >= 0 && n_frames[v] >= 0, AVS_ERR_INVALID, "video %d: bad sizes", v);
        AVS_CHECK(cps_start[v + 1] >= cps_start[v], AVS_ERR_INVALID, "cps_start not monotone at %d", v);
        rows = std::max<int64_t>(rows, static_cast<int64_t>(row_start[v]) + lengths[v]);
        int prev_end = -1;
        for (int s = cps_start[v]; s < cps_start[v + 1]; ++s) {
            const int a = cps[2 * s], b = cps[2 * s + 1];
            AVS_CHECK(a > prev_end && b >= a && b < n_frames[v], AVS_ERR_INVALID,
                      "video %d shot %d = [%d, %d]: change points must be sorted, disjoint, inclusive and inside "
                      "[0, n_frames)", v, s - cps_start[v], a, b);
            prev_end = b;
        }
        const long long cap = (static_cast<long long>(n_frames[v]) * prop_num) / prop_den;
        AVS_CHECK(cap < (1ll << 30), AVS_ERR_UNSUPPORTED, "video %d: capacity too large", v);
        max_cap = std::max(max_cap, static_cast<int>(cap));
        const int S = cps_start[v + 1] - cps_start[v];
        off[v + 1] = off[v] + static_cast<int64_t>(S) * knapsack_keep_words(cap);
        off[n + 1 + v + 1] = off[n + 1 + v] + 2 * (cap + 1);
        if (want_summary) AVS_CHECK(summary_start[v + 1] - summary_start[v] >= n_frames[v], AVS_ERR_INVALID,
                                    "summary_start leaves fewer than n_frames bytes for video %d", v);
    }
    const int64_t keep_words = off[n];
    const bool dp_global = 2ull * (static_cast<size_t>(max_cap) + 1) * 8 > 200 * 1024;
    const int64_t dp_elems = dp_global ? off[2 * n + 1] : 0;
    const int64_t sum_bytes = want_summary ? summary_start[n] : 0;

    // int32 descriptor block: row_start | lengths | n_frames | cps_start | cps
    std::vector<int32_t>& desc = P->desc;
    desc.clear();
    desc.insert(desc.end(), row_start, row_start + n);
    desc.insert(desc.end(), lengths, lengths + n);
    desc.insert(desc.end(), n_frames, n_frames + n);
    desc.insert(desc.end(), cps_start, cps_start + n + 1);
    if (desc.size() % 2) desc.push_back(0);  // keep cps 8-byte aligned for int2 loads
    const size_t cps_off = desc.size();
    desc.insert(desc.end(), cps, cps + 2 * static_cast<size_t>(total_S));
    std::vector<int64_t>& off64 = P->off64;
    off64 = off;
    if (want_summary) off64.insert(off64.end(), summary_start, summary_start + n + 1);

    size_t need = desc.size() * 4 + off64.size() * 8 + static_cast<size_t>(total_S) * (8 + 8 + 1) +
                  static_cast<size_t>(keep_words) * 4 + static_cast<size_t>(dp_elems) * 8 + 64 * 256;
    if (in_space == AVS_HOST) need += static_cast<size_t>(rows) * 8 + 512;
    if (space == AVS_HOST) need += static_cast<size_t>(sum_bytes) + static_cast<size_t>(total_S) + 512;
    AVS_TRY(A.reserve(need));
    A.reset();
    int64_t* off_dev = A.take<int64_t>(off64.size());
    int32_t* desc_dev = A.take<int32_t>(desc.size());
    P->seg_sum = A.take<unsigned long long>(std::max(total_S, 1));
    P->seg_mean_dev = A.take<long long>(std::max(total_S, 1));
    P->keep = A.take<uint32_t>(std::max<int64_t>(keep_words, 1));
    P->dp_ws = A.take<long long>(std::max<int64_t>(dp_elems, 1));
    if (in_space == AVS_HOST) {
        P->sc_stage = A.take<float>(rows);
        P->pos_stage = A.take<int32_t>(rows);
    }
    if (space == AVS_HOST) {
        P->picks_dev = A.take<uint8_t>(std::max(total_S, 1));
        if (want_summary) P->summary_dev = A.take<uint8_t>(sum_bytes);
    }
    P->off_dev = off_dev;
    P->desc_dev = desc_dev;

    SummaryBatch& sb = P->sb;
    sb.row_start = desc_dev;
    sb.lengths = desc_dev + n;
    sb.n_frames = desc_dev + 2 * n;
    sb.cps_start = desc_dev + 3 * n;
    sb.cps = desc_dev + cps_off;
    sb.keep_start = off_dev;
    sb.dp_start = off_dev + (n + 1);
    sb.summary_start = want_summary ? off_dev + 2 * (n + 1) : nullptr;
    sb.n = n;
    sb.prop_num = prop_num;
    sb.prop_den = prop_den;
    sb.max_cap = max_cap;
    sb.max_S = 0;
    for (int v = 0; v < n; ++v) sb.max_S = std::max(sb.max_S, cps_start[v + 1] - cps_start[v]);
    sb.max_wt = 0;
    for (int i = 0; i < total_S; ++i) sb.max_wt = std::max(sb.max_wt, cps[2 * i + 1] - cps[2 * i] + 1);
    // K7 + K8 in one launch when every video's shots fit in shared memory (always, for TVSum/SumMe-sized inputs)
    P->fuse_pool = knapsack_can_fuse_pool(sb);
    P->total_S = total_S;
    P->rows = rows;
    P->sum_bytes = sum_bytes;
    return AVS_OK;
}

static avs_status summarize_upload(const SummaryPlan& P, cudaStream_t st) {
    AVS_TRY(upload_small(P.off_dev, P.off64.data(), P.off64.size() * 8, st));
    return upload_small(P.desc_dev, P.desc.data(), P.desc.size() * 4, st);
}

// the uploads run on the side stream as soon as the arena's previous user has finished; st then waits for them
static avs_status summarize_upload_side(avs_model* m, avs_model::SumArena& sa, const SummaryPlan& P, cudaStream_t st) {
    if (sa.used) AVS_CUDA(cudaStreamWaitEvent(m->sum_stream, sa.done, 0));
    AVS_TRY(summarize_upload(P, m->sum_stream));
    AVS_CUDA(cudaEventRecord(m->ev_sum_out, m->sum_stream));
    AVS_CUDA(cudaStreamWaitEvent(st, m->ev_sum_out, 0));
    return AVS_OK;
}
static avs_status summarize_mark_done(avs_model::SumArena& sa, cudaStream_t st) {
    AVS_CUDA(cudaEventRecord(sa.done, st));
    sa.used = true;
    return AVS_OK;
}

// sync: end with a stream synchronisation (host space); otherwise the caller records its own completion event
static avs_status summarize_run(const SummaryPlan& P, const float* scores, const int32_t* positions, uint8_t* picks,
                                int64_t* seg_mean, uint8_t* summary, int in_space, int space, cudaStream_t st,
                                bool sync) {
    AVS_CHECK(scores && positions && picks, AVS_ERR_INVALID, "a required pointer is null");
    const float* sc = scores;
    const int32_t* pos = positions;
    if (in_space == AVS_HOST) {
        AVS_CUDA(cudaMemcpyAsync(P.sc_stage, scores, P.rows * 4, cudaMemcpyHostToDevice, st));
        AVS_CUDA(cudaMemcpyAsync(P.pos_stage, positions, P.rows * 4, cudaMemcpyHostToDevice, st));
        sc = P.sc_stage;
        pos = P.pos_stage;
    }
    uint8_t* picks_dev = space == AVS_HOST ? P.picks_dev : picks;
    uint8_t* summary_dev = space == AVS_HOST ? (summary ? P.summary_dev : nullptr) : summary;
    if (!P.fuse_pool) {
        AVS_CUDA(cudaMemsetAsync(P.seg_sum, 0, static_cast<size_t>(std::max(P.total_S, 1)) * 8, st));
        StageTimer tm(ST_POOL, st);
        AVS_TRY(shot_pool(sc, pos, P.sb, P.seg_sum, st));
    }
    long long* seg_mean_target = space == AVS_HOST ? (seg_mean ? P.seg_mean_dev : nullptr)
                                                    : reinterpret_cast<long long*>(seg_mean);
    {
        StageTimer tm(ST_KNAPSACK, st);
        AVS_TRY(knapsack_select(P.sb, P.seg_sum, seg_mean_target, picks_dev, summary_dev, P.keep, P.dp_ws, st,
                                P.fuse_pool ? sc : nullptr, pos));
    }
    const bool trace = g_e2e.enabled() && space == AVS_HOST;
    if (trace) cudaEventRecord(g_e2e.knap, st);
    if (space == AVS_HOST) {
        AVS_CUDA(cudaMemcpyAsync(picks, picks_dev, P.total_S, cudaMemcpyDeviceToHost, st));
        if (seg_mean) AVS_CUDA(cudaMemcpyAsync(seg_mean, P.seg_mean_dev, static_cast<size_t>(P.total_S) * 8,
                                               cudaMemcpyDeviceToHost, st));
        if (summary) AVS_CUDA(cudaMemcpyAsync(summary, summary_dev, P.sum_bytes, cudaMemcpyDeviceToHost, st));
        if (trace) {
            cudaEventRecord(g_e2e.end, st);
            g_e2e.host[3] = E2ETrace::now();
        }
        if (sync) {
            AVS_CUDA(cudaStreamSynchronize(st));
            if (trace) g_e2e.host[4] = E2ETrace::now();
        }
    }
    return AVS_OK;
}

avs_status avs_summarize(avs_model* m, const float* scores, const int32_t* positions, int32_t n_videos,
                         const int32_t* row_start, const int32_t* lengths, const int32_t* n_frames, const int32_t* cps,
                         const int32_t* cps_start, int32_t prop_num, int32_t prop_den, uint8_t* picks,
                         int64_t* seg_mean, uint8_t* summary, const int64_t* summary_start, int space,
                         void* cuda_stream) {
    AVS_CHECK(m != nullptr, AVS_ERR_INVALID, "model handle is null");
    AVS_CHECK(n_videos >= 0, AVS_ERR_INVALID, "n_videos negative");
    if (n_videos == 0) return AVS_OK;
    Guard g(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    avs_model::SumArena& sa = m->sum_ws[avs_model::SLOTS + (m->sum_flip++ & 1)];
    SummaryPlan P;
    AVS_TRY(summarize_prepare(m, sa.a, n_videos, row_start, lengths, n_frames, cps, cps_start, prop_num, prop_den,
                              summary != nullptr, summary_start, space, space, &P));
    AVS_TRY(summarize_upload_side(m, sa, P, st));
    AVS_TRY(summarize_run(P, scores, positions, picks, seg_mean, summary, space, space, st, true));
    return summarize_mark_done(sa, st);
}

static avs_status forward_summarize_impl(avs_model* m, const float* visual, const float* audio,
                                         const int32_t* positions, int64_t total_rows, int32_t n_videos,
                                         const int32_t* row_start, const int32_t* lengths, int attn_axis,
                                         int precision, const int32_t* n_frames, const int32_t* cps,
                                         const int32_t* cps_start, int32_t prop_num, int32_t prop_den, float* scores,
                                         uint8_t* picks, int64_t* seg_mean, uint8_t* summary,
                                         const int64_t* summary_start, int space, void* cuda_stream, int slot,
                                         bool async) {
    AVS_CHECK(m != nullptr, AVS_ERR_INVALID, "model handle is null");
    AVS_CHECK(space == AVS_HOST || space == AVS_DEVICE, AVS_ERR_INVALID, "bad memory space %d", space);
    AVS_CHECK(positions != nullptr && scores != nullptr, AVS_ERR_INVALID, "positions / scores pointer is null");
    AVS_CHECK(n_videos >= 0, AVS_ERR_INVALID, "n_videos negative");
    AVS_CHECK(slot >= 0 && slot < avs_model::SLOTS, AVS_ERR_INVALID, "bad slot %d", slot);
    AVS_CHECK(!m->slot_busy[slot], AVS_ERR_INVALID, "slot %d is in flight: call avs_slot_wait first", slot);
    AVS_CHECK(!async || space == AVS_HOST, AVS_ERR_INVALID, "the asynchronous call is for host-space (pinned) buffers");
    Guard g(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (g_e2e.enabled() && space == AVS_HOST) g_e2e.host[0] = E2ETrace::now();
    int max_len = 0;
    AVS_TRY(validate_videos(total_rows, n_videos, row_start, lengths, &max_len));
    avs_model::SumArena& sa = m->sum_ws[async ? slot : avs_model::SLOTS + (m->sum_flip++ & 1)];
    SummaryPlan P;
    if (n_videos > 0)   // validate everything before any work is queued
        AVS_TRY(summarize_prepare(m, sa.a, n_videos, row_start, lengths, n_frames, cps, cps_start, prop_num, prop_den,
                                  summary != nullptr, summary_start, AVS_DEVICE, space, &P));
    if (space == AVS_DEVICE) {
        AVS_TRY(forward_entry(m, visual, audio, total_rows, n_videos, row_start, lengths, attn_axis, precision, scores,
                              AVS_DEVICE, cuda_stream, nullptr, nullptr, nullptr));
        if (n_videos == 0) return AVS_OK;
        AVS_TRY(summarize_upload_side(m, sa, P, st));
        AVS_TRY(summarize_run(P, scores, positions, picks, seg_mean, summary, AVS_DEVICE, AVS_DEVICE, st, true));
        return summarize_mark_done(sa, st);
    }
    // host space: features in, scores stay on the device for pooling + knapsack, everything comes back with ONE
    // synchronisation at the end (no D2H -> H2D round trip of the scores between the two halves)
    if (n_videos == 0) return AVS_OK;
    float* sc_dev = nullptr;
    int32_t* pos_dev = nullptr;
    if (total_rows == 0 || max_len == 0) {   // nothing to score: every shot pools to zero
        Arena& HI = m->host_in[slot];
        AVS_TRY(HI.reserve(1024));
        HI.reset();
        sc_dev = HI.take<float>(1);
        pos_dev = HI.take<int32_t>(1);
    } else {
        AVS_TRY(forward_entry(m, visual, audio, total_rows, n_videos, row_start, lengths, attn_axis, precision, nullptr,
                              AVS_HOST, cuda_stream, &sc_dev, positions, &pos_dev, slot, async));
        AVS_CUDA(cudaStreamWaitEvent(st, m->ev_chunk[0], 0));   // positions travelled on the copy stream (grouped path)
        AVS_CUDA(cudaMemcpyAsync(scores, sc_dev, static_cast<size_t>(total_rows) * 4, cudaMemcpyDeviceToHost, st));
    }
    // the descriptor uploads go to the side stream, queued (on the host) after the feature copies and the groups'
    // kernels: they need nothing but the workspace
    AVS_TRY(summarize_upload_side(m, sa, P, st));
    AVS_TRY(summarize_run(P, sc_dev, pos_dev, picks, seg_mean, summary, AVS_DEVICE, AVS_HOST, st, !async));
    AVS_TRY(summarize_mark_done(sa, st));
    if (async) {
        AVS_CUDA(cudaEventRecord(m->ev_slot[slot], st));
        m->slot_busy[slot] = true;
        return AVS_OK;     // avs_slot_wait reports the range watch
    }
    return range_report(m);   // the stream is synchronised: the forward of this call has completed
}

avs_status avs_forward_summarize(avs_model* m, const float* visual, const float* audio, const int32_t* positions,
                                 int64_t total_rows, int32_t n_videos, const int32_t* row_start,
                                 const int32_t* lengths, int attn_axis, int precision, const int32_t* n_frames,
                                 const int32_t* cps, const int32_t* cps_start, int32_t prop_num, int32_t prop_den,
                                 float* scores, uint8_t* picks, int64_t* seg_mean, uint8_t* summary,
                                 const int64_t* summary_start, int space, void* cuda_stream) {
    return forward_summarize_impl(m, visual, audio, positions, total_rows, n_videos, row_start, lengths, attn_axis,
                                  precision, n_frames, cps, cps_start, prop_num, prop_den, scores, picks, seg_mean,
                                  summary, summary_start, space, cuda_stream, 0, false);
}

avs_status avs_forward_summarize_async(avs_model* m, const float* visual, const float* audio,
                                       const int32_t* positions, int64_t total_rows, int32_t n_videos,
                                       const int32_t* row_start, const int32_t* lengths, int attn_axis, int precision,
                                       const int32_t* n_frames, const int32_t* cps, const int32_t* cps_start,
                                       int32_t prop_num, int32_t prop_den, float* scores, uint8_t* picks,
                                       int64_t* seg_mean, uint8_t* summary, const int64_t* summary_start, int slot,
                                       void* cuda_stream) {
    return forward_summarize_impl(m, visual, audio, positions, total_rows, n_videos, row_start, lengths, attn_axis,
                                  precision, n_frames, cps, cps_start, prop_num, prop_den, scores, picks, seg_mean,
                                  summary, summary_start, AVS_HOST, cuda_stream, slot, true);
}

/* fp16 range watch for DEVICE-space callers (host-space calls report by themselves): *saturated = 1 when, since the last
 * report, an fc activation of a forward on this handle reached the fp16 limit and was clamped.  Reads and clears the
 * flag; the caller has synchronised the stream(s) of the forwards it asks about.  No reference counterpart (the
 * reference computes in fp32). */
avs_status avs_model_range_status(avs_model* m, int32_t* saturated) {
    AVS_CHECK(m != nullptr && saturated != nullptr, AVS_ERR_INVALID, "avs_model_range_status: null pointer");
    *saturated = 0;
    if (m->range_flag != nullptr) {
        volatile unsigned int* f = m->range_flag;
        *saturated = *f != 0u ? 1 : 0;
        *f = 0u;
    }
    return AVS_OK;
}

avs_status avs_slot_wait(avs_model* m, int slot) {
    AVS_CHECK(m != nullptr, AVS_ERR_INVALID, "model handle is null");
    AVS_CHECK(slot >= 0 && slot < avs_model::SLOTS, AVS_ERR_INVALID, "bad slot %d", slot);
    if (!m->slot_busy[slot]) return AVS_OK;
    Guard g(m->device);
    AVS_CUDA(cudaEventSynchronize(m->ev_slot[slot]));
    m->slot_busy[slot] = false;
    return range_report(m);
}

avs_status avs_linear(const float* A, const float* W, const float* bias, int64_t M, int32_t N, int32_t K, int relu,
                      int precision, float* C, void* cuda_stream) {
    AVS_CHECK(A && W && C, AVS_ERR_INVALID, "avs_linear: null pointer");
    AVS_CHECK(precision_ok(precision), AVS_ERR_INVALID, "bad precision %d", precision);
    GemmEpilogue e;
    e.bias = bias;
    e.C = C;
    e.ldc = N;
    e.relu = relu;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (precision == AVS_PREC_FP32_SIMT) return gemm_simt(A, K, W, K, M, N, K, e, st);
    if (precision == AVS_PREC_TF32) {
        // ONE operand policy for kind::tf32 everywhere (avs_forward's fc layers use the same): the weight is rounded
        // to nearest (a copy made for this call), the activation is read as it is -- the tensor core truncates its
        // low 13 mantissa bits, x -> x (1 - d) with d in [0, 2^-10) -- and the -2^-11 mean of that truncation is
        // scaled out of the accumulator.  Assumption: the low mantissa bits of A are uniformly distributed (any
        // fp32 data that is not already exact in tf32); data that IS tf32-exact (fp16 / bf16 values upcast to fp32)
        // sees +4.9e-4 relative on the product instead of 0, inside the 1e-3 budget
        // (tests/test_gpu_parity.py::test_forward_with_fp16_exact_features).
        AVS_CHECK(M >= 0 && N > 0 && K > 0, AVS_ERR_INVALID, "avs_linear: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
        if (M == 0) return AVS_OK;
        float* w_rn = nullptr;
        AVS_CUDA(cudaMallocAsync(&w_rn, static_cast<size_t>(N) * K * 4, st));
        avs_status s = convert_f32(W, w_rn, static_cast<int64_t>(N) * K, DT_F32, 1, st);
        e.acc_scale = 1.0f + 1.0f / 2048.0f;
        if (s == AVS_OK) s = gemm_tc(A, K, w_rn, K, DT_F32, M, N, K, e, st);
        cudaFreeAsync(w_rn, st);
        return s;
    }
    // bf16 operands: cast both sides for this call
    AVS_CHECK(M >= 0 && N > 0 && K > 0 && K % 8 == 0, AVS_ERR_UNSUPPORTED,
              "avs_linear[bf16]: K=%d must be a multiple of 8 (16-byte TMA row pitch)", K);
    if (M == 0) return AVS_OK;
    uint16_t* tmp = nullptr;
    AVS_CUDA(cudaMallocAsync(&tmp, (static_cast<size_t>(M) + N) * K * 2, st));
    avs_status s = convert_f32(A, tmp, M * K, DT_BF16, 0, st);
    if (s == AVS_OK) s = convert_f32(W, tmp + M * K, static_cast<int64_t>(N) * K, DT_BF16, 0, st);
    if (s == AVS_OK) s = gemm_tc(tmp, K, tmp + M * K, K, DT_BF16, M, N, K, e, st);
    cudaFreeAsync(tmp, st);
    return s;
}

static avs_status bilstm_pair_impl(avs_model* m, const float* v_emb, const float* a_emb, int64_t total_rows,
                                   int32_t n_videos, const int32_t* row_start, const int32_t* lengths, int precision,
                                   float* fused, void* save_pre, float* save_c, void* cuda_stream) {
    AVS_CHECK(m && v_emb && a_emb && fused, AVS_ERR_INVALID, "avs_bilstm_pair: null pointer");
    AVS_CHECK(precision_ok(precision), AVS_ERR_INVALID, "bad precision %d", precision);
    int max_len = 0;
    AVS_TRY(validate_videos(total_rows, n_videos, row_start, lengths, &max_len));
    if (total_rows == 0 || max_len == 0) return AVS_OK;
    Guard g(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int64_t R = total_rows;
    const bool simt = precision == AVS_PREC_FP32_SIMT;
    const int act = simt ? DT_F32 : (precision == AVS_PREC_BF16 ? DT_BF16 : DT_F16);
    LstmPlan plan = plan_lstm(n_videos, row_start, lengths, !simt);
    AVS_TRY(m->ws.reserve(static_cast<size_t>(R) * (2 * 2 * G4 + 2 * H) * 4 + plan.host.size() * 4 + 16 * 256));
    m->ws.reset();
    float* xg_v = m->ws.take<float>(R * 2 * G4);
    float* xg_a = m->ws.take<float>(R * 2 * G4);
    int32_t* plan_dev = m->ws.take<int32_t>(plan.host.size());
    const void* xv = v_emb;
    const void* xa = a_emb;
    if (!simt) {   // the embeddings reach the input projection as 16-bit operands, exactly as in avs_forward
        uint16_t* rv = m->ws.take<uint16_t>(R * H);
        uint16_t* ra = m->ws.take<uint16_t>(R * H);
        AVS_TRY(convert_f32(v_emb, rv, R * H, act, 0, st));
        AVS_TRY(convert_f32(a_emb, ra, R * H, act, 0, st));
        xv = rv;
        xa = ra;
    }
    // kernel-parameter upload: no pageable-memory copy (host-blocking, and not capturable in a CUDA graph)
    AVS_TRY(upload_small(plan_dev, plan.host.data(), plan.host.size() * 4, st));
    const GemmW w_ih_v{m->ih_v_x, m->ih_v_t, m->ih_v_l}, w_ih_a{m->ih_a_x, m->ih_a_t, m->ih_a_l};
    GemmEpilogue e2;
    e2.ldc = 2 * G4;
    // the two input projections are independent: the audio one runs on the branch stream (fork / join, capturable)
    cudaStream_t sa = m->branch_stream[0];
    AVS_CUDA(cudaEventRecord(m->ev_branch_in[0], st));
    AVS_CUDA(cudaStreamWaitEvent(sa, m->ev_branch_in[0], 0));
    e2.bias = m->ih_v_b;
    e2.C = xg_v;
    avs_status sv = run_gemm(precision, xv, act, H, w_ih_v, 0, H, R, 2 * G4, H, e2, st);
    e2.bias = m->ih_a_b;
    e2.C = xg_a;
    avs_status sa_rc = run_gemm(precision, xa, act, H, w_ih_a, 0, H, R, 2 * G4, H, e2, sa);
    cudaEventRecord(m->ev_branch_out[0], sa);
    cudaStreamWaitEvent(st, m->ev_branch_out[0], 0);
    AVS_TRY(sv);
    AVS_TRY(sa_rc);
    const int slots = plan.n_groups * plan.nb;
    LstmBatch lb{plan_dev, plan_dev + slots, plan_dev + 2 * slots, plan.n_groups, plan.nb};
    if (!simt) return lstm_recurrence_tc(xg_v, xg_a, DT_F32, m->whh, lb, act, fused, DT_F32, 0, st, save_pre, save_c);
    AVS_CHECK(save_pre == nullptr, AVS_ERR_UNSUPPORTED, "the training forward needs a tensor-core precision");
    return lstm_recurrence(xg_v, xg_a, m->whh, lb, fused, 0, nullptr, 0, st);
}

avs_status avs_bilstm_pair(avs_model* m, const float* v_emb, const float* a_emb, int64_t total_rows, int32_t n_videos,
                           const int32_t* row_start, const int32_t* lengths, int precision, float* fused,
                           void* cuda_stream) {
    return bilstm_pair_impl(m, v_emb, a_emb, total_rows, n_videos, row_start, lengths, precision, fused, nullptr, nullptr,
                            cuda_stream);
}

// ---- training step building blocks (scripts/train_av_model.py:86-96) ---------------------------------
avs_status avs_bilstm_pair_train(avs_model* m, const float* v_emb, const float* a_emb, int64_t total_rows,
                                 int32_t n_videos, const int32_t* row_start, const int32_t* lengths, float* fused,
                                 float* save_pre, float* save_c, void* cuda_stream) {
    AVS_CHECK(save_pre && save_c, AVS_ERR_INVALID, "avs_bilstm_pair_train: null save buffer");
    return bilstm_pair_impl(m, v_emb, a_emb, total_rows, n_videos, row_start, lengths, AVS_PREC_TF32, fused, save_pre,
                            save_c, cuda_stream);
}

namespace {
// Side streams for the backward pass: its GEMMs have few output tiles (weight gradients: 16-32 CTAs each) and are
// independent of one another, so they run side by side instead of one after the other.  Fork / join with events, all
// inside one call, so the pattern is capturable in a CUDA graph (training.TrainStep).  One set per device; the
// library serialises calls per handle / per calling thread.
struct SideStreams {
    static constexpr int N = 4;
    cudaStream_t s[N] = {};
    cudaEvent_t fork = nullptr, join[N] = {}, aux[2] = {};
    bool ok = false, tried = false;
};
SideStreams& side_streams() {
    static SideStreams per_device[kMaxDevices];
    SideStreams& c = per_device[current_device()];
    if (!c.tried) {
        c.tried = true;
        bool good = cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < SideStreams::N && good; ++i)
            good = cudaStreamCreateWithFlags(&c.s[i], cudaStreamNonBlocking) == cudaSuccess &&
                   cudaEventCreateWithFlags(&c.join[i], cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < 2 && good; ++i) good = cudaEventCreateWithFlags(&c.aux[i], cudaEventDisableTiming) == cudaSuccess;
        c.ok = good && getenv("AVS_BWD_ONE_STREAM") == nullptr;
    }
    return c;
}
// C[M, N] = A[M, K] * Wt[N, K]^T on the tensor cores (kind::tf32), all operands pre-rounded fp32
avs_status gemm_nt_tf32(const float* A, int64_t lda, const float* Wt, int64_t ldw, int64_t M, int N, int K, float* C,
                        int64_t ldc, cudaStream_t st) {
    GemmEpilogue e;
    e.C = C;
    e.ldc = ldc;
    return gemm_tc(A, lda, Wt, ldw, DT_F32, M, N, K, e, st);
}
int64_t pad4(int64_t x) { return (x + 3) / 4 * 4; }
}  // namespace

avs_status avs_linear_bwd(const float* dY, const float* X, const float* W, int64_t M, int32_t N, int32_t K, float* dX,
                          float* dW, float* db, void* cuda_stream) {
    AVS_CHECK(dY != nullptr && M >= 0 && N > 0 && K > 0, AVS_ERR_INVALID, "avs_linear_bwd: bad arguments");
    AVS_CHECK(M < (1ll << 31), AVS_ERR_UNSUPPORTED, "avs_linear_bwd: too many rows");
    AVS_CHECK((dX == nullptr || W != nullptr) && (dW == nullptr || X != nullptr), AVS_ERR_INVALID,
              "avs_linear_bwd: dX needs W and dW needs X");
    AVS_CHECK(N % 4 == 0 && K % 4 == 0, AVS_ERR_UNSUPPORTED,
              "avs_linear_bwd: N=%d and K=%d must be multiples of 4 (16-byte TMA row pitch)", N, K);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (M == 0) {
        if (dW) AVS_CUDA(cudaMemsetAsync(dW, 0, static_cast<size_t>(N) * K * 4, st));
        if (db) AVS_CUDA(cudaMemsetAsync(db, 0, static_cast<size_t>(N) * 4, st));
        return AVS_OK;
    }
    const int64_t Mp = pad4(M);
    size_t bytes = 0;
    const size_t o_dyr = bytes;  bytes += dX ? align_up(static_cast<size_t>(M) * N * 4, 256) : 0;
    const size_t o_wt = bytes;   bytes += dX ? align_up(static_cast<size_t>(K) * N * 4, 256) : 0;
    const size_t o_dyt = bytes;  bytes += dW ? align_up(static_cast<size_t>(N) * Mp * 4, 256) : 0;
    const size_t o_xt = bytes;   bytes += dW ? align_up(static_cast<size_t>(K) * Mp * 4, 256) : 0;
    char* ws = nullptr;
    if (bytes) AVS_CUDA(cudaMallocAsync(&ws, bytes, st));
    avs_status s = AVS_OK;
    // the weight-gradient branch (dW, db) runs beside the input-gradient branch (dX)
    SideStreams& sd = side_streams();
    const bool par = sd.ok && dX != nullptr && (dW != nullptr || db != nullptr);
    cudaStream_t sw = par ? sd.s[0] : st;
    if (par) {
        AVS_CUDA(cudaEventRecord(sd.fork, st));
        AVS_CUDA(cudaStreamWaitEvent(sw, sd.fork, 0));
    }
    if (dX) {   // dX[M, K] = dY[M, N] * W[N, K]  ==  dY * (W^T)^T
        float* dyr = reinterpret_cast<float*>(ws + o_dyr);
        float* wt = reinterpret_cast<float*>(ws + o_wt);
        if (s == AVS_OK) s = convert_f32(dY, dyr, M * N, DT_F32, 1, st);
        if (s == AVS_OK) s = transpose_f32(W, K, N, K, wt, N, 0, 1, st);
        if (s == AVS_OK) s = gemm_nt_tf32(dyr, N, wt, N, M, K, N, dX, K, st);
    }
    if (dW) {   // dW[N, K] = dY^T[N, M] * X[M, K]  ==  (dY^T) * (X^T)^T
        float* dyt = reinterpret_cast<float*>(ws + o_dyt);
        float* xt = reinterpret_cast<float*>(ws + o_xt);
        if (s == AVS_OK) s = transpose_f32(dY, N, static_cast<int>(M), N, dyt, Mp, 0, 1, sw);
        if (s == AVS_OK) s = transpose_f32(X, K, static_cast<int>(M), K, xt, Mp, 0, 1, sw);
        if (s == AVS_OK) s = gemm_nt_tf32(dyt, Mp, xt, Mp, N, K, static_cast<int>(M), dW, K, sw);
    }
    if (db && s == AVS_OK) s = colsum_f32(dY, N, static_cast<int>(M), N, db, 0, sw);
    if (par) {   // join even after an error: a forked stream must never be left dangling (graph capture)
        cudaEventRecord(sd.join[0], sw);
        cudaStreamWaitEvent(st, sd.join[0], 0);
    }
    if (ws) cudaFreeAsync(ws, st);
    return s;
}

avs_status avs_bilstm_pair_bwd(avs_model* m, const float* d_fused, const float* save_pre, const float* save_c,
                               const float* fused, const float* v_emb, const float* a_emb, int64_t total_rows,
                               int32_t n_videos, const int32_t* row_start, const int32_t* lengths, float* d_v_emb,
                               float* d_a_emb, float* const* dW_ih, float* const* dW_hh, float* const* db,
                               void* cuda_stream) {
    AVS_CHECK(m && d_fused && save_pre && save_c && fused && v_emb && a_emb && d_v_emb && d_a_emb && dW_ih && dW_hh && db,
              AVS_ERR_INVALID, "avs_bilstm_pair_bwd: null pointer");
    for (int i = 0; i < 4; ++i)
        AVS_CHECK(dW_ih[i] && dW_hh[i] && db[i], AVS_ERR_INVALID, "avs_bilstm_pair_bwd: null gradient pointer");
    int max_len = 0;
    AVS_TRY(validate_videos(total_rows, n_videos, row_start, lengths, &max_len));
    Guard g(m->device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const int64_t R = total_rows;
    const size_t uR = static_cast<size_t>(R);
    if (R == 0 || max_len == 0) {
        for (int i = 0; i < 4; ++i) {
            AVS_CUDA(cudaMemsetAsync(dW_ih[i], 0, static_cast<size_t>(G4) * H * 4, st));
            AVS_CUDA(cudaMemsetAsync(dW_hh[i], 0, static_cast<size_t>(G4) * HC * 4, st));
            AVS_CUDA(cudaMemsetAsync(db[i], 0, static_cast<size_t>(G4) * 4, st));
        }
        return AVS_OK;
    }
    // ---- plan: up to 8 videos per cluster, longest first (any number of groups: clusters are independent)
    std::vector<int> order;
    for (int b = 0; b < n_videos; ++b)
        if (lengths[b] > 0) order.push_back(b);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return lengths[a] > lengths[b]; });
    const int B = static_cast<int>(order.size());
    // Videos per cluster: as few as still fit the batch into ONE wave of clusters.  A group is 4 recurrences x 8 CTAs
    // with one CTA per SM (the kernel keeps its W_hh^T slice in 256 of the SM's 512 tensor-memory columns), and
    // clusters are placed inside one GPC: 12 such clusters are resident at once on B200, 16 are not (measured: with
    // 8 videos as 4 groups of 2 the kernel took two waves, 580 us instead of 300) -- so at most 3 groups.  The
    // exchange of a step is bound by the number of st.async messages an SM receives (64 per video of the
    // cluster), which is why fewer videos per cluster are preferred when they fit.
    int nb = 8;
    for (int c : {1, 2, 4, 8})
        if ((B + c - 1) / c <= 3) { nb = c; break; }
    if (const char* e = getenv("AVS_BPTT_NB")) {   // experiment switch
        const int c = atoi(e);
        if (c == 1 || c == 2 || c == 4 || c == 8) nb = c;
    }
    const int n_groups = (B + nb - 1) / nb;
    const int slots = n_groups * nb;
    std::vector<int32_t> plan(2 * slots + n_groups + 2 * static_cast<size_t>(n_videos), 0);
    for (int i = 0; i < B; ++i) {
        plan[i] = row_start[order[i]];
        plan[slots + i] = lengths[order[i]];
    }
    for (int gidx = 0; gidx < n_groups; ++gidx) plan[2 * slots + gidx] = lengths[order[gidx * nb]];
    const size_t desc_off = 2 * slots + n_groups;
    for (int b = 0; b < n_videos; ++b) {
        plan[desc_off + b] = row_start[b];
        plan[desc_off + n_videos + b] = lengths[b];
    }
    // ---- workspace
    const int64_t Rp = pad4(R);
    const size_t need = 2 * uR * 2 * G4 * 4      /* d_xg_v, d_xg_a */
                        + 2 * uR * 2 * G4 * 4      /* rounded copies of d_xg (one per modality) */
                        + uR * E * 4               /* hprev */
                        + 2 * 2ull * G4 * Rp * 4   /* d_xg^T (un-permuted rows), per modality */
                        + static_cast<size_t>(E) * Rp * 4      /* hprev^T */
                        + 2 * static_cast<size_t>(H) * Rp * 4  /* emb^T, per modality */
                        + 2 * static_cast<size_t>(H) * 2 * G4 * 4 /* W_ih^T, per modality */
                        + 2 * 2 * G4 * 4 + plan.size() * 4 + 48 * 256;
    AVS_TRY(m->ws.reserve(need));
    m->ws.reset();
    float* d_xg_v = m->ws.take<float>(uR * 2 * G4);
    float* d_xg_a = m->ws.take<float>(uR * 2 * G4);
    float* hprev = m->ws.take<float>(uR * E);
    float* hprev_t = m->ws.take<float>(static_cast<size_t>(E) * Rp);
    float *d_xg_r[2], *dxg_t[2], *emb_t[2], *wih_t[2], *db_tmp[2];
    for (int mod = 0; mod < 2; ++mod) {   // one set per modality: the two branches run concurrently
        d_xg_r[mod] = m->ws.take<float>(uR * 2 * G4);
        dxg_t[mod] = m->ws.take<float>(2ull * G4 * Rp);
        emb_t[mod] = m->ws.take<float>(static_cast<size_t>(H) * Rp);
        wih_t[mod] = m->ws.take<float>(static_cast<size_t>(H) * 2 * G4);
        db_tmp[mod] = m->ws.take<float>(2 * G4);
    }
    int32_t* plan_dev = m->ws.take<int32_t>(plan.size());
    AVS_TRY(upload_small(plan_dev, plan.data(), plan.size() * 4, st));
    AVS_CUDA(cudaMemsetAsync(d_xg_v, 0, 2 * uR * 2 * G4 * 4 + 256, st));   // rows no video owns contribute nothing

    LstmBatch lb{plan_dev, plan_dev + slots, plan_dev + 2 * slots, n_groups, nb};
    AVS_TRY(lstm_backward(d_fused, save_pre, save_c, m->whh, lb, d_xg_v, d_xg_a, st));
    AVS_TRY(shift_h(fused, plan_dev + desc_off, plan_dev + desc_off + n_videos, n_videos, max_len, hprev, st));
    AVS_TRY(transpose_f32(hprev, E, static_cast<int>(R), E, hprev_t, Rp, 0, 1, st));
    // The ten GEMMs that follow have 16 - 80 output tiles each and are independent: per modality, branch X (d_emb,
    // then the two dW_hh) and branch Y (the transposed operands, then the two dW_ih) run on four side streams
    // (fork / join with events: capturable).  One after the other they were ~0.3 ms of the 2.5-ms training step.
    SideStreams& sd = side_streams();
    const bool par = sd.ok;
    if (par) AVS_CUDA(cudaEventRecord(sd.fork, st));
    avs_status rc = AVS_OK;
    auto step = [&](avs_status x) { if (rc == AVS_OK) rc = x; };
    for (int mod = 0; mod < 2 && rc == AVS_OK; ++mod) {
        cudaStream_t sx = par ? sd.s[2 * mod] : st, sy = par ? sd.s[2 * mod + 1] : st;
        if (par) {
            AVS_CUDA(cudaStreamWaitEvent(sx, sd.fork, 0));
            AVS_CUDA(cudaStreamWaitEvent(sy, sd.fork, 0));
        }
        const float* d_xg = mod ? d_xg_a : d_xg_v;
        const float* emb = mod ? a_emb : v_emb;
        const float* wih = mod ? m->ih_a_x : m->ih_v_x;      // packed [2048, 512], exact fp32
        float* d_emb = mod ? d_a_emb : d_v_emb;
        // branch Y: weight-gradient operands in the reference's row order (rows of d_xg^T are un-permuted on the way)
        step(transpose_f32(d_xg, 2 * G4, static_cast<int>(R), 2 * G4, dxg_t[mod], Rp, 1, 1, sy));
        if (par && rc == AVS_OK) {
            cudaEventRecord(sd.aux[mod], sy);
            cudaStreamWaitEvent(sx, sd.aux[mod], 0);
        }
        step(transpose_f32(emb, H, static_cast<int>(R), H, emb_t[mod], Rp, 0, 1, sy));
        step(colsum_f32(d_xg, 2 * G4, static_cast<int>(R), 2 * G4, db_tmp[mod], 1, sy));
        // branch X: d_emb[R, 512] = d_xg[R, 2048] * W_ih_packed[2048, 512]  (both directions at once)
        step(convert_f32(d_xg, d_xg_r[mod], R * 2 * G4, DT_F32, 1, sx));
        step(transpose_f32(wih, H, 2 * G4, H, wih_t[mod], 2 * G4, 0, 1, sx));
        step(gemm_nt_tf32(d_xg_r[mod], 2 * G4, wih_t[mod], 2 * G4, R, H, 2 * G4, d_emb, H, sx));
        for (int dir = 0; dir < 2 && rc == AVS_OK; ++dir) {
            const int ld = mod * 2 + dir;
            const float* a_t = dxg_t[mod] + static_cast<size_t>(dir) * G4 * Rp;
            step(gemm_nt_tf32(a_t, Rp, emb_t[mod], Rp, G4, H, static_cast<int>(R), dW_ih[ld], H, sy));
            step(gemm_nt_tf32(a_t, Rp, hprev_t + static_cast<size_t>(ld) * HC * Rp, Rp, G4, HC, static_cast<int>(R),
                              dW_hh[ld], HC, sx));
            if (rc == AVS_OK && cudaMemcpyAsync(db[ld], db_tmp[mod] + dir * G4, G4 * 4, cudaMemcpyDeviceToDevice, sy) != cudaSuccess) {
                set_error("avs_bilstm_pair_bwd: bias gradient copy failed");
                rc = AVS_ERR_CUDA;
            }
        }
    }
    if (par)   // join every forked stream, also after an error (graph capture must not be left with dangling forks)
        for (int i = 0; i < SideStreams::N; ++i) {
            cudaEventRecord(sd.join[i], sd.s[i]);
            cudaStreamWaitEvent(st, sd.join[i], 0);
        }
    return rc;
}

avs_status avs_attention(const float* qkv, int64_t rows, int32_t E_, int32_t num_heads, int32_t n_seqs,
                         const int32_t* seq_base, const int32_t* seq_stride, const int32_t* seq_len, int precision,
                         float* ctx, void* cuda_stream) {
    AVS_CHECK(qkv && ctx && (n_seqs == 0 || (seq_base && seq_stride && seq_len)), AVS_ERR_INVALID,
              "avs_attention: null pointer");
    AVS_CHECK(precision_ok(precision), AVS_ERR_INVALID, "bad precision %d", precision);
    if (n_seqs == 0) return AVS_OK;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    std::vector<int32_t> sd(3 * static_cast<size_t>(n_seqs));
    int mx = 0;
    for (int s = 0; s < n_seqs; ++s) {
        AVS_CHECK(seq_len[s] >= 0 && seq_base[s] >= 0 && seq_stride[s] >= 1 &&
                      (seq_len[s] == 0 ||
                       static_cast<int64_t>(seq_base[s]) + static_cast<int64_t>(seq_len[s] - 1) * seq_stride[s] < rows),
                  AVS_ERR_INVALID, "avs_attention: sequence %d outside the %lld rows", s, static_cast<long long>(rows));
        sd[s] = seq_base[s];
        sd[n_seqs + s] = seq_stride[s];
        sd[2 * n_seqs + s] = seq_len[s];
        mx = std::max(mx, seq_len[s]);
    }
    int32_t* dev = nullptr;
    AVS_CUDA(cudaMallocAsync(&dev, sd.size() * 4, st));
    AVS_CUDA(cudaMemcpyAsync(dev, sd.data(), sd.size() * 4, cudaMemcpyHostToDevice, st));
    SeqDesc seqs{dev, dev + n_seqs, dev + 2 * n_seqs, n_seqs, mx};
    bool contiguous = true;
    for (int i = 0; i < n_seqs; ++i) contiguous &= seq_stride[i] == 1;
    avs_status s;
    if (precision != AVS_PREC_FP32_SIMT && contiguous && E_ == num_heads * 256) {
        const int dt = precision == AVS_PREC_BF16 ? DT_BF16 : DT_F16;
        void* qkv_h = nullptr;   // 16-bit copy of q | k | v for the tcgen05 kernel
        AVS_CUDA(cudaMallocAsync(&qkv_h, static_cast<size_t>(rows) * 3 * E_ * 2, st));
        s = convert_f32(qkv, qkv_h, rows * 3 * E_, dt, 0, st);
        if (s == AVS_OK) s = attention_tc(qkv_h, dt, rows, E_, num_heads, seqs, ctx, E_, DT_F32, 0, st);
        cudaFreeAsync(qkv_h, st);
    } else {
        s = attention_simt(qkv, 3ll * E_, E_, num_heads, seqs, ctx, E_, DT_F32, 0, st);
    }
    cudaFreeAsync(dev, st);
    return s;
}

avs_status avs_temporal_f1(const int32_t* pred, const int32_t* pred_start, const int32_t* gt, const int32_t* gt_start,
                           int32_t n_videos, double* f1_host, void* cuda_stream) {
    AVS_CHECK(n_videos >= 0, AVS_ERR_INVALID, "n_videos negative");
    if (n_videos == 0) return AVS_OK;
    AVS_CHECK(pred_start && gt_start && f1_host, AVS_ERR_INVALID, "avs_temporal_f1: null pointer");
    const int n = n_videos;
    const int P = pred_start[n], G = gt_start[n];
    AVS_CHECK(pred_start[0] == 0 && gt_start[0] == 0 && P >= 0 && G >= 0, AVS_ERR_INVALID, "bad offsets");
    AVS_CHECK((P == 0 || pred) && (G == 0 || gt), AVS_ERR_INVALID, "avs_temporal_f1: null shot list");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    // layout (8-byte aligned pieces): f1[n] | pred[2P] | gt[2G] | pred_start[n+1] | gt_start[n+1]
    const size_t bytes = static_cast<size_t>(n) * 8 + (2ull * P + 2ull * G + 2ull * (n + 1)) * 4;
    char* dev = nullptr;
    AVS_CUDA(cudaMallocAsync(&dev, bytes, st));
    double* f1_dev = reinterpret_cast<double*>(dev);
    int32_t* pred_dev = reinterpret_cast<int32_t*>(dev + static_cast<size_t>(n) * 8);
    int32_t* gt_dev = pred_dev + 2ull * P;
    int32_t* ps_dev = gt_dev + 2ull * G;
    int32_t* gs_dev = ps_dev + (n + 1);
    if (P) AVS_CUDA(cudaMemcpyAsync(pred_dev, pred, 2ull * P * 4, cudaMemcpyHostToDevice, st));
    if (G) AVS_CUDA(cudaMemcpyAsync(gt_dev, gt, 2ull * G * 4, cudaMemcpyHostToDevice, st));
    AVS_CUDA(cudaMemcpyAsync(ps_dev, pred_start, (n + 1) * 4, cudaMemcpyHostToDevice, st));
    AVS_CUDA(cudaMemcpyAsync(gs_dev, gt_start, (n + 1) * 4, cudaMemcpyHostToDevice, st));
    avs_status s = temporal_f1_device(pred_dev, ps_dev, gt_dev, gs_dev, n, f1_dev, st);
    if (s == AVS_OK) {
        cudaError_t e = cudaMemcpyAsync(f1_host, f1_dev, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_error("avs_temporal_f1: %s", cudaGetErrorString(e));
            s = AVS_ERR_CUDA;
        }
    }
    cudaFreeAsync(dev, st);
    return s;
}

// ---- scripts/evaluate.py:25-36 metric block for a batch of videos ----------------------------------
avs_status avs_eval_metrics(const float* pred, const void* target, int target_is_f64, int32_t n_videos,
                            const int32_t* row_start, const int32_t* lengths, double* metrics, int64_t* counts,
                            int space, void* cuda_stream) {
    AVS_CHECK(space == AVS_HOST || space == AVS_DEVICE, AVS_ERR_INVALID, "bad memory space %d", space);
    AVS_CHECK(n_videos >= 0, AVS_ERR_INVALID, "n_videos negative");
    if (n_videos == 0) return AVS_OK;
    AVS_CHECK(pred && target && row_start && lengths && metrics && counts, AVS_ERR_INVALID,
              "avs_eval_metrics: null pointer");
    int64_t rows = 0;
    for (int v = 0; v < n_videos; ++v) {
        AVS_CHECK(row_start[v] >= 0 && lengths[v] >= 0, AVS_ERR_INVALID, "video %d: bad row range", v);
        rows = std::max<int64_t>(rows, static_cast<int64_t>(row_start[v]) + lengths[v]);
    }
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const size_t tsz = target_is_f64 ? 8 : 4;
    const size_t n = static_cast<size_t>(n_videos);
    // one async allocation: descriptors | (host space: pred, target copies, outputs)
    size_t bytes = align_up(2 * n * 4, 256);
    const size_t off_pred = bytes;
    if (space == AVS_HOST) bytes += align_up(rows * 4, 256);
    const size_t off_tgt = bytes;
    if (space == AVS_HOST) bytes += align_up(rows * tsz, 256);
    const size_t off_f = bytes;
    if (space == AVS_HOST) bytes += align_up(n * 4 * 8, 256);
    const size_t off_i = bytes;
    if (space == AVS_HOST) bytes += align_up(n * 8 * 8, 256);
    char* dev = nullptr;
    AVS_CUDA(cudaMallocAsync(&dev, bytes, st));
    int32_t* desc = reinterpret_cast<int32_t*>(dev);
    avs_status s = AVS_OK;
    auto fail = [&](cudaError_t e) {
        if (e != cudaSuccess && s == AVS_OK) {
            set_error("avs_eval_metrics: %s", cudaGetErrorString(e));
            s = AVS_ERR_CUDA;
        }
    };
    fail(cudaMemcpyAsync(desc, row_start, n * 4, cudaMemcpyHostToDevice, st));
    fail(cudaMemcpyAsync(desc + n, lengths, n * 4, cudaMemcpyHostToDevice, st));
    const float* p_dev = pred;
    const void* t_dev = target;
    double* f_dev = metrics;
    long long* i_dev = reinterpret_cast<long long*>(counts);
    if (space == AVS_HOST) {
        fail(cudaMemcpyAsync(dev + off_pred, pred, rows * 4, cudaMemcpyHostToDevice, st));
        fail(cudaMemcpyAsync(dev + off_tgt, target, rows * tsz, cudaMemcpyHostToDevice, st));
        p_dev = reinterpret_cast<const float*>(dev + off_pred);
        t_dev = dev + off_tgt;
        f_dev = reinterpret_cast<double*>(dev + off_f);
        i_dev = reinterpret_cast<long long*>(dev + off_i);
    }
    if (s == AVS_OK) s = eval_metrics_device(p_dev, t_dev, target_is_f64, desc, desc + n, n_videos, f_dev, i_dev, st);
    if (s == AVS_OK && space == AVS_HOST) {
        fail(cudaMemcpyAsync(metrics, f_dev, n * 4 * 8, cudaMemcpyDeviceToHost, st));
        fail(cudaMemcpyAsync(counts, i_dev, n * 8 * 8, cudaMemcpyDeviceToHost, st));
        fail(cudaStreamSynchronize(st));
    }
    cudaFreeAsync(dev, st);
    return s;
}

// ---- features/fusion.py helpers -------------------------------------------------------------------
avs_status avs_cdist(const float* a, const float* b, int32_t na, int32_t nb, int32_t D, double* out, int space,
                     void* cuda_stream) {
    AVS_CHECK(space == AVS_HOST || space == AVS_DEVICE, AVS_ERR_INVALID, "bad memory space %d", space);
    AVS_CHECK(na >= 0 && nb >= 0 && D >= 0, AVS_ERR_INVALID, "avs_cdist: negative size");
    if (na == 0 || nb == 0) return AVS_OK;
    AVS_CHECK(a && b && out, AVS_ERR_INVALID, "avs_cdist: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (space == AVS_DEVICE) return cdist_device(a, b, na, nb, D, out, st);
    const size_t ba = align_up(static_cast<size_t>(na) * D * 4, 256), bb = align_up(static_cast<size_t>(nb) * D * 4, 256);
    const size_t bo = static_cast<size_t>(na) * nb * 8;
    char* dev = nullptr;
    AVS_CUDA(cudaMallocAsync(&dev, ba + bb + bo + 256, st));
    avs_status s = AVS_OK;
    cudaError_t e = cudaMemcpyAsync(dev, a, static_cast<size_t>(na) * D * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dev + ba, b, static_cast<size_t>(nb) * D * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        s = cdist_device(reinterpret_cast<float*>(dev), reinterpret_cast<float*>(dev + ba), na, nb, D,
                         reinterpret_cast<double*>(dev + ba + bb), st);
    if (e == cudaSuccess && s == AVS_OK) e = cudaMemcpyAsync(out, dev + ba + bb, bo, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && s == AVS_OK) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_error("avs_cdist: %s", cudaGetErrorString(e));
        s = AVS_ERR_CUDA;
    }
    cudaFreeAsync(dev, st);
    return s;
}

avs_status avs_interpolate(const float* features, int64_t n_rows, int32_t D, const int32_t* idx, const float* weights,
                           int32_t U, float* out, int space, void* cuda_stream) {
    AVS_CHECK(space == AVS_HOST || space == AVS_DEVICE, AVS_ERR_INVALID, "bad memory space %d", space);
    AVS_CHECK(n_rows >= 0 && D >= 0 && U >= 0, AVS_ERR_INVALID, "avs_interpolate: negative size");
    if (U == 0 || D == 0) return AVS_OK;
    AVS_CHECK(features && idx && weights && out, AVS_ERR_INVALID, "avs_interpolate: null pointer");
    for (int k = 0; k < U; ++k)
        AVS_CHECK(idx[k] >= 0 && idx[k] < n_rows, AVS_ERR_INVALID, "avs_interpolate: index %d = %d outside [0, %lld)", k,
                  idx[k], static_cast<long long>(n_rows));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const size_t b_idx = align_up(static_cast<size_t>(U) * 4, 256);
    const size_t b_feat = space == AVS_HOST ? align_up(static_cast<size_t>(n_rows) * D * 4, 256) : 0;
    const size_t b_out = space == AVS_HOST ? static_cast<size_t>(U) * D * 4 : 0;
    char* dev = nullptr;
    AVS_CUDA(cudaMallocAsync(&dev, 2 * b_idx + b_feat + b_out + 256, st));
    avs_status s = AVS_OK;
    cudaError_t e = cudaMemcpyAsync(dev, idx, static_cast<size_t>(U) * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dev + b_idx, weights, static_cast<size_t>(U) * 4, cudaMemcpyHostToDevice, st);
    const float* f_dev = features;
    float* o_dev = out;
    if (space == AVS_HOST) {
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(dev + 2 * b_idx, features, static_cast<size_t>(n_rows) * D * 4, cudaMemcpyHostToDevice, st);
        f_dev = reinterpret_cast<float*>(dev + 2 * b_idx);
        o_dev = reinterpret_cast<float*>(dev + 2 * b_idx + b_feat);
    }
    if (e == cudaSuccess)
        s = gather_scale_device(f_dev, reinterpret_cast<int32_t*>(dev), reinterpret_cast<float*>(dev + b_idx), U, D, o_dev, st);
    if (space == AVS_HOST && e == cudaSuccess && s == AVS_OK) {
        e = cudaMemcpyAsync(out, o_dev, b_out, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (e != cudaSuccess) {
        set_error("avs_interpolate: %s", cudaGetErrorString(e));
        s = AVS_ERR_CUDA;
    }
    cudaFreeAsync(dev, st);
    return s;
}

avs_status avs_dtw_path(const double* cost, int32_t n, int32_t m, int32_t* path, int32_t* path_len, double* total,
                        void* cuda_stream) {
    AVS_CHECK(cost && path && path_len && total, AVS_ERR_INVALID, "avs_dtw_path: null pointer");
    AVS_CHECK(n > 0 && m > 0, AVS_ERR_INVALID, "avs_dtw_path: the cost matrix must be non-empty (got %d x %d)", n, m);
    AVS_CHECK(static_cast<int64_t>(n) * m < (1ll << 31), AVS_ERR_UNSUPPORTED, "avs_dtw_path: matrix too large");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const size_t cells = static_cast<size_t>(n) * m;
    const size_t b_cost = align_up(cells * 8, 256), b_choice = align_up(cells, 256);
    const size_t b_path = align_up(static_cast<size_t>(n + m) * 2 * 4, 256);
    char* dev = nullptr;
    AVS_CUDA(cudaMallocAsync(&dev, 2 * b_cost + b_choice + b_path + 512, st));
    double* d_cost = reinterpret_cast<double*>(dev);
    double* d_acc = reinterpret_cast<double*>(dev + b_cost);
    uint8_t* d_choice = reinterpret_cast<uint8_t*>(dev + 2 * b_cost);
    int32_t* d_path = reinterpret_cast<int32_t*>(dev + 2 * b_cost + b_choice);
    int32_t* d_len = reinterpret_cast<int32_t*>(dev + 2 * b_cost + b_choice + b_path);
    double* d_total = reinterpret_cast<double*>(dev + 2 * b_cost + b_choice + b_path + 256);
    avs_status s = AVS_OK;
    cudaError_t e = cudaMemcpyAsync(d_cost, cost, cells * 8, cudaMemcpyDefault, st);
    if (e == cudaSuccess) s = dtw_device(d_cost, n, m, d_acc, d_choice, d_path, d_len, d_total, st);
    if (e == cudaSuccess && s == AVS_OK) e = cudaMemcpyAsync(path_len, d_len, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && s == AVS_OK) e = cudaMemcpyAsync(total, d_total, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && s == AVS_OK)
        e = cudaMemcpyAsync(path, d_path, static_cast<size_t>(n + m - 1) * 2 * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && s == AVS_OK) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_error("avs_dtw_path: %s", cudaGetErrorString(e));
        s = AVS_ERR_CUDA;
    }
    cudaFreeAsync(dev, st);
    return s;
}

}  // extern "C"
