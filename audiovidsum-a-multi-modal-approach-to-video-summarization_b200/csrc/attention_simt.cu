// CUDA-core fp32 scaled-dot-product attention over arbitrary row sequences.
//
// Arithmetic of /root/reference/models/attention.py:17-23 (== the core of nn.MultiheadAttention
// used at /root/reference/models/av_model.py:26,44): contiguous dh-wide head slices, scale
// 1/sqrt(dh), softmax over keys, no mask beyond the sequence length.
//
// Sequence s = rows base[s] + i*stride[s], i < len[s]; this covers both the temporal axis
// (stride 1, one sequence per video) and the reference's literal axis quirk (one sequence per
// frame index, stride T, length B).  It serves the literal B>1 mode, where sequences have length
// B (tiny), the AVS_PREC_FP32_SIMT mode, and as the on-device cross-check of the tcgen05 kernel.
// One warp per (query, head); online softmax; K/V streamed from L2.
#include "common.cuh"
#include "ptx.cuh"

namespace avs {

namespace {

constexpr int MAX_DPL = 8;  // head dim <= 256

__global__ void __launch_bounds__(128) attention_simt_kernel(const float* __restrict__ qkv, int64_t ld_qkv, int E,
                                                             int H, SeqDesc seqs, void* __restrict__ ctx,
                                                             int64_t ld_ctx, int out_dtype, int round_tf32,
                                                             float scale) {
    const int seq = blockIdx.z;
    const int head = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int len = seqs.len[seq];
    const int qi = blockIdx.x * 4 + warp;
    if (qi >= len) return;
    const int64_t base = seqs.base[seq], stride = seqs.stride[seq];
    const int dh = E / H;
    const int dpl = dh / 32;
    const int col = head * dh + lane * dpl;

    float q[MAX_DPL], o[MAX_DPL];
    const float* qrow = qkv + (base + static_cast<int64_t>(qi) * stride) * ld_qkv + col;
#pragma unroll
    for (int d = 0; d < MAX_DPL; ++d) {
        q[d] = d < dpl ? qrow[d] * scale : 0.f;
        o[d] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < len; ++j) {
        const float* krow = qkv + (base + static_cast<int64_t>(j) * stride) * ld_qkv + E + col;
        const float* vrow = krow + E;
        float dot = 0.f;
#pragma unroll
        for (int d = 0; d < MAX_DPL; ++d)
            if (d < dpl) dot = fmaf(q[d], __ldg(krow + d), dot);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
        const float m_new = fmaxf(m, dot);
        const float corr = expf(m - m_new);
        const float p = expf(dot - m_new);
        l = l * corr + p;
#pragma unroll
        for (int d = 0; d < MAX_DPL; ++d)
            if (d < dpl) o[d] = fmaf(p, __ldg(vrow + d), o[d] * corr);
        m = m_new;
    }
    const float inv = 1.0f / l;
    const int64_t ooff = (base + static_cast<int64_t>(qi) * stride) * ld_ctx + col;
#pragma unroll
    for (int d = 0; d < MAX_DPL; ++d) {
        if (d < dpl) {
            if (out_dtype != DT_F32)
                reinterpret_cast<uint16_t*>(ctx)[ooff + d] = to_lowp_bits(o[d] * inv, out_dtype);
            else
                reinterpret_cast<float*>(ctx)[ooff + d] = round_tf32 ? to_tf32_rn(o[d] * inv) : o[d] * inv;
        }
    }
}

}  // namespace

avs_status attention_simt(const float* qkv, int64_t ld_qkv, int E, int H, const SeqDesc& seqs, void* ctx,
                          int64_t ld_ctx, int out_dtype, int round_tf32, cudaStream_t stream) {
    if (seqs.n_seqs == 0 || seqs.max_len == 0) return AVS_OK;
    AVS_CHECK(H > 0 && E % H == 0, AVS_ERR_INVALID, "attention: embed dim %d not divisible by %d heads", E, H);
    const int dh = E / H;
    AVS_CHECK(dh % 32 == 0 && dh <= 32 * MAX_DPL, AVS_ERR_UNSUPPORTED,
              "attention: head dim %d must be a multiple of 32 and <= 256", dh);
    AVS_CHECK(seqs.n_seqs <= 65535 && H <= 65535, AVS_ERR_UNSUPPORTED, "attention: too many sequences in one launch");
    dim3 grid((seqs.max_len + 3) / 4, H, seqs.n_seqs);
    attention_simt_kernel<<<grid, 128, 0, stream>>>(qkv, ld_qkv, E, H, seqs, ctx, ld_ctx, out_dtype, round_tf32,
                                                    1.0f / sqrtf(static_cast<float>(dh)));
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

}  // namespace avs
