// CUDA-core fp32 GEMM with the same epilogues as gemm_tc.cu.  This is the AVS_PREC_FP32_SIMT
// path: an exact-fp32 debugging aid used to separate tensor-core rounding from logic errors;
// it is not the performance path.  Also holds the elementwise conversion kernel.
#include "common.cuh"
#include "ptx.cuh"

namespace avs {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, int64_t lda,
                                                        const float* __restrict__ W, int64_t ldw, int M, int N,
                                                        int K, GemmEpilogue epi) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Ws[TK][TN + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += TK) {
        for (int i = threadIdx.x; i < TM * TK; i += 256) {
            const int r = i / TK, c = i % TK;
            const int gm = m0 + r, gk = k0 + c;
            As[c][r] = (gm < M && gk < K) ? A[static_cast<int64_t>(gm) * lda + gk] : 0.f;
            const int gn = n0 + r;
            Ws[c][r] = (gn < N && gk < K) ? W[static_cast<int64_t>(gn) * ldw + gk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Ws[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    // epilogue
    __shared__ float srow[TM][17];
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        float part = 0.f;
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            float x = acc[i][j] * epi.acc_scale;
            if (n < N) {
                if (epi.bias) x += epi.bias[n];
                if (epi.relu) x = fmaxf(x, 0.f);
                if (epi.scores) {
                    part = fmaf(x, epi.score_w2[n], part);
                } else if (m < M) {
                    if (epi.out_dtype == DT_F32)
                        reinterpret_cast<float*>(epi.C)[static_cast<int64_t>(m) * epi.ldc + n] =
                            epi.round_tf32 ? to_tf32_rn(x) : x;
                    else if (epi.out_dtype == DT_F16)
                        reinterpret_cast<__half*>(epi.C)[static_cast<int64_t>(m) * epi.ldc + n] = __float2half_rn(x);
                    else
                        reinterpret_cast<__nv_bfloat16*>(epi.C)[static_cast<int64_t>(m) * epi.ldc + n] =
                            __float2bfloat16_rn(x);
                }
            }
        }
        srow[ty * 4 + i][tx] = part;
    }
    if (epi.scores) {  // N == 64 -> a single n-block holds the full row
        __syncthreads();
        if (threadIdx.x < TM) {
            const int m = m0 + threadIdx.x;
            float s = 0.f;
            for (int t = 0; t < 16; ++t) s += srow[threadIdx.x][t];
            if (m < M) epi.scores[m] = 1.0f / (1.0f + expf(-(s + epi.score_b2[0])));
        }
    }
}

__global__ void convert_kernel(const float* __restrict__ src, void* __restrict__ dst, int64_t n4, int dst_dtype,
                               int round_tf32) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
        if (dst_dtype == DT_F32) {
            float4 o = round_tf32 ? make_float4(to_tf32_rn(v.x), to_tf32_rn(v.y), to_tf32_rn(v.z), to_tf32_rn(v.w)) : v;
            reinterpret_cast<float4*>(dst)[i] = o;
        } else if (dst_dtype == DT_F16) {
            __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
            reinterpret_cast<uint2*>(dst)[i] =
                make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
        } else {
            __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
            reinterpret_cast<uint2*>(dst)[i] =
                make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
        }
    }
}

}  // namespace

avs_status gemm_simt(const float* A, int64_t lda, const float* W, int64_t ldw, int64_t M, int N, int K,
                     const GemmEpilogue& epi, cudaStream_t stream) {
    if (M == 0) return AVS_OK;
    AVS_CHECK(M > 0 && N > 0 && K > 0 && M < (1ll << 31), AVS_ERR_INVALID, "gemm_simt: bad shape");
    if (epi.scores) AVS_CHECK(N == 64, AVS_ERR_INVALID, "score epilogue needs N == 64");
    dim3 grid((N + TN - 1) / TN, static_cast<unsigned>((M + TM - 1) / TM));
    gemm_simt_kernel<<<grid, 256, 0, stream>>>(A, lda, W, ldw, static_cast<int>(M), N, K, epi);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

// Range check of fp16 activations (AVS_CHECK_RANGE=1): counts the elements whose magnitude is at the saturation value
// of the fp32 -> fp16 casts (65504) or that are not finite.
__global__ void count_saturated_f16_kernel(const uint16_t* __restrict__ x, int64_t n, unsigned int* __restrict__ count) {
    unsigned int local = 0;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        local += ((x[i] & 0x7FFFu) >= 0x7BFFu) ? 1u : 0u;   // |x| == 65504, inf or NaN
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}
avs_status count_saturated_f16(const void* x, int64_t n, unsigned int* count_dev, cudaStream_t stream) {
    if (n == 0) return AVS_OK;
    const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 8));
    count_saturated_f16_kernel<<<blocks, 256, 0, stream>>>(static_cast<const uint16_t*>(x), n, count_dev);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

avs_status convert_f32(const float* src, void* dst, int64_t n, int dst_dtype, int round_tf32, cudaStream_t stream) {
    if (n == 0) return AVS_OK;
    AVS_CHECK(n % 4 == 0, AVS_ERR_INVALID, "convert: element count must be a multiple of 4");
    const int64_t n4 = n / 4;
    const int blocks = static_cast<int>(std::min<int64_t>((n4 + 255) / 256, 148 * 8));
    convert_kernel<<<blocks, 256, 0, stream>>>(src, dst, n4, dst_dtype, round_tf32);
    AVS_LAUNCH_CHECK();
    return AVS_OK;
}

}  // namespace avs
