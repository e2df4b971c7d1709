// Row LayerNorm over the model dimension (d = 256): one warp per row, 128-bit loads, two-pass
// statistics in registers (mean, then centred variance) — the numerically safe form.
//
// Replaces `torch.nn.LayerNorm(size, eps=1e-5)` at encoder.py:64-72,115,122,141,153,161,342 and
// convolution.py:66,124.   HBM-bound: 2 x 4 bytes per element.
#include <cuda_fp16.h>

#include "common.cuh"

namespace masr {

// SPLIT: write the result as the fp16 (h, l) pair consumed by the tensor-core GEMM (tc_gemm.cu) instead of fp32.
template <int D, bool SPLIT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int64_t ldx,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ y, __half* __restrict__ yh,
                                                        __half* __restrict__ yl, int64_t ldy, int M, float eps) {
    static_assert(D % 128 == 0, "D must be a multiple of 128");
    constexpr int V = D / 128;                 // float4 per lane
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* xr = x + (int64_t)row * ldx;
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        v[i] = ldg_f4(xr + (i * 32 + lane) * 4);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c = (i * 32 + lane) * 4;
        float4 g = ldg_f4(gamma + c), b = ldg_f4(beta + c);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (SPLIT) {
            __half hh[4], ll[4];
            const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hh[j] = __float2half_rn(ov[j]);
                ll[j] = __float2half_rn((ov[j] - __half2float(hh[j])) * 2048.0f);
            }
            *reinterpret_cast<uint2*>(yh + (int64_t)row * ldy + c) = *reinterpret_cast<const uint2*>(hh);
            *reinterpret_cast<uint2*>(yl + (int64_t)row * ldy + c) = *reinterpret_cast<const uint2*>(ll);
        } else {
            *reinterpret_cast<float4*>(y + (int64_t)row * ldy + c) = o;
        }
    }
}

}  // namespace masr

using namespace masr;

extern "C" int masr_layernorm_f32(const float* x, int64_t ldx, const float* gamma, const float* beta, float* y,
                                  int64_t ldy, int M, int D, float eps, void* stream) {
    if (M == 0) return MASR_OK;
    MASR_REQUIRE(x && gamma && beta && y, "masr_layernorm_f32: null pointer");
    MASR_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, "masr_layernorm_f32: leading dimensions must be multiples of 4");
    dim3 grid((M + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    switch (D) {
        case 256: layernorm_kernel<256, false><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, y, nullptr, nullptr, ldy, M, eps); break;
        case 512: layernorm_kernel<512, false><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, y, nullptr, nullptr, ldy, M, eps); break;
        case 1024: layernorm_kernel<1024, false><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, y, nullptr, nullptr, ldy, M, eps); break;
        case 2048: layernorm_kernel<2048, false><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, y, nullptr, nullptr, ldy, M, eps); break;
        default:
            set_last_error("masr_layernorm_f32: unsupported width D=%d (256/512/1024/2048)", D);
            return MASR_ERR_INVALID_ARGUMENT;
    }
    return check_launch("layernorm_kernel");
}

// LayerNorm whose output is the fp16 (h, l) operand pair of masr_gemm_tc_f16x2.
extern "C" int masr_layernorm_split_f16(const float* x, int64_t ldx, const float* gamma, const float* beta, void* yh,
                                        void* yl, int64_t ldy, int M, int D, float eps, void* stream) {
    if (M == 0) return MASR_OK;
    MASR_REQUIRE(x && gamma && beta && yh && yl, "masr_layernorm_split_f16: null pointer");
    MASR_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, "masr_layernorm_split_f16: leading dimensions must be multiples of 4");
    dim3 grid((M + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    switch (D) {
        case 256: layernorm_kernel<256, true><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, nullptr, (__half*)yh, (__half*)yl, ldy, M, eps); break;
        case 1024: layernorm_kernel<1024, true><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, nullptr, (__half*)yh, (__half*)yl, ldy, M, eps); break;
        case 2048: layernorm_kernel<2048, true><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, nullptr, (__half*)yh, (__half*)yl, ldy, M, eps); break;
        default:
            set_last_error("masr_layernorm_split_f16: unsupported width D=%d (256/1024/2048)", D);
            return MASR_ERR_INVALID_ARGUMENT;
    }
    return check_launch("layernorm_kernel<split>");
}
