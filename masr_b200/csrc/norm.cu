// Row LayerNorm over the model dimension (d = 256): one warp per row, 128-bit loads, two-pass
// statistics in registers (mean, then centred variance) — the numerically safe form.
//
// Replaces `torch.nn.LayerNorm(size, eps=1e-5)` at encoder.py:64-72,115,122,141,153,161,342 and
// convolution.py:66,124.   HBM-bound: 2 x 4 bytes per element.
#include "common.cuh"

namespace masr {

template <int D>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int64_t ldx,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ y, int64_t ldy, int M, float eps) {
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
    float* yr = y + (int64_t)row * ldy;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c = (i * 32 + lane) * 4;
        float4 g = ldg_f4(gamma + c), b = ldg_f4(beta + c);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        *reinterpret_cast<float4*>(yr + c) = o;
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
        case 256: layernorm_kernel<256><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, y, ldy, M, eps); break;
        case 512: layernorm_kernel<512><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, y, ldy, M, eps); break;
        case 1024: layernorm_kernel<1024><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, y, ldy, M, eps); break;
        case 2048: layernorm_kernel<2048><<<grid, 256, 0, st>>>(x, ldx, gamma, beta, y, ldy, M, eps); break;
        default:
            set_last_error("masr_layernorm_f32: unsupported width D=%d (256/512/1024/2048)", D);
            return MASR_ERR_INVALID_ARGUMENT;
    }
    return check_launch("layernorm_kernel");
}
