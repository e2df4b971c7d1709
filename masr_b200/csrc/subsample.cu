// First convolution of Conv2dSubsampling4 with the global CMVN folded into the load:
//   c1[b, t, f, co] = relu(b1[co] + sum_{kh,kw} w1[co, kh, kw] * cmvn(feats[b, 2t+kh, 2f+kw]))
// written channels-last so the second convolution (implicit GEMM, gemm.cu) reads contiguous
// 256-channel K-slices.
//
// Replaces GlobalCMVN.forward (utils/cmvn.py:29-31) and conv #1 + ReLU (conformer/subsampling.py:81-82,108).
// Write-bound: 4*W1*C bytes out per (b,t) row against 3*idim*4 bytes in — one CTA per output row,
// one thread per output channel, 1 KB coalesced stores per (t,f).
#include <cuda_fp16.h>

#include "common.cuh"

namespace masr {

__global__ void __launch_bounds__(256) conv1_cmvn_relu_kernel(const float* __restrict__ feats,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ istd,
                                                              const float* __restrict__ w1, const float* __restrict__ b1,
                                                              float* __restrict__ out, __half* __restrict__ ph,
                                                              __half* __restrict__ pl, int Fmax, int idim, int F1max,
                                                              int W1, int C) {
    extern __shared__ float s_in[];            // [3][idim] normalised input rows
    const int b = blockIdx.y, t = blockIdx.x;
    for (int i = threadIdx.x; i < 3 * idim; i += blockDim.x) {
        int r = i / idim, c = i - r * idim;
        float v = __ldg(feats + ((int64_t)b * Fmax + 2 * t + r) * idim + c);
        if (mean) v = (v - __ldg(mean + c)) * __ldg(istd + c);
        s_in[i] = v;
    }
    __syncthreads();
    for (int co = threadIdx.x; co < C; co += blockDim.x) {
        float w[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = __ldg(w1 + co * 9 + k);
        const float bias = __ldg(b1 + co);
        float* o = out ? out + (((int64_t)b * F1max + t) * W1) * C + co : nullptr;
        const int B = gridDim.y, TH = (F1max + 1) >> 1;
        for (int f = 0; f < W1; ++f) {
            float acc = bias;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) acc = fmaf(w[kh * 3 + kw], s_in[kh * idim + 2 * f + kw], acc);
            acc = fmaxf(acc, 0.f);
            if (o) o[(int64_t)f * C] = acc;
            if (ph) {
                // (t,f)-parity planes [4][B][TH][20][C] for the stride-2 implicit GEMM (tc_gemm.cu)
                const int plane = (t & 1) * 2 + (f & 1);
                const int64_t idx = ((((int64_t)plane * B + b) * TH + (t >> 1)) * 20 + (f >> 1)) * C + co;
                const __half hh = __float2half_rn(acc);
                ph[idx] = hh;
                pl[idx] = __float2half_rn((acc - __half2float(hh)) * 2048.0f);
            }
        }
    }
}

}  // namespace masr

using namespace masr;

extern "C" int masr_conv1_cmvn_relu_f32(const float* feats, const float* mean, const float* istd, const float* w1,
                                        const float* b1, float* out, int B, int Fmax, int idim, int F1max, int W1,
                                        int C, void* stream) {
    if (B == 0 || F1max == 0) return MASR_OK;
    MASR_REQUIRE(feats && w1 && b1 && out, "masr_conv1_cmvn_relu_f32: null pointer");
    MASR_REQUIRE((mean == nullptr) == (istd == nullptr), "masr_conv1_cmvn_relu_f32: mean/istd must both be set or both null");
    MASR_REQUIRE(2 * (F1max - 1) + 2 < Fmax && 2 * (W1 - 1) + 2 < idim, "masr_conv1_cmvn_relu_f32: window exceeds input");
    conv1_cmvn_relu_kernel<<<dim3(F1max, B), 256, 3 * idim * sizeof(float), (cudaStream_t)stream>>>(
        feats, mean, istd, w1, b1, out, nullptr, nullptr, Fmax, idim, F1max, W1, C);
    return check_launch("conv1_cmvn_relu_kernel");
}

// Same convolution, output as fp16 (h,l) pairs in four (t,f)-parity planes [4][B][(F1max+1)/2][20][C]
// (the operand layout of masr_conv2_tc_f16x2).  W1 must be <= 40.
extern "C" int masr_conv1_cmvn_relu_planes_f16(const float* feats, const float* mean, const float* istd, const float* w1,
                                               const float* b1, void* planes_h, void* planes_l, int B, int Fmax,
                                               int idim, int F1max, int W1, int C, void* stream) {
    if (B == 0 || F1max == 0) return MASR_OK;
    MASR_REQUIRE(feats && w1 && b1 && planes_h && planes_l, "masr_conv1_cmvn_relu_planes_f16: null pointer");
    MASR_REQUIRE((mean == nullptr) == (istd == nullptr), "masr_conv1_cmvn_relu_planes_f16: mean/istd must both be set or both null");
    MASR_REQUIRE(2 * (F1max - 1) + 2 < Fmax && 2 * (W1 - 1) + 2 < idim && W1 <= 40, "masr_conv1_cmvn_relu_planes_f16: bad geometry");
    conv1_cmvn_relu_kernel<<<dim3(F1max, B), 256, 3 * idim * sizeof(float), (cudaStream_t)stream>>>(
        feats, mean, istd, w1, b1, nullptr, (__half*)planes_h, (__half*)planes_l, Fmax, idim, F1max, W1, C);
    return check_launch("conv1_cmvn_relu_kernel<planes>");
}
