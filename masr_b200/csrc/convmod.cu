// Conformer convolution-module middle, fused:  depthwise Conv1d (k = 7/15/31) -> LayerNorm over the
// channels -> SiLU.   Input is the GLU output of pointwise_conv1 (GEMM epilogue in gemm.cu).
//
// Replaces convolution.py:121-126 (`depthwise_conv`, `norm`, `activation`).  Padding semantics are
// per utterance (B=1 API semantics):
//   * causal model, whole utterance: the reference left-pads the *pre-pointwise* input with zeros
//     (convolution.py:103), so positions t<0 see GLU(pointwise bias) — passed here as `pad_vec`;
//   * non-causal model: symmetric zero padding of the GLU output (`padding=(k-1)//2`, :57-65);
//     positions >= the utterance length are zeros, never a neighbour's or padding frames' data;
//   * streaming chunk: the caller runs pointwise_conv1 over [cache ++ chunk] and calls with lpad=0.
//
// HBM-bound: 4*(1 + 1) bytes per element, (TT+K-1)/TT read amplification served by L1/L2.
#include <cuda_fp16.h>

#include "common.cuh"

namespace masr {

template <int KS, int TT, int STRIDE, bool AFFINE>
__global__ void __launch_bounds__(256) dwconv_ln_silu_kernel(const float* __restrict__ g, int64_t ldg,
                                                             int64_t g_bstride, const float* __restrict__ w,
                                                             const float* __restrict__ bias,
                                                             const float* __restrict__ ln_g,
                                                             const float* __restrict__ ln_b,
                                                             const float* __restrict__ pad_vec, float* __restrict__ y,
                                                             __half* __restrict__ yh, __half* __restrict__ yl, int64_t ldy, int64_t y_bstride,
                                                             const int* __restrict__ in_lens, int lpad, int out_rows,
                                                             float eps) {
    // y[t] = sum_k w[k] * g[t*STRIDE - lpad + k]   (STRIDE 2 = the strided block of the EfficientConformer)
    constexpr int C = 256;
    const int c = threadIdx.x, b = blockIdx.y, t0 = blockIdx.x * TT;
    const int in_len = in_lens[b];
    float wk[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) wk[k] = __ldg(w + c * KS + k);
    const float bs = __ldg(bias + c);
    const float pv = pad_vec ? __ldg(pad_vec + c) : 0.f;
    float acc[TT];
#pragma unroll
    for (int j = 0; j < TT; ++j) acc[j] = bs;
    const float* gb = g + (int64_t)b * g_bstride * ldg + c;
#pragma unroll
    for (int i = 0; i < (TT - 1) * STRIDE + KS; ++i) {
        const int tau = t0 * STRIDE - lpad + i;
        float v;
        if (tau < 0) v = pv;
        else if (tau >= in_len) v = 0.f;
        else v = __ldg(gb + (int64_t)tau * ldg);
#pragma unroll
        for (int j = 0; j < TT; ++j) {
            const int k = i - j * STRIDE;
            if (k >= 0 && k < KS) acc[j] = fmaf(wk[k], v, acc[j]);
        }
    }
    // LayerNorm over the 256 channels of each of the TT frames (two-pass, block-wide)
    __shared__ float red[8][TT];
    __shared__ float stat[TT];
    const int warp = c >> 5, lane = c & 31;
#pragma unroll
    for (int j = 0; j < TT; ++j) {
        float s = warp_sum(acc[j]);
        if (lane == 0) red[warp][j] = s;
    }
    __syncthreads();
    if (c < TT) {
        float s = 0.f;
        for (int q = 0; q < 8; ++q) s += red[q][c];
        stat[c] = s * (1.0f / C);
    }
    __syncthreads();
    float dev[TT];
#pragma unroll
    for (int j = 0; j < TT; ++j) {
        dev[j] = acc[j] - stat[j];
        float s = warp_sum(dev[j] * dev[j]);
        if (lane == 0) red[warp][j] = s;      // safe: all reads of red[] above completed before the barrier
    }
    __syncthreads();
    if (c < TT) {
        float s = 0.f;
        for (int q = 0; q < 8; ++q) s += red[q][c];
        stat[c] = rsqrtf(s * (1.0f / C) + eps);
    }
    __syncthreads();
    const float gg = __ldg(ln_g + c), bb = __ldg(ln_b + c);
    if (AFFINE) {
        // BatchNorm1d(eval) folded by the caller: ln_g = gamma / sqrt(running_var + eps), ln_b = beta - running_mean * ln_g
#pragma unroll
        for (int j = 0; j < TT; ++j) { dev[j] = acc[j]; }
    }
    const int64_t yoff = (int64_t)b * y_bstride * ldy + c;
#pragma unroll
    for (int j = 0; j < TT; ++j) {
        const int t = t0 + j;
        if (t < out_rows) {
            const float o = AFFINE ? silu_f(dev[j] * gg + bb) : silu_f(dev[j] * stat[j] * gg + bb);
            if (y) y[yoff + (int64_t)t * ldy] = o;
            if (yh) {
                const __half hh = __float2half_rn(o);
                yh[yoff + (int64_t)t * ldy] = hh;
                yl[yoff + (int64_t)t * ldy] = __float2half_rn((o - __half2float(hh)) * 2048.0f);
            }
        }
    }
}

}  // namespace masr

using namespace masr;

extern "C" int masr_dwconv_ln_silu_f32(const float* g, int64_t ldg, int64_t g_bstride, const float* w,
                                       const float* bias, const float* ln_gamma, const float* ln_beta,
                                       const float* pad_vec, float* y, void* yh, void* yl, int64_t ldy, int64_t y_bstride,
                                       const int* in_lens, int B, int C, int kernel_size, int lpad, int out_rows,
                                       float eps, void* stream) {
    return masr_dwconv_ln_silu_strided_f32(g, ldg, g_bstride, w, bias, ln_gamma, ln_beta, pad_vec, y, yh, yl, ldy, y_bstride,
                                           in_lens, B, C, kernel_size, lpad, 1, out_rows, eps, stream);
}

extern "C" int masr_dwconv_ln_silu_strided_f32(const float* g, int64_t ldg, int64_t g_bstride, const float* w,
                                               const float* bias, const float* ln_gamma, const float* ln_beta,
                                               const float* pad_vec, float* y, void* yh, void* yl, int64_t ldy,
                                               int64_t y_bstride, const int* in_lens, int B, int C, int kernel_size,
                                               int lpad, int stride, int out_rows, float eps, void* stream) {
    if (B == 0 || out_rows == 0) return MASR_OK;
    MASR_REQUIRE(g && w && bias && ln_gamma && ln_beta && (y || (yh && yl)) && in_lens, "masr_dwconv_ln_silu_f32: null pointer");
    MASR_REQUIRE(C == 256, "masr_dwconv_ln_silu_f32: C=%d unsupported (this build: 256)", C);
    constexpr int TT = 8;
    dim3 grid((out_rows + TT - 1) / TT, B);
    cudaStream_t st = (cudaStream_t)stream;
#define MASR_DW_LAUNCH(KS, S)                                                                                       \
    dwconv_ln_silu_kernel<KS, TT, S, false><<<grid, 256, 0, st>>>(g, ldg, g_bstride, w, bias, ln_gamma, ln_beta, pad_vec, y, \
                                                           (__half*)yh, (__half*)yl, ldy, y_bstride, in_lens, lpad, out_rows, eps)
    MASR_REQUIRE(stride == 1 || stride == 2, "masr_dwconv_ln_silu: stride %d unsupported (1/2)", stride);
    if (stride == 2) {
        MASR_REQUIRE(kernel_size == 15, "masr_dwconv_ln_silu: stride 2 is built for kernel size 15 only");
        MASR_DW_LAUNCH(15, 2);
    } else
    switch (kernel_size) {
        case 7: MASR_DW_LAUNCH(7, 1); break;
        case 15: MASR_DW_LAUNCH(15, 1); break;
        case 31: MASR_DW_LAUNCH(31, 1); break;
        default:
            set_last_error("masr_dwconv_ln_silu_f32: unsupported kernel size %d (7/15/31)", kernel_size);
            return MASR_ERR_INVALID_ARGUMENT;
    }
#undef MASR_DW_LAUNCH
    return check_launch("dwconv_ln_silu_kernel");
}

// Squeezeformer conv-module middle (squeezeformer/convolution.py:136-142): depthwise Conv1d(k) -> BatchNorm1d (eval mode,
// folded by the caller into per-channel scale/shift) -> SiLU.  Same padding contract as masr_dwconv_ln_silu_f32.
extern "C" int masr_dwconv_bn_silu_f32(const float* g, int64_t ldg, int64_t g_bstride, const float* w, const float* bias,
                                       const float* bn_scale, const float* bn_shift, const float* pad_vec, float* y, void* yh,
                                       void* yl, int64_t ldy, int64_t y_bstride, const int* in_lens, int B, int C,
                                       int kernel_size, int lpad, int out_rows, void* stream) {
    if (B == 0 || out_rows == 0) return MASR_OK;
    MASR_REQUIRE(g && w && bias && bn_scale && bn_shift && (y || (yh && yl)) && in_lens, "masr_dwconv_bn_silu_f32: null pointer");
    MASR_REQUIRE(C == 256, "masr_dwconv_bn_silu_f32: C=%d unsupported (this build: 256)", C);
    constexpr int TT = 8;
    dim3 grid((out_rows + TT - 1) / TT, B);
    cudaStream_t st = (cudaStream_t)stream;
    switch (kernel_size) {
        case 15:
            dwconv_ln_silu_kernel<15, TT, 1, true><<<grid, 256, 0, st>>>(g, ldg, g_bstride, w, bias, bn_scale, bn_shift, pad_vec, y,
                (__half*)yh, (__half*)yl, ldy, y_bstride, in_lens, lpad, out_rows, 0.f);
            break;
        case 31:
            dwconv_ln_silu_kernel<31, TT, 1, true><<<grid, 256, 0, st>>>(g, ldg, g_bstride, w, bias, bn_scale, bn_shift, pad_vec, y,
                (__half*)yh, (__half*)yl, ldy, y_bstride, in_lens, lpad, out_rows, 0.f);
            break;
        default:
            set_last_error("masr_dwconv_bn_silu_f32: unsupported kernel size %d (15/31)", kernel_size);
            return MASR_ERR_INVALID_ARGUMENT;
    }
    return check_launch("dwconv_bn_silu_kernel");
}
