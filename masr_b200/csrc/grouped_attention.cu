// Grouped relative-position attention of the EfficientConformer (blocks 0-3), fp32.
//
// Reference: GroupedRelPositionMultiHeadedAttention (masr/model_utils/efficient_conformer/attention.py:35-69,
// 120-182).  `pad4group` zero-pads time to a multiple of `group` and then *views* the [t, h*d_k] memory of
// `group` consecutive frames as h heads of width group*d_k (:58-60) — i.e. head hh' of group j' is the flat
// range [hh'*G*dk, (hh'+1)*G*dk) of the G*h*dk floats of frames G*j' .. G*j'+G-1.  Scores use 1/sqrt(G*dk),
// no rel_shift; the output goes back through the same view and the padding is dropped (:113-114).
// Here the view is folded into the addressing: q/k/v/p are plain [T, d_model] row-major buffers, an element of
// frame >= T_b (the zero padding of the reference) is read as 0, and outputs of padded frames are not stored.
//
// One warp per (utterance, head, query group); 16 queries per CTA share K/P/V tiles staged in shared memory.
// Cost is small (T/G queries x T/G keys): this is not a roofline kernel.
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace masr {

constexpr int GQ = 16;     // query groups (warps) per CTA
constexpr int GK = 8;      // key groups per shared-memory tile

template <int DG>          // DG = group * d_k (192)
__global__ void __launch_bounds__(GQ * 32) grouped_attention_kernel(
    const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, const float* __restrict__ P,
    int64_t ld, int64_t bstride, int64_t ldk, int64_t k_bstride, const float* __restrict__ pos_u,
    const float* __restrict__ pos_v, float* __restrict__ O, __half* __restrict__ Oh, __half* __restrict__ Ol,
    const int* __restrict__ lens, const int* __restrict__ k_lens, int H, int group, int d_model, float scale) {
    constexpr int R = DG / 32;                       // values per lane
    __shared__ float sK[GK][DG], sP[GK][DG], sV[GK][DG];
    const int b = blockIdx.z, hh = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // queries: T frames at Q + b*bstride rows; keys/values: Tk frames at K/V + b*k_bstride rows of pitch ldk (the chunk
    // path reads [cache ++ chunk] keys; both groupings start at their own first frame, attention.py:44-60)
    const int T = lens[b];
    const int Tk = k_lens[b];
    const int Tq = (T + group - 1) / group;          // query groups incl. the zero-padded last one
    const int Tg = (Tk + group - 1) / group;         // key groups
    const int qi = blockIdx.x * GQ + warp;
    const int64_t base = (int64_t)b * bstride * ld;
    const int64_t kbase = (int64_t)b * k_bstride * ldk;
    // flat offset f inside a group -> (frame offset, column)
    int foff[R], fcol[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int f = hh * DG + lane + 32 * r;
        foff[r] = f / d_model;
        fcol[r] = f - foff[r] * d_model;
    }
    float qu[R], qv[R], o[R];
    const bool q_ok = qi < Tq;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int frame = qi * group + foff[r];
        float q = (q_ok && frame < T) ? __ldg(Q + base + (int64_t)frame * ld + fcol[r]) : 0.f;
        qu[r] = q + __ldg(pos_u + hh * DG + lane + 32 * r);
        qv[r] = q + __ldg(pos_v + hh * DG + lane + 32 * r);
        o[r] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < Tg; k0 += GK) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < GK * DG; idx += GQ * 32) {
            const int kk = idx / DG, e = idx - kk * DG;
            const int f = hh * DG + e;
            const int fo = f / d_model, fc = f - fo * d_model;
            const int frame = (k0 + kk) * group + fo;
            const bool ok = (k0 + kk) < Tg && frame < Tk;
            sK[kk][e] = ok ? __ldg(K + kbase + (int64_t)frame * ldk + fc) : 0.f;
            sV[kk][e] = ok ? __ldg(V + kbase + (int64_t)frame * ldk + fc) : 0.f;
            sP[kk][e] = ok ? __ldg(P + (int64_t)frame * d_model + fc) : 0.f;   // P is [T, d_model] dense, shared by the batch
        }
        __syncthreads();
        const int nk = min(GK, Tg - k0);
        for (int kk = 0; kk < nk; ++kk) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) s = fmaf(qu[r], sK[kk][lane + 32 * r], fmaf(qv[r], sP[kk][lane + 32 * r], s));
            s = warp_sum(s) * scale;
            const float m_new = fmaxf(m, s);
            const float alpha = expf(m - m_new), pw = expf(s - m_new);
            l = l * alpha + pw;
            m = m_new;
#pragma unroll
            for (int r = 0; r < R; ++r) o[r] = fmaf(pw, sV[kk][lane + 32 * r], o[r] * alpha);
        }
    }
    if (!q_ok) return;
    const float inv = 1.0f / l;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int frame = qi * group + foff[r];
        if (frame >= T) continue;                    // the reference drops the padded frames (attention.py:113-114)
        const float v = o[r] * inv;
        const int64_t off = base + (int64_t)frame * ld + fcol[r];
        if (O) O[off] = v;
        if (Oh) {
            const __half h = __float2half_rn(v);
            Oh[off] = h;
            Ol[off] = __float2half_rn((v - __half2float(h)) * 2048.0f);
        }
    }
}

// residual path of the strided block: AvgPool1d(kernel=stride=2, ceil_mode=True, count_include_pad=False) over time
// (efficient_conformer/encoder.py:173-175,520-523), per utterance.
__global__ void __launch_bounds__(256) avgpool2_time_kernel(const float* __restrict__ x, int64_t in_bstride,
                                                            float* __restrict__ y, int64_t out_bstride,
                                                            const int* __restrict__ lens, int out_rows, int D) {
    const int b = blockIdx.y, t = blockIdx.x;
    const int T = lens[b];
    const float* x0 = x + ((int64_t)b * in_bstride + 2 * t) * D;
    float* yo = y + ((int64_t)b * out_bstride + t) * D;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float v = 0.f;
        if (2 * t < T) {
            v = x0[c];
            if (2 * t + 1 < T) v = (v + x0[D + c]) * 0.5f;
        }
        yo[c] = v;
    }
}

}  // namespace masr

using namespace masr;

extern "C" int masr_grouped_attention_f32(const float* Q, const float* K, const float* V, const float* P, int64_t ld,
                                          int64_t bstride, const float* pos_u, const float* pos_v, float* O, void* Oh,
                                          void* Ol, const int* lens, int B, int H, int d_k, int group, int max_t,
                                          void* stream) {
    if (B == 0 || max_t == 0) return MASR_OK;
    MASR_REQUIRE(Q && K && V && P && pos_u && pos_v && (O || (Oh && Ol)) && lens, "masr_grouped_attention_f32: null pointer");
    MASR_REQUIRE(group * d_k == 192 && H * d_k == 256, "masr_grouped_attention_f32: this build supports group*d_k=192, d_model=256");
    const int max_g = (max_t + group - 1) / group;
    dim3 grid((max_g + GQ - 1) / GQ, H, B);
    grouped_attention_kernel<192><<<grid, GQ * 32, 0, (cudaStream_t)stream>>>(
        Q, K, V, P, ld, bstride, ld, bstride, pos_u, pos_v, O, (__half*)Oh, (__half*)Ol, lens, lens, H, group, H * d_k,
        1.0f / sqrtf((float)(group * d_k)));
    return check_launch("grouped_attention_kernel");
}

extern "C" int masr_grouped_attention_cache_f32(const float* Q, int64_t ldq, int64_t q_bstride, const float* K, const float* V,
                                                int64_t ldk, int64_t k_bstride, const float* P, const float* pos_u,
                                                const float* pos_v, float* O, void* Oh, void* Ol, const int* q_lens,
                                                const int* k_lens, int B, int H, int d_k, int group, int max_q, void* stream) {
    if (B == 0 || max_q == 0) return MASR_OK;
    MASR_REQUIRE(Q && K && V && P && pos_u && pos_v && (O || (Oh && Ol)) && q_lens && k_lens,
                 "masr_grouped_attention_cache_f32: null pointer");
    MASR_REQUIRE(group * d_k == 192 && H * d_k == 256, "masr_grouped_attention_cache_f32: this build supports group*d_k=192, d_model=256");
    const int max_g = (max_q + group - 1) / group;
    dim3 grid((max_g + GQ - 1) / GQ, H, B);
    grouped_attention_kernel<192><<<grid, GQ * 32, 0, (cudaStream_t)stream>>>(
        Q, K, V, P, ldq, q_bstride, ldk, k_bstride, pos_u, pos_v, O, (__half*)Oh, (__half*)Ol, q_lens, k_lens, H, group, H * d_k,
        1.0f / sqrtf((float)(group * d_k)));
    return check_launch("grouped_attention_kernel<cache>");
}

extern "C" int masr_avgpool2_time_f32(const float* x, int64_t in_bstride, float* y, int64_t out_bstride, const int* lens,
                                      int B, int out_rows, int D, void* stream) {
    if (B == 0 || out_rows == 0) return MASR_OK;
    MASR_REQUIRE(x && y && lens, "masr_avgpool2_time_f32: null pointer");
    avgpool2_time_kernel<<<dim3(out_rows, B), 256, 0, (cudaStream_t)stream>>>(x, in_bstride, y, out_bstride, lens, out_rows, D);
    return check_launch("avgpool2_time_kernel");
}
