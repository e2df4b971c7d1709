// CTC head tail: per-frame softmax statistics + argmax over the vocabulary, then greedy (best-path)
// collapse per utterance.
//
// Replaces
//   CTCLoss.softmax (loss/ctc.py:70, the softmax half; the ctc_lo GEMM is gemm.cu)
//   greedy_decoder (decoders/ctc_greedy_decoder.py:21-30): argmax over the probabilities (first index
//   on ties), score = mean of the max-probabilities of the non-blank frames (left-to-right float32 sum),
//   collapse repeats, drop blank.
// The reference copies the whole [B,T,V] posterior to the host (inference_predictor.py:64) and decodes
// with numpy; here only ids/score leave the GPU (the posterior can still be requested for the
// `InferencePredictor.predict` seam).
//
// HBM-bound: one read of the logits row (V*4 bytes per frame).
#include <math.h>

#include "common.cuh"

namespace masr {

// One CTA per frame.  argmax(softmax(x)) == argmax(x) (exp is monotone; ties keep the lowest index),
// max-prob = 1 / sum_j exp(x_j - max).
__global__ void __launch_bounds__(256) ctc_frame_argmax_kernel(const float* __restrict__ logits, int64_t ldl, int V,
                                                               int* __restrict__ ids, float* __restrict__ maxp,
                                                               float* __restrict__ probs, int64_t ldp) {
    const int row = blockIdx.x;
    const float* x = logits + (int64_t)row * ldl;
    float m = -INFINITY;
    int mi = 0x7fffffff;
    for (int j = threadIdx.x; j < V; j += 256) {
        float v = __ldg(x + j);
        if (v > m) { m = v; mi = j; }              // ascending j per thread: strict > keeps the first max
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float om = __shfl_xor_sync(0xffffffffu, m, o);
        int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
    }
    __shared__ float sm[8];
    __shared__ int si[8];
    __shared__ float ssum[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sm[warp] = m; si[warp] = mi; }
    __syncthreads();
    m = sm[0]; mi = si[0];
#pragma unroll
    for (int w = 1; w < 8; ++w)
        if (sm[w] > m || (sm[w] == m && si[w] < mi)) { m = sm[w]; mi = si[w]; }
    float s = 0.f;
    for (int j = threadIdx.x; j < V; j += 256) s += expf(__ldg(x + j) - m);
    s = warp_sum(s);
    if (lane == 0) ssum[warp] = s;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += ssum[w];
    if (threadIdx.x == 0) { ids[row] = mi; maxp[row] = 1.0f / tot; }
    if (probs) {
        float* pr = probs + (int64_t)row * ldp;
        for (int j = threadIdx.x; j < V; j += 256) pr[j] = expf(__ldg(x + j) - m) / tot;
    }
}

// One thread per utterance: sequential scan (T <= a few thousand frames; latency-trivial).
__global__ void ctc_greedy_collapse_kernel(const int* __restrict__ ids, const float* __restrict__ maxp, int64_t bstride,
                                           const int* __restrict__ lens, int B, int blank, int prev_id_in,
                                           int* __restrict__ tokens, int64_t tok_stride, int* __restrict__ ntok,
                                           float* __restrict__ psum, int* __restrict__ pcount) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int T = lens[b];
    const int* id = ids + (int64_t)b * bstride;
    const float* mp = maxp + (int64_t)b * bstride;
    int* tk = tokens + (int64_t)b * tok_stride;
    int prev = prev_id_in, n = 0, cnt = 0;
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
        const int i = id[t];
        if (i != blank) { acc += mp[t]; ++cnt; }   // float32 running sum, in frame order
        if (i != prev && i != blank) tk[n++] = i;
        prev = i;
    }
    ntok[b] = n;
    psum[b] = acc;
    pcount[b] = cnt;
}

}  // namespace masr

using namespace masr;

extern "C" int masr_ctc_frame_argmax_f32(const float* logits, int64_t ldl, int M, int V, int* ids, float* maxp,
                                         float* probs, int64_t ldp, void* stream) {
    if (M == 0) return MASR_OK;
    MASR_REQUIRE(logits && ids && maxp && V > 0, "masr_ctc_frame_argmax_f32: bad argument");
    ctc_frame_argmax_kernel<<<M, 256, 0, (cudaStream_t)stream>>>(logits, ldl, V, ids, maxp, probs, ldp);
    return check_launch("ctc_frame_argmax_kernel");
}

extern "C" int masr_ctc_greedy_collapse(const int* ids, const float* maxp, int64_t bstride, const int* lens, int B,
                                        int blank, int* tokens, int64_t tok_stride, int* ntok, float* psum,
                                        int* pcount, void* stream) {
    if (B == 0) return MASR_OK;
    MASR_REQUIRE(ids && maxp && lens && tokens && ntok && psum && pcount, "masr_ctc_greedy_collapse: null pointer");
    ctc_greedy_collapse_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(ids, maxp, bstride, lens, B, blank, -1,
                                                                               tokens, tok_stride, ntok, psum, pcount);
    return check_launch("ctc_greedy_collapse_kernel");
}
