// LSTM recurrence of DeepSpeech2 (masr/model_utils/deepspeech2/encoder.py:36-45 -> torch.nn.LSTM), one launch per
// time step, fp32:
//     gates = gates_x[b, t] + W_hh . h_{t-1}[b]            (gates_x = W_ih x_t + b_ih + b_hh, a GEMM done beforehand)
//     i, f, g, o = sigmoid, sigmoid, tanh, sigmoid   (PyTorch gate order)
//     c_t = f * c_{t-1} + i * g ;  h_t = o * tanh(c_t)
// Ragged batches follow pack_padded_sequence semantics (encoder.py:42-44): utterance b is active for steps s < len_b and
// reads/writes time index t = s (forward) or len_b - 1 - s (reverse); afterwards its state is frozen.
//
// One warp per hidden unit (its 4 gate rows), one lane per utterance: no cross-lane reduction, W_hh rows are warp-uniform
// 128-bit loads, the state is kept transposed ([H][32]) so the lanes' reads are one 128-byte line per k.
// FMA-pipe bound (B*4H*H MACs per step); round-1 implementation (DeepSpeech2 is the lowest-priority model).
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace masr {

constexpr int LSTM_UNITS = 8;     // hidden units (warps) per CTA
constexpr int LSTM_BP = 32;       // batch lanes per pass

__global__ void __launch_bounds__(LSTM_UNITS * 32) lstm_step_kernel(
    const float* __restrict__ gates_x, int64_t ldg, int64_t bstride, const float* __restrict__ Whh,
    const float* __restrict__ h_in_T, float* __restrict__ h_out_T, float* __restrict__ c_state, float* __restrict__ out,
    __half* __restrict__ outh, __half* __restrict__ outl, int64_t ld_out, int col_off, const int* __restrict__ lens, int B,
    int H, int step, int reverse) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = blockIdx.x * LSTM_UNITS + warp;
    if (u >= H) return;
    const float* w_i = Whh + (int64_t)(0 * H + u) * H;
    const float* w_f = Whh + (int64_t)(1 * H + u) * H;
    const float* w_g = Whh + (int64_t)(2 * H + u) * H;
    const float* w_o = Whh + (int64_t)(3 * H + u) * H;
    const int nb = (B + LSTM_BP - 1) / LSTM_BP;
    for (int bc = 0; bc < nb; ++bc) {
        const int b = bc * LSTM_BP + lane;
        const float* hT = h_in_T + (int64_t)bc * H * LSTM_BP + lane;     // [chunk][H][32]
        float ai = 0.f, af = 0.f, ag = 0.f, ao = 0.f;
#pragma unroll 4
        for (int k = 0; k < H; k += 4) {
            const float4 wi = ldg_f4(w_i + k), wf = ldg_f4(w_f + k), wg = ldg_f4(w_g + k), wo = ldg_f4(w_o + k);
            const float h0 = __ldg(hT + (int64_t)(k + 0) * LSTM_BP), h1 = __ldg(hT + (int64_t)(k + 1) * LSTM_BP);
            const float h2 = __ldg(hT + (int64_t)(k + 2) * LSTM_BP), h3 = __ldg(hT + (int64_t)(k + 3) * LSTM_BP);
            ai = fmaf(wi.x, h0, ai); ai = fmaf(wi.y, h1, ai); ai = fmaf(wi.z, h2, ai); ai = fmaf(wi.w, h3, ai);
            af = fmaf(wf.x, h0, af); af = fmaf(wf.y, h1, af); af = fmaf(wf.z, h2, af); af = fmaf(wf.w, h3, af);
            ag = fmaf(wg.x, h0, ag); ag = fmaf(wg.y, h1, ag); ag = fmaf(wg.z, h2, ag); ag = fmaf(wg.w, h3, ag);
            ao = fmaf(wo.x, h0, ao); ao = fmaf(wo.y, h1, ao); ao = fmaf(wo.z, h2, ao); ao = fmaf(wo.w, h3, ao);
        }
        float* hTo = h_out_T + (int64_t)bc * H * LSTM_BP + (int64_t)u * LSTM_BP + lane;
        const float h_prev = __ldg(hT + (int64_t)u * LSTM_BP);
        const int len = b < B ? lens[b] : 0;
        if (step >= len) { *hTo = h_prev; continue; }                   // finished (or padding lane): state frozen
        const int t = reverse ? len - 1 - step : step;
        const float* gx = gates_x + ((int64_t)b * bstride + t) * ldg;
        const float gi = sigmoid_f(gx[0 * H + u] + ai), gf = sigmoid_f(gx[1 * H + u] + af);
        const float gg = tanhf(gx[2 * H + u] + ag), go = sigmoid_f(gx[3 * H + u] + ao);
        float* cp = c_state + (int64_t)b * H + u;
        const float c = gf * (*cp) + gi * gg;
        const float h = go * tanhf(c);
        *cp = c;
        *hTo = h;
        const int64_t o = ((int64_t)b * bstride + t) * ld_out + col_off + u;
        if (out) out[o] = h;
        if (outh) {
            const __half hh = __float2half_rn(h);
            outh[o] = hh;
            outl[o] = __float2half_rn((h - __half2float(hh)) * 2048.0f);
        }
    }
}

}  // namespace masr

using namespace masr;

extern "C" int masr_lstm_step_f32(const float* gates_x, int64_t ldg, int64_t bstride, const float* Whh, const float* h_in_T,
                                  float* h_out_T, float* c_state, float* out, void* outh, void* outl, int64_t ld_out,
                                  int col_off, const int* lens, int B, int H, int step, int reverse, void* stream) {
    if (B == 0) return MASR_OK;
    MASR_REQUIRE(gates_x && Whh && h_in_T && h_out_T && c_state && lens && (out || (outh && outl)), "masr_lstm_step_f32: null pointer");
    MASR_REQUIRE(H % 4 == 0 && h_in_T != h_out_T, "masr_lstm_step_f32: H %% 4 == 0 and distinct in/out state buffers required");
    lstm_step_kernel<<<(H + LSTM_UNITS - 1) / LSTM_UNITS, LSTM_UNITS * 32, 0, (cudaStream_t)stream>>>(
        gates_x, ldg, bstride, Whh, h_in_T, h_out_T, c_state, out, (__half*)outh, (__half*)outl, ld_out, col_off, lens, B, H, step,
        reverse);
    return check_launch("lstm_step_kernel");
}
