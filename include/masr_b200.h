/* masr_b200 — C ABI of the B200-native MASR inference hot path.
 *
 * The reference (yeyupiaoling/MASR) is pure Python and has NO FFI of its own: its hot path is
 * `MASRPredictor.predict / predict_stream` (masr/predict.py:167,237) -> `AudioFeaturizer.featurize`
 * (masr/data_utils/featurizer/audio_featurizer.py:37) -> `InferencePredictor.predict[_chunk_*]`
 * (masr/infer_utils/inference_predictor.py:52,80) -> TorchScript `get_encoder_out[_chunk]`
 * (masr/model_utils/conformer/model.py:152,169) -> `greedy_decoder` (masr/decoders/ctc_greedy_decoder.py:6).
 * Every library call on that path (torchaudio kaldi.fbank, ATen linear/conv/layer_norm/softmax, numpy
 * argmax) is replaced by one of the entry points below; the Python host code in `masr_b200/` binds them
 * with ctypes and keeps the reference's class/method interface.  INTEGRATION.md shows the stub a MASR
 * maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (caller-owned, allocated e.g. with torch.empty(device='cuda'))
 *     unless its name ends in `_host`; nothing is allocated, freed or retained by the library;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it;
 *   - float32 everywhere ("f32" suffix), row-major, leading dimensions (`ld*`) in elements;
 *   - returns MASR_OK (0) or an error code; `masr_last_error()` has the message (thread-local);
 *   - there is no CPU fallback: without an sm_100 device the calls fail.
 */
#ifndef MASR_B200_H_
#define MASR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MASR_ABI_VERSION 1

enum {
    MASR_OK = 0,
    MASR_ERR_INVALID_ARGUMENT = 10001,
    MASR_ERR_UNSUPPORTED_DEVICE = 10002,
    MASR_ERR_INTERNAL = 10003,
};

/* per-utterance status flags written by masr_wave_gain_f32 */
enum { MASR_STATUS_GAIN_EXCEEDED = 1 };

/* GEMM epilogues */
enum {
    MASR_EPI_BIAS = 0,      /* C = A.W^T + bias                                              */
    MASR_EPI_BIAS_SILU = 1, /* C = silu(A.W^T + bias)                  positionwise.py:37    */
    MASR_EPI_BIAS_RELU = 2, /* C = relu(A.W^T + bias)                  subsampling.py:82,84  */
    MASR_EPI_BIAS_GLU = 3,  /* C[:, j] = v[2j] * sigmoid(v[2j+1])      convolution.py:118    */
    MASR_EPI_BIAS_SCALE = 4,/* C = (A.W^T + bias) * alpha              embedding.py:98       */
    MASR_EPI_RESIDUAL = 5,  /* C = residual + alpha * (A.W^T + bias)   encoder.py:117,131,145,155 */
};

const char* masr_last_error(void);
int masr_abi_version(void);
int masr_check_device(void);

/* ---- audio front-end -------------------------------------------------------------------------- */

/* Host-side staging of a batch: copy B separate float32 host arrays (what MASRPredictor.predict receives one at a
 * time, masr/predict.py:147-164; waves[b] has lengths[b] samples) back to back into the page-locked buffer `pinned`
 * and from there to `dev` (both sum(lengths) floats), packing on up to `nthreads` host threads and issuing each part's
 * cudaMemcpyAsync on `stream` as soon as it is packed.  Returns once the last copy has been issued. */
int masr_stage_waves_f32(const void* const* waves, const int64_t* lengths, int B, float* pinned, float* dev, int nthreads,
                         void* stream);

/* Bytes of scratch masr_wave_gain_f32 needs for B utterances of at most max_samples samples. */
int masr_fbank_workspace_bytes(int B, int64_t max_samples, int64_t* bytes_host);

/* dB normalisation factor per utterance: AudioSegment.normalize / rms_db / gain_db
 * (masr/data_utils/audio.py:287-304,519-529,256-264).  wave: packed float32 samples in [-1,1);
 * offsets: int64[B+1] sample offsets.  gain[b] = 10^((target_db - 10*log10(mean(x^2)))/20);
 * status[b] = MASR_STATUS_GAIN_EXCEEDED where the reference raises ValueError (gain > max_gain_db). */
int masr_wave_gain_f32(const float* wave, const int64_t* offsets, int B, int64_t max_samples, float target_db,
                       float max_gain_db, float* gain, int* status, void* workspace, void* stream);

/* AudioSegment.to('int16') + torchaudio.compliance.kaldi.fbank(num_mel_bins=80, frame_length=25,
 * frame_shift=10, dither=0, sample_frequency=16000) (audio.py:244-254,549-574;
 * audio_featurizer.py:120-138; torchaudio kaldi.py:514-645).  gain may be NULL (no dB normalisation).
 * feats: [B, Fmax, 80] raw log-mel, rows >= the utterance's frame count are zero;
 * num_frames[b] = 1 + (n_b - 400) / 160 (0 if n_b < 400), may be NULL. */
int masr_fbank_f32(const float* wave, const int64_t* offsets, const float* gain, int B, int Fmax, float* feats,
                   int* num_frames, void* stream);

/* ---- encoder building blocks ------------------------------------------------------------------ */

/* GlobalCMVN (masr/model_utils/utils/cmvn.py:29-31) + Conv2d(1,C,3,2) + ReLU
 * (masr/model_utils/conformer/subsampling.py:81-82).  feats [B,Fmax,idim] -> out [B,F1max,W1,C]
 * channels-last; w1 [C,1,3,3] as stored by the reference; mean/istd may both be NULL. */
int masr_conv1_cmvn_relu_f32(const float* feats, const float* mean, const float* istd, const float* w1,
                             const float* b1, float* out, int B, int Fmax, int idim, int F1max, int W1, int C,
                             void* stream);

/* Conv2d(C,C,3,2) + ReLU (subsampling.py:83-84) as an implicit GEMM over the channels-last conv-1
 * activation.  w2p [C, 3, 3, C] = reference weight [co,ci,kh,kw] permuted to [co,kh,kw,ci].
 * out [B, T2max, W2, C] (== the [B*T2max, W2*C] input of the `out` linear, subsampling.py:110). */
int masr_conv2_s2_relu_f32(const float* c1, const float* w2p, const float* b2, float* out, int B, int F1max,
                           int W1, int T2max, int W2, int C, void* stream);

/* torch.nn.functional.linear / Conv1d(k=1) with a fused epilogue: C[M,N] = epi(A[M,K] . W[N,K]^T).
 * K % 16 == 0, lda % 4 == 0.  For MASR_EPI_BIAS_GLU the weight/bias rows are interleaved
 * (row 2j = value j, row 2j+1 = gate j) and C has N/2 columns. */
int masr_gemm_f32(const float* A, int64_t lda, const float* W, const float* bias, const float* residual,
                  int64_t ldr, float* C, int64_t ldc, int M, int N, int K, int epilogue, float alpha,
                  void* stream);

/* Tensor-core (tcgen05/TMEM/TMA) variant of masr_gemm_f32 with fp32-grade results: operands are fp16
 * (h, l) pairs, h = fp16(x), l = fp16((x - h) * 2^11) (masr_split_f16); C ~= Ah.Wh^T + 2^-11 (Ah.Wl^T + Al.Wh^T)
 * accumulated in fp32.  Output: fp32 C and/or the (Ch, Cl) pair the next GEMM consumes (either may be NULL,
 * not both).  K % 64 == 0, lda % 8 == 0, ldc % 8 == 0 (ldc % 4 when only fp32 is written); W is [N, K] dense. */
int masr_gemm_tc_f16x2(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl,
                       const float* bias, const float* residual, int64_t ldr, float* C, void* Ch, void* Cl,
                       int64_t ldc, int M, int N, int K, int epilogue, float alpha, void* stream);

/* Sub-layer output projection (N = 256) + residual add + the LayerNorm(s) that follow it, in ONE tensor-core kernel:
 *   x_new = residual + alpha * (A.W^T + bias)
 *   gamma2 == NULL:  X <- x_new,                      (Yh, Yl) <- LN(x_new; gamma1, beta1)
 *                    (conformer/encoder.py:117 -> 122, 131 -> 141, 145 -> 153: the next sub-layer's pre-norm)
 *   gamma2 != NULL:  X <- LN(x_new; gamma1, beta1),   (Yh, Yl) <- LN(X; gamma2, beta2)
 *                    (encoder.py:155 -> 161 `norm_final` -> the next block's 106 `norm_ff_macaron`, or :342 `after_norm`)
 *   Y2 (optional): fp32 copy of what the pair holds.  X, Y2, Yh, Yl share the row pitch ldx; X may alias residual.
 * Launched as clusters of 2 CTAs (the two 128-column tiles of a row block); the row statistics (two-pass mean / centred
 * variance) cross the pair through distributed shared memory.  Same results as masr_gemm_tc_f16x2(MASR_EPI_RESIDUAL)
 * followed by masr_layernorm_split_f16 / masr_layernorm2_split_f16 up to the summation order of the statistics. */
int masr_gemm_tc_residual_ln_f16x2(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl,
                                   const float* bias, const float* residual, int64_t ldr, float alpha, float* X,
                                   const float* gamma1, const float* beta1, const float* gamma2, const float* beta2,
                                   float* Y2, void* Yh, void* Yl, int64_t ldx, int M, int N, int K, float eps, void* stream);

/* Post-norm form of the same cluster kernel (Squeezeformer blocks, squeezeformer/encoder.py:412-463):
 *   X <- LN(residual + alpha * (A.W^T + bias); gamma, beta)   becomes the stream,
 *   (Yh, Yl) <- ada_scale * X + ada_bias   (the next sub-module's adaptive scale, positionwise.py:57-58; NULL: the pair of X).
 * Replaces masr_gemm_tc_f16x2(MASR_EPI_RESIDUAL) + masr_layernorm_ada_split_f16 (used by the stream pools, where every launch
 * saved counts: a chunk step is latency-bound). */
int masr_gemm_tc_residual_postln_f16x2(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl,
                                       const float* bias, const float* residual, int64_t ldr, float alpha, float* X,
                                       const float* gamma, const float* beta, const float* ada_scale, const float* ada_bias,
                                       void* Yh, void* Yl, int64_t ldx, int M, int N, int K, float eps, void* stream);

/* LayerNorm + Linear in one launch (K = D = 256): C / (Ch, Cl) = epilogue(LN(x; gamma, beta) . W^T + bias) — the pre-norm
 * sub-layer inputs: norm_mha -> linear_q/k/v (encoder.py:122, attention.py:72-74), norm_conv -> pointwise_conv1 + GLU
 * (encoder.py:141, convolution.py:117-118), norm_ff / norm_ff_macaron -> w_1 + SiLU (encoder.py:153/106, positionwise.py:37).
 * Every CTA (pair) takes a contiguous range of output tiles and first normalises the rows of the <= 2 row blocks that range
 * touches into (Ah, Al) — an [M, 256] fp16 (h, l) scratch pair the caller provides; on return it holds LN(x) exactly as
 * masr_layernorm_split_f16 would have written it — then multiplies from it.  epilogue: MASR_EPI_BIAS .. MASR_EPI_BIAS_SCALE.
 * Same results as masr_layernorm_split_f16 + masr_gemm_tc_f16x2, one launch less. */
int masr_gemm_tc_lnpre_f16x2(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* Ah, void* Al,
                             int64_t lda, const void* Wh, const void* Wl, const float* bias, float* C, void* Ch, void* Cl,
                             int64_t ldc, int M, int N, int K, int epilogue, float alpha, void* stream);

/* CTC head without the [M, V] logits: ctc_lo Linear (loss/ctc.py:70) with a GEMM epilogue that keeps, per frame and per
 * 32-column group, (max logit, first argmax, sum exp(x - max)), then a combine kernel -> per-frame argmax id (first
 * maximum, ctc_greedy_decoder.py:21) and max-probability 1 / sum_j exp(x_j - max) (the softmax value of the argmax).
 * workspace: 3 * ceil(V/32) * M * 4 bytes.  Same outputs as masr_gemm_tc_f16x2 + masr_ctc_frame_argmax_f32. */
int masr_ctc_head_argmax_tc_f16x2(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl,
                                  const float* bias, int M, int V, int K, void* workspace, int64_t workspace_bytes,
                                  int* ids, float* maxp, void* stream);

/* Conv2dSubsampling4's first conv (as masr_conv1_cmvn_relu_f32) written as fp16 (h,l) pairs in four
 * (t,f)-parity planes [4][B][(F1max+1)/2][20][C], and its second conv + ReLU (subsampling.py:83-84) as a
 * tensor-core implicit GEMM over those planes: the stride-2 window of tap (kh,kw) is a dense TMA box of plane
 * (kh&1, kw&1).  Wh/Wl: split of the [C, 3,3,C]-permuted weight.  Output rows ((b*T2 + t)*19 + f) x C. */
int masr_conv1_cmvn_relu_planes_f16(const float* feats, const float* mean, const float* istd, const float* w1,
                                    const float* b1, void* planes_h, void* planes_l, int B, int Fmax, int idim,
                                    int F1max, int W1, int C, void* stream);
int masr_conv2_tc_f16x2(const void* c1h, const void* c1l, const void* Wh, const void* Wl, const float* bias,
                        float* out, void* outh, void* outl, int B, int F1, int T2, int C, void* stream);

/* fp32 -> fp16 (h, l) pair, elementwise over n contiguous values. */
int masr_split_f16(const float* x, void* h, void* l, int64_t n, void* stream);

/* torch.nn.LayerNorm(D, eps) over the last dimension (encoder.py:64-72; convolution.py:66). */
int masr_layernorm_f32(const float* x, int64_t ldx, const float* gamma, const float* beta, float* y, int64_t ldy,
                       int M, int D, float eps, void* stream);

/* LayerNorm writing the fp16 (h, l) operand pair of masr_gemm_tc_f16x2 instead of fp32. */
int masr_layernorm_split_f16(const float* x, int64_t ldx, const float* gamma, const float* beta, void* yh, void* yl,
                             int64_t ldy, int M, int D, float eps, void* stream);

/* Two LayerNorms back to back in one pass: y1 = LN(x; gamma1, beta1) as fp32 (row pitch ldx; may alias x), then
 * LN(y1; gamma2, beta2) as fp32 y2 (optional) and as the fp16 pair (row pitch ldy) — `norm_final` of one encoder block
 * followed by the next block's `norm_ff_macaron`, or by `after_norm` after the last block (conformer/encoder.py:161,106,342).
 * Bit-identical to masr_layernorm_f32 followed by masr_layernorm_split_f16. */
int masr_layernorm2_split_f16(const float* x, int64_t ldx, const float* gamma1, const float* beta1, float* y1,
                              const float* gamma2, const float* beta2, float* y2, void* yh, void* yl, int64_t ldy, int M,
                              int D, float eps, void* stream);

/* Squeezeformer helpers.
 *  masr_layernorm_ada_split_f16: y = LayerNorm(x) (optional fp32 copy) and the fp16 pair of ada_scale*y+ada_bias — the
 *    post-norm + adaptive scale of the next sub-module input (squeezeformer/encoder.py:412-463, positionwise.py:57-58);
 *  masr_affine_split_f16: pair of scale*x+bias (scale/bias optional), elementwise;
 *  masr_dwconv_bn_silu_f32: conv-module middle with BatchNorm1d(eval) folded to bn_scale/bn_shift (convolution.py:136-142);
 *  masr_time_reduce_dw_split_f16 / masr_upsample2_add_f32: time reduction depthwise conv (stride 2, kernel 1 or 5) and
 *    recovery `saved[t] + z[t/2]` (time_reduction.py:53-76,174-197; encoder.py:198-204). */
int masr_layernorm_ada_split_f16(const float* x, int64_t ldx, const float* gamma, const float* beta, float* y,
                                 const float* ada_scale, const float* ada_bias, void* yh, void* yl, int64_t ldy, int M, int D,
                                 float eps, void* stream);
int masr_affine_split_f16(const float* x, const float* scale, const float* bias, void* yh, void* yl, int64_t M, int D,
                          void* stream);
int masr_dwconv_bn_silu_f32(const float* g, int64_t ldg, int64_t g_bstride, const float* w, const float* bias,
                            const float* bn_scale, const float* bn_shift, const float* pad_vec, float* y, void* yh, void* yl,
                            int64_t ldy, int64_t y_bstride, const int* in_lens, int B, int C, int kernel_size, int lpad,
                            int out_rows, void* stream);
int masr_time_reduce_dw_split_f16(const float* x, int64_t in_bstride, const float* w, const float* bias, void* yh, void* yl,
                                  int64_t out_bstride, const int* lens, int B, int out_rows, int k, int pad, int D,
                                  void* stream);
int masr_upsample2_add_f32(const float* saved, const float* z, float* out, int64_t full_bstride, int64_t half_bstride, int B,
                           int rows, int D, void* stream);

/* RelPositionMultiHeadedAttention core (masr/model_utils/conformer/attention.py:230-251,107-118):
 * Q rows (b*q_bstride + i), K/V rows (b*k_bstride + j), head h at column h*d_k; P [>=max klen, ldp] =
 * linear_pos(pos_emb) rows aligned with key index j; pos_u/pos_v [H,d_k]; O like Q.
 * q_lens/k_lens int32[B]: valid queries / keys per utterance (rows beyond q_lens are written as 0).
 * Output: fp32 O and/or the fp16 (Oh, Ol) operand pair of masr_gemm_tc_f16x2 (same ldo; either may be NULL). */
int masr_relpos_attention_f32(const float* Q, int64_t ldq, int64_t q_bstride, const float* K, const float* V,
                              int64_t ldk, int64_t k_bstride, const float* P, int64_t ldp, const float* pos_u,
                              const float* pos_v, float* O, void* Oh, void* Ol, int64_t ldo, int64_t o_bstride,
                              const int* q_lens, const int* k_lens, int B, int H, int d_k, int max_q, void* stream);

/* Same result as masr_relpos_attention_f32, computed on the tensor cores (mma.sync m16n8k16) with the FP16x2 operand
 * split (fp32-grade); used by the batched path.  Q is fp32 (the positional biases are added before the split); K, V
 * (row stride ldk halves, head h at column h*d_k) and P = linear_pos(pe) (ldp) arrive as fp16 (h,l) pairs — the qkv
 * GEMM epilogue and the weight loader produce them — so key tiles are 16-byte cp.async copies. */
int masr_relpos_attention_tc(const float* Q, int64_t ldq, int64_t q_bstride, const void* Kh, const void* Kl, const void* Vh,
                             const void* Vl, int64_t ldk, int64_t k_bstride, const void* Ph, const void* Pl, int64_t ldp,
                             const float* pos_u, const float* pos_v, float* O, void* Oh, void* Ol, int64_t ldo,
                             int64_t o_bstride, const int* q_lens, const int* k_lens, int B, int H, int d_k, int max_q,
                             void* stream);

/* masr_relpos_attention_tc on the 5th-generation tensor cores (tcgen05.mma, S and O accumulators in TMEM, K / linear_pos(pe) / V
 * tiles by TMA, V consumed as an MN-major operand, softmax between the two products inside the kernel): one CTA per
 * (utterance, head), for utterances of up to 256 frames — max_q <= 256 and every k_lens[b] <= 256 (the batched whole-utterance
 * path of 10 s audio: T = 248).  Same arguments plus `table_rows` (rows of the P table), same results (fp32-grade). */
int masr_relpos_attention_tc5(const float* Q, int64_t ldq, int64_t q_bstride, const void* Kh, const void* Kl, const void* Vh,
                              const void* Vl, int64_t ldk, int64_t k_bstride, const void* Ph, const void* Pl, int64_t ldp,
                              int64_t table_rows, const float* pos_u, const float* pos_v, float* O, void* Oh, void* Ol,
                              int64_t ldo, int64_t o_bstride, const int* q_lens, const int* k_lens, int B, int H, int d_k,
                              int max_q, void* stream);

/* ConvolutionModule middle (masr/model_utils/conformer/convolution.py:121-126): depthwise Conv1d(k)
 * -> LayerNorm(C) -> SiLU.  y[b,t,:] for t < out_rows from g[b, t - lpad + k, :], k < kernel_size;
 * g rows < 0 read pad_vec (NULL = 0), rows >= in_lens[b] read 0.  w [C,k] (reference [C,1,k]).
 * Output: fp32 y and/or the fp16 (yh, yl) pair (either may be NULL). */
int masr_dwconv_ln_silu_f32(const float* g, int64_t ldg, int64_t g_bstride, const float* w, const float* bias,
                            const float* ln_gamma, const float* ln_beta, const float* pad_vec, float* y, void* yh,
                            void* yl, int64_t ldy, int64_t y_bstride, const int* in_lens, int B, int C,
                            int kernel_size, int lpad, int out_rows, float eps, void* stream);

/* masr_dwconv_ln_silu_f32 with a time stride (1 or 2): y[t] reads g[t*stride - lpad + k] — the strided depthwise conv of
 * the EfficientConformer's block 3 (masr/model_utils/efficient_conformer/convolution.py:40-48, encoder.py:160-175). */
int masr_dwconv_ln_silu_strided_f32(const float* g, int64_t ldg, int64_t g_bstride, const float* w, const float* bias,
                                    const float* ln_gamma, const float* ln_beta, const float* pad_vec, float* y, void* yh,
                                    void* yl, int64_t ldy, int64_t y_bstride, const int* in_lens, int B, int C,
                                    int kernel_size, int lpad, int stride, int out_rows, float eps, void* stream);

/* GroupedRelPositionMultiHeadedAttention core (masr/model_utils/efficient_conformer/attention.py:35-69,120-182): q/k/v
 * [B*bstride, ld] and p [>=max_t, H*d_k] row-major; `group` consecutive frames are viewed as H heads of width group*d_k;
 * frames >= lens[b] read as the reference's zero padding; pos_u/pos_v [H, group*d_k]; outputs for frames < lens[b]. */
int masr_grouped_attention_f32(const float* Q, const float* K, const float* V, const float* P, int64_t ld, int64_t bstride,
                               const float* pos_u, const float* pos_v, float* O, void* Oh, void* Ol, const int* lens,
                               int B, int H, int d_k, int group, int max_t, void* stream);

/* The same with a K|V cache (``forward`` with ``cache``, attention.py:151-158): q_lens[b] query frames at Q rows
 * b*q_bstride.. (pitch ldq), k_lens[b] key/value frames = [cache ++ chunk] at K/V rows b*k_bstride.. (pitch ldk); P row j
 * belongs to key j; queries are grouped from the first chunk frame, keys from key 0.  Outputs in Q's layout. */
int masr_grouped_attention_cache_f32(const float* Q, int64_t ldq, int64_t q_bstride, const float* K, const float* V,
                                     int64_t ldk, int64_t k_bstride, const float* P, const float* pos_u, const float* pos_v,
                                     float* O, void* Oh, void* Ol, const int* q_lens, const int* k_lens, int B, int H,
                                     int d_k, int group, int max_q, void* stream);

/* AvgPool1d(2, 2, ceil_mode=True, count_include_pad=False) over time per utterance (efficient_conformer/encoder.py:
 * 173-175): y[b,t] = mean(x[b,2t], x[b,2t+1]) (single element at an odd tail), rows >= ceil(len/2) are 0. */
int masr_avgpool2_time_f32(const float* x, int64_t in_bstride, float* y, int64_t out_bstride, const int* lens, int B,
                           int out_rows, int D, void* stream);

/* One time step of torch.nn.LSTM as DeepSpeech2 uses it (masr/model_utils/deepspeech2/encoder.py:36-45, packed ragged
 * batches): gates_x [B*bstride, 4H] = W_ih x + b_ih + b_hh (gate order i,f,g,o); Whh [4H, H]; states transposed and
 * batch-chunked h_*_T [ceil(B/32)][H][32] (ping-pong: in != out), c_state [B][H]; utterance b is active while
 * step < lens[b] and uses time index step (forward) or lens[b]-1-step (reverse); h_t is written to
 * out[(b*bstride + t), col_off + u] as fp32 and/or fp16 pair. */
int masr_lstm_step_f32(const float* gates_x, int64_t ldg, int64_t bstride, const float* Whh, const float* h_in_T,
                       float* h_out_T, float* c_state, float* out, void* outh, void* outl, int64_t ld_out, int col_off,
                       const int* lens, int B, int H, int step, int reverse, void* stream);

/* All T steps of one LSTM layer / direction in ONE persistent launch: every CTA keeps its slice of W_hh (the four gate rows
 * of 8 hidden units) resident in shared memory, h_{t-1} is streamed through a shared-memory window and the steps are
 * separated by a grid-wide barrier.  Same results as T calls of masr_lstm_step_f32 up to the order of the K summation.
 * h0_T / hN_T: initial / final hidden state, transposed [ceil(B/32)][H][32]; c_state [B][H] updated in place; workspace
 * from masr_lstm_seq_workspace_bytes.  H % 128 == 0, H <= 1024 (the shipped configs: 1024). */
int masr_lstm_seq_workspace_bytes(int B, int H, int64_t* bytes);
int masr_lstm_seq_f32(const float* gates_x, int64_t ldg, int64_t bstride, const float* Whh, const float* h0_T, float* hN_T,
                      float* c_state, float* out, void* outh, void* outl, int64_t ld_out, int col_off, const int* lens, int B,
                      int H, int T, int reverse, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- batched chunk (streaming) state ------------------------------------------------------------ */

/* Append the chunk's new rows to every slot's cache (the `torch.cat` on time of the attention K|V cache, conformer/
 * attention.py:218-225), on the device: for slot s and t < cnt[s]
 *   dst{0,1}[(s*cap + base[s] + t) * dst_pitch + c] = src{0,1}[(s*rows_per_slot + t) * src_pitch + col0 + c],  c < row_bytes
 * (all sizes in BYTES, multiples of 16; src1/dst1 may both be NULL; the pair form moves the fp16 (h,l) operand halves). */
int masr_stream_append_rows(const void* src0, const void* src1, int64_t src_pitch_bytes, int64_t col0_bytes, int row_bytes,
                            void* dst0, void* dst1, int64_t dst_pitch_bytes, int64_t cap, const int* base, const int* cnt,
                            int rows_per_slot, int S, void* stream);

/* Slide every slot's conv-module left context (convolution.py:105-109, `new_cache = x[:, :, -lorder:]`): for slot s with
 * n = cnt[s] > 0, rows [0, lorder) <- rows [n, n + lorder) of its [rows_per_slot, row_bytes] block of x0 (and x1). */
int masr_stream_shift_cache(void* x0, void* x1, int64_t rows_per_slot, int lorder, int row_bytes, const int* cnt, int S,
                            void* stream);

/* ---- CTC head / greedy decode ------------------------------------------------------------------- */

/* softmax statistics of CTCLoss.softmax (masr/model_utils/loss/ctc.py:70) fused with the argmax of
 * greedy_decoder (masr/decoders/ctc_greedy_decoder.py:21-22): ids[m] = first argmax_v, maxp[m] =
 * softmax(logits[m])[ids[m]].  probs (optional, may be NULL): full posterior [M, ldp]. */
int masr_ctc_frame_argmax_f32(const float* logits, int64_t ldl, int M, int V, int* ids, float* maxp, float* probs,
                              int64_t ldp, void* stream);

/* greedy_decoder's collapse (ctc_greedy_decoder.py:23-30): per utterance b over frames t < lens[b]
 * (rows b*bstride + t): tokens = ids with consecutive repeats merged and `blank` dropped;
 * psum/pcount = float32 left-to-right sum / count of maxp over non-blank frames (score = 100*psum/pcount). */
int masr_ctc_greedy_collapse(const int* ids, const float* maxp, int64_t bstride, const int* lens, int B, int blank,
                             int* tokens, int64_t tok_stride, int* ntok, float* psum, int* pcount, void* stream);

/* ---- CTC prefix beam search (no LM) --------------------------------------------------------------------
 * Replaces the external paddlespeech_ctcdecoders call behind masr/decoders/swig_wrapper.py:35-64
 * (`ctc_beam_search_decoding(probs, vocab, beam_size, cutoff_prob, cutoff_top_n, scorer, blank_id)`), scorer = None.
 * PARITY UNPINNED (library absent): semantics per SURVEY.md Appendix D, checked against oracle/beam.py.
 *  masr_ctc_topk_f32:     per frame the <= top_n (<= 40) most probable tokens, cut where the cumulative probability
 *                         reaches cutoff_prob: cand_id / cand_logp [M, 40] (best first), cand_cnt [M];
 *  masr_ctc_prefix_beam:  per utterance (rows b*bstride + t, t < lens[b]) the best prefix: out_tok [B, tok_stride],
 *                         out_n [B], out_score [B] = log P(prefix); beam_size <= 512.  Scratch sizes from
 *                         masr_ctc_prefix_beam_workspace (pool floats total, trie ints per utterance for each of
 *                         trie_parent / trie_tok). */
int masr_ctc_topk_f32(const float* logits, int64_t ldl, int M, int V, int top_n, float cutoff_prob, int* cand_id,
                      float* cand_logp, int* cand_cnt, void* stream);
int masr_ctc_prefix_beam_workspace(int B, int Tmax, int64_t* pool_floats_host, int64_t* trie_ints_per_utt_host);
int masr_ctc_prefix_beam(const int* cand_id, const float* cand_logp, const int* cand_cnt, int64_t bstride, const int* lens,
                         int B, int beam_size, int blank, float* pool, int* trie_parent, int* trie_tok, int64_t trie_cap,
                         int* out_tok, int64_t tok_stride, int* out_n, float* out_score, void* stream);

/* Streaming form of masr_ctc_prefix_beam (beam_search_decoder.py:75-96: CTCBeamSearchDecoder.next() + decode(),
 * reset_state()): the same search fed chunk by chunk.  lens[b] = frames of THIS chunk; resume = 0 starts a new utterance
 * (reset_decoder), != 0 continues from state_i / state_f (sizes per utterance from masr_ctc_prefix_beam_state_size);
 * trie_parent / trie_tok are sized for the whole stream (masr_ctc_prefix_beam_workspace with Tmax = its frame count) and
 * persist between calls.  Outputs: the best prefix and its score after all frames seen so far — identical to one
 * masr_ctc_prefix_beam call over the concatenated chunks. */
int masr_ctc_prefix_beam_state_size(int64_t* ints_per_utt, int64_t* floats_per_utt);
int masr_ctc_prefix_beam_stream(const int* cand_id, const float* cand_logp, const int* cand_cnt, int64_t bstride,
                                const int* lens, int B, int beam_size, int blank, float* pool, int* trie_parent,
                                int* trie_tok, int64_t trie_cap, int* state_i, float* state_f, int resume, int* out_tok,
                                int64_t tok_stride, int* out_n, float* out_score, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MASR_B200_H_ */
