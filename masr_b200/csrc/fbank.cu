// Fused audio front-end: dB normalisation -> int16 quantisation -> Kaldi fbank.
//
// Replaces, for a packed batch of utterances resident in HBM:
//   AudioSegment.rms_db/normalize/gain_db          masr/data_utils/audio.py:519-529,287-304,256-264
//   AudioSegment.to('int16')                       masr/data_utils/audio.py:244-254,549-574
//   AudioFeaturizer._compute_fbank                 masr/data_utils/featurizer/audio_featurizer.py:120-138
//   torchaudio.compliance.kaldi.fbank              kaldi.py:514-645 (_get_strided :44-83, _get_window :154-217,
//                                                  get_mel_banks :436-511)
//
// Kernels
//   wave_sumsq_kernel   per-utterance sum of squares, fixed-order double partials (deterministic)
//   wave_gain_kernel    mean square -> gain factor (float32 chain as numpy does it), error flag if the
//                       required gain exceeds 300 dB (the reference raises ValueError, audio.py:302)
//   fbank_kernel        one warp per frame: quantise 400 samples on the fly, DC removal, pre-emphasis,
//                       povey window, 512-point real FFT (as a 256-point complex FFT in shared memory),
//                       power spectrum, sparse triangular mel filters, log, coalesced 80-float store.
//
// Roofline: HBM-bound by design — algorithmic bytes per 10 s utterance = 640 KB float32 samples in +
// 319 KB features out; each sample is re-read 2.5x by overlapping frames, served by L1/L2.
#include <math.h>
#include <mutex>
#include <stdarg.h>

#include "common.cuh"

namespace masr {

constexpr int kFrameLen = 400;
constexpr int kFrameShift = 160;
constexpr int kNfft = 512;
constexpr int kHalf = 256;          // complex FFT size
constexpr int kMel = 80;
constexpr int kMaxMelW = 640;       // non-zero mel weights (measured: 510 for 80 bins @16 kHz/512)
constexpr int kSumChunk = 8192;

struct FbankTables {
    float window[kFrameLen];
    float2 twiddle[kHalf / 2];      // exp(-2*pi*i*k/256)
    float2 post[kHalf + 1];         // exp(-2*pi*i*k/512), k = 0..256
    int mel_start[kMel];
    int mel_len[kMel];
    int mel_off[kMel];
    float mel_w[kMaxMelW];
};

__device__ FbankTables g_tab;   // read through L1 with __ldg: lanes index it divergently

static std::mutex g_tab_mu;
static bool g_tab_ready[64] = {false};
static int g_tab_status = 0;

static double mel_scale_d(double f) { return 1127.0 * log(1.0 + f / 700.0); }

static void build_tables() {
    static FbankTables h;
    g_tab_status = 0;
    // povey window: hann(400, periodic=False) ** 0.85 in float32  (kaldi.py:100)
    for (int i = 0; i < kFrameLen; ++i) {
        float hann = (float)(0.5 - 0.5 * cos(2.0 * M_PI * i / (kFrameLen - 1)));
        h.window[i] = powf(hann, 0.85f);
    }
    for (int k = 0; k < kHalf / 2; ++k) {
        double a = -2.0 * M_PI * k / kHalf;
        h.twiddle[k] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int k = 0; k <= kHalf; ++k) {
        double a = -2.0 * M_PI * k / kNfft;
        h.post[k] = make_float2((float)cos(a), (float)sin(a));
    }
    // mel filterbank in float32, the way get_mel_banks computes it (kaldi.py:463-499)
    const float mlo = (float)mel_scale_d(20.0), mhi = (float)mel_scale_d(8000.0);
    const float delta = (mhi - mlo) / (kMel + 1);
    const float bin_w = 16000.0f / kNfft;
    int off = 0;
    for (int m = 0; m < kMel; ++m) {
        float left = mlo + m * delta, center = mlo + (m + 1.0f) * delta, right = mlo + (m + 2.0f) * delta;
        int start = -1, len = 0;
        for (int k = 0; k < kHalf; ++k) {
            float mel = 1127.0f * logf(1.0f + (bin_w * k) / 700.0f);
            float up = (mel - left) / (center - left), down = (right - mel) / (right - center);
            float w = fmaxf(0.f, fminf(up, down));
            if (w > 0.f) {
                if (start < 0) start = k;
                len = k - start + 1;
            }
        }
        h.mel_start[m] = start < 0 ? 0 : start;
        h.mel_len[m] = len;
        h.mel_off[m] = off;
        for (int k = h.mel_start[m]; k < h.mel_start[m] + len; ++k) {
            float mel = 1127.0f * logf(1.0f + (bin_w * k) / 700.0f);
            float up = (mel - left) / (center - left), down = (right - mel) / (right - center);
            if (off >= kMaxMelW) { g_tab_status = MASR_ERR_INTERNAL; return; }
            h.mel_w[off++] = fmaxf(0.f, fminf(up, down));
        }
    }
    cudaError_t e = cudaMemcpyToSymbol(g_tab, &h, sizeof(h));
    if (e != cudaSuccess) {
        set_last_error("fbank tables: %s", cudaGetErrorString(e));
        g_tab_status = (int)e;
    }
}

// ---- dB normalisation ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wave_sumsq_kernel(const float* __restrict__ wave,
                                                         const int64_t* __restrict__ offs, double* __restrict__ partial,
                                                         int max_chunks) {
    const int b = blockIdx.y, c = blockIdx.x;
    const int64_t beg = offs[b], n = offs[b + 1] - beg;
    const int64_t s0 = (int64_t)c * kSumChunk;
    if (s0 >= n) return;
    const int64_t s1 = min(n, s0 + (int64_t)kSumChunk);
    double acc = 0.0;
    for (int64_t i = s0 + threadIdx.x; i < s1; i += 256) {
        float x = __ldg(wave + beg + i);
        float sq = x * x;                      // `samples ** 2` is a float32 array in the reference
        acc += (double)sq;
    }
    acc = warp_sum(acc);
    __shared__ double sw[8];
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sw[w];
        partial[(int64_t)b * max_chunks + c] = t;
    }
}

__global__ void wave_gain_kernel(const int64_t* __restrict__ offs, const double* __restrict__ partial, int max_chunks,
                                 int B, float target_db, float max_gain_db, float* __restrict__ gain,
                                 int* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int64_t n = offs[b + 1] - offs[b];
    const int chunks = (int)((n + kSumChunk - 1) / kSumChunk);
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += partial[(int64_t)b * max_chunks + c];
    float ms = n > 0 ? (float)(s / (double)n) : 0.f;       // np.mean -> float32
    if (ms == 0.f) ms = 1.f;                               // audio.py:526-527
    float rms_db = 10.f * (float)log10((double)ms);        // float32 scalar chain (audio.py:529)
    float g = target_db - rms_db;
    int bad = g > max_gain_db;
    if (bad) g = max_gain_db;
    float e = g / 20.f;
    gain[b] = (float)pow(10.0, (double)e);                 // 10. ** float32 -> float32 (audio.py:264)
    status[b] = bad ? MASR_STATUS_GAIN_EXCEEDED : 0;
}

// ---- fbank ----------------------------------------------------------------------------------------
constexpr int kWarpsPerBlock = 4;
constexpr int kFramesPerWarp = 2;

__global__ void __launch_bounds__(kWarpsPerBlock * 32) fbank_kernel(
    const float* __restrict__ wave, const int64_t* __restrict__ offs, const float* __restrict__ gain,
    float* __restrict__ feats, int* __restrict__ nframes_out, int Fmax) {
    __shared__ float2 s_z[kWarpsPerBlock][kHalf + 1];      // FFT work area (+1: Z[256] alias slot)
    __shared__ float s_p[kWarpsPerBlock][kFrameLen];      // frame samples, later the power spectrum P[0..256]

    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t beg = offs[b], n = offs[b + 1] - beg;
    const int F = n < kFrameLen ? 0 : 1 + (int)((n - kFrameLen) / kFrameShift);
    if (blockIdx.x == 0 && threadIdx.x == 0 && nframes_out) nframes_out[b] = F;
    const float g = gain ? gain[b] : 1.f;
    float2* z = s_z[warp];
    float* pw = s_p[warp];

    for (int fi = 0; fi < kFramesPerWarp; ++fi) {
        const int f = (blockIdx.x * kWarpsPerBlock + warp) * kFramesPerWarp + fi;
        if (f >= Fmax) break;                              // warp-uniform
        float* out = feats + ((int64_t)b * Fmax + f) * kMel;
        if (f >= F) {                                      // padded frame: deterministic zeros
            for (int m = lane; m < kMel; m += 32) out[m] = 0.f;
            continue;
        }
        const float* src = wave + beg + (int64_t)f * kFrameShift;
        // 1. load + quantise (audio.py:264,566-574: fl(x*g), *2^15, clip, truncate) and frame mean
        float q[13];
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 13; ++j) {
            int i = lane + 32 * j;
            float v = 0.f;
            if (i < kFrameLen) {
                v = __ldg(src + i) * g;
                v = v * 32768.0f;
                v = fminf(fmaxf(v, -32768.0f), 32767.0f);
                v = truncf(v);
            }
            q[j] = v;
            sum += v;
        }
        sum = warp_sum(sum);
        const float mean = sum / (float)kFrameLen;         // kaldi.py:183-186
#pragma unroll
        for (int j = 0; j < 13; ++j) {
            int i = lane + 32 * j;
            if (i < kFrameLen) pw[i] = q[j] - mean;
        }
        __syncwarp();
        // 2. pre-emphasis (replicate-left), window, pack even/odd samples into complex points,
        //    stored bit-reversed for the in-place radix-2 DIT FFT
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int c = lane + 32 * j;                         // complex point index 0..255
            int i0 = 2 * c, i1 = 2 * c + 1;
            float re = 0.f, im = 0.f;
            if (i0 < kFrameLen) {
                float cur = pw[i0], prev = pw[i0 > 0 ? i0 - 1 : 0];
                re = (cur - 0.97f * prev) * __ldg(&g_tab.window[i0]);
                float cur1 = pw[i1];
                im = (cur1 - 0.97f * cur) * __ldg(&g_tab.window[i1]);
            }
            z[__brev((unsigned)c) >> 24] = make_float2(re, im);
        }
        __syncwarp();
        // 3. 256-point complex FFT, 8 radix-2 stages, 4 butterflies per lane per stage
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int half = 1 << s;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                int j = lane + 32 * r;                     // butterfly 0..127
                int pos = j & (half - 1);
                int i0 = ((j >> s) << (s + 1)) + pos;
                int i1 = i0 + half;
                float2 w = __ldg(&g_tab.twiddle[pos << (7 - s)]);
                float2 a = z[i0], bb = z[i1];
                float tr = bb.x * w.x - bb.y * w.y;
                float ti = bb.x * w.y + bb.y * w.x;
                z[i0] = make_float2(a.x + tr, a.y + ti);
                z[i1] = make_float2(a.x - tr, a.y - ti);
            }
            __syncwarp();
        }
        if (lane == 0) z[kHalf] = z[0];
        __syncwarp();
        // 4. real-FFT post-processing -> power spectrum P[k], k = 0..256  (kaldi.py:616-618)
        for (int k = lane; k <= kHalf; k += 32) {
            float2 zk = z[k], zn = z[kHalf - k];
            float er = 0.5f * (zk.x + zn.x), ei = 0.5f * (zk.y - zn.y);    // even part
            float orr = 0.5f * (zk.y + zn.y), oi = -0.5f * (zk.x - zn.x);  // odd part
            float2 w = __ldg(&g_tab.post[k]);
            float xr = er + (orr * w.x - oi * w.y);
            float xi = ei + (orr * w.y + oi * w.x);
            pw[k] = xr * xr + xi * xi;
        }
        __syncwarp();
        // 5. mel filterbank (sparse triangles), log floor, store  (kaldi.py:630-633)
        for (int m = lane; m < kMel; m += 32) {
            const int st = __ldg(&g_tab.mel_start[m]), ln = __ldg(&g_tab.mel_len[m]), of = __ldg(&g_tab.mel_off[m]);
            float e = 0.f;
            for (int k = 0; k < ln; ++k) e = fmaf(pw[st + k], __ldg(&g_tab.mel_w[of + k]), e);
            out[m] = logf(fmaxf(e, 1.1920928955078125e-07f));
        }
        __syncwarp();
    }
}

}  // namespace masr

using namespace masr;

extern "C" int masr_fbank_workspace_bytes(int B, int64_t max_samples, int64_t* bytes) {
    MASR_REQUIRE(bytes && B >= 0 && max_samples >= 0, "masr_fbank_workspace_bytes: bad argument");
    int64_t chunks = (max_samples + kSumChunk - 1) / kSumChunk;
    if (chunks < 1) chunks = 1;
    *bytes = (int64_t)B * chunks * (int64_t)sizeof(double);
    return MASR_OK;
}

extern "C" int masr_wave_gain_f32(const float* wave, const int64_t* offsets, int B, int64_t max_samples,
                                  float target_db, float max_gain_db, float* gain, int* status, void* workspace,
                                  void* stream) {
    if (B == 0) return MASR_OK;
    MASR_REQUIRE(wave && offsets && gain && status && workspace, "masr_wave_gain_f32: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int chunks = (int)((max_samples + kSumChunk - 1) / kSumChunk);
    if (chunks < 1) chunks = 1;
    wave_sumsq_kernel<<<dim3(chunks, B), 256, 0, st>>>(wave, offsets, (double*)workspace, chunks);
    int rc = check_launch("wave_sumsq_kernel");
    if (rc) return rc;
    wave_gain_kernel<<<(B + 127) / 128, 128, 0, st>>>(offsets, (const double*)workspace, chunks, B, target_db,
                                                       max_gain_db, gain, status);
    return check_launch("wave_gain_kernel");
}

extern "C" int masr_fbank_f32(const float* wave, const int64_t* offsets, const float* gain, int B, int Fmax,
                              float* feats, int* num_frames, void* stream) {
    if (B == 0 || Fmax == 0) return MASR_OK;
    MASR_REQUIRE(wave && offsets && feats, "masr_fbank_f32: null pointer");
    {
        std::lock_guard<std::mutex> lk(g_tab_mu);
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) dev = 0;
        if (!g_tab_ready[dev]) {
            build_tables();
            if (g_tab_status) return g_tab_status;
            g_tab_ready[dev] = true;
        }
    }
    const int frames_per_block = kWarpsPerBlock * kFramesPerWarp;
    dim3 grid((Fmax + frames_per_block - 1) / frames_per_block, B);
    fbank_kernel<<<grid, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(wave, offsets, gain, feats, num_frames, Fmax);
    return check_launch("fbank_kernel");
}
