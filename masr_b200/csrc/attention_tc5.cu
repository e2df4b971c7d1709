// Relative-position attention core on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by
// TMA / written in the UMMA shared-memory layout), for utterances of up to 256 encoder frames — the batched whole-utterance
// path of the headline config (T = 248).  Same contract and results (fp32-grade, FP16x2 operand split) as
// relpos_attention_mma_kernel (attention_mma.cu, legacy mma.sync), which stays the kernel for longer utterances.
//
//   S      = [q+u | q+v] . [k | p]^T / sqrt(d_k)          (128-wide contraction; no rel_shift, attention.py:245-247)
//   out    = softmax_j(S[:, j < klen]) . v                 (attention.py:107-118)
//   x.y   ~= xh.yh + 2^-11 (xh.yl + xl.yh)                 (two fp32 TMEM accumulators, as in tc_gemm.cu)
//
// One CTA per (utterance, head); the utterance's queries are processed as one or two 128-row tiles.  Per tile:
//   1. TMA: K and linear_pos(pe) rows 0..255 of this head (fp16 (h,l) pairs, written by the qkv GEMM epilogue / at load time)
//      -> B operand [256 keys x (64 | 64)] in region R1; meanwhile all warps build A = [q+u | q+v] (fp32 add, then split) for
//      the tile's 128 rows in region R2, directly in the 128-byte-swizzled K-major UMMA layout;
//   2. 24 tcgen05.mma (M=128, N=256, K=16): S main -> TMEM columns [0,256), S correction -> [256,512);
//   3. 8 softmax warps (TMEM lane quarter x column half): pass 1 row maximum, pass 2 p = 2^(s - max) (keys >= klen -> 0), row
//      sums, and p as fp16 (h,l) pairs written into R1 as the A operand of the second product (un-normalised, flash-style);
//      meanwhile TMA loads V [256 keys x 64] into R2 — consumed as an MN-major B operand, so no transpose is needed;
//   4. 48 tcgen05.mma (M=128, N=64, K=16): O main -> TMEM [0,64), O correction -> [64,128);
//   5. epilogue: O / rowsum -> fp32 and/or the fp16 (h,l) pair the output projection consumes; rows >= qlen are zeros.
// Replaces attention.py:230-251,107-118 like the other attention kernels.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <mutex>
#include <string.h>

#include "common.cuh"

namespace masr {
namespace at5 {

constexpr int AT_D = 64;                       // d_k
constexpr int AT_KEYS = 256;                   // keys per CTA (one tile)
constexpr int AT_ROWS = 128;                   // query rows per tile
constexpr int AT_THREADS = 384;                // warp 0: TMA + MMA issue; warps 4..11: softmax / epilogue; all: Q staging
constexpr int AT_R1 = 128 * 1024;              // K|P (h,l): 4 x 32 KB   /  P-matrix (h,l): 2 x 4 x 16 KB
constexpr int AT_R2 = 64 * 1024;               // Qcat (h,l): 4 x 16 KB  /  V (h,l): 2 x 32 KB
constexpr size_t kAttnTc5Smem = AT_R1 + AT_R2 + 1024 /*align*/ + 2 * 2 * 128 * 4 /*row max / sum exchange*/ + 256;
constexpr float kLoS = 2048.0f, kLoSInv = 1.0f / 2048.0f;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "AT_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra AT_DONE;\n\t"
        "bra AT_WAIT;\n\t"
        "AT_DONE:\n\t"
        "}" ::"r"(s_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tm_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tm_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 %%rx;\n\t"
        ".reg .pred %%px;\n\t"
        "elect.sync %%rx|%%px, %1;\n\t"
        "@%%px mov.s32 %0, 1;\n\t"
        "}"
        : "+r"(pred)
        : "r"(0xFFFFFFFFu));
    return pred != 0;
}
// 128-byte-swizzled operand tile, rows of 64 halves (128 B), 8-row groups 1024 B apart.  Valid both for a K-major operand
// (rows = M/N index, the 128 B = 64 contraction elements) and for an MN-major one (rows = contraction index, the 128 B = 64
// M/N elements; SBO = distance between 8-row groups): start address >> 4 | LBO = 1 (unused) | SBO = 1024 B | version 1 | SWIZZLE_128B.
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// byte offset of the 16-byte chunk `c16` (0..7) of row `r` inside a [rows x 64 halves] swizzled tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c16) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// 8 fp32 values -> (h, l) halves, one 16-byte piece each
__device__ __forceinline__ void split8(const float (&v)[8], uint32_t (&h)[4], uint32_t (&l)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half2 hh = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn((v[2 * j] - hf.x) * kLoS, (v[2 * j + 1] - hf.y) * kLoS);
        h[j] = h2_bits(hh); l[j] = h2_bits(ll);
    }
}

struct AttnTc5Params {
    const float* Q; int64_t ldq, q_bstride;
    const float* pos_u; const float* pos_v;
    float* O; __half* Oh; __half* Ol; int64_t ldo, o_bstride;
    const int* q_lens; const int* k_lens;
    int64_t k_bstride;
    float scale;
    int max_q;
};
struct AttnTc5Maps { CUtensorMap kh, kl, vh, vl, ph, pl; };

__global__ void __launch_bounds__(AT_THREADS, 1) relpos_attention_tc5_kernel(const __grid_constant__ AttnTc5Maps maps, AttnTc5Params p) {
    extern __shared__ uint8_t at_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* R1 = smem;                         // K|P tiles: [kb][hl] x 32 KB   /  P-matrix tiles: [hl][kb4] x 16 KB
    uint8_t* R2 = smem + AT_R1;                 // Qcat tiles: [kb][hl] x 16 KB  /  V tiles: [hl] x 32 KB
    float* xch = reinterpret_cast<float*>(R2 + AT_R2);          // [2 column halves][128 rows] max, then [2][128] sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 2 * 2 * 128);   // kv, s, v, o
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int h = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mb_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.kh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.ph) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.vh) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                                 // programmatic dependent launch: the qkv GEMM has completed
    pdl_launch_dependents();

    const int qlen = min(p.q_lens[b], p.max_q), klen = min(p.k_lens[b], AT_KEYS);
    const int64_t krow0 = (int64_t)b * p.k_bstride;
    const int ntile = (qlen > 0 && klen > 0) ? (qlen + AT_ROWS - 1) / AT_ROWS : 0;

    for (int mt = 0; mt < ntile; ++mt) {
        const uint32_t par = mt & 1;
        const int r0 = mt * AT_ROWS;
        // ---- 1. K | P via TMA into R1; [q+u | q+v] built by all threads into R2 ----
        if (warp == 0 && elect_one()) {
            mb_expect_tx(&bars[0], 4 * 32768);
            tma_2d(&maps.kh, &bars[0], R1 + 0 * 32768, h * AT_D, (int)krow0);       // kb 0 (keys), h
            tma_2d(&maps.kl, &bars[0], R1 + 1 * 32768, h * AT_D, (int)krow0);       // kb 0, l
            tma_2d(&maps.ph, &bars[0], R1 + 2 * 32768, h * AT_D, 0);                // kb 1 (positions), h
            tma_2d(&maps.pl, &bars[0], R1 + 3 * 32768, h * AT_D, 0);                // kb 1, l
        }
        {
            // every thread owns one 16-byte column chunk (c16 = tid & 7: 384 % 8 == 0) of rows tid/8, +48, +96 (the last only for
            // tid < 256): the positional biases are loaded once, the three rows' query loads are issued together (the r02 capture
            // showed the dependent load -> add chain of a one-row-at-a-time loop as the kernel's largest stall)
            const float* qsrc = p.Q + ((int64_t)b * p.q_bstride + r0) * p.ldq + h * AT_D;
            const uint32_t base = s_u32(R2);
            const int c16 = tid & 7;
            const float4 u0 = ldg_f4(p.pos_u + h * AT_D + c16 * 8), u1 = ldg_f4(p.pos_u + h * AT_D + c16 * 8 + 4);
            const float4 v0 = ldg_f4(p.pos_v + h * AT_D + c16 * 8), v1 = ldg_f4(p.pos_v + h * AT_D + c16 * 8 + 4);
            float4 qa[3], qb[3];
#pragma unroll
            for (int it = 0; it < 3; ++it) {
                const int r = (tid >> 3) + it * (AT_THREADS / 8);
                qa[it] = make_float4(0.f, 0.f, 0.f, 0.f); qb[it] = qa[it];
                if (r < AT_ROWS && r0 + r < qlen) {
                    qa[it] = ldg_f4(qsrc + (int64_t)r * p.ldq + c16 * 8);
                    qb[it] = ldg_f4(qsrc + (int64_t)r * p.ldq + c16 * 8 + 4);
                }
            }
#pragma unroll
            for (int it = 0; it < 3; ++it) {
                const int r = (tid >> 3) + it * (AT_THREADS / 8);
                if (r >= AT_ROWS) break;
                const bool ok = r0 + r < qlen;
                const float4 a0 = qa[it], a1 = qb[it];
                float qu[8], qv[8];
                qu[0] = a0.x + u0.x; qu[1] = a0.y + u0.y; qu[2] = a0.z + u0.z; qu[3] = a0.w + u0.w;
                qu[4] = a1.x + u1.x; qu[5] = a1.y + u1.y; qu[6] = a1.z + u1.z; qu[7] = a1.w + u1.w;
                qv[0] = a0.x + v0.x; qv[1] = a0.y + v0.y; qv[2] = a0.z + v0.z; qv[3] = a0.w + v0.w;
                qv[4] = a1.x + v1.x; qv[5] = a1.y + v1.y; qv[6] = a1.z + v1.z; qv[7] = a1.w + v1.w;
                if (!ok) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { qu[j] = 0.f; qv[j] = 0.f; }
                }
                uint32_t hh[4], ll[4];
                const uint32_t off = sw128_off(r, c16);
                split8(qu, hh, ll);
                sts128(base + 0 * 16384 + off, hh[0], hh[1], hh[2], hh[3]);          // kb 0 (q+u), h
                sts128(base + 1 * 16384 + off, ll[0], ll[1], ll[2], ll[3]);          // kb 0, l
                split8(qv, hh, ll);
                sts128(base + 2 * 16384 + off, hh[0], hh[1], hh[2], hh[3]);          // kb 1 (q+v), h
                sts128(base + 3 * 16384 + off, ll[0], ll[1], ll[2], ll[3]);          // kb 1, l
            }
        }
        fence_async_smem();                     // the generic-proxy stores above must be visible to the tensor core
        __syncthreads();
        // ---- 2. S = Qcat . Kcat^T : main -> TMEM [0,256), correction -> [256,512) ----
        if (warp == 0 && elect_one()) {
            mb_wait(&bars[0], par);
            fence_after();
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(AT_KEYS >> 3) << 17) | ((uint32_t)(AT_ROWS >> 4) << 24);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
                const uint64_t dAh = desc_sw128(s_u32(R2 + (2 * kb + 0) * 16384)), dAl = desc_sw128(s_u32(R2 + (2 * kb + 1) * 16384));
                const uint64_t dBh = desc_sw128(s_u32(R1 + (2 * kb + 0) * 32768)), dBl = desc_sw128(s_u32(R1 + (2 * kb + 1) * 32768));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);                          // 16 halves = 32 B = 2 x 16-byte units
                    mma_f16(tmem + 0, dAh + adv, dBh + adv, idesc, (kb | ks) ? 1u : 0u);
                    mma_f16(tmem + 256, dAh + adv, dBl + adv, idesc, (kb | ks) ? 1u : 0u);
                    mma_f16(tmem + 256, dAl + adv, dBh + adv, idesc, 1u);
                }
            }
            mma_commit(&bars[1]);
            // ---- V via TMA into R2 as soon as the S products have consumed Qcat ----
            mb_wait(&bars[1], par);
            mb_expect_tx(&bars[2], 2 * 32768);
            tma_2d(&maps.vh, &bars[2], R2 + 0 * 32768, h * AT_D, (int)krow0);
            tma_2d(&maps.vl, &bars[2], R2 + 1 * 32768, h * AT_D, (int)krow0);
        }
        // ---- 3. softmax: 8 warps = TMEM lane quarter (warp & 3) x column half ((warp - 4) >> 2) ----
        float inv_sum = 0.f;
        if (warp >= 4) {
            const int q = warp & 3, ch = (warp - 4) >> 2;
            const int row = q * 32 + lane;
            const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 128);
            mb_wait(&bars[1], par);
            fence_after();
            float mx = -INFINITY;
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t a[32], c[32];
                tm_ld32(tl + cc * 32, a);
                tm_ld32(tl + 256 + cc * 32, c);
                tm_ld_wait();
                const int col0 = ch * 128 + cc * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float s = fmaf(__uint_as_float(c[j]), kLoSInv, __uint_as_float(a[j])) * p.scale;
                    mx = fmaxf(mx, col0 + j < klen ? s : -INFINITY);
                }
            }
            xch[ch * 128 + row] = mx;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mx = fmaxf(xch[row], xch[128 + row]);                                    // (klen >= 1: finite)
            float sum = 0.f;
            const uint32_t pbase = s_u32(R1);
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                uint32_t a[32], c[32];
                tm_ld32(tl + cc * 32, a);
                tm_ld32(tl + 256 + cc * 32, c);
                tm_ld_wait();
                const int col0 = ch * 128 + cc * 32;
                const int kb = col0 >> 6, c16b = (col0 & 63) >> 3;                   // 64-key K-block and first 16-byte chunk in it
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    float pv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int jj = g8 * 8 + j;
                        const float s = fmaf(__uint_as_float(c[jj]), kLoSInv, __uint_as_float(a[jj])) * p.scale;
                        pv[j] = col0 + jj < klen ? ex2_approx(s - mx) : 0.f;
                        sum += pv[j];
                    }
                    uint32_t hh[4], ll[4];
                    split8(pv, hh, ll);
                    const uint32_t off = sw128_off(row, c16b + g8);
                    sts128(pbase + (0 * 4 + kb) * 16384 + off, hh[0], hh[1], hh[2], hh[3]);
                    sts128(pbase + (1 * 4 + kb) * 16384 + off, ll[0], ll[1], ll[2], ll[3]);
                }
            }
            xch[256 + ch * 128 + row] = sum;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            inv_sum = 1.0f / (xch[256 + row] + xch[256 + 128 + row]);
            fence_before();
        }
        fence_async_smem();
        __syncthreads();                        // P-matrix complete in R1; S has been read out of TMEM
        // ---- 4. O = P . V : main -> TMEM [0,64), correction -> [64,128) ----
        if (warp == 0 && elect_one()) {
            mb_wait(&bars[2], par);
            fence_after();
            constexpr uint32_t idesc = (1u << 4) | (1u << 16) /* B is MN-major */ | ((uint32_t)(AT_D >> 3) << 17) | ((uint32_t)(AT_ROWS >> 4) << 24);
            const uint64_t dVh = desc_sw128(s_u32(R2)), dVl = desc_sw128(s_u32(R2 + 32768));
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
                const uint64_t dPh = desc_sw128(s_u32(R1 + (0 * 4 + kb) * 16384)), dPl = desc_sw128(s_u32(R1 + (1 * 4 + kb) * 16384));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t adva = (uint64_t)(ks * 2);                         // A (K-major): 16 halves = 32 B
                    const uint64_t advb = (uint64_t)((kb * 64 + ks * 16) * 128 >> 4); // B (MN-major): 16 key rows of 128 B
                    mma_f16(tmem + 0, dPh + adva, dVh + advb, idesc, (kb | ks) ? 1u : 0u);
                    mma_f16(tmem + 64, dPh + adva, dVl + advb, idesc, (kb | ks) ? 1u : 0u);
                    mma_f16(tmem + 64, dPl + adva, dVh + advb, idesc, 1u);
                }
            }
            mma_commit(&bars[3]);
        }
        // ---- 5. epilogue: O / rowsum -> global (rows >= qlen: zeros) ----
        if (warp >= 4) {
            const int q = warp & 3, cg = (warp - 4) >> 2;                            // 32 output columns per warp
            const int row = q * 32 + lane;
            mb_wait(&bars[3], par);
            fence_after();
            uint32_t a[32], c[32];
            const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 32);
            tm_ld32(tl, a);
            tm_ld32(tl + 64, c);
            tm_ld_wait();
            fence_before();
            const int grow = r0 + row;
            if (grow < p.max_q) {
                const bool valid = grow < qlen;
                const int64_t off = ((int64_t)b * p.o_bstride + grow) * p.ldo + h * AT_D + cg * 32;
                float o[32];
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    o[j] = valid ? fmaf(__uint_as_float(c[j]), kLoSInv, __uint_as_float(a[j])) * inv_sum : 0.f;
                if (p.O) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(p.O + off + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                }
                if (p.Oh) {
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        float t8[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) t8[j] = o[g8 * 8 + j];
                        uint32_t hh[4], ll[4];
                        split8(t8, hh, ll);
                        *reinterpret_cast<uint4*>(p.Oh + off + g8 * 8) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                        *reinterpret_cast<uint4*>(p.Ol + off + g8 * 8) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                    }
                }
            }
        }
        fence_before();
        __syncthreads();                        // TMEM, R1 and R2 are free for the next tile
        fence_after();
    }
    // query rows no tile covered (padded rows of a short utterance, empty utterances): deterministic zeros
    {
        const int first = ntile * AT_ROWS;
        for (int idx = tid; idx < (p.max_q - first) * 16; idx += AT_THREADS) {
            const int r = first + (idx >> 4), c = (idx & 15) * 4;
            if (r >= p.max_q) break;
            const int64_t off = ((int64_t)b * p.o_bstride + r) * p.ldo + h * AT_D + c;
            if (p.O) *reinterpret_cast<float4*>(p.O + off) = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.Oh) {
                *reinterpret_cast<uint2*>(p.Oh + off) = make_uint2(0u, 0u);
                *reinterpret_cast<uint2*>(p.Ol + off) = make_uint2(0u, 0u);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        fence_after();
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

typedef CUresult (*EncodeTiledFnA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFnA encode_fn() {
    static EncodeTiledFnA fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFnA>(ptr);
    });
    return fn;
}
// [rows, cols] fp16 row-major (ld halves), box = 64 columns x 256 rows, 128-byte swizzle, zero fill out of bounds
int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld) {
    EncodeTiledFnA fn = encode_fn();
    if (!fn) { set_last_error("cuTensorMapEncodeTiled entry point unavailable"); return MASR_ERR_INTERNAL; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)AT_D, (cuuint32_t)AT_KEYS};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled(attention) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows, (long long)cols, (long long)ld); return MASR_ERR_INTERNAL; }
    return MASR_OK;
}

}  // namespace at5
}  // namespace masr

using namespace masr;
using namespace masr::at5;

// Same arguments and results as masr_relpos_attention_tc, computed on tcgen05 / TMEM / TMA.  Restrictions of this kernel:
// max_q <= 256, every k_lens[b] <= 256, d_k = 64, table_rows >= 1 rows in the linear_pos(pe) table (rows beyond it read as zero and
// only meet masked keys).  Callers fall back to masr_relpos_attention_tc otherwise.
extern "C" int masr_relpos_attention_tc5(const float* Q, int64_t ldq, int64_t q_bstride, const void* Kh, const void* Kl,
                                         const void* Vh, const void* Vl, int64_t ldk, int64_t k_bstride, const void* Ph,
                                         const void* Pl, int64_t ldp, int64_t table_rows, const float* pos_u, const float* pos_v,
                                         float* O, void* Oh, void* Ol, int64_t ldo, int64_t o_bstride, const int* q_lens,
                                         const int* k_lens, int B, int H, int d_k, int max_q, void* stream) {
    if (B == 0 || max_q == 0) return MASR_OK;
    MASR_REQUIRE(Q && Kh && Kl && Vh && Vl && Ph && Pl && pos_u && pos_v && (O || (Oh && Ol)) && q_lens && k_lens,
                 "masr_relpos_attention_tc5: null pointer");
    MASR_REQUIRE(d_k == AT_D, "masr_relpos_attention_tc5: d_k=%d unsupported (this build: 64)", d_k);
    MASR_REQUIRE(max_q <= 2 * AT_ROWS, "masr_relpos_attention_tc5: max_q=%d > 256 (use masr_relpos_attention_tc)", max_q);
    MASR_REQUIRE(ldq % 4 == 0 && ldk % 8 == 0 && ldp % 8 == 0 && ldo % 8 == 0, "masr_relpos_attention_tc5: leading dimensions misaligned");
    MASR_REQUIRE(((reinterpret_cast<uintptr_t>(Kh) | reinterpret_cast<uintptr_t>(Kl) | reinterpret_cast<uintptr_t>(Vh) |
                   reinterpret_cast<uintptr_t>(Vl) | reinterpret_cast<uintptr_t>(Ph) | reinterpret_cast<uintptr_t>(Pl) |
                   reinterpret_cast<uintptr_t>(Oh) | reinterpret_cast<uintptr_t>(Ol)) & 15) == 0,
                 "masr_relpos_attention_tc5: pair pointers must be 16-byte aligned");
    AttnTc5Maps maps;
    memset(&maps, 0, sizeof(maps));
    const int64_t kv_rows = (int64_t)(B - 1) * k_bstride + max_q;        // rows of the K / V matrices that exist
    int rc;
    if ((rc = make_map(&maps.kh, Kh, kv_rows, (int64_t)H * AT_D, ldk))) return rc;
    if ((rc = make_map(&maps.kl, Kl, kv_rows, (int64_t)H * AT_D, ldk))) return rc;
    if ((rc = make_map(&maps.vh, Vh, kv_rows, (int64_t)H * AT_D, ldk))) return rc;
    if ((rc = make_map(&maps.vl, Vl, kv_rows, (int64_t)H * AT_D, ldk))) return rc;
    if ((rc = make_map(&maps.ph, Ph, table_rows, (int64_t)H * AT_D, ldp))) return rc;
    if ((rc = make_map(&maps.pl, Pl, table_rows, (int64_t)H * AT_D, ldp))) return rc;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(relpos_attention_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnTc5Smem);
        if (e != cudaSuccess) { set_last_error("attention_tc5 smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set[dev] = true;
    }
    AttnTc5Params p{Q, ldq, q_bstride, pos_u, pos_v, O, (__half*)Oh, (__half*)Ol, ldo, o_bstride, q_lens, k_lens, k_bstride,
                    1.4426950408889634f / sqrtf((float)d_k), max_q};
    launch_pdl(relpos_attention_tc5_kernel, dim3(H, B), dim3(AT_THREADS), kAttnTc5Smem, (cudaStream_t)stream, maps, p);
    return check_launch("relpos_attention_tc5_kernel");
}
