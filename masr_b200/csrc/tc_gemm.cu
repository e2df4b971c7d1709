// Tensor-core GEMM with fp32-grade results:  C[M,N] = epilogue(A[M,K] * W[N,K]^T)  on tcgen05 / TMEM,
// operands staged by TMA, accumulators in tensor memory.
//
// Precision scheme ("FP16x2 split", DESIGN.md §4 precision policy): every fp32 operand x is carried as
// two fp16 numbers  h = fp16(x),  l = fp16((x - h) * 2^11)  (22 significand bits), and the product is
//     A.W^T  ~=  Ah.Wh^T  +  2^-11 * (Ah.Wl^T + Al.Wh^T)            (the Al.Wl term is < 2^-22 relative)
// with both sums accumulated in fp32 in two TMEM accumulators.  3 MMAs per K-step => one third of the
// fp16/bf16 tensor peak, but the greedy ids stay bit-exact against the fp32 reference (a single-pass
// bf16/tf32/fp16 GEMM flips argmaxes).
//
// Replaces the same reference call sites as gemm.cu (positionwise.py:37, attention.py:72-74,119,
// convolution.py:117-118,127, subsampling.py:110, loss/ctc.py:70).
//
// Structure (persistent: one CTA per SM walks 128x128 output tiles; 64 + 32*EW threads, EW = 16 epilogue warps by default):
//   warp 0   TMA producer: 4 boxes per K-block (Ah, Al, Wh, Wl; 64 halves = one 128-byte swizzle row)
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer (12 MMAs per K-block); main accumulator ping-pong by
//            256-wide K chunk, correction accumulator ping-pong by tile (512 TMEM columns), so the MMAs of tile i+1
//            run while the epilogue drains tile i
//   warps 2.. epilogue: tcgen05.ld 32x32b -> registers -> fused bias/SiLU/ReLU/GLU/scale/residual -> transposed
//            through shared memory -> row-contiguous 128-bit stores (fp32 and/or the fp16 (h,l) pair the next GEMM consumes)
//   smem ring of 3 x 64 KB stages with full/empty mbarriers; tcgen05.commit releases slots and signals the epilogue.
//   Launched with programmatic dependent launch: the prologue overlaps the producer kernel's tail.
// PAIR form (the default): clusters of 2 CTAs own 256 x 128 tiles through tcgen05.mma.cta_group::2 — each CTA
// stages its 128 A rows and half of the W tile, the pair's leader issues the MMAs and multicasts the commits (4 x 48 KB stages).
// Round-2 additions: residual epilogues fetch the residual at tile start into the running sum (flags bit 4);
// EPI_CTC_PARTIAL keeps per (row, 32 columns) softmax partials instead of logits (+ ctc_partial_combine_kernel);
// the LNC variant (clusters of 2 CTAs) fuses the LayerNorm(s) that follow a residual projection, row statistics over DSMEM
// (pre-norm LN / LN2 and the Squeezeformer's post-norm + adaptive scale) — measured, not the default.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace masr {

constexpr int TBM = 128, TBN = 128, TBK = 64;
constexpr int TSTAGES = 3;
constexpr int TILE_BYTES = TBM * TBK * 2;              // 16 KB: one operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;            // Ah, Al, Wh, Wl
// PAIR kernels (cta_group::2, a 256 x 128 tile per pair of CTAs): a CTA stages its own 128 A rows and HALF of the W tile
// (64 of the 128 output columns; the tensor cores of both SMs read both halves) -> 48 KB per K-block, 4 stages
template <bool PAIR> struct RingCfg {
    static constexpr int STAGES = PAIR ? 4 : TSTAGES;
    static constexpr int W_BYTES = PAIR ? TILE_BYTES / 2 : TILE_BYTES;
    static constexpr int BYTES = 2 * TILE_BYTES + 2 * W_BYTES;
};
constexpr int MAX_STAGES = 4;
static_assert(RingCfg<true>::STAGES * RingCfg<true>::BYTES == TSTAGES * STAGE_BYTES, "both rings take 192 KB");
constexpr int EPI_STAGE_TOTAL = 32768;                 // epilogue store staging, split evenly over the epilogue warps
constexpr int tc_threads(int ew) { return 64 + 32 * ew; }   // TMA warp + MMA warp + EW epilogue warps
constexpr uint32_t TMEM_COLS = 512;                    // 2 x main (ping-pong) + correction accumulator, 128 fp32 columns each (384 -> 512)
constexpr float kLoScale = 2048.0f, kLoInv = 1.0f / 2048.0f;

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem_dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// One lane of a converged warp (CUTLASS's elect_one_sync): the compiler then knows the region runs on a single thread and
// issues the uniform-datapath instructions (UTMALDG, UTCHMMA, UTCBAR) directly.  With `if (lane == 0)` every one of them
// was wrapped in an ELECT / vote / branch retry sequence (~9 SASS instructions per MMA): the issuing thread needed about as
// long to issue a K-block's 12 MMAs as the tensor core to execute them, and the pipe idled half of the time (ncu r01).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0, laneid = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 %%rx;\n\t"
        ".reg .pred %%px;\n\t"
        "elect.sync %%rx|%%px, %2;\n\t"
        "@%%px mov.s32 %1, 1;\n\t"
        "mov.s32 %0, %%rx;\n\t"
        "}"
        : "+r"(laneid), "+r"(pred)
        : "r"(0xFFFFFFFFu));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- cta_group::2 forms (PAIR kernels): issued by the leader CTA (cluster rank 0) for both SMs of the pair ----
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier at this CTA-relative address in BOTH CTAs of the pair once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// TMA into this CTA's shared memory, transaction bytes reported to an mbarrier of the pair's leader (shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cluster, void* smem_dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, uint32_t bar_cluster, void* smem_dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile (rows of 64 halves, 8-row groups 1024 B apart):
// start address >> 4 | LBO (ignored for swizzled K-major) | SBO = 1024 B | version 1 | SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// fp32 -> (h, l) with l pre-scaled by 2^11
// packed fp32 arithmetic (Blackwell f32x2: two independent IEEE round-to-nearest operations per issue slot; the results are
// bit-identical to the scalar forms).  The epilogues of the K <= 256 GEMMs are bound by issue slots and pipe time, not by data.
#ifdef MASR_TC_SCALAR_EPI      // A/B builds only: the scalar forms
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) { d0 = fmaf(a0, b0, c0); d1 = fmaf(a1, b1, c1); }
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) { d0 = a0 + b0; d1 = a1 + b1; }
__device__ __forceinline__ void sub2(float& d0, float& d1, float a0, float a1, float b0, float b1) { d0 = a0 - b0; d1 = a1 - b1; }
__device__ __forceinline__ void mul2(float& d0, float& d1, float a0, float a1, float b0, float b1) { d0 = a0 * b0; d1 = a1 * b1; }
#else
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
    asm("{\n\t.reg .b64 a, b, c, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%6, %7};\n\t"
        "fma.rn.f32x2 d, a, b, c;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void sub2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tsub.rn.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void mul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 a, b, d;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmul.rn.f32x2 d, a, b;\n\tmov.b64 {%0, %1}, d;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
#endif
__device__ __forceinline__ void split_f16(float x, __half& h, __half& l) {
    h = __float2half_rn(x);
    l = __float2half_rn((x - __half2float(h)) * kLoScale);
}
// two values at a time: cvt.rn.f16x2.f32 packs a pair per instruction
__device__ __forceinline__ void split_f16x2(float x0, float x1, __half2& h, __half2& l) {
    h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    float d0, d1;
    sub2(d0, d1, x0, x1, hf.x, hf.y);
    mul2(d0, d1, d0, d1, kLoScale, kLoScale);
    l = __floats2half2_rn(d0, d1);
}
struct TcParams {
    const float* bias;
    const float* residual;
    float* C;
    __half* Ch;
    __half* Cl;
    int64_t ldr, ldc;
    int M, N, K;
    int epi;
    float alpha;
    // conv mode (implicit GEMM over parity planes of the conv-1 activation)
    int conv_T2;       // output rows per utterance (T2max)
    int flags;         // bit 0: stage epilogue stores through shared memory (row-contiguous global writes)
    // MASR_EPI_RESIDUAL_LN / _LN2 (cluster kernel): LayerNorm(s) of the finished 256-wide row fused behind the residual add
    const float* ln_g;
    const float* ln_b;
    const float* ln_g2;
    const float* ln_b2;
    float* y2;         // optional fp32 copy of the last LayerNorm's output (row pitch ldc)
    float ln_eps;
    // MASR_EPI_CTC_PARTIAL: per (row, 32-column group) softmax partials [group][M] instead of logits
    float* part_m;
    float* part_s;
    int* part_i;
    float inv_alpha;   // flags bit 4 (residual prefetch): 1 / alpha, exact (alpha is a power of two)
    const float* ada_s;   // EPI_RESIDUAL_POSTLN: optional per-channel scale / bias applied to the LayerNorm output for the pair
    const float* ada_b;
    // LayerNorm prologue (masr_gemm_tc_lnpre_f16x2, K = 256): the A operand is LayerNorm(lnp_x; ln_g, ln_b) — every CTA
    // normalises the rows of its own tiles into the pair buffer (lnp_h, lnp_l: the A tensor maps point at it) before loading them
    const float* lnp_x;
    int64_t lnp_ldx, lnp_ld;
    __half* lnp_h;
    __half* lnp_l;
};

// One LayerNorm row of width 256 by one warp -> the fp16 (h, l) operand pair.  Lane l holds columns (i*32 + l)*4 .. +3, i = 0, 1:
// the arithmetic, and its order, of layernorm_kernel<256, true> (norm.cu), so the fused path is bit-identical to the separate one.
__device__ __forceinline__ void ln_row256_split(const float4 (&v)[2], const float* __restrict__ gamma, const float* __restrict__ beta,
                                                float eps, int lane, __half* __restrict__ yh, __half* __restrict__ yl) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / 256);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / 256) + eps);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = (i * 32 + lane) * 4;
        float4 g = ldg_f4(gamma + c), b = ldg_f4(beta + c);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        __half hh[4], ll[4];
        const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hh[j] = __float2half_rn(ov[j]);
            ll[j] = __float2half_rn((ov[j] - __half2float(hh[j])) * 2048.0f);
        }
        *reinterpret_cast<uint2*>(yh + c) = *reinterpret_cast<const uint2*>(hh);
        *reinterpret_cast<uint2*>(yl + c) = *reinterpret_cast<const uint2*>(ll);
    }
}

// internal epilogue codes (continuing include/masr_b200.h's MASR_EPI_*)
constexpr int EPI_RESIDUAL_LN = 6, EPI_RESIDUAL_LN2 = 7, EPI_CTC_PARTIAL = 8, EPI_RESIDUAL_POSTLN = 9;

struct TcMaps {
    CUtensorMap a[8];  // GEMM: a[0]=Ah, a[1]=Al.  CONV: a[2*plane + {0:h,1:l}], plane = (kh&1)*2 + (kw&1)
    CUtensorMap w[2];  // Wh, Wl
};

constexpr int CHUNK_KB = 4;            // K-blocks per accumulation chunk (K = 256): see "accumulation" below
constexpr int CONV_TR = 6, CONV_W2 = 19, CONV_ROWS = CONV_TR * CONV_W2;   // 114 of the 128 tile rows are real

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* smem_dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// ---- epilogue stores ------------------------------------------------------------------------------
// A TMEM lane is an output row, so each epilogue thread holds 32 consecutive columns of ONE row: storing
// straight from registers makes every warp store touch 32 different 128-byte lines, 16 bytes each (ncu:
// the store pipe, not the tensor pipe, bounded the K=256 GEMMs).  Each warp therefore transposes its
// 32 x 32 block through a private, XOR-swizzled 4 KB shared-memory buffer (conflict-free both ways) and
// writes it back row-contiguous: every store instruction covers whole lines (4 rows x 128 B for fp32).
struct EpiCtx {
    uint32_t sb;        // shared-space address of this warp's staging buffer (4 KB with 8 epilogue warps, 2 KB with 16)
    int lane;
    int64_t row0;       // global output row of lane 0
    int nvalid;         // rows of this warp's 32 that exist
};

// explicit shared-space accesses: through a generic pointer these compiled to LD.E/ST.E (generic), whose
// latency the 2-warps-per-scheduler epilogue could not hide (ncu: 25 % of its samples waited on them)
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

// regs: CH 16-byte pieces = this thread's row segment (needs CH * 512 bytes of staging).
// g0: address of (row0, first column of the segment).
template <int CH>
__device__ __forceinline__ void staged_store(const EpiCtx& c, const uint4 (&regs)[CH], uint8_t* g0, int64_t pitch_bytes) {
    constexpr int RSH = (CH == 8) ? 0 : (CH == 4) ? 1 : 2;
    constexpr int RPI = 32 / CH;                                  // rows per store instruction
    __syncwarp();                                                 // the previous block has been read back
#pragma unroll
    for (int k = 0; k < CH; ++k) sts128(c.sb + c.lane * (CH * 16) + ((k ^ ((c.lane >> RSH) & (CH - 1))) << 4), regs[k]);
    __syncwarp();
    const int sub = c.lane / CH, k = c.lane % CH;
    uint4 v[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int row = i * RPI + sub;
        v[i] = lds128(c.sb + row * (CH * 16) + ((k ^ ((row >> RSH) & (CH - 1))) << 4));
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int row = i * RPI + sub;
        if (row < c.nvalid) *reinterpret_cast<uint4*>(g0 + row * pitch_bytes + (k << 4)) = v[i];
    }
}

// inverse of staged_store for CH = 8 (4 KB staging): fetch a 32-row x 128-byte block row-contiguous, hand each thread its row
__device__ __forceinline__ void staged_load(const EpiCtx& c, uint4 (&regs)[8], const uint8_t* g0, int64_t pitch_bytes) {
    __syncwarp();
    const int sub = c.lane >> 3, k = c.lane & 7;
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + sub;
        v[i] = make_uint4(0, 0, 0, 0);
        if (row < c.nvalid) v[i] = *reinterpret_cast<const uint4*>(g0 + row * pitch_bytes + (k << 4));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + sub;
        sts128(c.sb + row * 128 + ((k ^ (row & 7)) << 4), v[i]);
    }
    __syncwarp();
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) regs[k2] = lds128(c.sb + c.lane * 128 + ((k2 ^ (c.lane & 7)) << 4));
}

// W fp32 values of this thread's row (W = 32, or 16 after GLU) at output column n -> fp32 rows at C (pitch ld).
// STG = bytes of this warp's staging buffer (4096 / 2048 / 1024): a row segment goes through it in pieces of STG/512 chunks.
template <int W, int STG>
__device__ __forceinline__ void emit_f32(const EpiCtx& c, float* C, int64_t ld, int flags, const float (&o)[W], int n, int n_limit) {
    const bool full = n + W - 1 < n_limit;
    const bool row_ok = c.lane < c.nvalid;
    const int64_t my_row = c.row0 + c.lane;
    if ((flags & 1) && full && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) {
        constexpr int CH = (W / 4 < STG / 512) ? W / 4 : STG / 512;
#pragma unroll
        for (int part = 0; part < (W / 4) / CH; ++part) {
            uint4 regs[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int e = (part * CH + j) * 4;
                regs[j] = make_uint4(__float_as_uint(o[e]), __float_as_uint(o[e + 1]), __float_as_uint(o[e + 2]), __float_as_uint(o[e + 3]));
            }
            staged_store<CH>(c, regs, reinterpret_cast<uint8_t*>(C + c.row0 * ld + n + part * CH * 4), ld * 4);
        }
    } else if (row_ok && full && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) {
        float* cp = C + my_row * ld + n;
#pragma unroll
        for (int j = 0; j < W; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    } else if (row_ok) {
        float* cp = C + my_row * ld + n;
#pragma unroll
        for (int j = 0; j < W; ++j)
            if (n + j < n_limit) cp[j] = o[j];
    }
}

// ... -> the fp16 (h, l) operand pair the next GEMM consumes
template <int W, int STG>
__device__ __forceinline__ void emit_pair(const EpiCtx& c, __half* Ch, __half* Cl, int64_t ld, int flags, const float (&o)[W], int n,
                                          int n_limit) {
    const bool full = n + W - 1 < n_limit;
    const bool row_ok = c.lane < c.nvalid;
    const int64_t my_row = c.row0 + c.lane;
    uint32_t hh[W / 2], ll[W / 2];                       // packed half2 bit patterns (kept in registers)
#pragma unroll
    for (int j = 0; j < W / 2; ++j) {
        __half2 h2, l2;
        split_f16x2(o[2 * j], o[2 * j + 1], h2, l2);
        hh[j] = *reinterpret_cast<uint32_t*>(&h2);
        ll[j] = *reinterpret_cast<uint32_t*>(&l2);
    }
    const bool aligned = (ld & 7) == 0 && ((reinterpret_cast<uintptr_t>(Ch) | reinterpret_cast<uintptr_t>(Cl)) & 15) == 0;
    if ((flags & 1) && full && aligned) {
        constexpr int CH = (W / 8 < STG / 512) ? W / 8 : STG / 512;
#pragma unroll
        for (int part = 0; part < (W / 8) / CH; ++part) {
            uint4 regs[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) { const int e = 4 * (part * CH + j); regs[j] = make_uint4(hh[e], hh[e + 1], hh[e + 2], hh[e + 3]); }
            staged_store<CH>(c, regs, reinterpret_cast<uint8_t*>(Ch + c.row0 * ld + n + part * CH * 8), ld * 2);
#pragma unroll
            for (int j = 0; j < CH; ++j) { const int e = 4 * (part * CH + j); regs[j] = make_uint4(ll[e], ll[e + 1], ll[e + 2], ll[e + 3]); }
            staged_store<CH>(c, regs, reinterpret_cast<uint8_t*>(Cl + c.row0 * ld + n + part * CH * 8), ld * 2);
        }
    } else if (row_ok && full && aligned) {
        uint4* hp = reinterpret_cast<uint4*>(Ch + my_row * ld + n);
        uint4* lp = reinterpret_cast<uint4*>(Cl + my_row * ld + n);
#pragma unroll
        for (int j = 0; j < W / 8; ++j) {
            hp[j] = make_uint4(hh[4 * j], hh[4 * j + 1], hh[4 * j + 2], hh[4 * j + 3]);
            lp[j] = make_uint4(ll[4 * j], ll[4 * j + 1], ll[4 * j + 2], ll[4 * j + 3]);
        }
    } else if (row_ok) {
        unsigned short* hp = reinterpret_cast<unsigned short*>(Ch + my_row * ld + n);
        unsigned short* lp = reinterpret_cast<unsigned short*>(Cl + my_row * ld + n);
#pragma unroll
        for (int j = 0; j < W / 2; ++j) {
            if (n + 2 * j < n_limit) { hp[2 * j] = (unsigned short)(hh[j] & 0xffff); lp[2 * j] = (unsigned short)(ll[j] & 0xffff); }
            if (n + 2 * j + 1 < n_limit) { hp[2 * j + 1] = (unsigned short)(hh[j] >> 16); lp[2 * j + 1] = (unsigned short)(ll[j] >> 16); }
        }
    }
}

template <int W, int STG>
__device__ __forceinline__ void emit(const TcParams& p, const EpiCtx& c, const float (&o)[W], int n, int n_limit) {
    if (p.C) emit_f32<W, STG>(c, p.C, p.ldc, p.flags, o, n, n_limit);
    if (p.Ch) emit_pair<W, STG>(c, p.Ch, p.Cl, p.ldc, p.flags, o, n, n_limit);
}

// ---- LayerNorm fused behind the residual epilogue (cluster of 2 CTAs) --------------------------------------------------------
// A 256-wide output row is spread over 2 CTAs (the two 128-column tiles of one row block = one cluster) x 4 epilogue warps
// (32 columns each), one thread per (row, 32 columns).  Row statistics are exchanged through distributed shared memory:
// every thread stores its partial into BOTH CTAs' `red[buf][src cta][column group][row]`, one lane per warp arrives
// (release.cluster) on both CTAs' mbarrier, everybody waits on its own (acquire.cluster) and sums the 8 partials in a fixed
// order — both CTAs obtain bit-identical statistics.  Two-pass (mean, then centred variance) like norm.cu; rounds alternate
// between two buffers / two barriers, so a fast warp can never overwrite or complete a round that a slow one still reads.
struct LnCtx {
    uint32_t red, red_peer;      // shared::cta / shared::cluster byte addresses of red[2][2][4][128] float2
    uint32_t bar, bar_peer;      // ... of the two mbarriers
    uint32_t rank, cgrp, row, lane;
    uint32_t round;
};

// o = LayerNorm(v) * gamma + beta over the 256-wide row this thread holds 32 columns of (g, b: pointers to those columns).
// ONE exchange per LayerNorm: every thread sends the mean and the centred sum of squares of its own 32 values (two-pass,
// in registers); the 8 partials of a row are combined with the pairwise-update formula of Chan et al. (equal counts):
//   mean = sum(m_i) / 8,   M2 = sum(M2_i) + 32 * sum((m_i - mean)^2),   var = M2 / 256
// — as stable as the two-pass form of norm.cu.  Must be called BEFORE the thread's global stores of the tile: the
// release.cluster arrive orders all earlier writes of the thread, and waiting for outstanding global stores there cost
// several microseconds per exchange in the first version.
__device__ __forceinline__ void ln_apply(LnCtx& L, const float (&v)[32], float (&o)[32], const float* g, const float* b, float eps) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 4) s += (v[j] + v[j + 1]) + (v[j + 2] + v[j + 3]);
    const float m_loc = s * (1.0f / 32.0f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float a0 = v[j] - m_loc, a1 = v[j + 1] - m_loc, a2 = v[j + 2] - m_loc, a3 = v[j + 3] - m_loc;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    const uint32_t buf = L.round & 1;
    const uint32_t off = ((((buf * 2 + L.rank) * 4 + L.cgrp) * 128) + L.row) * 8;
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(L.red + off), "f"(m_loc), "f"(q) : "memory");
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(L.red_peer + off), "f"(m_loc), "f"(q) : "memory");
    __syncwarp();
    if (L.lane == 0) {
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(L.bar + buf * 8) : "memory");
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(L.bar_peer + buf * 8) : "memory");
    }
    const uint32_t parity = (L.round >> 1) & 1;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LN_WAIT:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra LN_DONE;\n\t"
        "bra LN_WAIT;\n\t"
        "LN_DONE:\n\t"
        "}" ::"r"(L.bar + buf * 8), "r"(parity) : "memory");
    float pm[8], pq[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)      // fixed order: both CTAs obtain bit-identical statistics
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(pm[k]), "=f"(pq[k]) : "r"(L.red + (((buf * 8 + k) * 128) + L.row) * 8) : "memory");
    ++L.round;
    const float mean = (((pm[0] + pm[1]) + (pm[2] + pm[3])) + ((pm[4] + pm[5]) + (pm[6] + pm[7]))) * 0.125f;
    float m2 = ((pq[0] + pq[1]) + (pq[2] + pq[3])) + ((pq[4] + pq[5]) + (pq[6] + pq[7]));
    float dev2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float dm = pm[k] - mean; dev2 += dm * dm; }
    m2 = fmaf(32.0f, dev2, m2);
    const float rstd = rsqrtf(m2 * (1.0f / 256.0f) + eps);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 gg = ldg_f4(g + j), bb = ldg_f4(b + j);          // warp-uniform addresses: one broadcast transaction
        o[j] = (v[j] - mean) * rstd * gg.x + bb.x;
        o[j + 1] = (v[j + 1] - mean) * rstd * gg.y + bb.y;
        o[j + 2] = (v[j + 2] - mean) * rstd * gg.z + bb.z;
        o[j + 3] = (v[j + 3] - mean) * rstd * gg.w + bb.w;
    }
}

// One 32-column slice of a finished output row: bias was already added; apply the epilogue and store.
// `n` is the global column of v[0] (warp-uniform).
template <int STG, bool LNC>
__device__ __forceinline__ void store_chunk(const TcParams& p, const EpiCtx& c, float (&v)[32], int n, LnCtx& L) {
    constexpr bool BIG = STG >= 4096;
    if (!LNC && p.epi == EPI_CTC_PARTIAL) {
        // CTC head (loss/ctc.py:70 softmax + ctc_greedy_decoder.py:21 argmax): keep only this (row, 32-column group)'s
        // softmax partials — max logit, its first column, sum of exp(x - max) — the [M, V] logits never reach HBM
        if (n + 31 >= p.N) {                                               // ragged last group (warp-uniform): mask once
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n + j >= p.N) v[j] = -INFINITY;
        }
        float m = v[0];
#pragma unroll
        for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
        int mj = 31;
#pragma unroll
        for (int j = 30; j >= 0; --j)
            if (v[j] == m) mj = j;                                         // descending scan: the FIRST maximum wins
        const int mi = n + mj;
        const float mneg = -m * 1.4426950408889634f;
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) sum += ex2_approx(fmaf(v[j], 1.4426950408889634f, mneg));   // exp(-inf) = 0 for masked columns
        if (c.lane < c.nvalid) {
            const int64_t idx = (int64_t)(n >> 5) * p.M + c.row0 + c.lane;  // [group][row]: a warp writes 32 consecutive entries
            p.part_m[idx] = m; p.part_s[idx] = sum; p.part_i[idx] = mi;
        }
        return;
    }
    if (p.epi == MASR_EPI_BIAS_GLU) {
        // interleaved (value, gate) columns -> 16 outputs at column n/2 of an N/2-wide output
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
            float e[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) e[k] = ex2_approx(v[2 * (j + k) + 1] * -1.4426950408889634f);
#pragma unroll
            for (int k = 0; k < 8; ++k) e[k] = rcp_approx(1.0f + e[k]);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[j + k] = v[2 * (j + k)] * e[k];
        }
        emit<16, STG>(p, c, o, n >> 1, p.N >> 1);
        return;
    }
    switch (LNC ? (int)MASR_EPI_RESIDUAL : p.epi) {
        case MASR_EPI_BIAS_SILU:
            // eight independent SFU chains at a time (a one-register serial chain was 5x slower than the MMA loop)
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                float e[8];
#pragma unroll
                for (int k = 0; k < 8; k += 2) mul2(e[k], e[k + 1], v[j + k], v[j + k + 1], -1.4426950408889634f, -1.4426950408889634f);
#pragma unroll
                for (int k = 0; k < 8; ++k) e[k] = ex2_approx(e[k]);
#pragma unroll
                for (int k = 0; k < 8; k += 2) add2(e[k], e[k + 1], e[k], e[k + 1], 1.0f, 1.0f);
#pragma unroll
                for (int k = 0; k < 8; ++k) e[k] = rcp_approx(e[k]);
#pragma unroll
                for (int k = 0; k < 8; k += 2) mul2(v[j + k], v[j + k + 1], v[j + k], v[j + k + 1], e[k], e[k + 1]);
            }
            break;
        case MASR_EPI_BIAS_RELU:
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            break;
        case MASR_EPI_BIAS_SCALE:
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
            break;
        case MASR_EPI_RESIDUAL: {
            const bool vec_r = (p.ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0;
            if (p.flags & 16) {                               // the residual seeded the running sum (see the kernel): v = (r/alpha + A.W + bias)
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
            } else if (BIG && (p.flags & 2) && n + 31 < p.N && vec_r) {
                uint4 rr[8];
                staged_load(c, rr, reinterpret_cast<const uint8_t*>(p.residual + c.row0 * p.ldr + n), p.ldr * 4);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[4 * j] = __uint_as_float(rr[j].x) + p.alpha * v[4 * j];
                    v[4 * j + 1] = __uint_as_float(rr[j].y) + p.alpha * v[4 * j + 1];
                    v[4 * j + 2] = __uint_as_float(rr[j].z) + p.alpha * v[4 * j + 2];
                    v[4 * j + 3] = __uint_as_float(rr[j].w) + p.alpha * v[4 * j + 3];
                }
            } else if (c.lane < c.nvalid && n + 31 < p.N && vec_r) {
                const float* r = p.residual + (c.row0 + c.lane) * p.ldr + n;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 rv = *reinterpret_cast<const float4*>(r + j);
                    v[j] = rv.x + p.alpha * v[j]; v[j + 1] = rv.y + p.alpha * v[j + 1];
                    v[j + 2] = rv.z + p.alpha * v[j + 2]; v[j + 3] = rv.w + p.alpha * v[j + 3];
                }
            } else if (c.lane < c.nvalid) {
                const float* r = p.residual + (c.row0 + c.lane) * p.ldr + n;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (n + j < p.N) v[j] = r[j] + p.alpha * v[j];
            }
            break;
        }
        default: break;
    }
    if (LNC) {
        // v = x + alpha * sublayer(x): the new residual stream.  EPI_RESIDUAL_LN: C <- v, pair <- LN(v) (the next sub-layer's
        // GEMM operand).  EPI_RESIDUAL_LN2: C <- LN1(v) (norm_final: the block output replaces x), pair <- LN2(LN1(v)).
        // (all statistics exchanges come before the first global store of the tile, see ln_apply)
        float o[32];
        ln_apply(L, v, o, p.ln_g + n, p.ln_b + n, p.ln_eps);
        if (p.epi == EPI_RESIDUAL_POSTLN) {
            // post-norm block (Squeezeformer): the stream becomes LN(v); the pair carries the next sub-module's adaptive
            // scale / bias applied to it (squeezeformer/positionwise.py:57-58), or the LayerNorm output itself
            emit_f32<32, STG>(c, p.C, p.ldc, p.flags, o, n, p.N);
            if (p.ada_s) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 as = ldg_f4(p.ada_s + n + j), ab = ldg_f4(p.ada_b + n + j);
                    o[j] = as.x * o[j] + ab.x; o[j + 1] = as.y * o[j + 1] + ab.y;
                    o[j + 2] = as.z * o[j + 2] + ab.z; o[j + 3] = as.w * o[j + 3] + ab.w;
                }
            }
            emit_pair<32, STG>(c, p.Ch, p.Cl, p.ldc, p.flags, o, n, p.N);
        } else if (p.epi == EPI_RESIDUAL_LN2) {
            ln_apply(L, o, v, p.ln_g2 + n, p.ln_b2 + n, p.ln_eps);
            emit_f32<32, STG>(c, p.C, p.ldc, p.flags, o, n, p.N);
            if (p.y2) emit_f32<32, STG>(c, p.y2, p.ldc, p.flags, v, n, p.N);
            emit_pair<32, STG>(c, p.Ch, p.Cl, p.ldc, p.flags, v, n, p.N);
        } else {
            emit_f32<32, STG>(c, p.C, p.ldc, p.flags, v, n, p.N);
            if (p.y2) emit_f32<32, STG>(c, p.y2, p.ldc, p.flags, o, n, p.N);
            emit_pair<32, STG>(c, p.Ch, p.Cl, p.ldc, p.flags, o, n, p.N);
        }
        return;
    }
    emit<32, STG>(p, c, v, n, p.N);
}

// Accumulation: the tensor core adds into its fp32 TMEM accumulator with truncation, so a long K loop
// drifts (measured: 1.4e-5 abs at K=2048 vs 2e-6 for an fp32 FMA loop).  The main product therefore
// accumulates in TMEM for at most CHUNK_KB K-blocks (K=256); the epilogue warps drain each chunk into
// round-to-nearest fp32 registers while the next chunk runs into the other TMEM buffer (ping-pong).
// The correction product is 2^-11 smaller, so its drift is irrelevant and it stays in TMEM for the tile.
//
// Persistent: grid = min(#tiles, #SMs); every CTA walks tiles blockIdx.x, +gridDim.x, ...  TMEM holds
// main[2] (ping-pong by chunk) and corr[2] (ping-pong by tile) = 512 columns, so the MMA warp runs tile
// i+1 while the 8 epilogue warps finish tile i.
// LNC: the LayerNorm-fused variant (N = 256, launched as clusters of 2 CTAs = the two column tiles of a row block); its
// per-warp store staging shrinks to 1 KB to make room for the 8 KB statistics exchange buffer.
constexpr int LN_RED_BYTES = 2 * 2 * 4 * 128 * 8;     // red[buffer][source CTA][column group][row] (mean, M2)
__host__ __device__ constexpr int epi_stage_bytes(int ew, bool lnc) { return lnc ? 1024 * ew : EPI_STAGE_TOTAL; }

// PAIR: the cta_group::2 form.  A cluster of 2 CTAs owns a 256 x 128 tile: CTA r stages A rows [128 r, 128 r + 128) and W rows
// (output columns) [64 r, 64 r + 64) of every K-block in its own shared memory, all transaction bytes land on the LEADER's
// (rank 0) full barrier, the leader's MMA thread issues M = 256 MMAs that read both CTAs' shared memory and write 128
// accumulator rows into each CTA's tensor memory, and its commits arrive (multicast) on the empty / accumulator-full barriers
// of both CTAs.  Each CTA's epilogue warps drain their own TMEM and release the accumulators on the leader's barriers.
// Per SM and K-block that is 48 KB of TMA writes + 72 KB of operand reads instead of 64 + 96 KB: the 128 x 128 single-CTA
// mainloop is bound by the 128 B/clk shared-memory port and the L2 -> SM fabric (profiles/r02_gemm_prof_summary.md).
template <bool CONV, int EW, bool LNC, bool PAIR = false>
__global__ void __launch_bounds__(tc_threads(EW), 1)
tc_gemm_kernel(const __grid_constant__ TcMaps maps, TcParams p, int num_tiles, int tiles_n, int tiles_t) {
    static_assert(!LNC || (EW == 16 && !CONV), "the LayerNorm epilogue needs one 32-column chunk per epilogue warp");
    static_assert(!(LNC && PAIR), "the LayerNorm-fused kernel pairs CTAs along N, the cta_group::2 kernel along M");
    using R = RingCfg<PAIR>;
    constexpr int EPI_BYTES = epi_stage_bytes(EW, LNC);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* epi_stage = smem + R::STAGES * R::BYTES;                 // EPI_BYTES / EW per warp (see staged_store)
    uint8_t* ln_red = epi_stage + EPI_BYTES;                          // LNC only
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ln_red + (LNC ? LN_RED_BYTES : 0));
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint64_t* main_full = empty_bar + MAX_STAGES;  // [2]
    uint64_t* main_empty = main_full + 2;          // [2]
    uint64_t* corr_full = main_empty + 2;          // [2]
    uint64_t* corr_empty = corr_full + 2;          // [2]
    uint64_t* ln_bar = corr_empty + 2;             // [2] (LNC)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ln_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.K / TBK;
    uint32_t crank = 0;                                               // PAIR: rank in the CTA pair (0 = leader)
    if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int tile_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    // Tile schedule: CTA (pair) u of U walks tiles u, u + U, ... — or, with the LayerNorm prologue, the contiguous range
    // [T u / U, T (u + 1) / U) (column tile fastest), so that it needs the rows of at most two row blocks and can normalise
    // them itself up front.
    const bool lnp = !LNC && !CONV && p.lnp_x != nullptr;
    const int tile_begin = lnp ? (int)((int64_t)num_tiles * tile_first / tile_step) : tile_first;
    const int tile_end = lnp ? (int)((int64_t)num_tiles * (tile_first + 1) / tile_step) : num_tiles;
    const int tile_stride = lnp ? 1 : tile_step;
    // K <= 256 without a prefetched residual: the "direct" epilogue (TMEM -> registers -> stores, no running sum).
    // (Tried and dropped, tools/step_ab.py: two groups of 8 epilogue warps taking alternate tiles, 64 columns per warp, to
    // overlap the FP32 / MUFU / store phases that 16 warps on one tile run in lockstep -> step +1.4 %, not faster.)
    const bool pre_res_k = (LNC || p.epi == MASR_EPI_RESIDUAL) && (p.flags & 16);
    const bool direct = nkb <= CHUNK_KB && !pre_res_k;
    const int acc_release = EW * (PAIR ? 2 : 1);                      // arrivals that hand an accumulator buffer back

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a[0]); tma_prefetch_desc(&maps.a[1]); tma_prefetch_desc(&maps.w[0]); tma_prefetch_desc(&maps.w[1]);
        for (int s = 0; s < R::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) {
            // PAIR: the accumulator-empty barriers that count are the leader's; the epilogue warps of both CTAs arrive there
            mbar_init(&main_full[s], 1); mbar_init(&main_empty[s], acc_release);
            mbar_init(&corr_full[s], 1); mbar_init(&corr_empty[s], acc_release);
            if (LNC) mbar_init(&ln_bar[s], 2 * EW);                  // one arrival per epilogue warp of both CTAs
            else if (lnp) mbar_init(&ln_bar[s], EW);                 // [0]: this CTA's rows are normalised
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {     // collective over the pair: the same warp of both CTAs, the same slot address
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    __syncthreads();
    // PAIR: nobody may signal a barrier of the other CTA (TMA transaction bytes, commits, accumulator releases) before that
    // CTA has initialised it
    if (PAIR) {
        asm volatile("barrier.cluster.arrive.release;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    }
    // the peer CTA must have initialised its barriers before anybody arrives on them remotely: arrive here, wait (long
    // since complete) right before the first exchange / after the producer and MMA loops
    if (LNC) asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // programmatic dependent launch: everything above (barriers, TMEM allocation, descriptor prefetch) overlapped the
    // producer kernel's tail; no global memory has been touched yet
    pdl_wait();
    pdl_launch_dependents();

    // tile -> coordinates.  PAIR: `tile` numbers 256-row (CONV: 12 time rows) pair tiles, this CTA owns half `crank` of it;
    // a half that lies beyond M / T2 loads zeros and stores nothing
    auto decode = [&](int tile, int& n0, int& m0, int& t0, int& b) {
        const int nt = tile % tiles_n;
        const int rest = tile / tiles_n;
        n0 = nt * TBN;
        if (CONV) { t0 = ((rest % tiles_t) * (PAIR ? 2 : 1) + (int)crank) * CONV_TR; b = rest / tiles_t; m0 = 0; }
        else { m0 = (rest * (PAIR ? 2 : 1) + (int)crank) * TBM; t0 = 0; b = 0; }
    };

    if (warp == 0) {
        if (elect_one_sync()) {
            constexpr uint32_t a_bytes = CONV ? CONV_ROWS * TBK * 2 : TILE_BYTES;
            constexpr uint32_t tx_bytes = (2 * a_bytes + 2 * R::W_BYTES) * (PAIR ? 2 : 1);   // PAIR: both CTAs' boxes
            uint32_t kg = 0;
            if (lnp) mbar_wait(&ln_bar[0], 0);      // the epilogue warps have written this CTA's A rows (LayerNorm prologue)
            for (int tile = tile_begin; tile < tile_end; tile += tile_stride) {
                int n0, m0, t0, b;
                decode(tile, n0, m0, t0, b);
                for (int kb = 0; kb < nkb; ++kb, ++kg) {
                    const uint32_t s = kg % R::STAGES;
                    mbar_wait(&empty_bar[s], ((kg / R::STAGES) & 1) ^ 1);
                    uint8_t* st = smem + s * R::BYTES;
                    if (p.flags & 32) {      // profiling switch (tools/gemm_bound_probe.py): no loads, the MMAs run on stale tiles
                        if (!PAIR || crank == 0) mbar_arrive(&full_bar[s]);
                        continue;
                    }
                    if (!PAIR || crank == 0) mbar_expect_tx(&full_bar[s], tx_bytes);
                    if (PAIR) {
                        const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);     // the leader's barrier
                        const int wrow = n0 + (int)crank * (TBN / 2);
                        if (CONV) {
                            const int tap = kb >> 2, cj = kb & 3;
                            const int kh = tap / 3, kw = tap - kh * 3;
                            const int plane = (kh & 1) * 2 + (kw & 1);
                            tma_load_4d_pair(&maps.a[2 * plane], fb, st, cj * TBK, kw >> 1, t0 + (kh >> 1), b);
                            tma_load_4d_pair(&maps.a[2 * plane + 1], fb, st + TILE_BYTES, cj * TBK, kw >> 1, t0 + (kh >> 1), b);
                        } else {
                            tma_load_2d_pair(&maps.a[0], fb, st, kb * TBK, m0);
                            tma_load_2d_pair(&maps.a[1], fb, st + TILE_BYTES, kb * TBK, m0);
                        }
                        tma_load_2d_pair(&maps.w[0], fb, st + 2 * TILE_BYTES, kb * TBK, wrow);      // box: 64 rows
                        tma_load_2d_pair(&maps.w[1], fb, st + 2 * TILE_BYTES + R::W_BYTES, kb * TBK, wrow);
                        continue;
                    }
                    if (CONV) {
                        const int tap = kb >> 2, cj = kb & 3;       // K index = tap*256 + cj*64  (C = 256)
                        const int kh = tap / 3, kw = tap - kh * 3;
                        const int plane = (kh & 1) * 2 + (kw & 1);
                        tma_load_4d(&maps.a[2 * plane], &full_bar[s], st, cj * TBK, kw >> 1, t0 + (kh >> 1), b);
                        tma_load_4d(&maps.a[2 * plane + 1], &full_bar[s], st + TILE_BYTES, cj * TBK, kw >> 1, t0 + (kh >> 1), b);
                    } else {
                        tma_load_2d(&maps.a[0], &full_bar[s], st, kb * TBK, m0);
                        tma_load_2d(&maps.a[1], &full_bar[s], st + TILE_BYTES, kb * TBK, m0);
                    }
                    tma_load_2d(&maps.w[0], &full_bar[s], st + 2 * TILE_BYTES, kb * TBK, n0);
                    tma_load_2d(&maps.w[1], &full_bar[s], st + 3 * TILE_BYTES, kb * TBK, n0);
                }
            }
        }
        if (LNC) asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    } else if (warp == 1) {
        if ((!PAIR || crank == 0) && elect_one_sync()) {
            // instruction descriptor: D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 @17, M>>4 @24 (PAIR: M = 256 over two CTAs)
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)((PAIR ? 2 * TBM : TBM) >> 4) << 24);
            auto mma = [](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
                if (PAIR) umma_f16_pair(d, da, db, idesc, acc); else umma_f16(d, da, db, idesc, acc);
            };
            auto commit = [](uint64_t* bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
            uint32_t kg = 0, cg = 0, tl = 0;
            for (int tile = tile_begin; tile < tile_end; tile += tile_stride, ++tl) {
                mbar_wait(&corr_empty[tl & 1], ((tl >> 1) & 1) ^ 1);    // epilogue has read corr of tile tl-2
                tc_fence_after();
                const uint32_t d_corr = tmem_base + 2 * TBN + (tl & 1) * TBN;
                for (int kb = 0; kb < nkb; ++kb, ++kg) {
                    const uint32_t s = kg % R::STAGES;
                    const bool first = (kb % CHUNK_KB) == 0;
                    const bool last = (kb % CHUNK_KB) == CHUNK_KB - 1 || kb == nkb - 1;
                    if (first) {
                        mbar_wait(&main_empty[cg & 1], ((cg >> 1) & 1) ^ 1);  // epilogue drained this buffer (chunk cg-2)
                        tc_fence_after();
                    }
                    const uint32_t d_main = tmem_base + (cg & 1) * TBN;
                    mbar_wait(&full_bar[s], (kg / R::STAGES) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * R::BYTES);
                    const uint64_t dAh = umma_desc_k_sw128(sa), dAl = umma_desc_k_sw128(sa + TILE_BYTES);
                    const uint64_t dWh = umma_desc_k_sw128(sa + 2 * TILE_BYTES), dWl = umma_desc_k_sw128(sa + 2 * TILE_BYTES + R::W_BYTES);
                    if (!(p.flags & 64))     // profiling switch: loads only, no MMAs
#pragma unroll
                    for (int ks = 0; ks < TBK / 16; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * 2);      // 16 halves = 32 B = 2 x 16-byte units
                        mma(d_main, dAh + adv, dWh + adv, (first && ks == 0) ? 0u : 1u);
                        mma(d_corr, dAh + adv, dWl + adv, (kb | ks) ? 1u : 0u);
                        mma(d_corr, dAl + adv, dWh + adv, 1u);
                    }
                    commit(&empty_bar[s]);                             // slot reusable once these MMAs retire
                    if (last) { commit(&main_full[cg & 1]); ++cg; }
                }
                commit(&corr_full[tl & 1]);                            // whole tile (incl. corrections) complete
            }
        }
        if (LNC) asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    } else {
        // ---- EW epilogue warps: TMEM lane quarter = warp % 4, column group = (warp - 2) / 4 (CW columns each) ----
        constexpr int CW = TBN / (EW / 4);                             // 64 columns per warp (EW = 8) or 32 (EW = 16)
        constexpr int NCH = CW / 32;                                   // 32-column chunks per warp
        constexpr int STG = EPI_BYTES / EW;
        const int q = warp & 3, cgrp = (warp - 2) >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cgrp * CW;
        const int nchunks = (nkb + CHUNK_KB - 1) / CHUNK_KB;
        EpiCtx ctx;
        ctx.sb = smem_u32(epi_stage) + (warp - 2) * STG;
        ctx.lane = lane;
        LnCtx lnx;
        if (LNC) {
            uint32_t rank;
            asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
            lnx.rank = rank; lnx.cgrp = (uint32_t)cgrp; lnx.row = (uint32_t)(q * 32 + lane); lnx.lane = (uint32_t)lane; lnx.round = 0;
            lnx.red = smem_u32(ln_red); lnx.bar = smem_u32(ln_bar);
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lnx.red_peer) : "r"(lnx.red), "r"(rank ^ 1u));
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(lnx.bar_peer) : "r"(lnx.bar), "r"(rank ^ 1u));
            asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
        }
        // accumulator release: PAIR -> the leader's barrier (a shared::cluster address; the leader's own for rank 0)
        auto release = [&](uint64_t* bar) {
            if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0)); else mbar_arrive(bar);
        };
        // row mapping: the warp's 32 tile rows are 32 consecutive output rows in both modes
        auto map_rows = [&](int m0, int t0, int b) {
            if (CONV) {
                // tile row r = ti*19 + f  ->  output row (b*T2 + t0)*19 + r, for r < 114 and t0 + ti < T2
                const int rows = min(CONV_ROWS, (p.conv_T2 - t0) * CONV_W2);
                ctx.row0 = ((int64_t)b * p.conv_T2 + t0) * CONV_W2 + q * 32;
                ctx.nvalid = max(0, min(32, rows - q * 32));
            } else {
                ctx.row0 = (int64_t)m0 + q * 32;
                ctx.nvalid = max(0, min(32, p.M - (m0 + q * 32)));
            }
        };
        // direct epilogue of one 32-column slice (tile columns [col, col + 32) = output columns [n, n + 32)) of accumulator
        // buffer `buf`: v = main + 2^-11 * correction + bias in packed fp32 pairs, the bias from warp-uniform 128-bit loads
        // (one broadcast transaction each, L1 hits after bias_prefetch) instead of 32 shuffles
        const uint32_t lane_tm = tmem_base + ((uint32_t)(q * 32) << 16);
        auto bias_prefetch = [&](int nw, int cols) {
            if (p.bias != nullptr && lane * 32 < cols && nw + lane * 32 < p.N)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bias + nw + lane * 32));
        };
        auto direct_slice = [&](uint32_t col, int n, bool last, uint32_t buf) {
            uint32_t r[32], rc[32];
            tmem_ld32(lane_tm + buf * TBN + col, r);
            tmem_ld32(lane_tm + 2 * TBN + buf * TBN + col, rc);
            tmem_ld_wait();
            if (last) {                                           // TMEM released: the rest runs from registers
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { release(&main_empty[buf]); release(&corr_empty[buf]); }
            }
            if (n >= p.N) return;                                  // warp-uniform
            float v[32];
            if (p.bias != nullptr && n + 31 < p.N && (reinterpret_cast<uintptr_t>(p.bias + n) & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bb = ldg_f4(p.bias + n + j);
                    fma2(v[j], v[j + 1], __uint_as_float(rc[j]), __uint_as_float(rc[j + 1]), kLoInv, kLoInv, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
                    fma2(v[j + 2], v[j + 3], __uint_as_float(rc[j + 2]), __uint_as_float(rc[j + 3]), kLoInv, kLoInv, __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                    add2(v[j], v[j + 1], v[j], v[j + 1], bb.x, bb.y);
                    add2(v[j + 2], v[j + 3], v[j + 2], v[j + 3], bb.z, bb.w);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float bj = (p.bias != nullptr && n + j < p.N) ? __ldg(p.bias + n + j) : 0.f;
                    v[j] = fmaf(__uint_as_float(rc[j]), kLoInv, __uint_as_float(r[j])) + bj;
                }
            }
            store_chunk<STG, LNC>(p, ctx, v, n, lnx);
        };
        if (lnp) {
            // LayerNorm prologue: the 16 epilogue warps normalise this CTA's 128 rows of every row block its tile range touches
            // (8 rows per warp and block, 4 rows in flight) into the pair buffer the A loads read.  A block shared with the
            // neighbouring CTA's range is written twice with identical values.
            if (tile_begin < tile_end) {
                const int b0 = tile_begin / tiles_n, b1 = (tile_end - 1) / tiles_n;
                for (int blk = b0; blk <= b1; ++blk) {
                    const int mrow0 = (blk * (PAIR ? 2 : 1) + (int)crank) * TBM + (warp - 2) * (TBM / EW);
#pragma unroll
                    for (int r4 = 0; r4 < TBM / EW; r4 += 4) {
                        float4 xv[4][2];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int row = mrow0 + r4 + k;
                            if (row < p.M) {
                                const float* xr = p.lnp_x + (int64_t)row * p.lnp_ldx;
                                xv[k][0] = ldg_f4(xr + lane * 4);
                                xv[k][1] = ldg_f4(xr + (32 + lane) * 4);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int row = mrow0 + r4 + k;
                            if (row < p.M)                          // warp-uniform
                                ln_row256_split(xv[k], p.ln_g, p.ln_b, p.ln_eps, lane, p.lnp_h + (int64_t)row * p.lnp_ld,
                                                p.lnp_l + (int64_t)row * p.lnp_ld);
                        }
                    }
                }
            }
            // generic-proxy global writes -> ordered before the TMA loads (async proxy) the producer issues after the barrier
            __threadfence();
            asm volatile("fence.proxy.async.global;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&ln_bar[0]);
        }
        uint32_t cg = 0, tl = 0;
        for (int tile = tile_begin; tile < tile_end; tile += tile_stride, ++tl) {
            int n0, m0, t0, b;
            decode(tile, n0, m0, t0, b);
            const int nw = n0 + cgrp * CW;                             // first column of this warp
            map_rows(m0, t0, b);
            if (direct) {
                // K <= 256: main and correction accumulators complete together (one chunk per tile: cg == tl); go straight
                // from TMEM to the stores 32 columns at a time (no 64-register running sum, so the activations keep their ILP)
                bias_prefetch(nw, CW);
                mbar_wait(&main_full[tl & 1], (tl >> 1) & 1);
                mbar_wait(&corr_full[tl & 1], (tl >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int cc = 0; cc < NCH; ++cc) direct_slice((uint32_t)(cgrp * CW + cc * 32), nw + cc * 32, cc == NCH - 1, tl & 1);
                ++cg;
                continue;
            }
            // bias of the warp's columns: lane l keeps columns l (and 32+l), broadcast by shuffle below;
            // fetched before the accumulator wait so its latency is hidden
            float bias0 = 0.f, bias1 = 0.f;
            if (p.bias != nullptr) {
                if (nw + lane < p.N) bias0 = __ldg(p.bias + nw + lane);
                if (NCH > 1 && nw + 32 + lane < p.N) bias1 = __ldg(p.bias + nw + 32 + lane);
            }
            // Residual epilogues (flags bit 4): this thread's residual values are fetched NOW, before the first accumulator
            // is ready, and seed the running sum as residual / alpha (alpha is a power of two: exact), so the finished row is
            // alpha * (sum + bias).  ncu r02: loaded after the last MMA, the row-strided residual read (8 x LDG.128 per
            // thread, 32 lines per instruction) was more than half of the exposed epilogue of the single-tile-per-CTA GEMMs
            // (w_2: ~11 of 34 us).
            const bool pre_res = (LNC || p.epi == MASR_EPI_RESIDUAL) && (p.flags & 16);
            float acc[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) acc[j] = 0.f;
            if (pre_res && lane < ctx.nvalid) {
#pragma unroll
                for (int cc = 0; cc < NCH; ++cc) {
                    const int n = nw + cc * 32;
                    const float* r = p.residual + (ctx.row0 + lane) * p.ldr + n;
                    if (n + 31 < p.N) {                                   // (host checked ldr % 4 == 0 and the 16-byte alignment)
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 rv = *reinterpret_cast<const float4*>(r + j);
                            acc[cc * 32 + j] = rv.x * p.inv_alpha; acc[cc * 32 + j + 1] = rv.y * p.inv_alpha;
                            acc[cc * 32 + j + 2] = rv.z * p.inv_alpha; acc[cc * 32 + j + 3] = rv.w * p.inv_alpha;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n + j < p.N) acc[cc * 32 + j] = r[j] * p.inv_alpha;
                    }
                }
            }
            for (int c = 0; c < nchunks; ++c, ++cg) {
                mbar_wait(&main_full[cg & 1], (cg >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int cc = 0; cc < NCH; ++cc) {
                    uint32_t r[32];
                    tmem_ld32(lane_base + (cg & 1) * TBN + cc * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += __uint_as_float(r[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release(&main_empty[cg & 1]);
            }
            mbar_wait(&corr_full[tl & 1], (tl >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < NCH; ++cc) {
                uint32_t rc[32];
                tmem_ld32(lane_base + 2 * TBN + (tl & 1) * TBN + cc * 32, rc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[cc * 32 + j] = fmaf(__uint_as_float(rc[j]), kLoInv, acc[cc * 32 + j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release(&corr_empty[tl & 1]);               // TMEM released: the rest runs from registers
#pragma unroll
            for (int cc = 0; cc < NCH; ++cc) {
                const int n = nw + cc * 32;
                if (n >= p.N) break;                                   // warp-uniform
                float v[32];
                const float bsrc = cc ? bias1 : bias0;
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = acc[cc * 32 + j] + __shfl_sync(0xffffffffu, bsrc, j);
                store_chunk<STG, LNC>(p, ctx, v, n, lnx);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (LNC || PAIR) {      // neither CTA may exit while its peer can still write into its shared memory / arrive on its barriers
        asm volatile("barrier.cluster.arrive.release;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    }
    if (warp == 1) {
        tc_fence_after();
        __syncwarp();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

constexpr size_t kTcSmem = TSTAGES * STAGE_BYTES + EPI_STAGE_TOTAL + 1024 /*align*/ + 256 /*barriers + tmem slot*/;
constexpr size_t kTcSmemLn = TSTAGES * STAGE_BYTES + epi_stage_bytes(16, true) + LN_RED_BYTES + 1024 + 256;
static_assert(kTcSmem <= 232448 && kTcSmemLn <= 232448, "tc_gemm shared memory exceeds the 227 KB per-CTA limit of sm_100");

// ---- fp32 -> (h,l) split, elementwise (weights at load time; activations produced by SIMT kernels) ----
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ x, __half* __restrict__ h,
                                                        __half* __restrict__ l, int64_t n) {
    int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 v = ldg_f4(x + i);
        __half hh[4], ll[4];
        split_f16(v.x, hh[0], ll[0]); split_f16(v.y, hh[1], ll[1]); split_f16(v.z, hh[2], ll[2]); split_f16(v.w, hh[3], ll[3]);
        *reinterpret_cast<uint2*>(h + i) = *reinterpret_cast<const uint2*>(hh);
        *reinterpret_cast<uint2*>(l + i) = *reinterpret_cast<const uint2*>(ll);
    } else {
        for (; i < n; ++i) split_f16(x[i], h[i], l[i]);
    }
}

// ---- CTC head: combine the per-(row, 32-column group) softmax partials of the EPI_CTC_PARTIAL epilogue -------------------------
// A CTA handles 32 frames: lane = frame (coalesced [group][row] reads), the 4 warps take interleaved quarters of the groups
// (8 independent loads in flight per thread) with an online max / sum-exp merge per thread (ascending groups, strict >, so a
// thread keeps the FIRST maximum of its groups); the four partials of a frame are then merged: the larger maximum wins, equal
// maxima resolve to the lower column index (numpy's argmax, ctc_greedy_decoder.py:21).  max-prob = 1 / sum_j exp(x_j - max).
__global__ void __launch_bounds__(128) ctc_partial_combine_kernel(const float* __restrict__ pm, const float* __restrict__ ps,
                                                                  const int* __restrict__ pi, int M, int groups,
                                                                  int* __restrict__ ids, float* __restrict__ maxp) {
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row = blockIdx.x * 32 + lane;
    float m = -INFINITY, s = 0.f;
    int mi = 0x7fffffff;
    if (row < M) {
#pragma unroll 8
        for (int g = w; g < groups; g += 4) {
            const int64_t idx = (int64_t)g * M + row;
            const float gm = __ldg(pm + idx), gs = __ldg(ps + idx);
            const int gi = __ldg(pi + idx);
            if (gm > m) { s = s * expf(m - gm) + gs; m = gm; mi = gi; }        // ascending groups within a thread: strict >
            else s += gs * expf(gm - m);
        }
    }
    __shared__ float sm_[4][32], ss_[4][32];
    __shared__ int si_[4][32];
    sm_[w][lane] = m; ss_[w][lane] = s; si_[w][lane] = mi;
    __syncthreads();
    if (w == 0 && row < M) {
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            const float gm = sm_[k][lane], gs = ss_[k][lane];
            const int gi = si_[k][lane];
            if (gm > m || (gm == m && gi < mi)) { s = s * expf(m - gm) + gs; m = gm; mi = gi; }
            else s += gs * expf(gm - m);
        }
        ids[row] = mi;
        maxp[row] = 1.0f / s;
    }
}

// ---- host: tensor maps ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// [rows, K] fp16 row-major (ld elements), box = 64 (K) x box_rows (128; 64 for the W halves of the PAIR kernels),
// 128-byte swizzle, zero OOB fill
static int make_map_2d(CUtensorMap* map, const void* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows = TBM) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_last_error("cuTensorMapEncodeTiled entry point unavailable"); return MASR_ERR_INTERNAL; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled failed (%d) rows=%lld K=%lld ld=%lld", (int)r, (long long)rows, (long long)K, (long long)ld); return MASR_ERR_INTERNAL; }
    return MASR_OK;
}

static bool g_tc_attr_set[64] = {false};

// conv-1 activation parity plane [B, TH, 20, C] fp16, box = 64 (C) x 19 (f) x 6 (t) x 1 (b)
static int make_map_plane(CUtensorMap* map, const void* ptr, int B, int TH, int C) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_last_error("cuTensorMapEncodeTiled entry point unavailable"); return MASR_ERR_INTERNAL; }
    cuuint64_t dims[4] = {(cuuint64_t)C, 20, (cuuint64_t)TH, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)20 * C * 2, (cuuint64_t)TH * 20 * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)TBK, (cuuint32_t)CONV_W2, (cuuint32_t)CONV_TR, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled(plane) failed (%d) B=%d TH=%d", (int)r, B, TH); return MASR_ERR_INTERNAL; }
    return MASR_OK;
}

// Kernel-variant flags, MASR_TC_FLAGS overrides for A/B runs (tools/gemm_bench.py; B200, M = 7936):
//   bit 0  staged (row-contiguous) epilogue stores      ffn_w1 58.5 -> 42.1 us, qkv 35.5 -> 23.5, ctc head 91 -> 63
//   bit 1  stage the residual READ as well              did not pay (w_2 31 -> 35 us); off
//   bit 2  16 epilogue warps x 32 columns instead of 8 x 64   ffn_w1 43.1 -> 35.0 us, step 4.88 -> 4.59 ms
//   bits 5, 6  profiling only (results are garbage): 32 = skip the TMA loads, 64 = skip the MMAs (tools/gemm_bound_probe.py)
static int tc_flags() {
    const char* e = getenv("MASR_TC_FLAGS");
    return e ? atoi(e) : 5;
}

// flags bit 4: prefetch the residual into the running sum — needs alpha to be a power of two (so that r / alpha and the final
// scaling are exact) and 16-byte aligned residual rows.  MASR_TC_PRERES=0 disables it.
static int preres_flag(const float* residual, int64_t ldr, float alpha) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("MASR_TC_PRERES"); enabled = e ? atoi(e) != 0 : 1; }
    int ex = 0;
    const float mant = frexpf(alpha, &ex);
    return (enabled && residual && mant == 0.5f && (ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0) ? 16 : 0;
}

static int num_sms() {
    static int n[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (n[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n[dev] = v;
    }
    return n[dev];
}

static int ensure_tc_attrs() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!g_tc_attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<false, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gemm_kernel<true, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gemm_kernel<false, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gemm_kernel<true, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gemm_kernel<false, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemLn);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gemm_kernel<false, 16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gemm_kernel<true, 16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
        if (e != cudaSuccess) { set_last_error("tc_gemm smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        g_tc_attr_set[dev] = true;
    }
    return MASR_OK;
}

// PAIR (cta_group::2) kernels: the default wherever there is more than one row block.  MASR_TC_PAIR=0 keeps the single-CTA
// kernel (A/B runs: tools/pair_gemm_check.py per GEMM, tools/step_ab.py for the whole step — B200, 32 x 10 s: 3.64 -> 3.54 ms
// with every GEMM paired; per GEMM the mainloop of a pair tile is ~15 % shorter, the pair's start-up costs ~0.4 us).
// Needs the 16-epilogue-warp variant (flags bit 2).
static bool pair_enabled(int flags, int K, int tiles, bool conv) {
    (void)K; (void)tiles; (void)conv;
    if (!(flags & 4)) return false;
    const char* e = getenv("MASR_TC_PAIR");             // read per call: the A/B tools flip it inside one process
    return e ? atoi(e) != 0 : true;
}

// Launch a PAIR kernel over `num_ptiles` pair tiles: clusters of 2 CTAs, as many pairs as fit on the device at one CTA per SM.
template <bool CONV>
static void launch_pair(const TcMaps& maps, const TcParams& p, int num_ptiles, int tiles_n, int tiles_t, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(tc_threads(16));
    cfg.dynamicSmemBytes = kTcSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static int max_pairs[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (max_pairs[dev] == 0) {
        int n = 0;
        cfg.gridDim = dim3(num_sms() & ~1);
        cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, tc_gemm_kernel<CONV, 16, false, true>, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = num_sms() / 2;
        }
        max_pairs[dev] = n < num_sms() / 2 ? n : num_sms() / 2;
    }
    const int pairs = num_ptiles < max_pairs[dev] ? num_ptiles : max_pairs[dev];
    cfg.gridDim = dim3(2 * pairs);
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaLaunchKernelEx(&cfg, tc_gemm_kernel<CONV, 16, false, true>, maps, p, num_ptiles, tiles_n, tiles_t);
}

}  // namespace masr

using namespace masr;

// Conv2d(C,C,3,2)+ReLU of Conv2dSubsampling4 (subsampling.py:83-84) as a tensor-core implicit GEMM.
// Input: the conv-1 activation stored as four (t,f)-parity planes of fp16 (h,l) pairs
// (masr_conv1_cmvn_relu_planes_f16), so the stride-2 window of tap (kh,kw) is a dense TMA box of plane
// ((kh&1),(kw&1)).  Output rows ((b*T2 + t)*19 + f), C columns, as fp32 and/or an fp16 pair.
extern "C" int masr_conv2_tc_f16x2(const void* c1h, const void* c1l, const void* Wh, const void* Wl, const float* bias,
                                   float* out, void* outh, void* outl, int B, int F1, int T2, int C, void* stream) {
    if (B == 0 || T2 == 0) return MASR_OK;
    MASR_REQUIRE(c1h && c1l && Wh && Wl && (out || (outh && outl)), "masr_conv2_tc_f16x2: null pointer");
    MASR_REQUIRE(C == 256, "masr_conv2_tc_f16x2: C=%d unsupported (this build: 256)", C);
    const int TH = (F1 + 1) / 2;
    TcMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc;
    const int64_t plane_elems = (int64_t)B * TH * 20 * C;
    for (int pl = 0; pl < 4; ++pl) {
        if ((rc = make_map_plane(&maps.a[2 * pl], (const __half*)c1h + pl * plane_elems, B, TH, C))) return rc;
        if ((rc = make_map_plane(&maps.a[2 * pl + 1], (const __half*)c1l + pl * plane_elems, B, TH, C))) return rc;
    }
    const bool pair = pair_enabled(tc_flags(), 9 * C, 0, true) && T2 > CONV_TR;
    if ((rc = make_map_2d(&maps.w[0], Wh, C, 9 * C, 9 * C, pair ? TBN / 2 : TBN))) return rc;
    if ((rc = make_map_2d(&maps.w[1], Wl, C, 9 * C, 9 * C, pair ? TBN / 2 : TBN))) return rc;
    if ((rc = ensure_tc_attrs())) return rc;
    TcParams p{bias, nullptr, out, (__half*)outh, (__half*)outl, 0, C, B * T2 * CONV_W2, C, 9 * C, MASR_EPI_BIAS_RELU, 1.f, T2, tc_flags()};
    const int tiles_n = (C + TBN - 1) / TBN, tiles_t = (T2 + CONV_TR - 1) / CONV_TR;
    if (pair) {
        const int ptiles_t = (tiles_t + 1) / 2;
        launch_pair<true>(maps, p, tiles_n * ptiles_t * B, tiles_n, ptiles_t, (cudaStream_t)stream);
        return check_launch("tc_gemm_kernel<conv, pair>");
    }
    const int num_tiles = tiles_n * tiles_t * B;
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    if (p.flags & 4) launch_pdl(tc_gemm_kernel<true, 16, false>, dim3(grid), dim3(tc_threads(16)), kTcSmem, (cudaStream_t)stream, maps, p, num_tiles, tiles_n, tiles_t);
    else launch_pdl(tc_gemm_kernel<true, 8, false>, dim3(grid), dim3(tc_threads(8)), kTcSmem, (cudaStream_t)stream, maps, p, num_tiles, tiles_n, tiles_t);
    return check_launch("tc_gemm_kernel<conv>");
}

extern "C" int masr_split_f16(const float* x, void* h, void* l, int64_t n, void* stream) {
    if (n == 0) return MASR_OK;
    MASR_REQUIRE(x && h && l, "masr_split_f16: null pointer");
    MASR_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(h) & 7) == 0 &&
                 (reinterpret_cast<uintptr_t>(l) & 7) == 0, "masr_split_f16: misaligned pointer");
    int64_t blocks = (n + 1023) / 1024;
    split_f16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, (__half*)h, (__half*)l, n);
    return check_launch("split_f16_kernel");
}

extern "C" int masr_gemm_tc_f16x2(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl,
                                  const float* bias, const float* residual, int64_t ldr, float* C, void* Ch, void* Cl,
                                  int64_t ldc, int M, int N, int K, int epilogue, float alpha, void* stream) {
    if (M == 0 || N == 0) return MASR_OK;
    MASR_REQUIRE(Ah && Al && Wh && Wl, "masr_gemm_tc_f16x2: null operand");
    MASR_REQUIRE(C || (Ch && Cl), "masr_gemm_tc_f16x2: no output");
    MASR_REQUIRE((Ch == nullptr) == (Cl == nullptr), "masr_gemm_tc_f16x2: Ch/Cl must come as a pair");
    MASR_REQUIRE(K > 0 && K % TBK == 0, "masr_gemm_tc_f16x2: K=%d must be a positive multiple of %d", K, TBK);
    MASR_REQUIRE(lda % 8 == 0, "masr_gemm_tc_f16x2: lda=%lld must be a multiple of 8", (long long)lda);
    MASR_REQUIRE(epilogue >= MASR_EPI_BIAS && epilogue <= MASR_EPI_RESIDUAL, "masr_gemm_tc_f16x2: bad epilogue %d", epilogue);
    MASR_REQUIRE(epilogue != MASR_EPI_RESIDUAL || residual, "masr_gemm_tc_f16x2: residual epilogue needs a residual");
    MASR_REQUIRE(epilogue != MASR_EPI_BIAS_GLU || N % 32 == 0, "masr_gemm_tc_f16x2: GLU epilogue needs N %% 32 == 0");
    MASR_REQUIRE(ldc % 8 == 0 || (C && !Ch && ldc % 4 == 0), "masr_gemm_tc_f16x2: ldc=%lld alignment", (long long)ldc);
    TcMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc;
    if ((rc = make_map_2d(&maps.a[0], Ah, M, K, lda))) return rc;
    if ((rc = make_map_2d(&maps.a[1], Al, M, K, lda))) return rc;
    const bool pair = M > TBM && pair_enabled(tc_flags(), K, ((N + TBN - 1) / TBN) * ((M + TBM - 1) / TBM), false);
    if ((rc = make_map_2d(&maps.w[0], Wh, N, K, K, pair ? TBN / 2 : TBN))) return rc;
    if ((rc = make_map_2d(&maps.w[1], Wl, N, K, K, pair ? TBN / 2 : TBN))) return rc;
    if ((rc = ensure_tc_attrs())) return rc;
    TcParams p{bias, residual, C, (__half*)Ch, (__half*)Cl, ldr, ldc, M, N, K, epilogue, alpha, 0, tc_flags()};
    if (epilogue == MASR_EPI_RESIDUAL) { p.flags |= preres_flag(residual, ldr, alpha); p.inv_alpha = 1.0f / alpha; }
    const int tiles_n = (N + TBN - 1) / TBN, tiles_m = (M + TBM - 1) / TBM;
    if (pair) {
        launch_pair<false>(maps, p, tiles_n * ((tiles_m + 1) / 2), tiles_n, 1, (cudaStream_t)stream);
        return check_launch("tc_gemm_kernel<pair>");
    }
    const int num_tiles = tiles_n * tiles_m;
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    if (p.flags & 4) launch_pdl(tc_gemm_kernel<false, 16, false>, dim3(grid), dim3(tc_threads(16)), kTcSmem, (cudaStream_t)stream, maps, p, num_tiles, tiles_n, 1);
    else launch_pdl(tc_gemm_kernel<false, 8, false>, dim3(grid), dim3(tc_threads(8)), kTcSmem, (cudaStream_t)stream, maps, p, num_tiles, tiles_n, 1);
    return check_launch("tc_gemm_kernel");
}

// LayerNorm + Linear in one launch (K = D = 256): C / (Ch, Cl) = epilogue(LN(x; gamma, beta) . W^T + bias).
//   encoder.py:122 -> attention.py:72-74 (norm_mha -> q/k/v), :141 -> convolution.py:117 (norm_conv -> pointwise_conv1 + GLU),
//   :153 / :106 -> positionwise.py:37 (norm_ff / norm_ff_macaron -> w_1 + SiLU)
// Every CTA (pair) takes a contiguous range of tiles, normalises the <= 2 row blocks that range touches into the operand
// pair buffer (Ah, Al: [M, 256] fp16, written here, same values as masr_layernorm_split_f16) and then runs the GEMM on it.
// Replaces masr_layernorm_split_f16 + masr_gemm_tc_f16x2: one launch and its ~5 us of fill / drain less per LayerNorm.
extern "C" int masr_gemm_tc_lnpre_f16x2(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* Ah,
                                        void* Al, int64_t lda, const void* Wh, const void* Wl, const float* bias, float* C,
                                        void* Ch, void* Cl, int64_t ldc, int M, int N, int K, int epilogue, float alpha,
                                        void* stream) {
    if (M == 0 || N == 0) return MASR_OK;
    MASR_REQUIRE(x && gamma && beta && Ah && Al && Wh && Wl, "masr_gemm_tc_lnpre_f16x2: null pointer");
    MASR_REQUIRE(C || (Ch && Cl), "masr_gemm_tc_lnpre_f16x2: no output");
    MASR_REQUIRE((Ch == nullptr) == (Cl == nullptr), "masr_gemm_tc_lnpre_f16x2: Ch/Cl must come as a pair");
    MASR_REQUIRE(K == 256, "masr_gemm_tc_lnpre_f16x2: K=%d unsupported (the LayerNorm width of this build is 256)", K);
    MASR_REQUIRE(lda % 8 == 0 && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
                 "masr_gemm_tc_lnpre_f16x2: lda=%lld ldx=%lld alignment", (long long)lda, (long long)ldx);
    MASR_REQUIRE(epilogue >= MASR_EPI_BIAS && epilogue <= MASR_EPI_BIAS_SCALE, "masr_gemm_tc_lnpre_f16x2: bad epilogue %d", epilogue);
    MASR_REQUIRE(epilogue != MASR_EPI_BIAS_GLU || N % 32 == 0, "masr_gemm_tc_lnpre_f16x2: GLU epilogue needs N %% 32 == 0");
    MASR_REQUIRE(ldc % 8 == 0 || (C && !Ch && ldc % 4 == 0), "masr_gemm_tc_lnpre_f16x2: ldc=%lld alignment", (long long)ldc);
    MASR_REQUIRE(tc_flags() & 4, "masr_gemm_tc_lnpre_f16x2: needs the 16-epilogue-warp kernel (MASR_TC_FLAGS bit 2)");
    TcMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc;
    const int tiles_n = (N + TBN - 1) / TBN, tiles_m = (M + TBM - 1) / TBM;
    const bool pair = M > TBM && pair_enabled(tc_flags(), K, tiles_n * tiles_m, false);
    if ((rc = make_map_2d(&maps.a[0], Ah, M, K, lda))) return rc;
    if ((rc = make_map_2d(&maps.a[1], Al, M, K, lda))) return rc;
    if ((rc = make_map_2d(&maps.w[0], Wh, N, K, K, pair ? TBN / 2 : TBN))) return rc;
    if ((rc = make_map_2d(&maps.w[1], Wl, N, K, K, pair ? TBN / 2 : TBN))) return rc;
    if ((rc = ensure_tc_attrs())) return rc;
    TcParams p{bias, nullptr, C, (__half*)Ch, (__half*)Cl, 0, ldc, M, N, K, epilogue, alpha, 0, tc_flags()};
    p.ln_g = gamma; p.ln_b = beta; p.ln_eps = eps;
    p.lnp_x = x; p.lnp_ldx = ldx; p.lnp_ld = lda; p.lnp_h = (__half*)Ah; p.lnp_l = (__half*)Al;
    if (pair) {
        launch_pair<false>(maps, p, tiles_n * ((tiles_m + 1) / 2), tiles_n, 1, (cudaStream_t)stream);
        return check_launch("tc_gemm_kernel<pair, ln prologue>");
    }
    const int num_tiles = tiles_n * tiles_m;
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    launch_pdl(tc_gemm_kernel<false, 16, false>, dim3(grid), dim3(tc_threads(16)), kTcSmem, (cudaStream_t)stream, maps, p, num_tiles, tiles_n, 1);
    return check_launch("tc_gemm_kernel<ln prologue>");
}

// ctc_lo Linear + softmax statistics + per-frame argmax (loss/ctc.py:70, ctc_greedy_decoder.py:21-27) without the [M, V]
// logits: the GEMM epilogue reduces every 32-column group of a row to (max, first argmax, sum exp) and a small combine
// kernel merges the ceil(V/32) groups of each frame.  workspace: 3 * ceil(V/32) * M * 4 bytes.
extern "C" int masr_ctc_head_argmax_tc_f16x2(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl,
                                             const float* bias, int M, int V, int K, void* workspace, int64_t workspace_bytes,
                                             int* ids, float* maxp, void* stream) {
    if (M == 0) return MASR_OK;
    MASR_REQUIRE(Ah && Al && Wh && Wl && workspace && ids && maxp, "masr_ctc_head_argmax_tc_f16x2: null pointer");
    MASR_REQUIRE(V > 0 && K > 0 && K % TBK == 0 && lda % 8 == 0, "masr_ctc_head_argmax_tc_f16x2: V=%d K=%d lda=%lld", V, K, (long long)lda);
    const int groups = (V + 31) / 32;
    const int64_t need = (int64_t)3 * groups * M * 4;
    MASR_REQUIRE(workspace_bytes >= need, "masr_ctc_head_argmax_tc_f16x2: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
    TcMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc;
    if ((rc = make_map_2d(&maps.a[0], Ah, M, K, lda))) return rc;
    if ((rc = make_map_2d(&maps.a[1], Al, M, K, lda))) return rc;
    const bool pair = M > TBM && pair_enabled(4, K, ((V + TBN - 1) / TBN) * ((M + TBM - 1) / TBM), false);
    if ((rc = make_map_2d(&maps.w[0], Wh, V, K, K, pair ? TBN / 2 : TBN))) return rc;
    if ((rc = make_map_2d(&maps.w[1], Wl, V, K, K, pair ? TBN / 2 : TBN))) return rc;
    if ((rc = ensure_tc_attrs())) return rc;
    TcParams p{bias, nullptr, nullptr, nullptr, nullptr, 0, 0, M, V, K, EPI_CTC_PARTIAL, 1.f, 0, tc_flags()};
    p.part_m = (float*)workspace;
    p.part_s = p.part_m + (int64_t)groups * M;
    p.part_i = (int*)(p.part_s + (int64_t)groups * M);
    const int tiles_n = (V + TBN - 1) / TBN, tiles_m = (M + TBM - 1) / TBM;
    const int num_tiles = tiles_n * tiles_m;
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    if (pair) launch_pair<false>(maps, p, tiles_n * ((tiles_m + 1) / 2), tiles_n, 1, (cudaStream_t)stream);
    else launch_pdl(tc_gemm_kernel<false, 16, false>, dim3(grid), dim3(tc_threads(16)), kTcSmem, (cudaStream_t)stream, maps, p, num_tiles, tiles_n, 1);
    if ((rc = check_launch("tc_gemm_kernel<ctc>"))) return rc;
    launch_pdl(ctc_partial_combine_kernel, dim3((M + 31) / 32), dim3(128), 0, (cudaStream_t)stream, (const float*)p.part_m,
               (const float*)p.part_s, (const int*)p.part_i, M, groups, ids, maxp);
    return check_launch("ctc_partial_combine_kernel");
}

// Sub-layer output projection (N = 256) + residual add + the LayerNorm(s) that follow it, in one kernel:
//   x_new = residual + alpha * (A.W^T + bias)
//   gamma2 == NULL:  X <- x_new,  (Yh, Yl) <- LN(x_new; gamma1, beta1)            encoder.py:117->122, 131->141, 145->153
//   gamma2 != NULL:  X <- LN(x_new; gamma1, beta1),  (Yh, Yl) <- LN(X; gamma2, beta2)   encoder.py:155->161->(next block) 106 / 342
//   Y2 (optional): fp32 copy of what the pair holds.
// Launched as clusters of 2 CTAs (the two 128-column tiles of a row block); row statistics cross the pair through
// distributed shared memory.  Replaces masr_gemm_tc_f16x2(MASR_EPI_RESIDUAL) + masr_layernorm[2]_split_f16.
static int launch_residual_ln(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl, const float* bias,
                              const float* residual, int64_t ldr, float alpha, float* X, const float* gamma1, const float* beta1,
                              const float* gamma2, const float* beta2, float* Y2, void* Yh, void* Yl, int64_t ldx, int M, int N,
                              int K, float eps, int epi_override, const float* ada_s, const float* ada_b, void* stream);

extern "C" int masr_gemm_tc_residual_ln_f16x2(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl,
                                              const float* bias, const float* residual, int64_t ldr, float alpha, float* X,
                                              const float* gamma1, const float* beta1, const float* gamma2, const float* beta2,
                                              float* Y2, void* Yh, void* Yl, int64_t ldx, int M, int N, int K, float eps,
                                              void* stream) {
    return launch_residual_ln(Ah, Al, lda, Wh, Wl, bias, residual, ldr, alpha, X, gamma1, beta1, gamma2, beta2, Y2, Yh, Yl, ldx, M, N,
                              K, eps, 0, nullptr, nullptr, stream);
}

// Post-norm form (Squeezeformer blocks, squeezeformer/encoder.py:412-463): X <- LN(residual + alpha * (A.W^T + bias); gamma, beta)
// becomes the stream, (Yh, Yl) <- ada_scale * X + ada_bias (the next sub-module's adaptive scale; NULL: the pair of X itself).
// Same kernel and restrictions as masr_gemm_tc_residual_ln_f16x2; replaces masr_gemm_tc_f16x2(RESIDUAL) + masr_layernorm_ada_split_f16.
extern "C" int masr_gemm_tc_residual_postln_f16x2(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl,
                                                  const float* bias, const float* residual, int64_t ldr, float alpha, float* X,
                                                  const float* gamma, const float* beta, const float* ada_scale,
                                                  const float* ada_bias, void* Yh, void* Yl, int64_t ldx, int M, int N, int K,
                                                  float eps, void* stream) {
    MASR_REQUIRE((ada_scale == nullptr) == (ada_bias == nullptr), "masr_gemm_tc_residual_postln_f16x2: ada_scale/ada_bias must come as a pair");
    return launch_residual_ln(Ah, Al, lda, Wh, Wl, bias, residual, ldr, alpha, X, gamma, beta, nullptr, nullptr, nullptr, Yh, Yl, ldx, M,
                              N, K, eps, EPI_RESIDUAL_POSTLN, ada_scale, ada_bias, stream);
}

static int launch_residual_ln(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl, const float* bias,
                              const float* residual, int64_t ldr, float alpha, float* X, const float* gamma1, const float* beta1,
                              const float* gamma2, const float* beta2, float* Y2, void* Yh, void* Yl, int64_t ldx, int M, int N,
                              int K, float eps, int epi_override, const float* ada_s, const float* ada_b, void* stream) {
    if (M == 0) return MASR_OK;
    MASR_REQUIRE(Ah && Al && Wh && Wl && residual && X && gamma1 && beta1 && Yh && Yl, "masr_gemm_tc_residual_ln_f16x2: null pointer");
    MASR_REQUIRE((gamma2 == nullptr) == (beta2 == nullptr), "masr_gemm_tc_residual_ln_f16x2: gamma2/beta2 must come as a pair");
    MASR_REQUIRE(N == 2 * TBN, "masr_gemm_tc_residual_ln_f16x2: N=%d unsupported (this build: %d)", N, 2 * TBN);
    MASR_REQUIRE(K > 0 && K % TBK == 0 && lda % 8 == 0 && ldx % 8 == 0 && ldr % 4 == 0, "masr_gemm_tc_residual_ln_f16x2: K=%d lda=%lld ldx=%lld ldr=%lld",
                 K, (long long)lda, (long long)ldx, (long long)ldr);
    TcMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc;
    if ((rc = make_map_2d(&maps.a[0], Ah, M, K, lda))) return rc;
    if ((rc = make_map_2d(&maps.a[1], Al, M, K, lda))) return rc;
    if ((rc = make_map_2d(&maps.w[0], Wh, N, K, K))) return rc;
    if ((rc = make_map_2d(&maps.w[1], Wl, N, K, K))) return rc;
    if ((rc = ensure_tc_attrs())) return rc;
    TcParams p{bias, residual, X, (__half*)Yh, (__half*)Yl, ldr, ldx, M, N, K,
               epi_override ? epi_override : (gamma2 ? EPI_RESIDUAL_LN2 : EPI_RESIDUAL_LN), alpha, 0, tc_flags() & ~2};
    p.ln_g = gamma1; p.ln_b = beta1; p.ln_g2 = gamma2; p.ln_b2 = beta2; p.y2 = Y2; p.ln_eps = eps;
    p.ada_s = ada_s; p.ada_b = ada_b;
    p.flags |= preres_flag(residual, ldr, alpha); p.inv_alpha = 1.0f / alpha;
    const int tiles_m = (M + TBM - 1) / TBM;
    const int num_tiles = 2 * tiles_m;
    const int sms = num_sms() & ~1;
    const int grid = num_tiles < sms ? num_tiles : sms;             // even: a cluster = the two column tiles of one row block
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(tc_threads(16));
    cfg.dynamicSmemBytes = kTcSmemLn;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaLaunchKernelEx(&cfg, tc_gemm_kernel<false, 16, true>, maps, p, num_tiles, 2, 1);
    return check_launch("tc_gemm_kernel<ln>");
}
