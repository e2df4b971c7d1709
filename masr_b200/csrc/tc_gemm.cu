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
// Structure (one 128x128 output tile per CTA, 192 threads):
//   warp 0   TMA producer: 4 boxes per K-block (Ah, Al, Wh, Wl; 64 halves = one 128-byte swizzle row)
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer (12 MMAs per K-block)
//   warps 2-5 epilogue: tcgen05.ld 32x32b -> registers -> fused bias/SiLU/ReLU/GLU/scale/residual ->
//            global (fp32, or the fp16 (h,l) pair the next GEMM consumes)
//   smem ring of STAGES x 64 KB with full/empty mbarriers; tcgen05.commit releases slots and signals
//   the epilogue.
#include <cuda.h>
#include <cuda_fp16.h>
#include <mutex>
#include <string.h>

#include "common.cuh"

namespace masr {

constexpr int TBM = 128, TBN = 128, TBK = 64;
constexpr int TSTAGES = 3;
constexpr int TILE_BYTES = TBM * TBK * 2;              // 16 KB: one operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;            // Ah, Al, Wh, Wl
constexpr int EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * EPI_WARPS;   // TMA warp + MMA warp + epilogue warps
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
__device__ __forceinline__ void split_f16(float x, __half& h, __half& l) {
    h = __float2half_rn(x);
    l = __float2half_rn((x - __half2float(h)) * kLoScale);
}
// two values at a time: cvt.rn.f16x2.f32 packs a pair per instruction
__device__ __forceinline__ void split_f16x2(float x0, float x1, __half2& h, __half2& l) {
    h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    l = __floats2half2_rn((x0 - hf.x) * kLoScale, (x1 - hf.y) * kLoScale);
}
// Epilogue activations on the SFU (ex2.approx / rcp.approx, <= 2 ulp each): the accurate expf + IEEE divide
// cost ~40 instructions per element and made the FFN w_1 epilogue 2.6x longer than its MMA main loop.
// Absolute error < 2e-7 on silu/sigmoid outputs, inside the fp32-grade budget (tests/test_gpu_tc_gemm.py).
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_silu(float x) { return x * fast_sigmoid(x); }

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
};

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

// One 32-column slice of a finished output row: bias was already added; apply the epilogue and store
// (fp32 and/or the fp16 (h,l) pair).  `n` is the global column of v[0]; warp-uniform except row_ok.
__device__ __forceinline__ void store_chunk(const TcParams& p, float (&v)[32], int n, int64_t out_row, bool row_ok) {
    if (n >= p.N || !row_ok) return;
    const bool vec_c = (p.ldc & 3) == 0;
    if (p.epi == MASR_EPI_BIAS_GLU) {
        // interleaved (value, gate) columns -> 16 outputs at column n/2
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = v[2 * j] * fast_sigmoid(v[2 * j + 1]);
        const int nn = n >> 1;
        if (p.C) {
            float* cp = p.C + out_row * p.ldc + nn;
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
        if (p.Ch) {
            __align__(16) __half2 hh[8], ll[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) split_f16x2(o[2 * j], o[2 * j + 1], hh[j], ll[j]);
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
                *reinterpret_cast<uint4*>(p.Ch + out_row * p.ldc + nn + j) = *reinterpret_cast<const uint4*>(&hh[j / 2]);
                *reinterpret_cast<uint4*>(p.Cl + out_row * p.ldc + nn + j) = *reinterpret_cast<const uint4*>(&ll[j / 2]);
            }
        }
        return;
    }
    switch (p.epi) {
        case MASR_EPI_BIAS_SILU:
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fast_silu(v[j]);
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
            const float* r = p.residual + out_row * p.ldr + n;
            const bool vec_r = (p.ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0;
            if (n + 31 < p.N && vec_r) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 rv = *reinterpret_cast<const float4*>(r + j);
                    v[j] = rv.x + p.alpha * v[j]; v[j + 1] = rv.y + p.alpha * v[j + 1];
                    v[j + 2] = rv.z + p.alpha * v[j + 2]; v[j + 3] = rv.w + p.alpha * v[j + 3];
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (n + j < p.N) v[j] = r[j] + p.alpha * v[j];
            }
            break;
        }
        default: break;
    }
    const bool full = n + 31 < p.N;
    if (p.C) {
        float* cp = p.C + out_row * p.ldc + n;
        if (full && vec_c) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n + j < p.N) cp[j] = v[j];
        }
    }
    if (p.Ch) {
        __align__(16) __half2 hh[16], ll[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) split_f16x2(v[2 * j], v[2 * j + 1], hh[j], ll[j]);
        __half* hp = p.Ch + out_row * p.ldc + n;
        __half* lp = p.Cl + out_row * p.ldc + n;
        if (full && (p.ldc & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                *reinterpret_cast<uint4*>(hp + j) = *reinterpret_cast<const uint4*>(&hh[j / 2]);
                *reinterpret_cast<uint4*>(lp + j) = *reinterpret_cast<const uint4*>(&ll[j / 2]);
            }
        } else {
            const __half* hs = reinterpret_cast<const __half*>(hh);
            const __half* ls = reinterpret_cast<const __half*>(ll);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n + j < p.N) { hp[j] = hs[j]; lp[j] = ls[j]; }
        }
    }
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
template <bool CONV>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ TcMaps maps, TcParams p, int num_tiles, int tiles_n, int tiles_t) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + TSTAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + TSTAGES;
    uint64_t* main_full = empty_bar + TSTAGES;     // [2]
    uint64_t* main_empty = main_full + 2;          // [2]
    uint64_t* corr_full = main_empty + 2;          // [2]
    uint64_t* corr_empty = corr_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(corr_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.K / TBK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a[0]); tma_prefetch_desc(&maps.a[1]); tma_prefetch_desc(&maps.w[0]); tma_prefetch_desc(&maps.w[1]);
        for (int s = 0; s < TSTAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&main_full[s], 1); mbar_init(&main_empty[s], EPI_WARPS);
            mbar_init(&corr_full[s], 1); mbar_init(&corr_empty[s], EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile -> coordinates
    auto decode = [&](int tile, int& n0, int& m0, int& t0, int& b) {
        const int nt = tile % tiles_n;
        const int rest = tile / tiles_n;
        n0 = nt * TBN;
        if (CONV) { t0 = (rest % tiles_t) * CONV_TR; b = rest / tiles_t; m0 = 0; }
        else { m0 = rest * TBM; t0 = 0; b = 0; }
    };

    if (warp == 0) {
        if (lane == 0) {
            constexpr uint32_t a_bytes = CONV ? CONV_ROWS * TBK * 2 : TILE_BYTES;
            uint32_t kg = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int n0, m0, t0, b;
                decode(tile, n0, m0, t0, b);
                for (int kb = 0; kb < nkb; ++kb, ++kg) {
                    const uint32_t s = kg % TSTAGES;
                    mbar_wait(&empty_bar[s], ((kg / TSTAGES) & 1) ^ 1);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], 2 * a_bytes + 2 * TILE_BYTES);
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
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 @17, M>>4 @24
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
            uint32_t kg = 0, cg = 0, tl = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
                mbar_wait(&corr_empty[tl & 1], ((tl >> 1) & 1) ^ 1);    // epilogue has read corr of tile tl-2
                tc_fence_after();
                const uint32_t d_corr = tmem_base + 2 * TBN + (tl & 1) * TBN;
                for (int kb = 0; kb < nkb; ++kb, ++kg) {
                    const uint32_t s = kg % TSTAGES;
                    const bool first = (kb % CHUNK_KB) == 0;
                    const bool last = (kb % CHUNK_KB) == CHUNK_KB - 1 || kb == nkb - 1;
                    if (first) {
                        mbar_wait(&main_empty[cg & 1], ((cg >> 1) & 1) ^ 1);  // epilogue drained this buffer (chunk cg-2)
                        tc_fence_after();
                    }
                    const uint32_t d_main = tmem_base + (cg & 1) * TBN;
                    mbar_wait(&full_bar[s], (kg / TSTAGES) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t dAh = umma_desc_k_sw128(sa), dAl = umma_desc_k_sw128(sa + TILE_BYTES);
                    const uint64_t dWh = umma_desc_k_sw128(sa + 2 * TILE_BYTES), dWl = umma_desc_k_sw128(sa + 3 * TILE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < TBK / 16; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * 2);      // 16 halves = 32 B = 2 x 16-byte units
                        umma_f16(d_main, dAh + adv, dWh + adv, idesc, (first && ks == 0) ? 0u : 1u);
                        umma_f16(d_corr, dAh + adv, dWl + adv, idesc, (kb | ks) ? 1u : 0u);
                        umma_f16(d_corr, dAl + adv, dWh + adv, idesc, 1u);
                    }
                    umma_commit(&empty_bar[s]);                        // slot reusable once these MMAs retire
                    if (last) { umma_commit(&main_full[cg & 1]); ++cg; }
                }
                umma_commit(&corr_full[tl & 1]);                       // whole tile (incl. corrections) complete
            }
        }
    } else {
        // ---- 8 epilogue warps: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 ----
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int r_tile = q * 32 + lane;                              // row of the 128-row tile
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)half * (TBN / 2);
        const int nchunks = (nkb + CHUNK_KB - 1) / CHUNK_KB;
        uint32_t cg = 0, tl = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
            int n0, m0, t0, b;
            decode(tile, n0, m0, t0, b);
            float acc[TBN / 2];
#pragma unroll
            for (int j = 0; j < TBN / 2; ++j) acc[j] = 0.f;
            for (int c = 0; c < nchunks; ++c, ++cg) {
                mbar_wait(&main_full[cg & 1], (cg >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int cc = 0; cc < TBN / 64; ++cc) {
                    uint32_t r[32];
                    tmem_ld32(lane_base + (cg & 1) * TBN + cc * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[cc * 32 + j] += __uint_as_float(r[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&main_empty[cg & 1]);
            }
            mbar_wait(&corr_full[tl & 1], (tl >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < TBN / 64; ++cc) {
                uint32_t rc[32];
                tmem_ld32(lane_base + 2 * TBN + (tl & 1) * TBN + cc * 32, rc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[cc * 32 + j] = fmaf(__uint_as_float(rc[j]), kLoInv, acc[cc * 32 + j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&corr_empty[tl & 1]);           // TMEM released: the rest runs from registers
            // row mapping
            int64_t out_row;
            bool row_ok;
            if (CONV) {
                const int ti = r_tile / CONV_W2, f = r_tile - ti * CONV_W2;
                const int t = t0 + ti;
                row_ok = r_tile < CONV_ROWS && t < p.conv_T2;
                out_row = ((int64_t)b * p.conv_T2 + t) * CONV_W2 + f;
            } else {
                out_row = m0 + r_tile;
                row_ok = out_row < p.M;
            }
#pragma unroll
            for (int cc = 0; cc < TBN / 64; ++cc) {
                const int n = n0 + half * (TBN / 2) + cc * 32;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float bb = (p.bias != nullptr && n + j < p.N) ? __ldg(p.bias + n + j) : 0.f;
                    v[j] = acc[cc * 32 + j] + bb;
                }
                store_chunk(p, v, n, out_row, row_ok);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

constexpr size_t kTcSmem = TSTAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers + tmem slot*/;

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

// [rows, K] fp16 row-major (ld elements), box = 64 (K) x 128 (rows), 128-byte swizzle, zero OOB fill
static int make_map_2d(CUtensorMap* map, const void* ptr, int64_t rows, int64_t K, int64_t ld) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_last_error("cuTensorMapEncodeTiled entry point unavailable"); return MASR_ERR_INTERNAL; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)TBM};
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
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
        if (e != cudaSuccess) { set_last_error("tc_gemm smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        g_tc_attr_set[dev] = true;
    }
    return MASR_OK;
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
    if ((rc = make_map_2d(&maps.w[0], Wh, C, 9 * C, 9 * C))) return rc;
    if ((rc = make_map_2d(&maps.w[1], Wl, C, 9 * C, 9 * C))) return rc;
    if ((rc = ensure_tc_attrs())) return rc;
    TcParams p{bias, nullptr, out, (__half*)outh, (__half*)outl, 0, C, B * T2 * CONV_W2, C, 9 * C, MASR_EPI_BIAS_RELU, 1.f, T2};
    const int tiles_n = (C + TBN - 1) / TBN, tiles_t = (T2 + CONV_TR - 1) / CONV_TR;
    const int num_tiles = tiles_n * tiles_t * B;
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    tc_gemm_kernel<true><<<grid, TC_THREADS, kTcSmem, (cudaStream_t)stream>>>(maps, p, num_tiles, tiles_n, tiles_t);
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
    if ((rc = make_map_2d(&maps.w[0], Wh, N, K, K))) return rc;
    if ((rc = make_map_2d(&maps.w[1], Wl, N, K, K))) return rc;
    if ((rc = ensure_tc_attrs())) return rc;
    TcParams p{bias, residual, C, (__half*)Ch, (__half*)Cl, ldr, ldc, M, N, K, epilogue, alpha, 0};
    const int tiles_n = (N + TBN - 1) / TBN, tiles_m = (M + TBM - 1) / TBM;
    const int num_tiles = tiles_n * tiles_m;
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    tc_gemm_kernel<false><<<grid, TC_THREADS, kTcSmem, (cudaStream_t)stream>>>(maps, p, num_tiles, tiles_n, 1);
    return check_launch("tc_gemm_kernel");
}
