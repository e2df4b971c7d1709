// Per-slot cache bookkeeping of the batched chunk (streaming) path, on the device so that a whole chunk step is a fixed
// sequence of launches (CUDA-graph replayable; no host-built index lists):
//   * append the chunk's new K|V rows to every slot's attention cache   (attention.py:218-225 `torch.cat` on time)
//   * slide every slot's conv-module left context                        (convolution.py:105-109 `new_cache = x[:, :, -lorder:]`)
// Pure data movement: bytes are copied unchanged.
#include "common.cuh"

namespace masr {

// dst[(s*cap + base[s] + t) * ld_dst + c] = src[(s*rows_per_slot + t) * ld_src + col0 + c]   for t < cnt[s], c < ncols
// (16-byte vectors; grid = (rows_per_slot, S))
__global__ void __launch_bounds__(128) stream_append_rows_kernel(const uint8_t* __restrict__ src0, const uint8_t* __restrict__ src1,
                                                                 int64_t src_pitch, int64_t col0_bytes, int row_bytes,
                                                                 uint8_t* __restrict__ dst0, uint8_t* __restrict__ dst1,
                                                                 int64_t dst_pitch, int64_t cap, const int* __restrict__ base,
                                                                 const int* __restrict__ cnt, int rows_per_slot) {
    const int s = blockIdx.y, t = blockIdx.x;
    if (t >= cnt[s]) return;
    const int64_t so = ((int64_t)s * rows_per_slot + t) * src_pitch + col0_bytes;
    const int64_t dof = ((int64_t)s * cap + base[s] + t) * dst_pitch;
    for (int c = threadIdx.x * 16; c < row_bytes; c += 128 * 16) {
        *reinterpret_cast<uint4*>(dst0 + dof + c) = *reinterpret_cast<const uint4*>(src0 + so + c);
        if (src1) *reinterpret_cast<uint4*>(dst1 + dof + c) = *reinterpret_cast<const uint4*>(src1 + so + c);
    }
}

// per slot s with n = cnt[s] > 0: rows [0, lorder) <- rows [n, n + lorder) of its [rows_per_slot, row_bytes] block.
// Each thread owns one 16-byte column and walks the rows upwards: the row it overwrites (r) is always below every row it
// still has to read (n + r' with r' >= r, n >= 1), so the overlapping move is safe without a scratch copy.
__global__ void __launch_bounds__(128) stream_shift_cache_kernel(uint8_t* __restrict__ x0, uint8_t* __restrict__ x1,
                                                                 int64_t rows_per_slot, int lorder, int row_bytes,
                                                                 const int* __restrict__ cnt) {
    const int s = blockIdx.x;
    const int n = cnt[s];
    if (n <= 0) return;
    uint8_t* b0 = x0 + (int64_t)s * rows_per_slot * row_bytes;
    uint8_t* b1 = x1 ? x1 + (int64_t)s * rows_per_slot * row_bytes : nullptr;
    for (int c = threadIdx.x * 16; c < row_bytes; c += 128 * 16)
        for (int r = 0; r < lorder; ++r) {
            *reinterpret_cast<uint4*>(b0 + (int64_t)r * row_bytes + c) = *reinterpret_cast<const uint4*>(b0 + (int64_t)(n + r) * row_bytes + c);
            if (b1) *reinterpret_cast<uint4*>(b1 + (int64_t)r * row_bytes + c) = *reinterpret_cast<const uint4*>(b1 + (int64_t)(n + r) * row_bytes + c);
        }
}

}  // namespace masr

using namespace masr;

extern "C" int masr_stream_append_rows(const void* src0, const void* src1, int64_t src_pitch_bytes, int64_t col0_bytes,
                                       int row_bytes, void* dst0, void* dst1, int64_t dst_pitch_bytes, int64_t cap,
                                       const int* base, const int* cnt, int rows_per_slot, int S, void* stream) {
    if (S == 0 || rows_per_slot == 0 || row_bytes == 0) return MASR_OK;
    MASR_REQUIRE(src0 && dst0 && base && cnt && (src1 == nullptr) == (dst1 == nullptr), "masr_stream_append_rows: bad pointers");
    MASR_REQUIRE(row_bytes % 16 == 0 && col0_bytes % 16 == 0 && src_pitch_bytes % 16 == 0 && dst_pitch_bytes % 16 == 0 &&
                 ((reinterpret_cast<uintptr_t>(src0) | reinterpret_cast<uintptr_t>(dst0) | reinterpret_cast<uintptr_t>(src1) |
                   reinterpret_cast<uintptr_t>(dst1)) & 15) == 0, "masr_stream_append_rows: 16-byte alignment required");
    stream_append_rows_kernel<<<dim3(rows_per_slot, S), 128, 0, (cudaStream_t)stream>>>(
        (const uint8_t*)src0, (const uint8_t*)src1, src_pitch_bytes, col0_bytes, row_bytes, (uint8_t*)dst0, (uint8_t*)dst1,
        dst_pitch_bytes, cap, base, cnt, rows_per_slot);
    return check_launch("stream_append_rows_kernel");
}

extern "C" int masr_stream_shift_cache(void* x0, void* x1, int64_t rows_per_slot, int lorder, int row_bytes, const int* cnt,
                                       int S, void* stream) {
    if (S == 0 || lorder == 0) return MASR_OK;
    MASR_REQUIRE(x0 && cnt, "masr_stream_shift_cache: null pointer");
    MASR_REQUIRE(row_bytes % 16 == 0 && ((reinterpret_cast<uintptr_t>(x0) | reinterpret_cast<uintptr_t>(x1)) & 15) == 0,
                 "masr_stream_shift_cache: 16-byte alignment required");
    stream_shift_cache_kernel<<<S, 128, 0, (cudaStream_t)stream>>>((uint8_t*)x0, (uint8_t*)x1, rows_per_slot, lorder, row_bytes, cnt);
    return check_launch("stream_shift_cache_kernel");
}
