// Host-side input staging for the batched path: gather the caller's utterances (separate float32 host arrays, what
// `MASRPredictor.predict` receives one at a time, masr/predict.py:147-164) into one pinned buffer and ship them to the
// device, with the packing spread over a few host threads and every finished part's H2D copy issued at once so that the
// PCIe transfer overlaps the packing of the following parts.  No arithmetic happens here.
#include <atomic>
#include <string.h>
#include <thread>
#include <vector>

#include "common.cuh"

using namespace masr;

// waves[b] -> pinned[offs[b] ..) -> dev[offs[b] ..), offs = exclusive prefix sum of lengths (in samples).
// `pinned` must be page-locked (cudaHostAlloc / torch pin_memory) and hold sum(lengths) floats; so must `dev`.
// Returns after the last cudaMemcpyAsync has been ISSUED on `stream` (the copies complete in stream order).
extern "C" int masr_stage_waves_f32(const void* const* waves, const int64_t* lengths, int B, float* pinned, float* dev,
                                    int nthreads, void* stream) {
    if (B == 0) return MASR_OK;
    MASR_REQUIRE(waves && lengths && pinned && dev, "masr_stage_waves_f32: null pointer");
    std::vector<int64_t> offs((size_t)B + 1, 0);
    for (int b = 0; b < B; ++b) {
        MASR_REQUIRE(lengths[b] >= 0 && (lengths[b] == 0 || waves[b]), "masr_stage_waves_f32: bad utterance %d", b);
        offs[b + 1] = offs[b] + lengths[b];
    }
    const int64_t total = offs[B];
    if (total == 0) return MASR_OK;
    cudaStream_t s = (cudaStream_t)stream;
    // parts: contiguous utterance ranges of roughly equal size (>= 256 K samples each, at most one per thread)
    int parts = nthreads < 1 ? 1 : nthreads;
    if (parts > B) parts = B;
    const int64_t min_part = 256 * 1024;
    if ((int64_t)parts * min_part > total) parts = (int)(total / min_part > 0 ? total / min_part : 1);
    std::vector<int> first((size_t)parts + 1, B);
    first[0] = 0;
    for (int p = 1, b = 0; p < parts; ++p) {
        const int64_t target = total * p / parts;
        while (b < B && offs[b] < target) ++b;
        first[p] = b;
    }
    auto pack = [&](int p) {
        for (int b = first[p]; b < first[p + 1]; ++b)
            if (lengths[b]) memcpy(pinned + offs[b], waves[b], (size_t)lengths[b] * sizeof(float));
    };
    auto ship = [&](int p) -> cudaError_t {
        const int64_t o0 = offs[first[p]], o1 = offs[first[p + 1]];
        if (o1 == o0) return cudaSuccess;
        return cudaMemcpyAsync(dev + o0, pinned + o0, (size_t)(o1 - o0) * sizeof(float), cudaMemcpyHostToDevice, s);
    };
    cudaError_t err = cudaSuccess;
    if (parts == 1) {
        pack(0);
        err = ship(0);
    } else {
        std::vector<std::atomic<int>> done((size_t)parts);
        for (auto& d : done) d.store(0, std::memory_order_relaxed);
        std::vector<std::thread> th;
        th.reserve((size_t)parts - 1);
        for (int p = 1; p < parts; ++p)
            th.emplace_back([&, p] { pack(p); done[p].store(1, std::memory_order_release); });
        pack(0);                                   // the calling thread packs the first part, then ships parts in order
        err = ship(0);
        for (int p = 1; p < parts; ++p) {
            while (!done[p].load(std::memory_order_acquire)) std::this_thread::yield();
            if (err == cudaSuccess) err = ship(p);
        }
        for (auto& t : th) t.join();
    }
    if (err != cudaSuccess) {
        set_last_error("masr_stage_waves_f32: cudaMemcpyAsync: %s", cudaGetErrorString(err));
        return (int)err;
    }
    return MASR_OK;
}
