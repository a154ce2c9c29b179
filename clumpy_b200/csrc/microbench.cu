// microbench.cu -- the atomic leg of the splat roofline ("the builder must micro-benchmark
// red.global.add ... to define the L2-atomic peak the target's '>= 60 % of HBM / L2-atomic roofline' refers to").
//
// The splat of advect_points (splat_points.cc:142-160) is, on the device, one fire-and-forget reduction per
// point into a counter image (advect.cu cell_add_f16 / atomicAdd).  What bounds such a stream of reductions is
// not in MEASURED_PEAKS.json, so it is measured here with the same instructions on the same access patterns:
//
//   op       0  red.global.add.noftz.f16x2 (two half cells per word -- the C4 path)     1  red.global.add.u32
//   pattern  0  uniform random cells of a region (worst case: every update its own 32-byte sector)
//            1  hot spot: all updates on `hot_cells` cells (contention at one L2 slice)
//            2  local: update i lands near cell i * ncells / nupdates (+- 64 cells) -- a cloud stored in tile
//               order like ours, each warp's updates within a few lines
//
// Each launch issues `nupdates` reductions; the caller times `iters` launches with CUDA events on the context's
// stream (clumpy_cuda_atomic_peak does it and returns the mean).  tools/atomic_peak.py prints the table.
#include "device_utils.cuh"
#include "internal.h"

#include <algorithm>

namespace clumpy {

__device__ __forceinline__ uint64_t mix64(uint64_t z) { // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <int OP, int PATTERN>
__global__ void __launch_bounds__(256)
atomic_peak_kernel(uint32_t* __restrict__ words, uint32_t ncells, uint64_t nupdates, uint32_t hot_cells, uint32_t salt) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nupdates; i += stride) {
        const uint32_t h = (uint32_t)(mix64(i + ((uint64_t)salt << 40)) >> 32);
        uint32_t cell;
        if (PATTERN == 0) cell = (uint32_t)(((uint64_t)h * ncells) >> 32);
        else if (PATTERN == 1) cell = (uint32_t)(((uint64_t)h * hot_cells) >> 32);
        else {
            const uint64_t centre = (uint64_t)((double)i * (double)ncells / (double)nupdates);
            cell = (uint32_t)min((uint64_t)ncells - 1u, centre + (h & 127u));
        }
        if (OP == 0) {
            const uint32_t one = (cell & 1u) ? 0x3C000000u : 0x00003C00u;
            asm volatile("red.global.add.noftz.f16x2 [%0], %1;" :: "l"(words + (cell >> 1)), "r"(one) : "memory");
        } else {
            asm volatile("red.global.add.u32 [%0], %1;" :: "l"(words + cell), "r"(1u) : "memory");
        }
    }
}

template <int OP>
static void launch_peak(int pattern, uint32_t grid, cudaStream_t s, uint32_t* words, uint32_t ncells,
                        uint64_t nupdates, uint32_t hot, uint32_t salt) {
    if (pattern == 0) atomic_peak_kernel<OP, 0><<<grid, 256, 0, s>>>(words, ncells, nupdates, hot, salt);
    else if (pattern == 1) atomic_peak_kernel<OP, 1><<<grid, 256, 0, s>>>(words, ncells, nupdates, hot, salt);
    else atomic_peak_kernel<OP, 2><<<grid, 256, 0, s>>>(words, ncells, nupdates, hot, salt);
}

} // namespace clumpy

using namespace clumpy;

CLUMPY_API int clumpy_cuda_atomic_peak(clumpy_cuda_ctx* ctx, int op, int pattern, size_t region_bytes,
                                       uint64_t nupdates, uint32_t hot_cells, int iters, float* ms_per_launch) {
    REQUIRE(ctx && ms_per_launch, "null argument");
    REQUIRE((op == 0 || op == 1) && pattern >= 0 && pattern <= 2, "bad op / pattern");
    REQUIRE(region_bytes >= 256 && nupdates > 0 && iters > 0, "bad size");
    const size_t cell_bytes = op == 0 ? 2 : 4;
    REQUIRE(region_bytes / cell_bytes < (1ull << 32), "region too large");
    CK(cudaSetDevice(ctx->device));
    const uint32_t ncells = (uint32_t)(region_bytes / cell_bytes);
    if (hot_cells == 0 || hot_cells > ncells) hot_cells = 1;
    uint32_t* words = nullptr;
    CK(cudaMalloc(&words, region_bytes));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = CLUMPY_OK;
    auto fail = [&](const char* what) { set_error("%s failed: %s", what, cudaGetErrorString(cudaGetLastError())); rc = CLUMPY_ERR_CUDA; };
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) fail("event creation");
    // as many threads as the advect kernel keeps resident: 8 CTAs of 256 per SM
    const uint32_t grid = (uint32_t)std::min<uint64_t>((nupdates + 255) / 256, (uint64_t)ctx->sm_count * 8);
    for (int it = -2; rc == CLUMPY_OK && it < iters; ++it) { // two warm-up launches
        // half cells saturate at 2048: start every launch from zero, outside the timed kernels' own work
        if (it == 0 && cudaEventRecord(e0, ctx->stream) != cudaSuccess) { fail("event record"); break; }
        if (cudaMemsetAsync(words, 0, region_bytes, ctx->stream) != cudaSuccess) { fail("memset"); break; }
        if (op == 0) launch_peak<0>(pattern, grid, ctx->stream, words, ncells, nupdates, hot_cells, (uint32_t)(it + 2));
        else launch_peak<1>(pattern, grid, ctx->stream, words, ncells, nupdates, hot_cells, (uint32_t)(it + 2));
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) { fail("launch"); break; }
    }
    float ms = 0.f, ms_clear = 0.f;
    if (rc == CLUMPY_OK) {
        if (cudaEventRecord(e1, ctx->stream) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess ||
            cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess) fail("timing");
    }
    // the clears are inside the bracket: time them alone and take them out
    if (rc == CLUMPY_OK) {
        cudaEventRecord(e0, ctx->stream);
        for (int it = 0; it < iters; ++it) cudaMemsetAsync(words, 0, region_bytes, ctx->stream);
        if (cudaEventRecord(e1, ctx->stream) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess ||
            cudaEventElapsedTime(&ms_clear, e0, e1) != cudaSuccess) fail("timing");
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(words);
    *ms_per_launch = rc == CLUMPY_OK ? std::max(0.f, ms - ms_clear) / (float)iters : 0.f;
    return rc;
}
