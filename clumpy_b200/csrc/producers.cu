// producers.cu -- the commands either side of the advection path: `pendulum_phase`
// (the vector field of README example6) and `cull_points` (the filter between bridson_points and
// advect_points in extras/test.py), for sm_100a.  `bridson_points` is inherently sequential dart
// throwing and stays on the host (host/commands/bridson_points.cc).
//
// pendulum_phase (commands/pendulum_phase.cc:38-71):
//     texel(row, col) = (omega, -friction * omega - g / L * sin(theta)),   g = L = 1,
//     theta = sx * (col / W - 0.5),  omega = sy * (row / H - 0.5),  sx = (float)(scale_x * M_PI / 0.5)
// theta depends on the column only and omega on the row only, and the reference takes `sin` from the
// host's libm (the call resolves to sinf in the compiled binary).  So the host evaluates the W column
// values sinf(theta_col) and the H row values omega_row exactly as the reference does -- W + H calls
// instead of W * H -- and the kernel combines them: two FP32 operations and one 8-byte store per texel,
// 8 B/px of HBM writes, bit-identical to the reference by construction.
//
// cull_points (commands/cull_points.cc:97-111): keep point i iff mask(nearest texel) > 0 (float ->
// uint32_t conversion of the coordinates as the x86 binary does it; outside the image reads 0), in the
// original order.  Three small kernels: flags + per-block counts, a single-block scan of the block
// counts, and the order-preserving scatter.
#include "device_utils.cuh"
#include "internal.h"

#include <cmath>
#include <vector>

namespace clumpy {

// ---- pendulum_phase ---------------------------------------------------------------------------------

constexpr int PEND_THREADS = 256;

// One thread per two adjacent texels of a row: a 16-byte store, 512 contiguous bytes per warp.
__global__ void __launch_bounds__(PEND_THREADS)
pendulum_kernel(const float* __restrict__ sin_col, const float* __restrict__ omega_row, uint32_t width,
                uint32_t nrows, float neg_friction, float g_over_l, float* __restrict__ out) {
    const uint32_t pair = blockIdx.x * PEND_THREADS + threadIdx.x;
    const uint32_t col = pair * 2;
    if (col >= width) return;
    const float s0 = __fmul_rn(g_over_l, sin_col[col]);
    const float s1 = col + 1 < width ? __fmul_rn(g_over_l, sin_col[col + 1]) : 0.f;
    const uint64_t pol = make_stream_policy();
    for (uint32_t row = blockIdx.y; row < nrows; row += gridDim.y) {
        const float omega = omega_row[row];
        const float damp = __fmul_rn(neg_friction, omega); // -friction * omega (:63)
        float* dst = out + ((size_t)row * width + col) * 2;
        if (col + 1 < width && (((uintptr_t)dst) & 15u) == 0)
            st_stream_f4(reinterpret_cast<float4*>(dst), make_float4(omega, __fsub_rn(damp, s0), omega, __fsub_rn(damp, s1)), pol);
        else {
            dst[0] = omega; dst[1] = __fsub_rn(damp, s0);
            if (col + 1 < width) { dst[2] = omega; dst[3] = __fsub_rn(damp, s1); }
        }
    }
}

// Host tables: exactly the reference's per-texel expressions (pendulum_phase.cc:52-62), evaluated once
// per column / row.  `0.5` and M_PI are doubles there, so the subtraction and the product run in double.
static void pendulum_tables(uint32_t width, uint32_t height, uint32_t row0, uint32_t nrows, float scalex, float scaley,
                            std::vector<float>& sin_col, std::vector<float>& omega_row) {
    const float gsx = (float)((double)scalex * M_PI / 0.5), gsy = scaley * 1;
    sin_col.resize(width);
    omega_row.resize(nrows);
    for (uint32_t col = 0; col < width; ++col) {
        const float theta = (float)((double)gsx * ((double)((float)col / (float)width) - 0.5));
        sin_col[col] = sinf(theta);
    }
    for (uint32_t r = 0; r < nrows; ++r)
        omega_row[r] = (float)((double)gsy * ((double)((float)(row0 + r) / (float)height) - 0.5));
}

static int pendulum_launch(clumpy_cuda_ctx* ctx, uint32_t width, uint32_t height, uint32_t row0, uint32_t nrows,
                           float friction, float scalex, float scaley, float* d_out) {
    std::vector<float> sin_col, omega_row;
    pendulum_tables(width, height, row0, nrows, scalex, scaley, sin_col, omega_row);
    float *d_sin = nullptr, *d_omega = nullptr;
    CK(cudaMalloc(&d_sin, (size_t)width * 4));
    if (cudaMalloc(&d_omega, (size_t)nrows * 4) != cudaSuccess) { cudaFree(d_sin); set_error("out of device memory"); return CLUMPY_ERR_NOMEM; }
    cudaMemcpyAsync(d_sin, sin_col.data(), (size_t)width * 4, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_omega, omega_row.data(), (size_t)nrows * 4, cudaMemcpyHostToDevice, ctx->stream);
    const dim3 grid(ceil_div(ceil_div(width, 2), PEND_THREADS),
                    std::min<uint32_t>(nrows, std::max<uint32_t>(1u, (uint32_t)ctx->sm_count * 8 / std::max<uint32_t>(1u, ceil_div(ceil_div(width, 2), PEND_THREADS)))));
    const float g = 1.0f, L = 1.0f;
    pendulum_kernel<<<grid, PEND_THREADS, 0, ctx->stream>>>(d_sin, d_omega, width, nrows, -friction, g / L, d_out);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    // the tables are host vectors that die with this frame: wait for their copies (and the kernel)
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_sin);
    cudaFree(d_omega);
    if (e != cudaSuccess) { set_error("pendulum_phase failed: %s", cudaGetErrorString(e)); return CLUMPY_ERR_CUDA; }
    return CLUMPY_OK;
}

// ---- cull_points ------------------------------------------------------------------------------------

constexpr int CULL_THREADS = 256;

__device__ __forceinline__ bool cull_keep(float2 p, const float* __restrict__ mask, uint32_t width, uint32_t height) {
    const uint32_t x = f2u32_x86(p.x), y = f2u32_x86(p.y); // cull_points.cc:21-23,102: implicit float -> uint32_t
    const float m = (x >= width || y >= height) ? 0.f : __ldg(mask + x + (size_t)width * y);
    return m > 0.f;
}

__global__ void __launch_bounds__(CULL_THREADS)
cull_count_kernel(const float2* __restrict__ pts, uint32_t npts, const float* __restrict__ mask, uint32_t width,
                  uint32_t height, uint32_t* __restrict__ block_counts) {
    const uint32_t i = blockIdx.x * CULL_THREADS + threadIdx.x;
    const bool keep = i < npts && cull_keep(pts[i], mask, width, height);
    const int n = __syncthreads_count(keep);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)n;
}

// exclusive scan of the block counts in place (any length), total kept into *total
__global__ void __launch_bounds__(1024)
cull_scan_kernel(uint32_t* __restrict__ counts, uint32_t n, uint32_t* __restrict__ total) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n; base += 1024u) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? counts[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const uint32_t pre = (warp ? warp_tot[warp - 1] : 0u) + carry_s;
        if (i < n) counts[i] = pre + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = pre + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

__global__ void __launch_bounds__(CULL_THREADS)
cull_scatter_kernel(const float2* __restrict__ pts, uint32_t npts, const float* __restrict__ mask, uint32_t width,
                    uint32_t height, const uint32_t* __restrict__ block_offsets, float2* __restrict__ out) {
    __shared__ uint32_t warp_base[CULL_THREADS / 32];
    const uint32_t i = blockIdx.x * CULL_THREADS + threadIdx.x;
    float2 p = make_float2(0.f, 0.f);
    bool keep = false;
    if (i < npts) { p = pts[i]; keep = cull_keep(p, mask, width, height); }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep);
    if (lane == 0) warp_base[warp] = (uint32_t)__popc(ballot);
    __syncthreads();
    uint32_t before = block_offsets[blockIdx.x];
    for (uint32_t w = 0; w < warp; ++w) before += warp_base[w];
    if (keep) out[before + (uint32_t)__popc(ballot & ((1u << lane) - 1u))] = p; // original order
}

} // namespace clumpy

using namespace clumpy;

CLUMPY_API int clumpy_cuda_pendulum_phase(clumpy_cuda_ctx* ctx, uint32_t width, uint32_t height,
                                          float friction, float scale_x, float scale_y, float* out_host) {
    REQUIRE(ctx && out_host, "null argument");
    REQUIRE(width > 0 && height > 0, "image dimensions must be positive");
    REQUIRE((uint64_t)width * height < (1ull << 32), "image too large");
    CK(cudaSetDevice(ctx->device));
    const size_t nbytes = (size_t)width * height * 8;
    float* d_out = nullptr;
    CK(cudaMalloc(&d_out, nbytes));
    int rc = pendulum_launch(ctx, width, height, 0, height, friction, scale_x, scale_y, d_out);
    cudaError_t e = cudaSuccess;
    if (rc == CLUMPY_OK) {
        e = cudaMemcpyAsync(out_host, d_out, nbytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(d_out);
    if (rc) return rc;
    if (e != cudaSuccess) { set_error("pendulum_phase copy failed: %s", cudaGetErrorString(e)); return CLUMPY_ERR_CUDA; }
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_pendulum_phase_dev(clumpy_cuda_ctx* ctx, uint32_t width, uint32_t height,
                                              uint32_t row0, uint32_t nrows, float friction, float scale_x,
                                              float scale_y, float* out_dev) {
    REQUIRE(ctx && (out_dev || nrows == 0), "null argument");
    REQUIRE(width > 0 && height > 0, "image dimensions must be positive");
    REQUIRE((uint64_t)row0 + nrows <= height, "row band exceeds the image");
    CK(cudaSetDevice(ctx->device));
    if (nrows == 0) return CLUMPY_OK;
    return pendulum_launch(ctx, width, height, row0, nrows, friction, scale_x, scale_y, out_dev);
}

CLUMPY_API int clumpy_cuda_cull_points(clumpy_cuda_ctx* ctx, const float* pts_host, uint32_t npts,
                                       const float* mask_host, uint32_t width, uint32_t height,
                                       float* out_host, uint32_t* nkept) {
    REQUIRE(ctx && nkept && mask_host && (pts_host || npts == 0) && (out_host || npts == 0), "null argument");
    REQUIRE(width > 0 && height > 0, "image dimensions must be positive");
    REQUIRE((uint64_t)width * height < (1ull << 32), "mask too large");
    CK(cudaSetDevice(ctx->device));
    *nkept = 0;
    if (npts == 0) return CLUMPY_OK;
    const uint32_t nblocks = ceil_div(npts, CULL_THREADS);
    float2 *d_pts = nullptr, *d_out = nullptr;
    float* d_mask = nullptr;
    uint32_t *d_counts = nullptr, *d_total = nullptr;
    cudaError_t e = cudaMalloc(&d_pts, (size_t)npts * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)npts * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_mask, (size_t)width * height * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_counts, (size_t)nblocks * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_total, 4);
    uint32_t total = 0;
    if (e == cudaSuccess) {
        cudaMemcpyAsync(d_pts, pts_host, (size_t)npts * 8, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d_mask, mask_host, (size_t)width * height * 4, cudaMemcpyHostToDevice, ctx->stream);
        cull_count_kernel<<<nblocks, CULL_THREADS, 0, ctx->stream>>>(d_pts, npts, d_mask, width, height, d_counts);
        cull_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, nblocks, d_total);
        cull_scatter_kernel<<<nblocks, CULL_THREADS, 0, ctx->stream>>>(d_pts, npts, d_mask, width, height, d_counts, d_out);
        ctx->launches += 3;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_total, 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess && total > 0) {
            e = cudaMemcpyAsync(out_host, d_out, (size_t)total * 8, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        }
    }
    cudaFree(d_pts); cudaFree(d_out); cudaFree(d_mask); cudaFree(d_counts); cudaFree(d_total);
    if (e != cudaSuccess) {
        set_error("cull_points failed: %s", cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? CLUMPY_ERR_NOMEM : CLUMPY_ERR_CUDA;
    }
    *nkept = total;
    return CLUMPY_OK;
}
