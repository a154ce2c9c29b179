// curl.cu -- forward-difference "curl" of a scalar potential (clumpy `curl_2d`) for sm_100a.
//
// Replaces the loop of commands/curl_2d.cc:57-71:  texel(row, col) = (dpdy, dpdx) with
//   dpdx = p[row][min(col+1, W-1)] - p[row][col],   dpdy = p[row][col] - p[min(row+1, H-1)][col].
// Two FP32 subtractions per pixel: results are bit-identical to the reference.
//
// HBM-bound stencil, 12 algorithmic bytes per pixel (4 read + 8 written).  Each thread owns two
// adjacent columns and walks CURL_ROWS rows downwards, keeping the current row in registers so that
// every input row is loaded once per strip (the one-row halo is the strip's extra load); the right
// neighbour comes from the next lane by shuffle.  One 16-byte store per thread and row, so a warp
// writes 512 contiguous bytes.
#include "device_utils.cuh"
#include "internal.h"

namespace clumpy {

constexpr int CURL_THREADS = 128; // 256 columns per CTA
constexpr int CURL_ROWS = 16;     // rows per strip

// std::min / std::max bookkeeping of curl_2d.cc:64-65 (min over dpdx only, max over both)
struct RangeAcc {
    float lo = 0.f, hi = 0.f;
    bool has_lo = false, has_hi = false;
    __device__ __forceinline__ void add(float dpdx, float dpdy) {
        if (!has_lo) { if (dpdx == dpdx) { lo = dpdx; has_lo = true; } }
        else if (dpdx < lo) lo = dpdx;
        const float m = (dpdx < dpdy) ? dpdy : dpdx;
        if (!has_hi) { if (m == m) { hi = m; has_hi = true; } }
        else if (hi < m) hi = m;
    }
};

__global__ void __launch_bounds__(CURL_THREADS)
curl_pair_kernel(const float* __restrict__ src, uint32_t width, uint32_t nrows,
                 const float* __restrict__ next_row, float* __restrict__ dst,
                 uint32_t* __restrict__ minmax) {
    const uint32_t pair = blockIdx.x * CURL_THREADS + threadIdx.x; // column pair index
    const uint32_t col = pair * 2;
    const uint32_t r0 = blockIdx.y * CURL_ROWS;
    const uint32_t r1 = min(r0 + CURL_ROWS, nrows);
    const bool active = col < width; // width is even on this path
    const bool last_pair = col + 2 >= width;
    const int lane = threadIdx.x & 31;
    RangeAcc acc;
    const uint64_t pol = make_stream_policy();

    float2 cur = make_float2(0.f, 0.f);
    if (active) cur = ld_stream_f2(reinterpret_cast<const float2*>(src + (size_t)r0 * width + col), pol);
    // element right of this thread's pair; lanes 0..30 get it from their neighbour
    float right = 0.f;
    for (uint32_t r = r0; r < r1; ++r) {
        float2 nxt = cur; // bottom edge clamps (curl_2d.cc:60)
        if (active) {
            if (r + 1 < nrows) nxt = ld_stream_f2(reinterpret_cast<const float2*>(src + (size_t)(r + 1) * width + col), pol);
            else if (next_row) nxt = ld_stream_f2(reinterpret_cast<const float2*>(next_row + col), pol);
        }
        float nb = __shfl_down_sync(0xFFFFFFFFu, cur.x, 1);
        if (active) {
            if (last_pair) right = cur.y; // right edge clamps (curl_2d.cc:59)
            else if (lane == 31) right = __ldg(src + (size_t)r * width + col + 2);
            else right = nb;
            const float dpdx0 = __fsub_rn(cur.y, cur.x), dpdx1 = __fsub_rn(right, cur.y);
            const float dpdy0 = __fsub_rn(cur.x, nxt.x), dpdy1 = __fsub_rn(cur.y, nxt.y);
            acc.add(dpdx0, dpdy0);
            acc.add(dpdx1, dpdy1);
            st_stream_f4(reinterpret_cast<float4*>(dst + ((size_t)r * width + col) * 2),
                         make_float4(dpdy0, dpdx0, dpdy1, dpdx1), pol);
        }
        cur = nxt;
    }
    block_minmax_commit(acc.lo, acc.has_lo, acc.hi, acc.has_hi, minmax);
}

// Any width / alignment: one thread per pixel.
__global__ void __launch_bounds__(256)
curl_scalar_kernel(const float* __restrict__ src, uint32_t width, uint32_t nrows,
                   const float* __restrict__ next_row, float* __restrict__ dst,
                   uint32_t* __restrict__ minmax) {
    RangeAcc acc;
    const uint64_t npix = (uint64_t)width * nrows;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t row = (uint32_t)(i / width), col = (uint32_t)(i - (uint64_t)row * width);
        const float p = src[i];
        const float pr = (col + 1 < width) ? src[i + 1] : p;
        float pd = p;
        if (row + 1 < nrows) pd = src[i + width];
        else if (next_row) pd = next_row[col];
        const float dpdx = __fsub_rn(pr, p), dpdy = __fsub_rn(p, pd);
        acc.add(dpdx, dpdy);
        dst[2 * i] = dpdy;
        dst[2 * i + 1] = dpdx;
    }
    block_minmax_commit(acc.lo, acc.has_lo, acc.hi, acc.has_hi, minmax);
}

static int run_curl(clumpy_cuda_ctx* ctx, const float* d_src, uint32_t width, uint32_t nrows,
                    const float* d_next, float* d_dst, float* minval, float* maxval) {
    int rc = minmax_reset(ctx);
    if (rc) return rc;
    if (width > 0 && nrows > 0) {
        const bool pair_ok = (width % 2 == 0) && ((uintptr_t)d_src % 8 == 0) &&
                             ((uintptr_t)d_dst % 16 == 0) && (!d_next || (uintptr_t)d_next % 8 == 0);
        if (pair_ok) {
            dim3 grid(ceil_div(width / 2, CURL_THREADS), ceil_div(nrows, CURL_ROWS));
            curl_pair_kernel<<<grid, CURL_THREADS, 0, ctx->stream>>>(d_src, width, nrows, d_next,
                                                                     d_dst, ctx->d_minmax);
        } else {
            const uint64_t npix = (uint64_t)width * nrows;
            uint64_t want = (npix + 255) / 256, cap = (uint64_t)ctx->sm_count * 32;
            curl_scalar_kernel<<<(uint32_t)(want < cap ? want : cap), 256, 0, ctx->stream>>>(
                d_src, width, nrows, d_next, d_dst, ctx->d_minmax);
        }
        CK_LAUNCH(ctx);
    }
    if (minval || maxval) return minmax_fetch(ctx, minval, maxval);
    return CLUMPY_OK;
}

} // namespace clumpy

using namespace clumpy;

CLUMPY_API int clumpy_cuda_curl_2d_dev(clumpy_cuda_ctx* ctx, const float* src_dev, uint32_t width,
                                       uint32_t nrows, const float* next_row_dev, float* dst_dev,
                                       float* minval, float* maxval) {
    REQUIRE(ctx, "ctx is NULL");
    REQUIRE((src_dev && dst_dev) || width == 0 || nrows == 0, "null device pointer");
    CK(cudaSetDevice(ctx->device));
    return run_curl(ctx, src_dev, width, nrows, next_row_dev, dst_dev, minval, maxval);
}

CLUMPY_API int clumpy_cuda_curl_2d(clumpy_cuda_ctx* ctx, const float* src_host, uint32_t width,
                                   uint32_t height, float* dst_host, float* minval, float* maxval) {
    REQUIRE(ctx && src_host && dst_host, "null argument");
    REQUIRE(width > 0 && height > 0, "image dimensions must be positive");
    CK(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)width * height;
    float *d_src = nullptr, *d_dst = nullptr;
    CK(cudaMalloc(&d_src, npix * sizeof(float)));
    if (cudaMalloc(&d_dst, npix * 2 * sizeof(float)) != cudaSuccess) {
        cudaFree(d_src);
        set_error("out of device memory for the curl output");
        return CLUMPY_ERR_NOMEM;
    }
    int rc = CLUMPY_OK;
    cudaError_t e = cudaMemcpyAsync(d_src, src_host, npix * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { set_error("H2D of the potential failed: %s", cudaGetErrorString(e)); rc = CLUMPY_ERR_CUDA; }
    if (rc == CLUMPY_OK) rc = run_curl(ctx, d_src, width, height, nullptr, d_dst, nullptr, nullptr);
    if (rc == CLUMPY_OK) {
        e = cudaMemcpyAsync(dst_host, d_dst, npix * 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error("D2H of the curl field failed: %s", cudaGetErrorString(e)); rc = CLUMPY_ERR_CUDA; }
    }
    if (rc == CLUMPY_OK && (minval || maxval)) rc = minmax_fetch(ctx, minval, maxval);
    cudaFree(d_src);
    cudaFree(d_dst);
    return rc;
}
