// splat_multi.cu -- `splat_points` (commands/splat_points.cc:44-104) on SEVERAL GPUs of one box from one process
// ("splat_points command: same as K4; the fast path is idempotent => reduce with max / OR").
//
//   points    sharded by index range; every GPU counts ITS points into a private accumulator of the whole image
//             (32-bit counts, or for sparse lists the 64-bit exact-order cells: count + sum of ORIGINAL indices --
//             both additive, so summing the GPUs' accumulators gives exactly the one-GPU accumulator);
//   exchange  a reduce-scatter by row bands done with peer loads: the GPU that owns a band sums that band's rows of
//             every GPU's accumulator (its own included) straight out of their memory over NVLink;
//   composite each GPU composites its band (splat.cu / composite_exact.cu, as on one GPU) onto its rows of the
//             caller's image and copies them back over its own PCIe link.
// The alpha = 1, k = 1 fast path (splat_points.cc:108-116) stores 255 per point: the private images are combined
// with OR.  Results are identical to the one-GPU entry points (the accumulators are).
#include "device_utils.cuh"
#include "internal.h"

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

namespace clumpy {

struct PeerPtrs { const void* p[8]; };

// dst[i] = sum over ranks of src_r[i] (T = uint32_t or unsigned long long), i < n
template <typename T>
__global__ void __launch_bounds__(256)
sum_peers_kernel(T* __restrict__ dst, PeerPtrs src, int nranks, size_t offset, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        T acc = 0;
        for (int r = 0; r < nranks; ++r) acc += static_cast<const T*>(src.p[r])[offset + i];
        dst[i] = acc;
    }
}

// img[i] |= OR over ranks of marks_r[offset + i]  (16 bytes at a time; n16 uint4 words)
__global__ void __launch_bounds__(256)
or_peers_kernel(uint4* __restrict__ img, PeerPtrs src, int nranks, size_t offset16, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 acc = img[i];
        for (int r = 0; r < nranks; ++r) {
            const uint4 v = static_cast<const uint4*>(src.p[r])[offset16 + i];
            acc.x |= v.x; acc.y |= v.y; acc.z |= v.z; acc.w |= v.w;
        }
        img[i] = acc;
    }
}

__global__ void __launch_bounds__(256)
mark_points_shard_kernel(const float2* __restrict__ pts, uint32_t npts, uint32_t width, uint32_t height, uint8_t* __restrict__ img) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += gridDim.x * blockDim.x) {
        const float2 p = pts[i];
        const uint32_t x = f2u32_x86(p.x), y = f2u32_x86(p.y); // splat_points.cc:110-111
        if (x < width && y < height) img[(size_t)width * y + x] = 255;
    }
}

} // namespace clumpy

using namespace clumpy;

namespace {
struct SplatRank {
    clumpy_cuda_ctx* ctx = nullptr;
    float* pts = nullptr;
    void* acc = nullptr;      // private accumulator of the whole image (cudaMalloc: the peers read it)
    void* band = nullptr;     // summed accumulator rows of this rank's band
    uint8_t* img = nullptr;   // this rank's rows of the image
    uint32_t lo = 0, hi = 0, y0 = 0, y1 = 0;
    int rc = CLUMPY_OK;
    std::string error;
};
} // namespace

CLUMPY_API int clumpy_cuda_splat_multi(const int* devices, int ndevices, const float* pts_host, uint32_t npts,
                                       uint32_t width, uint32_t height, uint8_t* img_host, float alpha,
                                       int kernel_size, int wrapx) {
    REQUIRE(devices && ndevices >= 1 && ndevices <= 8, "1 .. 8 devices");
    REQUIRE(img_host && (pts_host || npts == 0), "null argument");
    REQUIRE(width > 0 && height > 0 && width < (1u << 30) && height < (1u << 30), "bad image dimensions");
    const bool fast = alpha == 1.0f && kernel_size == 1; // splat_points.cc:93-94
    SpriteRings rings;
    if (!fast) {
        int rc = make_sprite_rings(kernel_size, alpha, &rings);
        if (rc) return rc;
    }
    const int world = ndevices;
    const HistGeom g = make_hist_geom(width, height, fast ? 1 : kernel_size);
    const size_t npix = (size_t)width * height;
    bool exact = !fast && exact_mode_supported(kernel_size, npts) && 8ull * npts <= (uint64_t)npix;
    if (const char* force = getenv("CLUMPY_CUDA_CELLS")) {
        if (!strcmp(force, "u32")) exact = false;
        if (!strcmp(force, "exact")) exact = !fast && exact_mode_supported(kernel_size, npts);
    }
    const size_t cell_bytes = exact ? 8 : 4;
    // the fast path ORs whole 16-byte words: its private images are padded to a multiple of 16 bytes per band
    const size_t acc_bytes = fast ? round_up(npix, 16) + 16 : g.cells() * cell_bytes;
    std::vector<SplatRank> ranks(world);
    auto cleanup = [&]() {
        for (auto& k : ranks) {
            if (!k.ctx) continue;
            cudaSetDevice(k.ctx->device);
            cudaStreamSynchronize(k.ctx->stream);
            cudaFree(k.pts); cudaFree(k.acc); cudaFree(k.band); cudaFree(k.img);
            clumpy_cuda_destroy(k.ctx);
            k.ctx = nullptr;
        }
    };
    int rc = CLUMPY_OK;
    for (int r = 0; r < world && rc == CLUMPY_OK; ++r) rc = clumpy_cuda_create(devices[r], nullptr, &ranks[r].ctx);
    for (int a = 0; a < world && rc == CLUMPY_OK; ++a)
        for (int b = 0; b < world && rc == CLUMPY_OK; ++b) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (!can) { set_error("device %d cannot access device %d's memory (no NVLink / P2P)", devices[a], devices[b]); rc = CLUMPY_ERR_CUDA; break; }
            cudaSetDevice(devices[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e)); rc = CLUMPY_ERR_CUDA; }
            cudaGetLastError();
        }
    if (rc) { cleanup(); return rc; }
    // row bands: whole 16-byte words of the image for the fast path's OR (rows x width bytes)
    for (int r = 0; r < world; ++r) {
        ranks[r].lo = (uint32_t)((uint64_t)npts * r / world);
        ranks[r].hi = (uint32_t)((uint64_t)npts * (r + 1) / world);
        ranks[r].y0 = (uint32_t)((uint64_t)height * r / world);
        ranks[r].y1 = (uint32_t)((uint64_t)height * (r + 1) / world);
    }
    auto run_phase = [&](auto&& body) {
        std::vector<std::thread> workers;
        for (int r = 0; r < world; ++r)
            workers.emplace_back([&, r] {
                SplatRank& k = ranks[r];
                if (k.rc) return;
                k.rc = body(r, k);
                if (k.rc) k.error = clumpy_cuda_last_error();
            });
        for (auto& t : workers) t.join();
        for (auto& k : ranks)
            if (k.rc) { set_error("%s", k.error.c_str()); return k.rc; }
        return (int)CLUMPY_OK;
    };
    // phase 1: every GPU counts (or marks) its shard of the points into its private accumulator
    rc = run_phase([&](int, SplatRank& k) -> int {
        clumpy_cuda_ctx* ctx = k.ctx;
        CK(cudaSetDevice(ctx->device));
        const uint32_t n = k.hi - k.lo;
        CK(cudaMalloc(&k.pts, std::max<size_t>(8, (size_t)n * 8)));
        CK(cudaMalloc(&k.acc, acc_bytes));
        CK(cudaMemcpyAsync(k.pts, pts_host + 2 * (size_t)k.lo, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(k.acc, 0, acc_bytes, ctx->stream));
        if (n) {
            if (fast) {
                mark_points_shard_kernel<<<std::min<uint32_t>(ceil_div(n, 256), (uint32_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(
                    reinterpret_cast<const float2*>(k.pts), n, width, height, static_cast<uint8_t*>(k.acc));
                CK_LAUNCH(ctx);
            } else if (exact) {
                int e = launch_exact_count(ctx, k.pts, n, g, wrapx, k.acc, k.lo);
                if (e) return e;
            } else {
                int e = launch_hist_points(ctx, k.pts, n, g, wrapx, static_cast<uint32_t*>(k.acc));
                if (e) return e;
            }
        }
        CK(cudaStreamSynchronize(ctx->stream));
        return CLUMPY_OK;
    });
    // phase 2: every GPU sums its band out of all the accumulators, composites it and copies its rows back
    if (rc == CLUMPY_OK)
        rc = run_phase([&](int, SplatRank& k) -> int {
            clumpy_cuda_ctx* ctx = k.ctx;
            CK(cudaSetDevice(ctx->device));
            const uint32_t nrows = k.y1 - k.y0;
            if (nrows == 0) return CLUMPY_OK;
            PeerPtrs src;
            for (int r = 0; r < 8; ++r) src.p[r] = r < world ? ranks[r].acc : nullptr;
            const size_t band_px = (size_t)nrows * width;
            uint8_t* host_rows = img_host + (size_t)k.y0 * width;
            const uint32_t blocks = (uint32_t)ctx->sm_count * 8;
            if (fast) {
                // bytes [y0 * W, y1 * W) of the image, widened to whole 16-byte words (the extra bytes belong to the
                // neighbouring bands, whose owners compute the same values for them: the host copy below takes only ours)
                const size_t b0 = (size_t)k.y0 * width, b1 = b0 + band_px;
                const size_t w0 = b0 / 16, w1 = (b1 + 15) / 16;
                CK(cudaMalloc(&k.img, (w1 - w0) * 16));
                CK(cudaMemsetAsync(k.img, 0, (w1 - w0) * 16, ctx->stream));
                CK(cudaMemcpyAsync(k.img + (b0 - w0 * 16), host_rows, band_px, cudaMemcpyHostToDevice, ctx->stream));
                or_peers_kernel<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<uint4*>(k.img), src, world, w0, w1 - w0);
                CK_LAUNCH(ctx);
                CK(cudaMemcpyAsync(host_rows, k.img + (b0 - w0 * 16), band_px, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
                return CLUMPY_OK;
            }
            // the band as an image of its own: padded rows [y0, y0 + rows_band) of the accumulators, zero beyond them
            HistGeom gb = g;
            gb.height = (int)nrows;
            gb.rows = (int)(round_up(nrows, 16) + 2 * g.halo);
            const size_t avail_rows = std::min<size_t>((size_t)gb.rows, (size_t)g.rows - k.y0);
            CK(cudaMalloc(&k.band, gb.cells() * cell_bytes));
            CK(cudaMemsetAsync(k.band, 0, gb.cells() * cell_bytes, ctx->stream));
            const size_t n = avail_rows * g.pitch, off = (size_t)k.y0 * g.pitch;
            if (exact) sum_peers_kernel<unsigned long long><<<blocks, 256, 0, ctx->stream>>>(static_cast<unsigned long long*>(k.band), src, world, off, n);
            else sum_peers_kernel<uint32_t><<<blocks, 256, 0, ctx->stream>>>(static_cast<uint32_t*>(k.band), src, world, off, n);
            CK_LAUNCH(ctx);
            CK(cudaMalloc(&k.img, band_px));
            CK(cudaMemcpyAsync(k.img, host_rows, band_px, cudaMemcpyHostToDevice, ctx->stream));
            int e = exact ? launch_composite_exact(ctx, gb, rings, k.band, wrapx, /*apply_decay=*/0, 1.0f, k.img, k.img)
                          : launch_composite(ctx, gb, rings, static_cast<const uint32_t*>(k.band), wrapx, /*apply_decay=*/0, 1.0f, k.img, k.img);
            if (e) return e;
            CK(cudaMemcpyAsync(host_rows, k.img, band_px, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            return CLUMPY_OK;
        });
    std::string keep = rc ? clumpy_cuda_last_error() : "";
    cleanup();
    if (rc) set_error("%s", keep.c_str());
    return rc;
}
