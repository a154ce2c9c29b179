// advect.cu -- clumpy `advect_points` for sm_100a: device-resident particle state, one fused
// Euler-step + reset + splat-count kernel per simulation frame, then the composite pass of splat.cu.
//
// Replaces commands/advect_points.cc:119-179.  Facts the kernels rely on:
//   * velocity lookup is NEAREST texel (advect_points.cc:27-36), index conversion via int64;
//   * pt += step * v is a rounded multiply then a rounded add (the reference binary has no FMA);
//   * the age / age_offset bookkeeping (:124-135,147-152,166-170) is equivalent to
//         "point i returns to its original position at the end of simulation frame s
//          iff s % nframes == age_offset_i",
//     so no per-point age is stored;
//   * warm-up frames with decay == 0 neither decay nor splat (:154);
//   * trajectories never depend on the framebuffer.
//
// Layout in HBM (SoA, all sized npts unless noted):
//   pos     float2   current positions            read + written every frame (streaming)
//   orig    float2   positions to reset to        read only when a point resets (1/nframes of them)
//   offset  uint32   reset phase of each point    read every frame (streaming)
//   perm    uint32   original index of each slot  only for read_points
//   vel     float2   (H, W) velocity field        random 8-byte gathers, kept L2-friendly by the sort
//   hist    uint32   padded splat counters        one RED per point, read + cleared by the composite
//   img[2]  uint8    framebuffer (ping-pong only while a frame copy is in flight)
//
// Points are bucket-sorted ONCE at set-up by the image tile of their original position.  Neighbouring
// points start together and, the field being smooth, travel together; resets send them home.  So a
// warp's velocity gathers and counter updates stay within a few cache lines for the whole run
// without any per-frame binning.  Results do not depend on the order (integer counts commute;
// trajectories are per point); read_points undoes the permutation.
#include "device_utils.cuh"
#include "internal.h"

#include <algorithm>
#include <cstdlib>
#include <random>
#include <vector>

namespace clumpy {

constexpr int ADVECT_THREADS = 256;
constexpr int ADVECT_PTS = 2; // points per thread: one 16-byte load/store of positions

// ---- the per-frame kernel -----------------------------------------------------------------------

// advect_points.cc:23-37 + :146
template <bool WRAP>
__device__ __forceinline__ float2 euler_step(float2 p, const float2* __restrict__ vel, uint32_t width,
                                             uint32_t height, float step) {
    uint32_t x = f2u32_x86(p.x);
    const uint32_t y = f2u32_x86(p.y);
    if (WRAP) x = (width + (x % width)) % width;
    float2 v = make_float2(0.f, 0.f);
    if (x < width && y < height) v = __ldg(vel + ((size_t)y * width + x));
    return make_float2(__fadd_rn(p.x, __fmul_rn(step, v.x)), __fadd_rn(p.y, __fmul_rn(step, v.y)));
}

template <bool WRAP, bool SPLAT>
__global__ void __launch_bounds__(ADVECT_THREADS)
advect_kernel(float2* __restrict__ pos, const float2* __restrict__ orig,
              const uint32_t* __restrict__ offset, uint32_t npts, const float2* __restrict__ vel,
              uint32_t width, uint32_t height, float step, uint32_t phase, GeomDev g,
              uint32_t* __restrict__ hist) {
    const uint32_t i0 = (blockIdx.x * ADVECT_THREADS + threadIdx.x) * ADVECT_PTS;
    if (i0 >= npts) return;
    float2 p[ADVECT_PTS];
    uint32_t off[ADVECT_PTS];
    const bool both = i0 + 1 < npts;
    const uint64_t pol = make_stream_policy();
    if (both) {
        const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(pos + i0), pol);
        p[0] = make_float2(v.x, v.y);
        p[1] = make_float2(v.z, v.w);
        const uint2 o = ld_stream_u2(reinterpret_cast<const uint2*>(offset + i0), pol);
        off[0] = o.x; off[1] = o.y;
    } else {
        p[0] = pos[i0];
        p[1] = p[0];
        off[0] = offset[i0];
        off[1] = 0xFFFFFFFFu;
    }
#pragma unroll
    for (int k = 0; k < ADVECT_PTS; ++k) {
        p[k] = euler_step<WRAP>(p[k], vel, width, height, step);
        if (off[k] == phase) p[k] = orig[i0 + k]; // reset AFTER the move (:147-152,166-170)
    }
    if (both) st_stream_f4(reinterpret_cast<float4*>(pos + i0), make_float4(p[0].x, p[0].y, p[1].x, p[1].y), pol);
    else pos[i0] = p[0];
    if (SPLAT) {
#pragma unroll
        for (int k = 0; k < ADVECT_PTS; ++k) {
            if (k == 1 && !both) break;
            // splat_points.cc:143-144: centre texel by (int32_t) truncation
            const long long at = hist_index<WRAP>(f2i32_x86(p[k].x), f2i32_x86(p[k].y), g);
            if (at >= 0) atomicAdd(hist + at, 1u);
        }
    }
}

// ---- one-time bucket sort by home tile ------------------------------------------------------------

struct SortGrid {
    int shift_x, shift_y; // tile = 2^shift_x x 2^shift_y texels
    uint32_t tiles_x, ntiles;
    uint32_t width, height;
};

__device__ __forceinline__ uint32_t home_tile(float2 p, const SortGrid& sg) {
    int32_t x = f2i32_x86(p.x), y = f2i32_x86(p.y);
    x = max(0, min(x, (int32_t)sg.width - 1));
    y = max(0, min(y, (int32_t)sg.height - 1));
    return (uint32_t)(y >> sg.shift_y) * sg.tiles_x + (uint32_t)(x >> sg.shift_x);
}

__global__ void __launch_bounds__(256)
sort_count_kernel(const float2* __restrict__ pts, uint32_t npts, SortGrid sg, uint32_t* __restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npts) atomicAdd(counts + home_tile(pts[i], sg), 1u);
}

constexpr int SCAN_BLOCK = 1024;

// exclusive scan of up to SCAN_BLOCK values per block; block totals to `sums`
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_local_kernel(uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t warp_tot[32];
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t v = i < n ? data[i] : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
        warp_tot[lane] = w; // inclusive totals
    }
    __syncthreads();
    const uint32_t base = warp ? warp_tot[warp - 1] : 0u;
    if (i < n) data[i] = base + inc - v;
    if (threadIdx.x == SCAN_BLOCK - 1) sums[blockIdx.x] = base + inc;
}

// single block: exclusive scan of `n` block totals in place, any n
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_sums_kernel(uint32_t* __restrict__ sums, uint32_t n) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base0 = 0; base0 < n; base0 += SCAN_BLOCK) {
        const uint32_t i = base0 + threadIdx.x;
        const uint32_t v = i < n ? sums[i] : 0u;
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
        const uint32_t carry = carry_s;
        const uint32_t pre = (warp ? warp_tot[warp - 1] : 0u) + carry;
        if (i < n) sums[i] = pre + inc - v;
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) carry_s = pre + inc;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_BLOCK)
scan_add_kernel(uint32_t* __restrict__ data, uint32_t n, const uint32_t* __restrict__ sums) {
    const uint32_t i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < n) data[i] += sums[blockIdx.x];
}

// cursors[] holds the exclusive scan; each point claims the next slot of its tile
__global__ void __launch_bounds__(256)
sort_scatter_kernel(const float2* __restrict__ pts, const uint32_t* __restrict__ offs, uint32_t npts,
                    SortGrid sg, uint32_t* __restrict__ cursors, float2* __restrict__ pos,
                    float2* __restrict__ orig, uint32_t* __restrict__ offset, uint32_t* __restrict__ perm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const float2 p = pts[i];
    const uint32_t slot = atomicAdd(cursors + home_tile(p, sg), 1u);
    pos[slot] = p;
    orig[slot] = p;
    offset[slot] = offs[i];
    perm[slot] = i;
}

__global__ void __launch_bounds__(256)
identity_order_kernel(const float2* __restrict__ pts, const uint32_t* __restrict__ offs, uint32_t npts,
                      float2* __restrict__ pos, float2* __restrict__ orig, uint32_t* __restrict__ offset,
                      uint32_t* __restrict__ perm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const float2 p = pts[i];
    pos[i] = p; orig[i] = p; offset[i] = offs[i]; perm[i] = i;
}

__global__ void __launch_bounds__(256)
unpermute_kernel(const float2* __restrict__ pos, const uint32_t* __restrict__ perm, uint32_t npts,
                 float2* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npts) out[perm[i]] = pos[i];
}

} // namespace clumpy

using namespace clumpy;

struct clumpy_advect {
    clumpy_cuda_ctx* ctx = nullptr;
    uint32_t npts = 0, width = 0, height = 0, nframes = 0, simframe = 0;
    int kernel_size = 1, wrapx = 0;
    float step_size = 0.f, decay = 0.f;
    HistGeom geom;
    SpriteRings rings;
    float2 *pos = nullptr, *orig = nullptr, *vel = nullptr;
    uint32_t *offset = nullptr, *perm = nullptr, *hist = nullptr;
    uint8_t* img[2] = {nullptr, nullptr};
    int cur = 0;                                  // which img holds the current framebuffer
    bool copy_pending[2] = {false, false};        // a frame D2H was issued from img[b]
    cudaEvent_t copy_done[2] = {nullptr, nullptr};
    cudaEvent_t frame_ready = nullptr;
};

static void advect_release(clumpy_advect* a) {
    if (!a) return;
    cudaFree(a->pos); cudaFree(a->orig); cudaFree(a->vel); cudaFree(a->offset); cudaFree(a->perm);
    cudaFree(a->hist); cudaFree(a->img[0]); cudaFree(a->img[1]);
    for (int b = 0; b < 2; ++b) if (a->copy_done[b]) cudaEventDestroy(a->copy_done[b]);
    if (a->frame_ready) cudaEventDestroy(a->frame_ready);
    delete a;
}

// advect_points.cc:127-135.  The distribution's algorithm is libstdc++'s; the reference and this
// library both take it from the host's <random>.
CLUMPY_API int clumpy_cuda_age_offsets(uint32_t nframes, uint32_t npts, uint32_t* out) {
    REQUIRE(out || npts == 0, "out is NULL");
    REQUIRE(nframes > 0, "nframes must be positive");
    std::mt19937 generator;
    generator.seed(0);
    std::uniform_int_distribution<> get_age_offset(0, (int)(nframes - 1));
    for (uint32_t i = 0; i < npts; ++i) out[i] = (uint32_t)get_age_offset(generator);
    return CLUMPY_OK;
}

static int sort_points(clumpy_advect* a, const float2* d_in, const uint32_t* d_offs) {
    clumpy_cuda_ctx* ctx = a->ctx;
    const uint32_t n = a->npts;
    if (n == 0) return CLUMPY_OK;
    const uint32_t pgrid = ceil_div(n, 256);
    const char* no_sort = getenv("CLUMPY_CUDA_NO_SORT");
    if (no_sort && no_sort[0] == '1') {
        identity_order_kernel<<<pgrid, 256, 0, ctx->stream>>>(d_in, d_offs, n, a->pos, a->orig, a->offset, a->perm);
        CK_LAUNCH(ctx);
        return CLUMPY_OK;
    }
    SortGrid sg;
    sg.width = a->width; sg.height = a->height;
    sg.shift_x = 3; sg.shift_y = 2; // 8 x 4 texels: one 64-byte run of the velocity field per row
    auto tiles = [&](int sx, int sy) {
        return (uint64_t)ceil_div(a->width, 1u << sx) * ceil_div(a->height, 1u << sy);
    };
    while (tiles(sg.shift_x, sg.shift_y) > (1u << 21)) {
        if (sg.shift_y < sg.shift_x) ++sg.shift_y; else ++sg.shift_x;
    }
    sg.tiles_x = ceil_div(a->width, 1u << sg.shift_x);
    sg.ntiles = (uint32_t)tiles(sg.shift_x, sg.shift_y);
    const uint32_t nblocks = ceil_div(sg.ntiles, SCAN_BLOCK);
    uint32_t *counts = nullptr, *sums = nullptr;
    CK(cudaMalloc(&counts, (size_t)sg.ntiles * 4));
    if (cudaMalloc(&sums, (size_t)nblocks * 4) != cudaSuccess) { cudaFree(counts); set_error("out of device memory"); return CLUMPY_ERR_NOMEM; }
    cudaMemsetAsync(counts, 0, (size_t)sg.ntiles * 4, ctx->stream);
    sort_count_kernel<<<pgrid, 256, 0, ctx->stream>>>(d_in, n, sg, counts);
    ctx->launches++;
    scan_local_kernel<<<nblocks, SCAN_BLOCK, 0, ctx->stream>>>(counts, sg.ntiles, sums);
    ctx->launches++;
    scan_sums_kernel<<<1, SCAN_BLOCK, 0, ctx->stream>>>(sums, nblocks);
    ctx->launches++;
    scan_add_kernel<<<nblocks, SCAN_BLOCK, 0, ctx->stream>>>(counts, sg.ntiles, sums);
    ctx->launches++;
    sort_scatter_kernel<<<pgrid, 256, 0, ctx->stream>>>(d_in, d_offs, n, sg, counts, a->pos, a->orig, a->offset, a->perm);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(counts);
    cudaFree(sums);
    if (e != cudaSuccess) { set_error("point sort failed: %s", cudaGetErrorString(e)); return CLUMPY_ERR_CUDA; }
    return CLUMPY_OK;
}

static int advect_begin_impl(clumpy_advect* a, const float* pts_host, const float* vel_host,
                             const uint32_t* age_offset) {
    clumpy_cuda_ctx* ctx = a->ctx;
    const size_t n = a->npts, npix = (size_t)a->width * a->height;
    const size_t nalloc = std::max<size_t>(n + 1, 2); // slack so that 16-byte accesses stay inside
    CK(cudaMalloc(&a->pos, nalloc * 8));
    CK(cudaMalloc(&a->orig, nalloc * 8));
    CK(cudaMalloc(&a->offset, nalloc * 4));
    CK(cudaMalloc(&a->perm, nalloc * 4));
    CK(cudaMalloc(&a->vel, npix * 8));
    CK(cudaMalloc(&a->hist, a->geom.bytes()));
    CK(cudaMalloc(&a->img[0], npix));
    CK(cudaMalloc(&a->img[1], npix));
    for (int b = 0; b < 2; ++b) CK(cudaEventCreateWithFlags(&a->copy_done[b], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&a->frame_ready, cudaEventDisableTiming));
    CK(cudaMemsetAsync(a->hist, 0, a->geom.bytes(), ctx->stream));
    CK(cudaMemsetAsync(a->img[0], 0, npix, ctx->stream)); // advect_points.cc:119: zero-initialised
    CK(cudaMemcpyAsync(a->vel, vel_host, npix * 8, cudaMemcpyHostToDevice, ctx->stream));

    std::vector<uint32_t> offs_own;
    if (!age_offset) {
        offs_own.resize(n);
        int rc = clumpy_cuda_age_offsets(a->nframes, a->npts, offs_own.data());
        if (rc) return rc;
        age_offset = offs_own.data();
    }
    float2* d_in = nullptr;
    uint32_t* d_offs = nullptr;
    CK(cudaMalloc(&d_in, nalloc * 8));
    if (cudaMalloc(&d_offs, nalloc * 4) != cudaSuccess) { cudaFree(d_in); set_error("out of device memory"); return CLUMPY_ERR_NOMEM; }
    cudaMemcpyAsync(d_in, pts_host, n * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_offs, age_offset, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    int rc = sort_points(a, d_in, d_offs);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in);
    cudaFree(d_offs);
    if (rc) return rc;
    if (e != cudaSuccess) { set_error("advect set-up failed: %s", cudaGetErrorString(e)); return CLUMPY_ERR_CUDA; }
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_begin(clumpy_cuda_ctx* ctx, const float* pts_host, uint32_t npts,
                                        const float* vel_host, uint32_t width, uint32_t height,
                                        float step_size, int kernel_size, float decay, int wrapx,
                                        uint32_t nframes, const uint32_t* age_offset,
                                        clumpy_advect** out) {
    REQUIRE(out, "out is NULL");
    *out = nullptr;
    REQUIRE(ctx && vel_host && (pts_host || npts == 0), "null argument");
    REQUIRE(width > 0 && height > 0, "image dimensions must be positive");
    REQUIRE(width < (1u << 30) && height < (1u << 30) && (uint64_t)width * height < (1ull << 32),
            "velocity field too large");
    REQUIRE(nframes > 0, "nframes must be positive");
    REQUIRE(decay >= 0.0f || decay != decay, "decay must be non-negative (pass wrapx for the negative-decay mode)");
    SpriteRings rings;
    // advect_points.cc:156,175: alpha is the constant 1.0f
    int rc = make_sprite_rings(kernel_size, 1.0f, &rings);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    clumpy_advect* a = new (std::nothrow) clumpy_advect();
    if (!a) { set_error("out of host memory"); return CLUMPY_ERR_NOMEM; }
    a->ctx = ctx; a->npts = npts; a->width = width; a->height = height; a->nframes = nframes;
    a->kernel_size = kernel_size; a->wrapx = wrapx ? 1 : 0; a->step_size = step_size; a->decay = decay;
    a->geom = make_hist_geom(width, height, kernel_size);
    a->rings = rings;
    rc = advect_begin_impl(a, pts_host, vel_host, age_offset);
    if (rc) { advect_release(a); return rc; }
    *out = a;
    return CLUMPY_OK;
}

template <bool WRAP, bool SPLAT>
static void launch_advect(clumpy_advect* a, uint32_t phase) {
    const uint32_t grid = ceil_div(a->npts, ADVECT_THREADS * ADVECT_PTS);
    const GeomDev g{a->geom.width, a->geom.height, a->geom.halo, a->geom.pitch};
    advect_kernel<WRAP, SPLAT><<<grid, ADVECT_THREADS, 0, a->ctx->stream>>>(
        a->pos, a->orig, a->offset, a->npts, a->vel, a->width, a->height, a->step_size, phase, g, a->hist);
}

// ev (optional): 4 events recorded before the advect kernel, after it, after the composite kernel and
// after the counter clear, for clumpy_cuda_advect_profile.
static int advect_one_frame(clumpy_advect* a, cudaEvent_t* ev = nullptr) {
    clumpy_cuda_ctx* ctx = a->ctx;
    const uint32_t s = a->simframe;
    const bool warm = s < a->nframes;
    const bool splat = !(warm && a->decay == 0.0f); // advect_points.cc:154
    const uint32_t phase = s % a->nframes;
    if (ev) CK(cudaEventRecord(ev[0], ctx->stream));
    if (a->npts) {
        if (a->wrapx) { if (splat) launch_advect<true, true>(a, phase); else launch_advect<true, false>(a, phase); }
        else          { if (splat) launch_advect<false, true>(a, phase); else launch_advect<false, false>(a, phase); }
        CK_LAUNCH(ctx);
    }
    if (ev) CK(cudaEventRecord(ev[1], ctx->stream));
    if (splat) {
        // write in place unless a frame copy is still reading the current buffer
        int dst = a->cur;
        if (a->copy_pending[a->cur]) {
            dst = a->cur ^ 1;
            if (a->copy_pending[dst]) {
                CK(cudaStreamWaitEvent(ctx->stream, a->copy_done[dst], 0));
                a->copy_pending[dst] = false;
            }
        }
        int rc = launch_composite(ctx, a->geom, a->rings, a->hist, a->wrapx, 1, a->decay, a->img[a->cur], a->img[dst]);
        if (rc) return rc;
        a->cur = dst;
        if (ev) CK(cudaEventRecord(ev[2], ctx->stream));
        CK(cudaMemsetAsync(a->hist, 0, a->geom.bytes(), ctx->stream));
    } else if (ev) {
        CK(cudaEventRecord(ev[2], ctx->stream));
    }
    if (ev) CK(cudaEventRecord(ev[3], ctx->stream));
    a->simframe = s + 1;
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_step(clumpy_advect* adv, uint32_t count) {
    REQUIRE(adv, "adv is NULL");
    CK(cudaSetDevice(adv->ctx->device));
    for (uint32_t i = 0; i < count; ++i) {
        int rc = advect_one_frame(adv);
        if (rc) return rc;
    }
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_profile(clumpy_advect* adv, uint32_t count, float* advect_ms,
                                          float* composite_ms, float* clear_ms) {
    REQUIRE(adv && count > 0, "bad argument");
    CK(cudaSetDevice(adv->ctx->device));
    std::vector<cudaEvent_t> ev(4 * (size_t)count, nullptr);
    int rc = CLUMPY_OK;
    for (auto& e : ev)
        if (cudaEventCreate(&e) != cudaSuccess) { set_error("event creation failed"); rc = CLUMPY_ERR_CUDA; break; }
    for (uint32_t i = 0; rc == CLUMPY_OK && i < count; ++i) rc = advect_one_frame(adv, &ev[4 * (size_t)i]);
    double t[3] = {0, 0, 0};
    if (rc == CLUMPY_OK && cudaStreamSynchronize(adv->ctx->stream) != cudaSuccess) { set_error("profile run failed"); rc = CLUMPY_ERR_CUDA; }
    for (uint32_t i = 0; rc == CLUMPY_OK && i < count; ++i)
        for (int k = 0; k < 3; ++k) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[4 * (size_t)i + k], ev[4 * (size_t)i + k + 1]);
            t[k] += ms;
        }
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (advect_ms) *advect_ms = (float)(t[0] / count);
    if (composite_ms) *composite_ms = (float)(t[1] / count);
    if (clear_ms) *clear_ms = (float)(t[2] / count);
    return rc;
}

CLUMPY_API int clumpy_cuda_advect_simframe(clumpy_advect* adv, uint32_t* simframe) {
    REQUIRE(adv && simframe, "null argument");
    *simframe = adv->simframe;
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_frame_async(clumpy_advect* adv, uint8_t* dst_host) {
    REQUIRE(adv && dst_host, "null argument");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    const int b = adv->cur;
    CK(cudaEventRecord(adv->frame_ready, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, adv->frame_ready, 0));
    CK(cudaMemcpyAsync(dst_host, adv->img[b], (size_t)adv->width * adv->height, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CK(cudaEventRecord(adv->copy_done[b], ctx->copy_stream));
    adv->copy_pending[b] = true;
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_wait_frames(clumpy_advect* adv) {
    REQUIRE(adv, "adv is NULL");
    CK(cudaSetDevice(adv->ctx->device));
    CK(cudaStreamSynchronize(adv->ctx->copy_stream));
    adv->copy_pending[0] = adv->copy_pending[1] = false;
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_read_points(clumpy_advect* adv, float* out_host) {
    REQUIRE(adv && (out_host || adv->npts == 0), "null argument");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    if (adv->npts == 0) return CLUMPY_OK;
    float2* tmp = nullptr;
    CK(cudaMalloc(&tmp, (size_t)adv->npts * 8));
    unpermute_kernel<<<ceil_div(adv->npts, 256), 256, 0, ctx->stream>>>(adv->pos, adv->perm, adv->npts, tmp);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, tmp, (size_t)adv->npts * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) { set_error("read_points failed: %s", cudaGetErrorString(e)); return CLUMPY_ERR_CUDA; }
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_read_image(clumpy_advect* adv, uint8_t* out_host) {
    REQUIRE(adv && out_host, "null argument");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out_host, adv->img[adv->cur], (size_t)adv->width * adv->height, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_end(clumpy_advect* adv) {
    if (!adv) return CLUMPY_OK;
    cudaSetDevice(adv->ctx->device);
    cudaStreamSynchronize(adv->ctx->stream);
    cudaStreamSynchronize(adv->ctx->copy_stream);
    advect_release(adv);
    return CLUMPY_OK;
}

// The whole command: advect_points.cc:137-179.
CLUMPY_API int clumpy_cuda_advect_points(clumpy_cuda_ctx* ctx, const float* pts_host, uint32_t npts,
                                         const float* vel_host, uint32_t width, uint32_t height,
                                         float step_size, int kernel_size, float decay, int wrapx,
                                         uint32_t nframes, const uint32_t* age_offset,
                                         clumpy_frame_fn on_frame, void* user) {
    REQUIRE(on_frame, "on_frame is NULL");
    clumpy_advect* a = nullptr;
    int rc = clumpy_cuda_advect_begin(ctx, pts_host, npts, vel_host, width, height, step_size,
                                      kernel_size, decay, wrapx, nframes, age_offset, &a);
    if (rc) return rc;
    const size_t npix = (size_t)width * height;
    uint8_t* pinned[2] = {(uint8_t*)clumpy_cuda_host_alloc(npix), (uint8_t*)clumpy_cuda_host_alloc(npix)};
    cudaEvent_t landed[2] = {nullptr, nullptr};
    auto cleanup = [&]() {
        clumpy_cuda_advect_end(a);
        for (int b = 0; b < 2; ++b) { clumpy_cuda_host_free(pinned[b]); if (landed[b]) cudaEventDestroy(landed[b]); }
    };
    if (!pinned[0] || !pinned[1]) { cleanup(); return CLUMPY_ERR_NOMEM; }
    for (int b = 0; b < 2; ++b)
        if (cudaEventCreateWithFlags(&landed[b], cudaEventDisableTiming) != cudaSuccess) { cleanup(); set_error("event creation failed"); return CLUMPY_ERR_CUDA; }

    // Warm-up frames, then for each recorded frame: simulate, start its copy, and only then hand the
    // PREVIOUS frame to the caller, so the callback (file writing) overlaps the GPU and the copy.
    rc = clumpy_cuda_advect_step(a, nframes);
    long long pending = -1;
    for (uint32_t f = 0; rc == CLUMPY_OK && f < nframes; ++f) {
        rc = clumpy_cuda_advect_step(a, 1);
        if (rc) break;
        rc = clumpy_cuda_advect_frame_async(a, pinned[f & 1]);
        if (rc) break;
        if (cudaEventRecord(landed[f & 1], ctx->copy_stream) != cudaSuccess) { set_error("event record failed"); rc = CLUMPY_ERR_CUDA; break; }
        if (pending >= 0) {
            if (cudaEventSynchronize(landed[pending & 1]) != cudaSuccess) { set_error("frame copy failed"); rc = CLUMPY_ERR_CUDA; break; }
            if (on_frame(user, (uint32_t)pending, pinned[pending & 1]) != 0) { set_error("frame callback aborted the run"); rc = CLUMPY_ERR_STATE; break; }
        }
        pending = f;
    }
    if (rc == CLUMPY_OK && pending >= 0) {
        if (cudaEventSynchronize(landed[pending & 1]) != cudaSuccess) { set_error("frame copy failed"); rc = CLUMPY_ERR_CUDA; }
        else if (on_frame(user, (uint32_t)pending, pinned[pending & 1]) != 0) { set_error("frame callback aborted the run"); rc = CLUMPY_ERR_STATE; }
    }
    cleanup();
    return rc;
}
