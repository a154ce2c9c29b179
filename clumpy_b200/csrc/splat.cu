// splat.cu -- point splatting (clumpy `splat_points` / splat_disks) for sm_100a.
//
// The reference (commands/splat_points.cc:118-161) walks the points in index order and, for each,
// read-modify-writes k*k u8 pixels with a truncating src-over blend.  Done literally on a GPU that is
// k*k contended byte atomics per point.  Here the work is split so that each point costs ONE
// integer atomic and every pixel is written once:
//
//   count pass      hist[centre texel of point] += 1                (launch_hist_points, or fused
//                                                                     into the advect kernel)
//   composite pass  for every pixel p:  d = u8 framebuffer value, optionally decayed
//                     for each sprite ring r (taps of equal opacity a_r, ascending distance):
//                        n_r = sum of hist at the centres whose ring-r tap lands on p
//                        d   = f_a(f_a(...f_a(d))) n_r times,  f_a(d) = (u8)((1-a)*d + a*255)
//                     write d
//
// f_a is the reference's per-hit blend (:153-155) with its two roundings and the truncation, so a
// pixel hit by taps of a single opacity is reproduced bit for bit (k = 1 always).  When several
// opacities hit the same pixel in one frame the reference interleaves them in point order; replaying
// ring by ring differs by at most 1 LSB (measured; tests/test_gpu_parity.py checks it).
// f_a is monotone on 0..255, so the replay stops as soon as it reaches a fixed point -- at most 255
// iterations however many points pile up on one pixel.
//
// `hist` is a padded image of uint32 counters: (H + 2h [+slack]) rows of `pitch` counters, texel
// (x, y) at [(y+h)*pitch + (x+h)], so centres up to h texels outside the image (whose sprites still
// reach it) are counted without bounds tests in the composite pass.  In x-wrap mode centres are
// reduced mod W first and the composite pass reads its halo columns mod W.
#include "device_utils.cuh"
#include "internal.h"
#include "rings.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace clumpy {

constexpr int TILE_W = 128; // pixels per CTA row: 32 lanes x 4 pixels
constexpr int TILE_H = 8;   // rows per CTA: one warp per row
constexpr int COMPOSITE_THREADS = 256;

// ---- host: geometry and sprite opacity classes ----------------------------------------------

HistGeom make_hist_geom(uint32_t width, uint32_t height, int kernel_size) {
    HistGeom g;
    g.width = (int)width;
    g.height = (int)height;
    g.halo = kernel_size / 2;
    g.pitch = (int)(round_up(width, TILE_W) + round_up(2 * g.halo, 4) + 4);
    g.rows = (int)(round_up(height, 16) + 2 * g.halo); // 16 = tallest composite tile
    return g;
}

// glm detail/func_common.inl:562-568
static float smoothstep_host(float e0, float e1, float x) {
    float t = (x - e0) / (e1 - e0);
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
    return t * t * (3.0f - 2.0f * t);
}

// splat_points.cc:126-136.  Opacity depends on the tap's squared distance d2 only and does not
// increase with it, so classes are numbered in ascending d2 = descending opacity, merging equal
// neighbours; taps with opacity exactly 0 leave the pixel unchanged and are skipped.
int make_sprite_rings(int kernel_size, float alpha, SpriteRings* out) {
    REQUIRE(kernel_size > 0 && kernel_size % 2 == 1, "Kernel size must be an odd integer.");
    REQUIRE(kernel_size <= CLUMPY_MAX_KERNEL_SIZE, "kernel size exceeds CLUMPY_MAX_KERNEL_SIZE");
    memset(out, 0, sizeof(*out));
    out->kernel_size = kernel_size;
    const int k = kernel_size, middle = k / 2;
    const float r2 = (float)(middle * middle);
    std::vector<int> d2s;
    for (int j = 0; j < k; ++j)
        for (int i = 0; i < k; ++i) d2s.push_back((i - middle) * (i - middle) + (j - middle) * (j - middle));
    std::vector<int> uniq(d2s);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    // distance rank -> opacity (used directly by the unrolled kernels for k <= 7)
    out->nring = (int)uniq.size();
    std::vector<float> rank_alpha(uniq.size());
    for (size_t r = 0; r < uniq.size(); ++r) {
        rank_alpha[r] = alpha * smoothstep_host(r2 + 5.0f, r2 - 5.0f, (float)uniq[r]);
        if (r < 16) out->ring_alpha[r] = rank_alpha[r];
    }
    // merged classes for the generic kernel
    std::vector<int> rank_class(uniq.size(), 255);
    int nclass = 0;
    for (size_t r = 0; r < uniq.size(); ++r) {
        const float a = rank_alpha[r];
        if (a == 0.0f) continue;
        if (nclass > 0 && out->class_alpha[nclass - 1] == a) { rank_class[r] = nclass - 1; continue; }
        if (nclass >= 16) { set_error("sprite has more than 16 opacity classes"); return CLUMPY_ERR_INVALID; }
        out->class_alpha[nclass] = a;
        rank_class[r] = nclass++;
    }
    out->nclass = nclass;
    for (int t = 0; t < k * k; ++t) {
        const size_t r = std::lower_bound(uniq.begin(), uniq.end(), d2s[t]) - uniq.begin();
        out->tap_class[t] = (uint8_t)rank_class[r];
    }
    return CLUMPY_OK;
}

// ---- device helpers ---------------------------------------------------------------------------

static GeomDev to_dev(const HistGeom& g) { return GeomDev{g.width, g.height, g.halo, g.pitch}; }

// One hit of splat_points.cc:153-155 with a = opacity: two products and a sum, each rounded
// (the reference binary has no FMA), truncated to u8 the way cvttss2si + byte store does.
__device__ __forceinline__ uint32_t blend_once(uint32_t d, float one_minus_a, float a255) {
    return f2u8_x86(__fadd_rn(__fmul_rn(one_minus_a, (float)d), a255));
}

__device__ __forceinline__ uint32_t replay(uint32_t d, float a, uint32_t n) {
    if (n == 0u || a == 0.0f) return d;
    const float one_minus_a = __fsub_rn(1.0f, a), a255 = __fmul_rn(a, 255.0f);
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t nd = blend_once(d, one_minus_a, a255);
        if (nd == d) break; // fixed point: further hits change nothing
        d = nd;
    }
    return d;
}

// Replays n[r] hits of opacity alpha[r], r < NR, INTERLEAVED in proportion: at step s the ring that
// is furthest behind its share s * n[r] / total goes next.  As far as one pixel is concerned the
// reference's point order is a random interleaving of the rings; ring-by-ring replay (all strong hits
// first, weak ones last) is the extreme case and accumulates the most truncation bias when the
// opacities are weak, the proportional interleave is the typical one.  With a single active ring this
// is the plain replay (bit-exact).  Stops when every ring still owed hits leaves the value unchanged.
template <int NR>
__device__ __forceinline__ uint32_t replay_mixed(uint32_t d, const float* __restrict__ alpha, const uint32_t (&n)[NR]) {
    uint32_t rem[NR], total = 0u, active = 0u;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        rem[r] = (alpha[r] != 0.0f) ? n[r] : 0u;
        total += rem[r];
        if (rem[r]) active |= 1u << r;
    }
    if (total == 0u) return d;
    if ((active & (active - 1u)) == 0u) { // one ring only
#pragma unroll
        for (int r = 0; r < NR; ++r)
            if (active == (1u << r)) d = replay(d, alpha[r], rem[r]);
        return d;
    }
    const float inv = 1.0f / (float)total;
    float share[NR], done[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) { share[r] = (float)rem[r] * inv; done[r] = 0.0f; }
    uint32_t stalled = 0u;
    for (uint32_t s = 1u; active; ++s) {
        int best = 0;
        float best_score = -3.0e38f;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const float score = fmaf((float)s, share[r], -done[r]);
            if (((active >> r) & 1u) && score > best_score) { best_score = score; best = r; }
        }
        float a = 0.0f;
#pragma unroll
        for (int r = 0; r < NR; ++r)
            if (best == r) { a = alpha[r]; done[r] += 1.0f; if (--rem[r] == 0u) active &= ~(1u << r); }
        const uint32_t nd = blend_once(d, __fsub_rn(1.0f, a), __fmul_rn(a, 255.0f));
        if (nd == d) stalled |= 1u << best; else { d = nd; stalled = 0u; }
        if ((active & ~stalled) == 0u) break;
    }
    return d;
}

// advect_points.cc:155,174: `v *= decay` on a uint8_t&
__device__ __forceinline__ uint32_t decay_once(uint32_t d, float decay) {
    return f2u8_x86(__fmul_rn((float)d, decay));
}

// ---- count pass for a plain (N, 2) point list -------------------------------------------------

template <bool WRAP>
__global__ void __launch_bounds__(256)
hist_points_kernel(const float2* __restrict__ pts, uint32_t npts, GeomDev g, uint32_t* __restrict__ hist) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += gridDim.x * blockDim.x) {
        const float2 p = pts[i];
        const long long at = hist_index<WRAP>(f2i32_x86(p.x), f2i32_x86(p.y), g);
        if (at >= 0) atomicAdd(hist + at, 1u);
    }
}

int launch_hist_points(clumpy_cuda_ctx* ctx, const float* d_pts, uint32_t npts, const HistGeom& g,
                       int wrapx, uint32_t* d_hist) {
    if (npts == 0) return CLUMPY_OK;
    const uint32_t grid = std::min<uint32_t>(ceil_div(npts, 256), (uint32_t)ctx->sm_count * 16);
    const float2* p = reinterpret_cast<const float2*>(d_pts);
    if (wrapx) hist_points_kernel<true><<<grid, 256, 0, ctx->stream>>>(p, npts, to_dev(g), d_hist);
    else hist_points_kernel<false><<<grid, 256, 0, ctx->stream>>>(p, npts, to_dev(g), d_hist);
    CK_LAUNCH(ctx);
    return CLUMPY_OK;
}

// ---- composite pass -----------------------------------------------------------------------------

struct RingAlpha { float a[16]; };

// Shared tile of counters: (TILE_H + 2h) rows, TILE_W + 2h columns rounded up for 16-byte reads.
template <int K> struct TileShape {
    static constexpr int H = K / 2;
    static constexpr int ROWS = TILE_H + 2 * H;
    static constexpr int WIN = 4 + 2 * H;           // columns one thread's 4 pixels look at
    static constexpr int WIN4 = (WIN + 3) / 4;      // in uint4 units
    static constexpr int COLS = 4 * 31 + 4 * WIN4;  // last lane's window end
};

// Fills the shared tile from the padded counter image.  Non-wrap: aligned uint4 loads (pitch and the
// tile origin are multiples of 4).  Wrap: halo columns are fetched modulo W, element by element.
template <int ROWS, int COLS, bool WRAP>
__device__ __forceinline__ void load_tile(uint32_t (*tile)[COLS], const uint32_t* __restrict__ hist,
                                          const GeomDev& g, int tx0, int ty0, int halo) {
    const int tid = threadIdx.x;
    if (!WRAP) {
        constexpr int C4 = COLS / 4;
        for (int idx = tid; idx < ROWS * C4; idx += COMPOSITE_THREADS) {
            const int r = idx / C4, c4 = idx - r * C4;
            const uint4 v = *reinterpret_cast<const uint4*>(hist + (size_t)(ty0 + r) * g.pitch + tx0 + 4 * c4);
            *reinterpret_cast<uint4*>(&tile[r][4 * c4]) = v;
        }
    } else {
        for (int idx = tid; idx < ROWS * COLS; idx += COMPOSITE_THREADS) {
            const int r = idx / COLS, c = idx - r * COLS;
            int x = (tx0 + c - halo) % g.width; // image column of this tile column, wrapped
            if (x < 0) x += g.width;
            tile[r][c] = hist[(size_t)(ty0 + r) * g.pitch + x + halo];
        }
    }
}

template <int K, bool WRAP>
__global__ void __launch_bounds__(COMPOSITE_THREADS)
composite_kernel(GeomDev g, RingAlpha ra, const uint32_t* __restrict__ hist, int apply_decay,
                 float decay, const uint8_t* __restrict__ img_in, uint8_t* __restrict__ img_out,
                 int vec_ok) {
    using TS = TileShape<K>;
    constexpr int H = TS::H;
    constexpr RingTable<K> RT{};
    __shared__ __align__(16) uint32_t tile[TS::ROWS][TS::COLS];

    const int tx0 = blockIdx.x * TILE_W, ty0 = blockIdx.y * TILE_H;
    load_tile<TS::ROWS, TS::COLS, WRAP>(tile, hist, g, tx0, ty0, H);
    __syncthreads();

    const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
    const int y = ty0 + wrow, x0 = tx0 + 4 * lane;
    if (y >= g.height || x0 >= g.width) return;

    // ring counts for the 4 pixels of this thread
    uint32_t cnt[4][RT.nring];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int r = 0; r < RT.nring; ++r) cnt[j][r] = 0u;
#pragma unroll
    for (int dy = 0; dy < K; ++dy) {
        uint32_t win[4 * TS::WIN4];
#pragma unroll
        for (int q = 0; q < TS::WIN4; ++q) {
            const uint4 v = *reinterpret_cast<const uint4*>(&tile[wrow + dy][4 * lane + 4 * q]);
            win[4 * q] = v.x; win[4 * q + 1] = v.y; win[4 * q + 2] = v.z; win[4 * q + 3] = v.w;
        }
        // pixel j sits at window column j + H; the centre at offset (dx-H, dy-H) from it hits it
        // with the mirrored tap, which has the same distance and therefore the same ring.
#pragma unroll
        for (int dx = 0; dx < K; ++dx)
#pragma unroll
            for (int j = 0; j < 4; ++j) cnt[j][RT.ring[dx + dy * K]] += win[j + dx];
    }

    const size_t at = (size_t)y * g.width + x0;
    uint32_t px[4];
    const bool full = vec_ok && (x0 + 3 < g.width);
    if (full) {
        const uchar4 v = *reinterpret_cast<const uchar4*>(img_in + at);
        px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) px[j] = (x0 + j < g.width) ? img_in[at + j] : 0u;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t d = px[j];
        if (apply_decay) d = decay_once(d, decay);
        px[j] = replay_mixed<RT.nring>(d, ra.a, cnt[j]);
    }
    if (full) {
        *reinterpret_cast<uchar4*>(img_out + at) = make_uchar4((uint8_t)px[0], (uint8_t)px[1], (uint8_t)px[2], (uint8_t)px[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (x0 + j < g.width) img_out[at + j] = (uint8_t)px[j];
    }
}

// Any odd kernel size up to CLUMPY_MAX_KERNEL_SIZE: taps and classes come from the parameter block,
// the tile lives in dynamic shared memory.  Correct but unoptimised; k <= 7 never comes here.
struct GenericSprite {
    int k, nclass;
    float class_alpha[16];
    uint8_t tap_class[CLUMPY_MAX_KERNEL_SIZE * CLUMPY_MAX_KERNEL_SIZE];
};

template <bool WRAP>
__global__ void __launch_bounds__(COMPOSITE_THREADS)
composite_generic_kernel(GeomDev g, GenericSprite sp, const uint32_t* __restrict__ hist,
                         int apply_decay, float decay, const uint8_t* __restrict__ img_in,
                         uint8_t* __restrict__ img_out) {
    extern __shared__ uint32_t tile_dyn[];
    const int h = sp.k / 2, rows = TILE_H + 2 * h, cols = TILE_W + 2 * h;
    const int tx0 = blockIdx.x * TILE_W, ty0 = blockIdx.y * TILE_H;
    for (int idx = threadIdx.x; idx < rows * cols; idx += COMPOSITE_THREADS) {
        const int r = idx / cols, c = idx - r * cols;
        int xc = tx0 + c; // padded column
        if (WRAP) {
            int x = (tx0 + c - h) % g.width;
            if (x < 0) x += g.width;
            xc = x + h;
        }
        tile_dyn[idx] = hist[(size_t)(ty0 + r) * g.pitch + xc];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
    const int y = ty0 + wrow;
    if (y >= g.height) return;
    for (int j = 0; j < 4; ++j) {
        const int xl = 4 * lane + j, x = tx0 + xl;
        if (x >= g.width) break;
        uint32_t cnt[16];
        for (int c = 0; c < 16; ++c) cnt[c] = 0u;
        for (int dy = 0; dy < sp.k; ++dy)
            for (int dx = 0; dx < sp.k; ++dx) {
                const uint32_t cls = sp.tap_class[dx + dy * sp.k];
                if (cls < 16u) cnt[cls] += tile_dyn[(wrow + dy) * cols + xl + dx];
            }
        const size_t at = (size_t)y * g.width + x;
        uint32_t d = img_in[at];
        if (apply_decay) d = decay_once(d, decay);
        img_out[at] = (uint8_t)replay_mixed<16>(d, sp.class_alpha, cnt); // unused classes have opacity 0
    }
}

template <int K>
static void launch_composite_k(cudaStream_t s, dim3 grid, const GeomDev& g, const SpriteRings& rings,
                               const uint32_t* hist, int wrapx, int apply_decay, float decay,
                               const uint8_t* in, uint8_t* out, int vec_ok) {
    RingAlpha ra;
    for (int i = 0; i < 16; ++i) ra.a[i] = rings.ring_alpha[i];
    if (wrapx) composite_kernel<K, true><<<grid, COMPOSITE_THREADS, 0, s>>>(g, ra, hist, apply_decay, decay, in, out, vec_ok);
    else composite_kernel<K, false><<<grid, COMPOSITE_THREADS, 0, s>>>(g, ra, hist, apply_decay, decay, in, out, vec_ok);
}

int launch_composite(clumpy_cuda_ctx* ctx, const HistGeom& hg, const SpriteRings& rings,
                     const uint32_t* d_hist, int wrapx, int apply_decay, float decay,
                     const uint8_t* d_img_in, uint8_t* d_img_out) {
    if (hg.width == 0 || hg.height == 0) return CLUMPY_OK;
    const GeomDev g = to_dev(hg);
    const dim3 grid(ceil_div(hg.width, TILE_W), ceil_div(hg.height, TILE_H));
    const int vec_ok = (hg.width % 4 == 0) && ((uintptr_t)d_img_in % 4 == 0) && ((uintptr_t)d_img_out % 4 == 0);
    cudaStream_t s = ctx->stream;
    switch (rings.kernel_size) {
    case 1: launch_composite_k<1>(s, grid, g, rings, d_hist, wrapx, apply_decay, decay, d_img_in, d_img_out, vec_ok); break;
    case 3: launch_composite_k<3>(s, grid, g, rings, d_hist, wrapx, apply_decay, decay, d_img_in, d_img_out, vec_ok); break;
    case 5: launch_composite_k<5>(s, grid, g, rings, d_hist, wrapx, apply_decay, decay, d_img_in, d_img_out, vec_ok); break;
    case 7: launch_composite_k<7>(s, grid, g, rings, d_hist, wrapx, apply_decay, decay, d_img_in, d_img_out, vec_ok); break;
    default: {
        GenericSprite sp;
        sp.k = rings.kernel_size;
        sp.nclass = rings.nclass;
        memcpy(sp.class_alpha, rings.class_alpha, sizeof(sp.class_alpha));
        memcpy(sp.tap_class, rings.tap_class, sizeof(sp.tap_class));
        const int h = sp.k / 2;
        const size_t smem = (size_t)(TILE_H + 2 * h) * (TILE_W + 2 * h) * sizeof(uint32_t);
        if (wrapx) composite_generic_kernel<true><<<grid, COMPOSITE_THREADS, smem, s>>>(g, sp, d_hist, apply_decay, decay, d_img_in, d_img_out);
        else composite_generic_kernel<false><<<grid, COMPOSITE_THREADS, smem, s>>>(g, sp, d_hist, apply_decay, decay, d_img_in, d_img_out);
    }
    }
    CK_LAUNCH(ctx);
    return CLUMPY_OK;
}

// ---- k = 1, alpha = 1 fast path of the splat_points command -----------------------------------

// splat_points.cc:108-116: the store is idempotent, so racing writers are harmless.
__global__ void __launch_bounds__(256)
mark_points_kernel(const float2* __restrict__ pts, uint32_t npts, uint32_t width, uint32_t height,
                   uint8_t* __restrict__ img) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += gridDim.x * blockDim.x) {
        const float2 p = pts[i];
        const uint32_t x = f2u32_x86(p.x), y = f2u32_x86(p.y);
        if (x < width && y < height) img[(size_t)width * y + x] = 255;
    }
}

} // namespace clumpy

using namespace clumpy;

namespace {
// Device copies of the caller's points and image for the one-shot commands.
struct SplatBuffers {
    float* pts = nullptr;
    uint8_t* img = nullptr;
    uint32_t* hist = nullptr;
    ~SplatBuffers() { cudaFree(pts); cudaFree(img); cudaFree(hist); }
};
} // namespace

CLUMPY_API int clumpy_cuda_splat_points(clumpy_cuda_ctx* ctx, const float* pts_host, uint32_t npts,
                                        uint32_t width, uint32_t height, uint8_t* img_host) {
    REQUIRE(ctx && img_host && (pts_host || npts == 0), "null argument");
    REQUIRE(width > 0 && height > 0, "image dimensions must be positive");
    CK(cudaSetDevice(ctx->device));
    SplatBuffers b;
    const size_t npix = (size_t)width * height;
    CK(cudaMalloc(&b.pts, std::max<size_t>(8, (size_t)npts * 8)));
    CK(cudaMalloc(&b.img, npix));
    CK(cudaMemcpyAsync(b.pts, pts_host, (size_t)npts * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b.img, img_host, npix, cudaMemcpyHostToDevice, ctx->stream));
    if (npts) {
        const uint32_t grid = std::min<uint32_t>(ceil_div(npts, 256), (uint32_t)ctx->sm_count * 16);
        mark_points_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const float2*>(b.pts), npts, width, height, b.img);
        CK_LAUNCH(ctx);
    }
    CK(cudaMemcpyAsync(img_host, b.img, npix, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_splat_disks(clumpy_cuda_ctx* ctx, const float* pts_host, uint32_t npts,
                                       uint32_t width, uint32_t height, uint8_t* img_host, float alpha,
                                       int kernel_size, int wrapx) {
    REQUIRE(ctx && img_host && (pts_host || npts == 0), "null argument");
    REQUIRE(width > 0 && height > 0, "image dimensions must be positive");
    REQUIRE(width < (1u << 30) && height < (1u << 30), "image dimensions too large");
    SpriteRings rings;
    int rc = make_sprite_rings(kernel_size, alpha, &rings);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const HistGeom g = make_hist_geom(width, height, kernel_size);
    SplatBuffers b;
    const size_t npix = (size_t)width * height;
    // Sparse list (8 * npts <= pixels): 64-bit cells that keep the reference's hit order
    // (composite_exact.cu).  Dense list: 32-bit counts, ring-order replay.  CLUMPY_CUDA_CELLS=u32 /
    // exact forces one of the two.
    bool exact = exact_mode_supported(kernel_size, npts) && 8ull * npts <= (uint64_t)npix;
    if (const char* force = getenv("CLUMPY_CUDA_CELLS")) {
        if (!strcmp(force, "u32")) exact = false;
        if (!strcmp(force, "exact")) exact = exact_mode_supported(kernel_size, npts);
    }
    const size_t hist_bytes = exact ? g.cells() * 8 : g.bytes();
    CK(cudaMalloc(&b.pts, std::max<size_t>(8, (size_t)npts * 8)));
    CK(cudaMalloc(&b.img, npix));
    CK(cudaMalloc(&b.hist, hist_bytes));
    CK(cudaMemcpyAsync(b.pts, pts_host, (size_t)npts * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b.img, img_host, npix, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(b.hist, 0, hist_bytes, ctx->stream));
    if (exact) {
        rc = launch_exact_count(ctx, b.pts, npts, g, wrapx, b.hist);
        if (rc) return rc;
        rc = launch_composite_exact(ctx, g, rings, b.hist, wrapx, /*apply_decay=*/0, 1.0f, b.img, b.img);
    } else {
        rc = launch_hist_points(ctx, b.pts, npts, g, wrapx, b.hist);
        if (rc) return rc;
        rc = launch_composite(ctx, g, rings, b.hist, wrapx, /*apply_decay=*/0, 1.0f, b.img, b.img);
    }
    if (rc) return rc;
    CK(cudaMemcpyAsync(img_host, b.img, npix, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CLUMPY_OK;
}
