// composite_exact.cu -- composite pass that reproduces the reference's hit ORDER (sparse clouds).
//
// The reference blends hits in point-index order (splat_points.cc:142-160).  Counting per texel loses
// that order; replaying ring by ring is within 1 LSB when hits are strong or many, but a pixel that
// receives a handful of hits of DIFFERENT opacity, some of them weak (the 0.028 ring of k = 5, 7),
// can end up 2 off.  For clouds that are sparse relative to the image -- README example6, the splat
// sweep at low density -- the order is cheap to keep: cells are 64-bit,
//
//      bits  0..23   number of points whose centre is this texel           (count)
//      bits 24..63   sum of their ORIGINAL indices, modulo 2^40            (index sum)
//
// updated with one 64-bit RED.ADD per point.  A cell holding one point knows that point's index
// exactly; a cell holding several keeps their mean.  The composite pass gathers the non-empty cells
// under a pixel's k x k window, sorts them by (mean) index and replays each cell's hits in that
// order.  When every contributing cell holds one point -- or its points are consecutive among the
// contributors, as duplicates in a point list are -- the pixel is bit-identical to the reference.
// Pixels with more than EXACT_MAX_CELLS contributing cells fall back to ring order (they are dense:
// many hits, which is where ring order converges).
#include "device_utils.cuh"
#include "internal.h"
#include "rings.cuh"

#include <cstring>

namespace clumpy {

constexpr int ETILE_W = 128, ETILE_H = 8, EXACT_THREADS = 256;
constexpr int EXACT_MAX_CELLS = 12;

struct ExactSprite {
    int k, nclass;
    float class_alpha[16];
    uint8_t tap_class[15 * 15]; // 255 = opacity 0, skipped
};

__device__ __forceinline__ uint32_t blend_n(uint32_t d, float a, uint32_t n) {
    const float one_minus_a = __fsub_rn(1.0f, a), a255 = __fmul_rn(a, 255.0f);
    for (uint32_t i = 0; i < n; ++i) {
        // splat_points.cc:153-155: two rounded products, a rounded sum, truncation to u8
        const uint32_t nd = f2u8_x86(__fadd_rn(__fmul_rn(one_minus_a, (float)d), a255));
        if (nd == d) break; // fixed point of a monotone map: the remaining hits change nothing
        d = nd;
    }
    return d;
}

template <bool WRAP>
__global__ void __launch_bounds__(EXACT_THREADS)
composite_exact_kernel(GeomDev g, ExactSprite sp, const unsigned long long* __restrict__ cells,
                       int apply_decay, float decay, const uint8_t* __restrict__ img_in,
                       uint8_t* __restrict__ img_out) {
    extern __shared__ unsigned long long etile[];
    const int h = sp.k / 2, rows = ETILE_H + 2 * h, cols = ETILE_W + 2 * h;
    const int tx0 = blockIdx.x * ETILE_W, ty0 = blockIdx.y * ETILE_H;
    unsigned long long any = 0ull;
    for (int idx = threadIdx.x; idx < rows * cols; idx += EXACT_THREADS) {
        const int r = idx / cols, c = idx - r * cols;
        int xc = tx0 + c; // padded column
        if (WRAP) {
            int x = (tx0 + c - h) % g.width;
            if (x < 0) x += g.width;
            xc = x + h;
        }
        const unsigned long long v = cells[(size_t)(ty0 + r) * g.pitch + xc];
        etile[idx] = v;
        any |= v;
    }
    const int tile_hit = __syncthreads_or(any != 0ull);

    const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
    const int y = ty0 + wrow;
    if (y >= g.height) return;
    for (int j = 0; j < 4; ++j) {
        const int xl = lane + 32 * j, x = tx0 + xl; // lanes take adjacent pixels: coalesced bytes
        if (x >= g.width) break;
        const size_t at = (size_t)y * g.width + x;
        uint32_t d = img_in[at];
        if (apply_decay) d = f2u8_x86(__fmul_rn((float)d, decay)); // advect_points.cc:155,174
        if (tile_hit) {
            // contributing cells, kept sorted by mean original index
            unsigned long long key[EXACT_MAX_CELLS];
            uint32_t cnt[EXACT_MAX_CELLS];
            uint8_t cls[EXACT_MAX_CELLS];
            int n = 0;
            bool overflow = false;
            for (int dy = 0; dy < sp.k && !overflow; ++dy)
                for (int dx = 0; dx < sp.k; ++dx) {
                    const unsigned long long c = etile[(wrow + dy) * cols + xl + dx];
                    if (c == 0ull) continue;
                    const uint32_t tc = sp.tap_class[dx + dy * sp.k];
                    if (tc == 255u) continue;
                    if (n == EXACT_MAX_CELLS) { overflow = true; break; }
                    const uint32_t cn = (uint32_t)(c & 0xFFFFFFull);
                    const unsigned long long kk = (c >> 24) / cn;
                    int pos = n++;
                    while (pos > 0 && key[pos - 1] > kk) { key[pos] = key[pos - 1]; cnt[pos] = cnt[pos - 1]; cls[pos] = cls[pos - 1]; --pos; }
                    key[pos] = kk; cnt[pos] = cn; cls[pos] = (uint8_t)tc;
                }
            if (!overflow) {
                for (int e = 0; e < n; ++e) d = blend_n(d, sp.class_alpha[cls[e]], cnt[e]);
            } else {
                // dense pixel: ring order (ascending distance), as the count-only paths do
                for (int c0 = 0; c0 < sp.nclass; ++c0) {
                    uint32_t total = 0;
                    for (int dy = 0; dy < sp.k; ++dy)
                        for (int dx = 0; dx < sp.k; ++dx)
                            if (sp.tap_class[dx + dy * sp.k] == c0)
                                total += (uint32_t)(etile[(wrow + dy) * cols + xl + dx] & 0xFFFFFFull);
                    if (total) d = blend_n(d, sp.class_alpha[c0], total);
                }
            }
        }
        img_out[at] = (uint8_t)d;
    }
}

// one 64-bit add per point of a plain (N, 2) list: count + original index
template <bool WRAP>
__global__ void __launch_bounds__(256)
exact_count_kernel(const float2* __restrict__ pts, uint32_t npts, uint32_t index_base, GeomDev g, unsigned long long* __restrict__ cells) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += gridDim.x * blockDim.x) {
        const float2 p = pts[i];
        const long long at = hist_index<WRAP>(f2i32_x86(p.x), f2i32_x86(p.y), g);
        if (at >= 0) atomicAdd(cells + at, ((unsigned long long)(index_base + i) << 24) | 1ull);
    }
}

bool exact_mode_supported(int kernel_size, uint64_t npts) { return kernel_size <= 15 && npts < (1ull << 24); }

int launch_exact_count(clumpy_cuda_ctx* ctx, const float* d_pts, uint32_t npts, const HistGeom& hg, int wrapx,
                       void* d_cells, uint32_t index_base) {
    if (npts == 0) return CLUMPY_OK;
    const GeomDev g{hg.width, hg.height, hg.halo, hg.pitch};
    const uint32_t grid = std::min<uint32_t>(ceil_div(npts, 256), (uint32_t)ctx->sm_count * 16);
    const float2* p = reinterpret_cast<const float2*>(d_pts);
    unsigned long long* c = reinterpret_cast<unsigned long long*>(d_cells);
    if (wrapx) exact_count_kernel<true><<<grid, 256, 0, ctx->stream>>>(p, npts, index_base, g, c);
    else exact_count_kernel<false><<<grid, 256, 0, ctx->stream>>>(p, npts, index_base, g, c);
    CK_LAUNCH(ctx);
    return CLUMPY_OK;
}

int launch_composite_exact(clumpy_cuda_ctx* ctx, const HistGeom& hg, const SpriteRings& rings,
                           const void* d_cells, int wrapx, int apply_decay, float decay,
                           const uint8_t* d_img_in, uint8_t* d_img_out) {
    if (hg.width == 0 || hg.height == 0) return CLUMPY_OK;
    if (rings.kernel_size > 15) { set_error("exact composite supports kernel sizes up to 15"); return CLUMPY_ERR_INVALID; }
    const GeomDev g{hg.width, hg.height, hg.halo, hg.pitch};
    ExactSprite sp;
    sp.k = rings.kernel_size;
    sp.nclass = rings.nclass;
    memcpy(sp.class_alpha, rings.class_alpha, sizeof(sp.class_alpha));
    memcpy(sp.tap_class, rings.tap_class, sizeof(sp.tap_class));
    const int h = sp.k / 2;
    const size_t smem = (size_t)(ETILE_H + 2 * h) * (ETILE_W + 2 * h) * sizeof(unsigned long long);
    const dim3 grid(ceil_div(hg.width, ETILE_W), ceil_div(hg.height, ETILE_H));
    const unsigned long long* c = reinterpret_cast<const unsigned long long*>(d_cells);
    if (wrapx) composite_exact_kernel<true><<<grid, EXACT_THREADS, smem, ctx->stream>>>(g, sp, c, apply_decay, decay, d_img_in, d_img_out);
    else composite_exact_kernel<false><<<grid, EXACT_THREADS, smem, ctx->stream>>>(g, sp, c, apply_decay, decay, d_img_in, d_img_out);
    CK_LAUNCH(ctx);
    return CLUMPY_OK;
}

} // namespace clumpy
