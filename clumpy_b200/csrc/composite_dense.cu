// composite_dense.cu -- the composite pass of advect_points for byte counters (see internal.h
// "dense path").  Per pixel: u8 decay (advect_points.cc:155,174), then for each sprite ring, in
// ascending distance, the reference's truncating blend (splat_points.cc:153-155) applied n times,
// n = number of points whose ring tap lands on the pixel -- as ONE lookup in T_r[min(n, depth_r)].
//
// Persistent kernel: one CTA per resident slot, each loads the replay tables into shared memory once
// and then walks 128 x 16 pixel tiles.  A tile's counters (+ halo) are staged in shared memory with
// aligned word loads; every thread owns 4 adjacent pixels of two rows, builds the ring sums for all 4
// at once in packed registers (byte lanes for k <= 3, 16-bit lanes for k = 5, 7) and does the table
// lookups.  All traffic to the counter image and the framebuffer is coalesced; both stay in L2
// between frames at the advect_points sizes (32 MiB each at 8192 x 4096).
#include "device_utils.cuh"
#include "internal.h"
#include "rings.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cstring>
#include <vector>

namespace clumpy {

constexpr int DTILE_W = 128;
constexpr int DTILE_H = 16;
constexpr int DENSE_THREADS = 256;

struct DenseParams {
    int nring, nrows;
    int depth[10];
    int row0[10];
};

// ---- host: replay tables ------------------------------------------------------------------------

int build_dense_tables(clumpy_cuda_ctx* ctx, const SpriteRings& rings, float decay, DenseTables* out) {
    *out = DenseTables();
    out->kernel_size = rings.kernel_size;
    const int k = rings.kernel_size;
    if (k > 7 || rings.nring > 10) return CLUMPY_OK; // not valid: caller falls back to u32 cells
    // ring multiplicities decide the lane width of the packed ring sums
    int max_mult = 1;
    {
        std::vector<int> mult(rings.nring, 0);
        const int h = k / 2;
        std::vector<int> uniq;
        for (int j = 0; j < k; ++j)
            for (int i = 0; i < k; ++i) uniq.push_back((i - h) * (i - h) + (j - h) * (j - h));
        std::vector<int> sorted(uniq);
        std::sort(sorted.begin(), sorted.end());
        sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
        for (int d2 : uniq) mult[std::lower_bound(sorted.begin(), sorted.end(), d2) - sorted.begin()]++;
        for (int m : mult) max_mult = std::max(max_mult, m);
    }
    std::vector<uint8_t> rows(256);
    for (int d = 0; d < 256; ++d) rows[d] = (uint8_t)decay_once_host((uint32_t)d, decay);
    int nrows = 1, max_depth = 0;
    out->nring = rings.nring;
    for (int r = 0; r < rings.nring; ++r) {
        const float a = rings.ring_alpha[r];
        out->row0[r] = nrows;
        out->depth[r] = 0;
        if (a == 0.0f) continue; // f_0 is the identity
        uint8_t prev[256], cur[256];
        for (int d = 0; d < 256; ++d) prev[d] = (uint8_t)d;
        rows.insert(rows.end(), prev, prev + 256); // row for n = 0: identity, so lookups need no branch
        int depth = 0;
        for (;;) {
            for (int d = 0; d < 256; ++d) cur[d] = (uint8_t)blend_once_host(prev[d], a);
            if (memcmp(cur, prev, 256) == 0) break; // stationary: more hits change nothing
            if (++depth > 255) return CLUMPY_OK;    // never settles within a byte counter: not valid
            rows.insert(rows.end(), cur, cur + 256);
            memcpy(prev, cur, 256);
        }
        out->depth[r] = depth;
        nrows += depth + 1;
        max_depth = std::max(max_depth, depth);
    }
    out->nrows = nrows;
    // cell cap: as large as the lanes allow, never below the deepest ring
    if (max_mult * 63 <= 255 && max_depth <= 63) { out->cell_cap = 63; out->byte_lanes = true; }
    else if (max_mult == 1) { out->cell_cap = 255; out->byte_lanes = true; }
    else { out->cell_cap = 255; out->byte_lanes = false; }
    if ((size_t)nrows * 256 > 40 * 1024) return CLUMPY_OK; // would not fit next to the tile in 48 KB
    CK(cudaMalloc(&out->d_rows, rows.size()));
    CK(cudaMemcpyAsync(out->d_rows, rows.data(), rows.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); // `rows` is a local
    out->valid = true;
    return CLUMPY_OK;
}

void free_dense_tables(DenseTables* t) {
    if (t->d_rows) cudaFree(t->d_rows);
    t->d_rows = nullptr;
    t->valid = false;
}

// ---- device ---------------------------------------------------------------------------------------

template <int K>
struct DenseShape {
    static constexpr int H = K / 2;
    static constexpr int NW = (4 + 2 * H + 3) / 4; // words one lane's window spans
    static constexpr int ROWW = 31 + NW;           // words per staged tile row
    static constexpr int ROWS = DTILE_H + 2 * H;
};

// bytes [dx, dx+4) of a window held in consecutive words w[0..NW)
template <int NW>
__device__ __forceinline__ uint32_t window_bytes(const uint32_t (&w)[NW + 1], int dx) {
    const int wi = dx >> 2, sh = (dx & 3) * 8;
    return sh ? __funnelshift_r(w[wi], w[wi + 1], sh) : w[wi];
}

template <int K, bool WRAP, bool BYTE_LANES>
__global__ void __launch_bounds__(DENSE_THREADS)
composite_dense_kernel(GeomDev g, DenseParams dp, const uint8_t* __restrict__ tables_g,
                       const uint8_t* __restrict__ cells, const uint8_t* __restrict__ img_in,
                       uint8_t* __restrict__ img_out, int tiles_x, int ntiles, int vec_ok) {
    using DS = DenseShape<K>;
    constexpr int H = DS::H;
    constexpr RingTable<K> RT{};
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* tab = smem;
    uint32_t* tile = reinterpret_cast<uint32_t*>(smem + (size_t)dp.nrows * 256);

    for (int i = threadIdx.x; i < dp.nrows * 64; i += DENSE_THREADS)
        reinterpret_cast<uint32_t*>(tab)[i] = reinterpret_cast<const uint32_t*>(tables_g)[i];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int ty = t / tiles_x, tx = t - ty * tiles_x;
        const int tx0 = tx * DTILE_W, ty0 = ty * DTILE_H;
        __syncthreads(); // tables loaded / previous tile fully consumed
        if (!WRAP) {
            for (int idx = threadIdx.x; idx < DS::ROWS * DS::ROWW; idx += DENSE_THREADS) {
                const int r = idx / DS::ROWW, c = idx - r * DS::ROWW;
                tile[idx] = __ldcg(reinterpret_cast<const uint32_t*>(cells + (size_t)(ty0 + r) * g.pitch + tx0) + c);
            }
        } else {
            uint8_t* tb = reinterpret_cast<uint8_t*>(tile);
            for (int idx = threadIdx.x; idx < DS::ROWS * DS::ROWW * 4; idx += DENSE_THREADS) {
                const int r = idx / (DS::ROWW * 4), c = idx - r * (DS::ROWW * 4);
                int x = (tx0 + c - H) % g.width; // image column of this tile byte, wrapped
                if (x < 0) x += g.width;
                tb[idx] = __ldcg(cells + (size_t)(ty0 + r) * g.pitch + x + H);
            }
        }
        __syncthreads();

#pragma unroll
        for (int half = 0; half < DTILE_H / 8; ++half) {
            const int row = warp + 8 * half;
            const int y = ty0 + row, x0 = tx0 + 4 * lane;
            if (y >= g.height || x0 >= g.width) continue;

            // packed ring sums of the 4 pixels: acc[r] (byte lanes) or acc[r], acc2[r] (16-bit lanes)
            uint32_t acc[RT.nring], acc2[RT.nring];
#pragma unroll
            for (int r = 0; r < RT.nring; ++r) { acc[r] = 0u; acc2[r] = 0u; }
            uint32_t any = 0u;
#pragma unroll
            for (int dy = 0; dy < K; ++dy) {
                uint32_t w[DS::NW + 1];
#pragma unroll
                for (int q = 0; q < DS::NW; ++q) w[q] = tile[(row + dy) * DS::ROWW + lane + q];
                w[DS::NW] = 0u;
#pragma unroll
                for (int dx = 0; dx < K; ++dx) {
                    const uint32_t s = window_bytes<DS::NW>(w, dx);
                    any |= s;
                    if (BYTE_LANES) {
                        acc[RT.ring[dx + dy * K]] += s;
                    } else {
                        acc[RT.ring[dx + dy * K]] += __byte_perm(s, 0u, 0x4140);  // pixels 0, 1
                        acc2[RT.ring[dx + dy * K]] += __byte_perm(s, 0u, 0x4342); // pixels 2, 3
                    }
                }
            }

            const size_t at = (size_t)y * g.width + x0;
            uint32_t px[4];
            const bool full = vec_ok && (x0 + 3 < g.width);
            if (full) {
                const uint32_t v = *reinterpret_cast<const uint32_t*>(img_in + at);
                px[0] = v & 0xFFu; px[1] = (v >> 8) & 0xFFu; px[2] = (v >> 16) & 0xFFu; px[3] = v >> 24;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) px[j] = (x0 + j < g.width) ? img_in[at + j] : 0u;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) px[j] = tab[px[j]]; // decay
            if (any) {
#pragma unroll
                for (int r = 0; r < RT.nring; ++r) {
                    const int depth = dp.depth[r];
                    if (depth == 0) continue;
                    const uint8_t* rows = tab + (size_t)dp.row0[r] * 256; // row n = f^n, row 0 = identity
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t n;
                        if (BYTE_LANES) n = (acc[r] >> (8 * j)) & 0xFFu;
                        else n = ((j < 2 ? acc[r] : acc2[r]) >> (16 * (j & 1))) & 0xFFFFu;
                        px[j] = rows[min(n, (uint32_t)depth) * 256 + px[j]];
                    }
                }
            }
            if (full) {
                *reinterpret_cast<uint32_t*>(img_out + at) = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (x0 + j < g.width) img_out[at + j] = (uint8_t)px[j];
            }
        }
    }
}

template <int K, bool WRAP, bool BYTE_LANES>
static int launch_dense_k(clumpy_cuda_ctx* ctx, const GeomDev& g, const DenseTables& t, const uint8_t* cells,
                          const uint8_t* in, uint8_t* out) {
    using DS = DenseShape<K>;
    DenseParams dp;
    dp.nring = t.nring;
    dp.nrows = t.nrows;
    for (int r = 0; r < 10; ++r) { dp.depth[r] = r < t.nring ? t.depth[r] : 0; dp.row0[r] = r < t.nring ? t.row0[r] : 0; }
    const size_t smem = (size_t)t.nrows * 256 + (size_t)DS::ROWS * DS::ROWW * 4;
    auto kern = composite_dense_kernel<K, WRAP, BYTE_LANES>;
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DENSE_THREADS, smem));
    if (per_sm < 1) { set_error("dense composite kernel does not fit on an SM"); return CLUMPY_ERR_CUDA; }
    const int tiles_x = (int)ceil_div(g.width, DTILE_W), tiles_y = (int)ceil_div(g.height, DTILE_H);
    const int ntiles = tiles_x * tiles_y;
    if (ctx->composite_ctas_per_sm > 0) per_sm = std::min(per_sm, ctx->composite_ctas_per_sm);
    const int grid = std::min(ntiles, per_sm * ctx->sm_count);
    const int vec_ok = (g.width % 4 == 0) && ((uintptr_t)in % 4 == 0) && ((uintptr_t)out % 4 == 0);
    kern<<<grid, DENSE_THREADS, smem, ctx->stream>>>(g, dp, t.d_rows, cells, in, out, tiles_x, ntiles, vec_ok);
    CK_LAUNCH(ctx);
    return CLUMPY_OK;
}

template <int K>
static int launch_dense_wrap(clumpy_cuda_ctx* ctx, const GeomDev& g, const DenseTables& t, const uint8_t* cells,
                             int wrapx, const uint8_t* in, uint8_t* out) {
    if (t.byte_lanes) {
        return wrapx ? launch_dense_k<K, true, true>(ctx, g, t, cells, in, out)
                     : launch_dense_k<K, false, true>(ctx, g, t, cells, in, out);
    }
    return wrapx ? launch_dense_k<K, true, false>(ctx, g, t, cells, in, out)
                 : launch_dense_k<K, false, false>(ctx, g, t, cells, in, out);
}

// ---- fp16 cells ---------------------------------------------------------------------------------
//
// Same pass for counters stored as IEEE half floats (two texels per word), which the advect kernel
// updates with a fire-and-forget red.global.add.noftz.f16x2.  A half counts exactly up to 2048 and
// then stays there (2048 + 1 rounds back to 2048), so it saturates on its own, never carries into its
// neighbour, and is exact far beyond the deepest replay table (88 rows).  Ring sums are formed with
// packed half adds (exact below 2048, monotone above, and anything >= depth selects the last row).

template <int K>
struct HalfShape {
    static constexpr int H = K / 2;
    static constexpr int NW = (4 + 2 * H + 1) / 2; // words (half2) one lane's window spans
    static constexpr int NW2 = (NW + 1) / 2 * 2;   // ... rounded up to whole 8-byte loads
    static constexpr int ROWW = 2 * 31 + NW2;      // words per staged tile row (even: rows stay 8-byte aligned)
    static constexpr int ROWS = DTILE_H + 2 * H;
};

template <int K, bool WRAP>
__global__ void __launch_bounds__(DENSE_THREADS)
composite_half_kernel(GeomDev g, DenseParams dp, const uint8_t* __restrict__ tables_g,
                      const __half* __restrict__ cells, const uint8_t* __restrict__ img_in,
                      uint8_t* __restrict__ img_out, int tiles_x, int ntiles, int vec_ok, int keep_cells) {
    using HS = HalfShape<K>;
    const uint64_t cell_policy = keep_cells ? make_keep_policy() : make_normal_policy();
    constexpr int H = HS::H;
    constexpr RingTable<K> RT{};
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* tab = smem;
    uint32_t* tile = reinterpret_cast<uint32_t*>(smem + (size_t)dp.nrows * 256);

    for (int i = threadIdx.x; i < dp.nrows * 64; i += DENSE_THREADS)
        reinterpret_cast<uint32_t*>(tab)[i] = reinterpret_cast<const uint32_t*>(tables_g)[i];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int LOADW = 2 * 31 + HS::NW; // words actually needed per row
    __half2 lim[RT.nring]; // replay depth of each ring, as the clamp for the packed half sums
#pragma unroll
    for (int r = 0; r < RT.nring; ++r) lim[r] = __float2half2_rn((float)dp.depth[r]);
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int ty = t / tiles_x, tx = t - ty * tiles_x;
        const int tx0 = tx * DTILE_W, ty0 = ty * DTILE_H;
        __syncthreads();
        if (!WRAP) {
            for (int r = warp; r < HS::ROWS; r += DENSE_THREADS / 32) {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(cells + (size_t)(ty0 + r) * g.pitch + tx0);
#pragma unroll
                for (int c = lane; c < LOADW; c += 32) tile[r * HS::ROWW + c] = ld_hint_u32(src + c, cell_policy);
            }
        } else {
            __half* th = reinterpret_cast<__half*>(tile);
            for (int idx = threadIdx.x; idx < HS::ROWS * LOADW * 2; idx += DENSE_THREADS) {
                const int r = idx / (LOADW * 2), c = idx - r * (LOADW * 2);
                int x = (tx0 + c - H) % g.width;
                if (x < 0) x += g.width;
                th[r * HS::ROWW * 2 + c] = __ldcg(cells + (size_t)(ty0 + r) * g.pitch + x + H);
            }
        }
        __syncthreads();

#pragma unroll
        for (int half = 0; half < DTILE_H / 8; ++half) {
            const int row = warp + 8 * half;
            const int y = ty0 + row, x0 = tx0 + 4 * lane;
            if (y >= g.height || x0 >= g.width) continue;

            __half2 accA[RT.nring], accB[RT.nring]; // pixels (0,1) and (2,3)
#pragma unroll
            for (int r = 0; r < RT.nring; ++r) { accA[r] = __float2half2_rn(0.f); accB[r] = accA[r]; }
            uint32_t any = 0u;
#pragma unroll
            for (int dy = 0; dy < K; ++dy) {
                uint32_t w[HS::NW2 + 1];
#pragma unroll
                for (int q = 0; q < HS::NW2; q += 2) { // 8-byte shared loads: conflict-free at stride 2
                    const uint2 v = *reinterpret_cast<const uint2*>(&tile[(row + dy) * HS::ROWW + 2 * lane + q]);
                    w[q] = v.x; w[q + 1] = v.y;
                }
#pragma unroll
                for (int q = 0; q < HS::NW; ++q) any |= w[q];
                w[HS::NW2] = 0u;
                // pair(c) = cells (c, c+1) of the window as a half2
                auto pair = [&](int c) -> __half2 {
                    const uint32_t v = (c & 1) ? __funnelshift_r(w[c >> 1], w[(c >> 1) + 1], 16) : w[c >> 1];
                    return *reinterpret_cast<const __half2*>(&v);
                };
#pragma unroll
                for (int dx = 0; dx < K; ++dx) {
                    accA[RT.ring[dx + dy * K]] = __hadd2(accA[RT.ring[dx + dy * K]], pair(dx));
                    accB[RT.ring[dx + dy * K]] = __hadd2(accB[RT.ring[dx + dy * K]], pair(dx + 2));
                }
            }

            const size_t at = (size_t)y * g.width + x0;
            uint32_t px[4];
            const bool full = vec_ok && (x0 + 3 < g.width);
            if (full) {
                const uint32_t v = *reinterpret_cast<const uint32_t*>(img_in + at);
                px[0] = v & 0xFFu; px[1] = (v >> 8) & 0xFFu; px[2] = (v >> 16) & 0xFFu; px[3] = v >> 24;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) px[j] = (x0 + j < g.width) ? img_in[at + j] : 0u;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) px[j] = tab[px[j]]; // decay
            if (any) {
#pragma unroll
                for (int r = 0; r < RT.nring; ++r) {
                    if (dp.depth[r] == 0) continue;
                    const uint8_t* rows = tab + (size_t)dp.row0[r] * 256; // row n = f^n, row 0 = identity
                    const __half2 a = __hmin2(accA[r], lim[r]), b = __hmin2(accB[r], lim[r]);
                    const uint32_t n[4] = {__half2uint_rz(__low2half(a)), __half2uint_rz(__high2half(a)),
                                           __half2uint_rz(__low2half(b)), __half2uint_rz(__high2half(b))};
#pragma unroll
                    for (int j = 0; j < 4; ++j) px[j] = rows[n[j] * 256 + px[j]];
                }
            }
            if (full) {
                *reinterpret_cast<uint32_t*>(img_out + at) = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (x0 + j < g.width) img_out[at + j] = (uint8_t)px[j];
            }
        }
    }
}

template <int K, bool WRAP>
static int launch_half_k(clumpy_cuda_ctx* ctx, const GeomDev& g, const DenseTables& t, const __half* cells,
                         const uint8_t* in, uint8_t* out) {
    using HS = HalfShape<K>;
    DenseParams dp;
    dp.nring = t.nring;
    dp.nrows = t.nrows;
    for (int r = 0; r < 10; ++r) { dp.depth[r] = r < t.nring ? t.depth[r] : 0; dp.row0[r] = r < t.nring ? t.row0[r] : 0; }
    const size_t smem = (size_t)t.nrows * 256 + (size_t)HS::ROWS * HS::ROWW * 4;
    auto kern = composite_half_kernel<K, WRAP>;
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, DENSE_THREADS, smem));
    if (per_sm < 1) { set_error("half-cell composite kernel does not fit on an SM"); return CLUMPY_ERR_CUDA; }
    const int tiles_x = (int)ceil_div(g.width, DTILE_W), tiles_y = (int)ceil_div(g.height, DTILE_H);
    const int ntiles = tiles_x * tiles_y;
    if (ctx->composite_ctas_per_sm > 0) per_sm = std::min(per_sm, ctx->composite_ctas_per_sm);
    const int grid = std::min(ntiles, per_sm * ctx->sm_count);
    const int vec_ok = (g.width % 4 == 0) && ((uintptr_t)in % 4 == 0) && ((uintptr_t)out % 4 == 0);
    kern<<<grid, DENSE_THREADS, smem, ctx->stream>>>(g, dp, t.d_rows, cells, in, out, tiles_x, ntiles, vec_ok, ctx->keep_cells);
    CK_LAUNCH(ctx);
    return CLUMPY_OK;
}

int launch_composite_half(clumpy_cuda_ctx* ctx, const HistGeom& hg, const DenseTables& t,
                          const void* d_cells, int wrapx, const uint8_t* d_img_in, uint8_t* d_img_out) {
    if (hg.width == 0 || hg.height == 0) return CLUMPY_OK;
    const GeomDev g{hg.width, hg.height, hg.halo, hg.pitch};
    const __half* c = reinterpret_cast<const __half*>(d_cells);
    switch (t.kernel_size) {
    case 1: return wrapx ? launch_half_k<1, true>(ctx, g, t, c, d_img_in, d_img_out) : launch_half_k<1, false>(ctx, g, t, c, d_img_in, d_img_out);
    case 3: return wrapx ? launch_half_k<3, true>(ctx, g, t, c, d_img_in, d_img_out) : launch_half_k<3, false>(ctx, g, t, c, d_img_in, d_img_out);
    case 5: return wrapx ? launch_half_k<5, true>(ctx, g, t, c, d_img_in, d_img_out) : launch_half_k<5, false>(ctx, g, t, c, d_img_in, d_img_out);
    case 7: return wrapx ? launch_half_k<7, true>(ctx, g, t, c, d_img_in, d_img_out) : launch_half_k<7, false>(ctx, g, t, c, d_img_in, d_img_out);
    default: set_error("half-cell composite supports kernel sizes 1, 3, 5, 7"); return CLUMPY_ERR_INVALID;
    }
}

int launch_composite_dense(clumpy_cuda_ctx* ctx, const HistGeom& hg, const DenseTables& t,
                           const uint8_t* d_cells, int wrapx, const uint8_t* d_img_in, uint8_t* d_img_out) {
    if (hg.width == 0 || hg.height == 0) return CLUMPY_OK;
    const GeomDev g{hg.width, hg.height, hg.halo, hg.pitch};
    switch (t.kernel_size) {
    case 1: return launch_dense_wrap<1>(ctx, g, t, d_cells, wrapx, d_img_in, d_img_out);
    case 3: return launch_dense_wrap<3>(ctx, g, t, d_cells, wrapx, d_img_in, d_img_out);
    case 5: return launch_dense_wrap<5>(ctx, g, t, d_cells, wrapx, d_img_in, d_img_out);
    case 7: return launch_dense_wrap<7>(ctx, g, t, d_cells, wrapx, d_img_in, d_img_out);
    default: set_error("dense composite supports kernel sizes 1, 3, 5, 7"); return CLUMPY_ERR_INVALID;
    }
}

} // namespace clumpy
