// internal.h -- shared host-side plumbing of libclumpy_cuda.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "clumpy_cuda.h"

#define CLUMPY_API extern "C" __attribute__((visibility("default")))

namespace clumpy {

void set_error(const char* fmt, ...);

// Returns CLUMPY_ERR_CUDA from the enclosing function when a runtime call fails.
#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            ::clumpy::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),       \
                                __FILE__, __LINE__);                                          \
            return (e_ == cudaErrorMemoryAllocation) ? CLUMPY_ERR_NOMEM : CLUMPY_ERR_CUDA;    \
        }                                                                                     \
    } while (0)

#define CK_LAUNCH(ctx)                                                                        \
    do {                                                                                      \
        (ctx)->launches++;                                                                    \
        CK(cudaGetLastError());                                                               \
    } while (0)

#define REQUIRE(cond, msg)                                                                    \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            ::clumpy::set_error("%s", msg);                                                   \
            return CLUMPY_ERR_INVALID;                                                        \
        }                                                                                     \
    } while (0)

static inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
static inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

} // namespace clumpy

struct clumpy_cuda_ctx {
    int device = 0;
    int sm_count = 0;
    size_t l2_bytes = 0;
    cudaStream_t stream = nullptr;      // compute stream (caller's or ours)
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr; // frame D2H side stream
    cudaStream_t aux_stream = nullptr;  // high-priority stream: composite + clear of frame s run here
                                        // while the advect kernel of frame s+1 runs on `stream`
    cudaStream_t aux2_stream = nullptr; // second high-priority stream: the multi-GPU path drains inboxes on
                                        // aux_stream and composites on this one, so that the drain of
                                        // frame s+1 overlaps the composite of frame s
    int keep_cells = 0;                 // counters are accessed with L2 evict-last hints
    int composite_ctas_per_sm = 0;      // >0: cap on resident composite CTAs per SM (leaves room for
                                        // the advect kernel when the two overlap); 0 = fill the SM
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    uint64_t launches = 0;
    uint32_t* d_minmax = nullptr; // 2 order-preserving encoded floats (min, max)
    uint32_t* h_minmax = nullptr; // pinned mirror
};

namespace clumpy {

// ---- min / max of a float stream, reduced on the device ------------------------------------
int minmax_reset(clumpy_cuda_ctx* ctx);
int minmax_fetch(clumpy_cuda_ctx* ctx, float* minval, float* maxval); // synchronises

// ---- splat accumulator ("histogram" of point centres) and the composite pass ---------------
//
// A splat of npts points with a k x k sprite is done in two steps (DESIGN.md "splat"):
//   1. count how many points have their centre texel at each pixel (one atomic per point) into a
//      padded image `hist` of (H + 2h + slack) x pitch counters, h = k/2;
//   2. one pass over the framebuffer: u8 decay, then for every sprite ring (set of taps with the
//      same opacity) replay the reference's truncating blend as many times as points sit at the
//      corresponding neighbouring centres; write the u8 pixel.
struct HistGeom {
    int width = 0, height = 0; // image
    int halo = 0;              // h = kernel_size / 2
    int pitch = 0;             // counters per padded row
    int rows = 0;              // padded rows
    size_t cells() const { return (size_t)pitch * rows; }
    size_t bytes() const { return cells() * sizeof(uint32_t); }
};
HistGeom make_hist_geom(uint32_t width, uint32_t height, int kernel_size);

// Opacity classes of the sprite (splat_points.cc:126-136) in replay order.
struct SpriteRings {
    int kernel_size;
    int nring;             // distinct squared tap distances of the k x k sprite
    float ring_alpha[16];  // opacity by distance rank (first 16 ranks; all of them for k <= 7)
    int nclass;            // distinct non-zero opacities, in ascending-distance order
    float class_alpha[16];
    uint8_t tap_class[CLUMPY_MAX_KERNEL_SIZE * CLUMPY_MAX_KERNEL_SIZE]; // per tap (i + j*k); 255 = skip
};
int make_sprite_rings(int kernel_size, float alpha, SpriteRings* out);

// ---- dense path: u8 saturating counters + replay tables (composite_dense.cu) ------------------
//
// For the sprites advect_points uses (opacity 1) every ring's truncating blend f_a becomes
// stationary after a small number of applications for ALL 256 start values (12 for k = 3, 88 at
// most).  So a counter only has to count up to that depth: cells are bytes that saturate at
// `cell_cap` (packed 4 per word, updated with a compare-and-swap loop), the counter image is W*H
// bytes and stays resident in L2, and the composite pass replaces the replay loop by one lookup per
// ring in a table of iterated blends T_r[n][d] = f_a^n(d) held in shared memory.
struct DenseTables {
    bool valid = false;      // false: some ring needs more than cell_cap applications -> use u32 cells
    int kernel_size = 0;
    int nring = 0;
    int depth[16] = {0};     // number of table rows of ring r (0: ring has opacity 0, skipped)
    int row0[16] = {0};      // first row of ring r; row 0 of the block is the decay table
    int nrows = 0;
    uint32_t cell_cap = 255; // saturation value of a cell
    bool byte_lanes = false; // ring sums of 4 pixels fit the 4 bytes of one register
    uint8_t* d_rows = nullptr; // nrows * 256 bytes, device
};
int build_dense_tables(clumpy_cuda_ctx* ctx, const SpriteRings& rings, float decay, DenseTables* out);
void free_dense_tables(DenseTables* t);
int launch_composite_dense(clumpy_cuda_ctx* ctx, const HistGeom& g, const DenseTables& t,
                           const uint8_t* d_cells, int wrapx, const uint8_t* d_img_in,
                           uint8_t* d_img_out);

// same pass for counters stored as half floats (fire-and-forget red.add.f16x2 on the advect side)
int launch_composite_half(clumpy_cuda_ctx* ctx, const HistGeom& g, const DenseTables& t,
                          const void* d_cells, int wrapx, const uint8_t* d_img_in, uint8_t* d_img_out);

// ---- exact-order path for sparse clouds (composite_exact.cu): 64-bit cells = count + index sum ----
bool exact_mode_supported(int kernel_size, uint64_t npts);
int launch_exact_count(clumpy_cuda_ctx* ctx, const float* d_pts, uint32_t npts, const HistGeom& g,
                       int wrapx, void* d_cells);
int launch_composite_exact(clumpy_cuda_ctx* ctx, const HistGeom& g, const SpriteRings& rings,
                           const void* d_cells, int wrapx, int apply_decay, float decay,
                           const uint8_t* d_img_in, uint8_t* d_img_out);

// counts AoS points (x, y) into hist
int launch_hist_points(clumpy_cuda_ctx* ctx, const float* d_pts, uint32_t npts, const HistGeom& g,
                       int wrapx, uint32_t* d_hist);
// img_out[p] = replay(decay(img_in[p]), counts around p); in-place allowed
int launch_composite(clumpy_cuda_ctx* ctx, const HistGeom& g, const SpriteRings& rings,
                     const uint32_t* d_hist, int wrapx, int apply_decay, float decay,
                     const uint8_t* d_img_in, uint8_t* d_img_out);

} // namespace clumpy
