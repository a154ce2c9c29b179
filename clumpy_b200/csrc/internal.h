// internal.h -- shared host-side plumbing of libclumpy_cuda.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

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

// Device memory of a command comes from the device's stream-ordered pool (cudaMallocAsync on the context's
// stream; clumpy_cuda_create raises the pool's release threshold so that what one command frees the next one
// gets back without going to the driver: allocation + release were 50 ms of a 200 ms advect_points run).
// Memory that other processes / devices map (the multi-GPU inbox) still comes from cudaMalloc.
#define DEV_ALLOC(ctx, pp, bytes) CK(cudaMallocAsync((void**)(pp), (bytes), (ctx)->stream))
static inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
static inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

} // namespace clumpy

struct clumpy_cuda_ctx {
    int device = 0;
    int sm_count = 0;
    size_t l2_bytes = 0;
    cudaStream_t stream = nullptr;      // compute stream (caller's or ours)
    bool own_stream = false;
    uint32_t* mt_checkpoints_dev = nullptr; // mt19937 states + per-stretch scratch of the reset-offset generator (age_offsets_dev.cu)
    cudaStream_t copy_stream = nullptr; // frame D2H side stream
    cudaStream_t aux_stream = nullptr;  // high-priority side stream: the receiving half of a multi-GPU group (drain +
                                        // band composite), and the composite work of a single-GPU group when asked
                                        // to overlap with the next advect launch
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    uint64_t launches = 0;
    uint32_t* d_minmax = nullptr; // 2 order-preserving encoded floats (min, max)
    uint32_t* h_minmax = nullptr; // pinned mirror
    // Pinned host buffers handed back by finished commands, kept for the next one (pinning 64 MB costs
    // ~40 ms, more than a whole 20-frame advect_points run at the C4 size); freed with the context.
    std::vector<std::pair<void*, size_t>> pinned_cache; // free buffers and the size each was allocated with
    std::vector<std::pair<void*, size_t>> pinned_out;   // buffers handed out
};

namespace clumpy {

static inline void dev_release(clumpy_cuda_ctx* ctx, void* p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}
void* ctx_pinned_get(clumpy_cuda_ctx* ctx, size_t bytes); // NULL on failure (error set)
void ctx_pinned_put(clumpy_cuda_ctx* ctx, void* p, size_t bytes);

// advect_points.cc:127-135 (age_offsets.cc)
void age_offsets_host(uint32_t nframes, uint32_t npts, uint32_t* dst);
// the same draws generated on the device (age_offsets_dev.cu), asynchronous on `stream`
int launch_age_offsets(clumpy_cuda_ctx* ctx, cudaStream_t stream, uint32_t nframes, uint32_t n, uint32_t* out_dev);

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

// ---- group composite for half-float cells, k <= 3 (composite_group.cu) ---------------------------
//
// For the sprites advect_points uses (opacity 1) every ring's truncating blend f_a becomes stationary
// after a small number of applications for ALL 256 start values (6 / 8 / 12 for the three rings of
// k = 3), so a frame's effect on a pixel is one lookup in a table indexed by the clamped ring counts and
// the pixel value.  Cells are IEEE halves, two per word, updated with red.global.add.noftz.f16x2: a half
// counts exactly to 2048 and then stays there, so it saturates by itself and never carries.
constexpr int GC_MAX_FRAMES = 8; // frames one composite launch can cover

struct GroupTable {
    bool valid = false;        // false: sprite not supported (k > 3) -> another cell type is used
    int kernel_size = 0;
    int depth[3] = {0, 0, 0};  // applications until ring r is stationary (0: ring absent / opacity 0)
    int nrows = 0;             // (depth0 + 1) * (depth1 + 1) * (depth2 + 1)
    size_t bytes = 0;          // nrows * 260 rounded up to 16
    uint8_t* d_table = nullptr;
};
int build_group_table(clumpy_cuda_ctx* ctx, const SpriteRings& rings, float decay, GroupTable* out);
void free_group_table(GroupTable* t);

// One launch: for f < nf, in order, img = composite(img, cells[f]); frame f's pixels go to frame_out[f]
// when that is non-null (frame_out[nf - 1] must be).  `cells[f]` points at padded row 0, column 0 of the
// region: the cell of image pixel (-halo, -halo); the region must have round_up(height, 8) + 2 * halo rows
// of `pitch` cells.  In-place (frame_out == img_in) is allowed.  zero[]: counter images this launch clears.
// flags != NULL: wait until flags[s] has reached `epoch` for every s < nsrc (multi-GPU frame fence).
struct GroupCompositeArgs {
    int width = 0, height = 0;   // pixels
    int pitch = 0;               // cells per padded row (a multiple of 4)
    int img_pitch = 0;           // bytes per image row
    int nf = 0;
    const void* cells[GC_MAX_FRAMES] = {};
    uint8_t* frame_out[GC_MAX_FRAMES] = {};
    const uint8_t* img_in = nullptr;
    int nzero = 0;
    uint4* zero[GC_MAX_FRAMES] = {};
    uint32_t zero_n16[GC_MAX_FRAMES] = {};
    const uint32_t* flags = nullptr;
    uint32_t epoch = 0;
    int nsrc = 0;
    uint32_t* errors = nullptr;
    unsigned long long timeout_ns = 0;
    bool alone = true;           // no other launch shares the SMs with this one: take the whole register file (prefetch)
    // filled in by the launcher
    int tiles_x = 0, ntiles = 0, vec_ok = 0;
    uint32_t zero_units = 0;     // 16-byte units per image to zero (all the same)
    int hints = 0; // CLUMPY_CUDA_HINTS (A/B): bit 2 = counter reads evict-first, bit 3 = zeroing stores evict-first
    float depth[3] = {0.f, 0.f, 0.f};
};
int launch_composite_group(clumpy_cuda_ctx* ctx, cudaStream_t stream, const GroupTable& t, GroupCompositeArgs& a);

// ---- exact-order path for sparse clouds (composite_exact.cu): 64-bit cells = count + index sum ----
bool exact_mode_supported(int kernel_size, uint64_t npts);
// index_base: the original index of d_pts[0] (a shard of a longer list: the cells carry ORIGINAL indices)
int launch_exact_count(clumpy_cuda_ctx* ctx, const float* d_pts, uint32_t npts, const HistGeom& g,
                       int wrapx, void* d_cells, uint32_t index_base = 0);
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
