// advect.cu -- host side of clumpy `advect_points` for sm_100a (kernels: advect_kernels.cuh,
// composite_group.cu, composite_exact.cu, splat.cu).  Replaces commands/advect_points.cc:83-179.
//
// Layout in HBM (SoA, all sized npts unless noted):
//   pos     float2   current positions            read + written once per GROUP of frames (streaming)
//   orig    float2   positions to reset to        read only when a point resets (1/nframes of them per frame)
//   offset  u16/u32  reset phase of each point    read once per group (streaming); u16 if nframes < 65535
//   perm    uint32   original index of each slot  only for read_points and regrouping
//   (a second copy of those four: regroup_points writes there and swaps)
//   vel     float2   (H, W) velocity field        random 8-byte gathers
//   hist    counter images (see CellMode): three sets of `fuse` images for half cells (one set being counted
//           into, one being composited, one being zeroed), a ring of 2 * fuse otherwise
//   imgs    uint8    framebuffers: the current one, + one per recorded frame whose copy is in flight
//
// One GROUP of up to 8 simulation frames is ONE advect launch (positions stay in registers, frame f counts
// into counter image f of the group's set) followed by ONE composite launch (pixels stay in registers,
// composite_group.cu) which also zeroes the images of the group before.  The composite of group g runs on
// a side stream while the advect launch of group g + 1 runs on the main stream.
//
// Points are kept in the order of the 8x4-texel tile they are in (counting sort at set-up, repeated
// every ~nframes/16 frames: see regroup_points), so a warp's velocity gathers and counter updates stay
// within a few cache lines.  Results do not depend on the order (integer counts commute; trajectories
// are per point); read_points undoes the permutation.  Sparse clouds stay in the caller's order: their
// cells carry the original point index (CELL_X64) so that the composite pass can keep the reference's
// hit order.
#include "advect_kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

using namespace clumpy;

namespace {

constexpr int NSETS = 3;                                // counter-image sets of the single-GPU group path: counted into | composited | being zeroed
constexpr int RSETS = ROUTE_SETS;                       // band-buffer sets of the routed path
constexpr int MAX_HIST = NSETS * ADVECT_MAX_FUSE;       // counter images at most
constexpr int MAX_IMAGES = 2 * ADVECT_MAX_FUSE + 2;     // framebuffers at most (recorded frames in flight)
constexpr int LEGACY_RING = 8;                          // counter images of the u32 / exact-order path

// CLUMPY_CUDA_TRACE=1: wall-clock phase marks of the set-up and the whole command on stderr (each mark
// synchronises the device first, so the trace perturbs the run it describes; measurement aid only).
void trace_mark(clumpy_cuda_ctx* ctx, const char* what) {
    static const bool on = [] { const char* t = getenv("CLUMPY_CUDA_TRACE"); return t && t[0] == '1'; }();
    if (!on) return;
    static const auto t0 = std::chrono::steady_clock::now();
    static double last = 0.0;
    if (ctx) { cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->aux_stream); cudaStreamSynchronize(ctx->copy_stream); }
    const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[clumpy_cuda trace] %8.2f ms (+%7.2f)  %s\n", t * 1e3, (t - last) * 1e3, what);
    last = t;
}

int env_int(const char* name, int fallback) {
    const char* v = getenv(name);
    return v && v[0] ? atoi(v) : fallback;
}

} // namespace

// Where the frames of a group go besides the framebuffer ring: rows [row0, row0 + nrows) of frame i of the
// group are copied to dst + (done + i) * frame_stride on the copy stream.  external: the caller issues the
// copies itself from the images slots[] (and records copy_done[] behind each of them).
struct RecordSink {
    uint8_t* dst = nullptr;
    size_t frame_stride = 0;
    uint32_t row0 = 0, nrows = 0;
    uint32_t done = 0; // frames already delivered
    bool external = false;
    int slots[ADVECT_MAX_FUSE] = {};
};

struct clumpy_advect {
    clumpy_cuda_ctx* ctx = nullptr;
    uint32_t npts = 0, width = 0, height = 0, nframes = 0, simframe = 0;
    int kernel_size = 1, wrapx = 0;
    float step_size = 0.f, decay = 0.f;
    HistGeom geom;
    SpriteRings rings;
    float2 *pos = nullptr, *orig = nullptr, *vel = nullptr;
    bool vel_owned = true;   // false: the caller's device field is used in place
    void* offset = nullptr;  // uint16_t[npts] when nframes <= 65536 (off16), else uint32_t[npts]
    bool off16 = false;
    uint32_t* perm = nullptr;
    bool sharded = false;    // the points are a selection of the caller's list: perm holds ORIGINAL indices
    // second copy of the per-point arrays (regroup writes there, then the roles swap) + sort scratch
    float2 *pos_alt = nullptr, *orig_alt = nullptr;
    void* offset_alt = nullptr;
    uint32_t *perm_alt = nullptr, *sort_counts = nullptr, *sort_sums = nullptr;
    SortGrid sort_grid = {};
    uint32_t regroup_every = 0, frames_since_regroup = 0; // 0: never regroup
    // counters
    int cell_mode = CELL_U32;
    size_t hist_bytes = 0, hist_stride = 0; // bytes of ONE counter image / between consecutive images
    uint32_t* hist = nullptr;               // the allocation
    int nhist = 0;
    uint32_t* hist_buf[MAX_HIST] = {};
    int fuse = 1;                           // simulation frames per advect launch
    bool overlap = false, side_dirty = false;
    int overlap_opt = -1;     // what the caller asked for: 1 on, 0 off, -1 automatic
    bool keep_pos = false;    // positions / phases are accessed with the normal L2 policy instead of
                              // evict-first (when the whole working set of a rank fits in L2)
    // group path (half cells): set g % 3 of group g; composite(g) zeroes the images group g - 1 used
    GroupTable gtab;
    uint64_t group_index = 0;
    int set_used[NSETS] = {0, 0, 0};
    cudaEvent_t set_composited[NSETS] = {};
    bool set_pending[NSETS] = {false, false, false};
    // legacy path (u32 / exact-order cells): a ring of images, per-frame composite + clear
    int cellbuf = 0;
    bool clear_pending[LEGACY_RING] = {};
    cudaEvent_t cleared[LEGACY_RING] = {};
    cudaEvent_t advected = nullptr, side_done = nullptr, frames_ready = nullptr;
    bool count_open = false;  // advect_count was called, advect_finish_frame not yet
    // framebuffers
    int cur = 0;
    uint8_t* imgs[MAX_IMAGES] = {};
    bool copy_pending[MAX_IMAGES] = {};
    uint64_t copy_seq[MAX_IMAGES] = {}, copy_counter = 0;
    cudaEvent_t copy_done[MAX_IMAGES] = {};
    cudaEvent_t frame_ready = nullptr;
    // multi-GPU routing (clumpy_cuda_advect_route_*): this rank's inbox, band buffers and peers
    struct Route {
        bool on = false;
        int rank = 0, world = 1, chunk = 0;
        int fuse = ROUTE_MAX_FUSE;   // frames per route launch
        uint32_t nregion = 0;        // regions of ROUTE_REGION entry slots per source segment
        bool inbox_owned = true;     // false: the caller's arena (kept across runs together with the peers' mappings of it)
        uint32_t* inbox = nullptr;   // [ROUTE_PARITIES][ROUTE_MAX_FUSE][world][nregion * ROUTE_REGION] cell indices, then
                                     // [ROUTE_PARITIES][ROUTE_MAX_FUSE][world][nregion] count words, then [world] epoch flags
        size_t entries_per_frame = 0, cnt_per_frame = 0, cnt_offset = 0, flags_offset = 0; // in uint32 units
        uint32_t* band = nullptr;    // RSETS x ROUTE_MAX_FUSE band buffers of [halo | chunk | halo | slack] rows
        size_t band_bytes = 0;       // bytes of ONE band buffer (a multiple of 256)
        uint32_t band_cells = 0;
        uint64_t group_index = 0;    // groups that splatted so far: parity = % ROUTE_PARITIES, set = % RSETS
        int set_used[RSETS] = {};
        cudaEvent_t routed = nullptr, composited[RSETS] = {};
        bool composited_pending[RSETS] = {};
        uint32_t* ticket = nullptr;
        uint32_t* peer_inbox[ROUTE_MAX_WORLD] = {nullptr};
        uint32_t* errors = nullptr;      // pinned, mapped: fence time-outs seen by this rank's drain kernels
        uint32_t* errors_dev = nullptr;  // its device address
        unsigned long long timeout_ns = 2000000000ull;
        uint32_t epoch = 0;          // epoch of the last group launched
        bool loopback = false;       // CLUMPY_ROUTE_LOOPBACK=1: see route_begin
    } route;
};

// ---- small helpers -------------------------------------------------------------------------------------

static size_t cell_bytes_of(int mode) { return mode == CELL_F16 ? 2 : mode == CELL_X64 ? 8 : 4; }

static bool frame_splats(const clumpy_advect* a) {
    return !(a->simframe < a->nframes && a->decay == 0.0f); // advect_points.cc:154
}

// A fence time-out is sticky: every later call on the run fails until it is closed.
static int route_check(const clumpy_advect* a) {
    if (a->route.on && a->route.errors && *reinterpret_cast<volatile uint32_t*>(a->route.errors) != 0u) {
        set_error("multi-GPU frame fence timed out (%u time-outs): a peer stopped publishing its epochs; the frames of this run are incomplete",
                  *reinterpret_cast<volatile uint32_t*>(a->route.errors));
        return CLUMPY_ERR_STATE;
    }
    return CLUMPY_OK;
}

static int ensure_image(clumpy_advect* a, int i) {
    if (a->imgs[i]) return CLUMPY_OK;
    clumpy_cuda_ctx* ctx = a->ctx;
    DEV_ALLOC(ctx, &a->imgs[i], (size_t)a->width * a->height);
    CK(cudaEventCreateWithFlags(&a->copy_done[i], cudaEventDisableTiming));
    // the allocation is ordered on the main stream; the composites run on the side stream
    CK(cudaEventRecord(a->advected, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->aux_stream, a->advected, 0));
    CK(cudaStreamWaitEvent(ctx->copy_stream, a->advected, 0));
    return CLUMPY_OK;
}

// An image other than the current one and those in `taken` that no copy is reading; when every image has
// a copy in flight, the one whose copy was issued first, after making `stream` wait for it.
static int pick_free_image(clumpy_advect* a, cudaStream_t stream, uint32_t taken_mask, int* out) {
    int oldest = -1;
    for (int i = 0; i < MAX_IMAGES; ++i) {
        if (i == a->cur || ((taken_mask >> i) & 1u)) continue;
        if (!a->copy_pending[i]) {
            int rc = ensure_image(a, i);
            if (rc) return rc;
            *out = i;
            return CLUMPY_OK;
        }
        if (oldest < 0 || a->copy_seq[i] < a->copy_seq[oldest]) oldest = i;
    }
    REQUIRE(oldest >= 0, "no framebuffer available");
    CK(cudaStreamWaitEvent(stream, a->copy_done[oldest], 0));
    a->copy_pending[oldest] = false;
    *out = oldest;
    return CLUMPY_OK;
}

static void advect_release(clumpy_advect* a) {
    if (!a) return;
    clumpy_cuda_ctx* ctx = a->ctx;
    cudaSetDevice(ctx->device);
    for (cudaStream_t s : {ctx->stream, ctx->aux_stream, ctx->copy_stream}) cudaStreamSynchronize(s);
    for (void* p : {(void*)a->pos, (void*)a->orig, a->offset, (void*)a->perm, (void*)a->pos_alt, (void*)a->orig_alt,
                    a->offset_alt, (void*)a->perm_alt, (void*)a->sort_counts, (void*)a->sort_sums, (void*)a->hist,
                    (void*)a->route.band, (void*)a->route.ticket})
        dev_release(ctx, p);
    if (a->vel_owned) dev_release(ctx, a->vel);
    for (int i = 0; i < MAX_IMAGES; ++i) {
        dev_release(ctx, a->imgs[i]);
        if (a->copy_done[i]) cudaEventDestroy(a->copy_done[i]);
    }
    if (a->route.inbox_owned) cudaFree(a->route.inbox); // peers map it: a plain allocation
    if (a->route.errors) cudaFreeHost(a->route.errors);
    for (int s = 0; s < NSETS; ++s) if (a->set_composited[s]) cudaEventDestroy(a->set_composited[s]);
    for (int s = 0; s < RSETS; ++s) if (a->route.composited[s]) cudaEventDestroy(a->route.composited[s]);
    if (a->route.routed) cudaEventDestroy(a->route.routed);
    free_group_table(&a->gtab);
    for (int b = 0; b < LEGACY_RING; ++b) if (a->cleared[b]) cudaEventDestroy(a->cleared[b]);
    for (cudaEvent_t e : {a->advected, a->side_done, a->frames_ready, a->frame_ready}) if (e) cudaEventDestroy(e);
    cudaStreamSynchronize(ctx->stream); // the releases above are ordered on the main stream
    cudaGetLastError();
    delete a;
}

// advect_points.cc:127-135.  The arithmetic is the standard's (mt19937) and libstdc++'s (distribution);
// age_offsets.cc restates both without going through <random> one draw at a time.
CLUMPY_API int clumpy_cuda_age_offsets(uint32_t nframes, uint32_t npts, uint32_t* out) {
    REQUIRE(out || npts == 0, "out is NULL");
    REQUIRE(nframes > 0 && nframes <= 0x80000000u, "nframes must be in 1 .. 2^31");
    age_offsets_host(nframes, npts, out);
    return CLUMPY_OK;
}

// ---- sorting ---------------------------------------------------------------------------------------------

// Counting sort of the per-point records by the image tile each point is in now (asynchronous, on
// the context's stream).  Used once at set-up (where "now" is "home") and then every
// `regroup_every` frames: points of one tile share a streamline but not an age (each has its own
// reset phase), so a fixed order slowly smears every warp's points along ~nframes*speed texels of
// streamline and the gathers / counter updates lose their locality; regrouping restores it for the
// price of one extra pass over the points every few dozen frames.  Results do not depend on the order.
static int regroup_points(clumpy_advect* a) {
    clumpy_cuda_ctx* ctx = a->ctx;
    const uint32_t n = a->npts;
    if (n == 0 || !a->sort_counts) return CLUMPY_OK;
    const SortGrid& sg = a->sort_grid;
    const uint32_t pgrid = ceil_div(n, 256), nblocks = ceil_div(sg.nbins, SCAN_BLOCK);
    CK(cudaMemsetAsync(a->sort_counts, 0, (size_t)sg.nbins * 4, ctx->stream));
    sort_count_kernel<<<pgrid, 256, 0, ctx->stream>>>(a->pos, n, sg, a->sort_counts);
    CK_LAUNCH(ctx);
    scan_local_kernel<<<nblocks, SCAN_BLOCK, 0, ctx->stream>>>(a->sort_counts, sg.nbins, a->sort_sums);
    CK_LAUNCH(ctx);
    scan_sums_kernel<<<1, SCAN_BLOCK, 0, ctx->stream>>>(a->sort_sums, nblocks);
    CK_LAUNCH(ctx);
    scan_add_kernel<<<nblocks, SCAN_BLOCK, 0, ctx->stream>>>(a->sort_counts, sg.nbins, a->sort_sums);
    CK_LAUNCH(ctx);
    if (a->off16)
        regroup_scatter_kernel<uint16_t><<<pgrid, 256, 0, ctx->stream>>>(
            a->pos, a->orig, (const uint16_t*)a->offset, a->perm, n, sg, a->sort_counts, a->pos_alt, a->orig_alt,
            (uint16_t*)a->offset_alt, a->perm_alt);
    else
        regroup_scatter_kernel<uint32_t><<<pgrid, 256, 0, ctx->stream>>>(
            a->pos, a->orig, (const uint32_t*)a->offset, a->perm, n, sg, a->sort_counts, a->pos_alt, a->orig_alt,
            (uint32_t*)a->offset_alt, a->perm_alt);
    CK_LAUNCH(ctx);
    std::swap(a->pos, a->pos_alt);
    std::swap(a->orig, a->orig_alt);
    std::swap(a->offset, a->offset_alt);
    std::swap(a->perm, a->perm_alt);
    a->frames_since_regroup = 0;
    return CLUMPY_OK;
}

// Sets up the sort grid and does the first regroup (the per-point arrays are filled in the caller's order).
static int sort_points(clumpy_advect* a, int regroup_opt) {
    clumpy_cuda_ctx* ctx = a->ctx;
    const uint32_t n = a->npts;
    if (n == 0) return CLUMPY_OK;
    if (env_int("CLUMPY_CUDA_NO_SORT", 0) == 1) return CLUMPY_OK;
    if (a->cell_mode == CELL_X64) return CLUMPY_OK; // exact-order path: slot == original index

    SortGrid sg;
    sg.width = a->width; sg.height = a->height;
    // 16 x 2 texels per tile: a tile row is one 32-byte sector of half cells and one 128-byte line of the field
    // (measured on B200, C4, steady state: 16x2 0.0775 ms per frame for the advect launch, 8x4 0.0820, 8x2 0.0800,
    // 16x4 0.0796, 32x1 0.0798, 4x4 0.0882; profiles/r2_sweep_single.jsonl).  CLUMPY_CUDA_SORT_TILE="sx,sy" (log2)
    // overrides (A/B runs).
    sg.shift_x = 4; sg.shift_y = 1;
    if (const char* t = getenv("CLUMPY_CUDA_SORT_TILE")) {
        int sx = 0, sy = 0;
        if (sscanf(t, "%d,%d", &sx, &sy) == 2 && sx >= 0 && sx < 12 && sy >= 0 && sy < 12) { sg.shift_x = sx; sg.shift_y = sy; }
    }
    auto tiles = [&](int sx, int sy) {
        return (uint64_t)ceil_div(a->width, 1u << sx) * ceil_div(a->height, 1u << sy);
    };
    while (tiles(sg.shift_x, sg.shift_y) > (1u << 21)) {
        if (sg.shift_y < sg.shift_x) ++sg.shift_y; else ++sg.shift_x;
    }
    sg.tiles_x = ceil_div(a->width, 1u << sg.shift_x);
    sg.ntiles = (uint32_t)tiles(sg.shift_x, sg.shift_y);
    sg.nbins = sg.ntiles + 2u * (a->width + 2u) + 2u * a->height; // + the ring of cells around the image (home_tile)
    a->sort_grid = sg;
    const size_t nalloc = std::max<size_t>((size_t)n + 1, 2);
    DEV_ALLOC(ctx, &a->sort_counts, (size_t)sg.nbins * 4);
    DEV_ALLOC(ctx, &a->sort_sums, (size_t)ceil_div(sg.nbins, SCAN_BLOCK) * 4);
    DEV_ALLOC(ctx, &a->pos_alt, nalloc * 8);
    DEV_ALLOC(ctx, &a->orig_alt, nalloc * 8);
    DEV_ALLOC(ctx, &a->offset_alt, nalloc * 4);
    DEV_ALLOC(ctx, &a->perm_alt, nalloc * 4);
    // Regroup period: by then about 1/16 of the points have been reset away from their group; a whole
    // number of launches.  Small clouds do not need it.  desc.regroup / CLUMPY_CUDA_REGROUP=<frames> override,
    // -1 / 0 disable.
    a->regroup_every = 0;
    if (n >= (1u << 16)) {
        const uint32_t f = (uint32_t)std::max(a->fuse, 1);
        a->regroup_every = std::max<uint32_t>(f, (std::max<uint32_t>(8u, a->nframes / 16u) + f / 2) / f * f);
    }
    if (regroup_opt > 0) a->regroup_every = (uint32_t)regroup_opt;
    if (regroup_opt < 0) a->regroup_every = 0;
    if (const char* r = getenv("CLUMPY_CUDA_REGROUP")) a->regroup_every = (uint32_t)atoi(r);
    return regroup_points(a);
}

// ---- set-up ------------------------------------------------------------------------------------------------

static int choose_cells(clumpy_advect* a, int cells_opt) {
    clumpy_cuda_ctx* ctx = a->ctx;
    const size_t npix = (size_t)a->width * a->height;
    //   auto (default)  sparse cloud (8 * npts <= pixels): exact-order 64-bit cells, points kept in the
    //                   caller's order; otherwise half-float cells + one-lookup replay for k <= 3 (every ring
    //                   contracts strongly: within 1 LSB), points regrouped by tile; 32-bit cells with the
    //                   proportionally interleaved replay for larger sprites (they have a 0.028 ring)
    //   desc.cells / CLUMPY_CUDA_CELLS = exact | f16 | u32 | count forces one (A/B runs, tests, multi-GPU)
    std::string want = cells_opt == CLUMPY_CELLS_F16 ? "f16" : cells_opt == CLUMPY_CELLS_U32 ? "u32"
                     : cells_opt == CLUMPY_CELLS_EXACT ? "exact" : cells_opt == CLUMPY_CELLS_COUNT ? "count" : "auto";
    if (const char* force = getenv("CLUMPY_CUDA_CELLS")) want = force;
    const bool can_exact = exact_mode_supported(a->kernel_size, a->npts);
    if (want == "auto")
        want = (can_exact && !a->sharded && 8ull * a->npts <= (uint64_t)npix) ? "exact" : "count";
    if (want == "count") want = a->kernel_size <= 3 ? "f16" : "u32"; // the count-only type of this sprite
    a->cell_mode = CELL_U32;
    if (want == "exact" && can_exact && !a->sharded) {
        a->cell_mode = CELL_X64;
    } else if (want == "f16") {
        int rc = build_group_table(ctx, a->rings, a->decay, &a->gtab);
        if (rc) return rc;
        if (a->gtab.valid) a->cell_mode = CELL_F16; // else (k > 3): 32-bit cells
    }
    return CLUMPY_OK;
}

static int advect_begin_impl(clumpy_advect* a, const clumpy_advect_desc& d) {
    clumpy_cuda_ctx* ctx = a->ctx;
    const size_t npix = (size_t)a->width * a->height;
    trace_mark(ctx, "advect_begin: start");
    for (cudaEvent_t* e : {&a->advected, &a->side_done, &a->frames_ready, &a->frame_ready})
        CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (int s = 0; s < NSETS; ++s) CK(cudaEventCreateWithFlags(&a->set_composited[s], cudaEventDisableTiming));
    for (int b = 0; b < LEGACY_RING; ++b) CK(cudaEventCreateWithFlags(&a->cleared[b], cudaEventDisableTiming));

    // Reset offsets: the caller's, or ours.  Ours are drawn on the device (age_offsets_dev.cu: one CTA, a few ms)
    // on the side stream, beside the uploads and the shard selection of the main stream.
    const uint32_t n_in = d.npts;
    float2* d_in_own = nullptr;
    uint32_t *d_offs_own = nullptr, *d_cursor = nullptr;
    struct Scratch { // scratch allocations go back to the pool on every path out
        clumpy_cuda_ctx* c; float2*& p; uint32_t*& o; uint32_t*& q;
        ~Scratch() { cudaStreamSynchronize(c->aux_stream); dev_release(c, p); dev_release(c, o); dev_release(c, q); }
    } scratch{ctx, d_in_own, d_offs_own, d_cursor};
    if (n_in && !(d.age_offset && d.age_on_device)) DEV_ALLOC(ctx, &d_offs_own, (size_t)n_in * 4);
    trace_mark(ctx, "advect_begin: events, offsets buffer");
    if (!d.age_offset && n_in) {
        CK(cudaEventRecord(a->advected, ctx->stream)); // the allocation is ordered on the main stream
        CK(cudaStreamWaitEvent(ctx->aux_stream, a->advected, 0));
        int rc = launch_age_offsets(ctx, ctx->aux_stream, a->nframes, n_in, d_offs_own);
        if (rc) return rc;
        CK(cudaEventRecord(a->side_done, ctx->aux_stream));
    }
    trace_mark(ctx, "advect_begin: events, reset offsets enqueued");

    // the caller's points on the device.  They go up BEFORE the field: the set-up kernels that need them (shard
    // selection, point records, sort) then run beside the field's upload, which has a stream of its own.
    const float2* d_in = reinterpret_cast<const float2*>(d.pts);
    if (!d.pts_on_device && n_in) {
        DEV_ALLOC(ctx, &d_in_own, (size_t)n_in * 8);
        CK(cudaMemcpyAsync(d_in_own, d.pts, (size_t)n_in * 8, cudaMemcpyHostToDevice, ctx->stream));
        d_in = d_in_own;
        trace_mark(ctx, "advect_begin: point upload");
    }
    // the field: used in place when it is already on the device
    bool field_in_flight = false;
    if (d.vel_on_device) {
        a->vel = const_cast<float2*>(reinterpret_cast<const float2*>(d.vel));
        a->vel_owned = false;
    } else {
        DEV_ALLOC(ctx, &a->vel, npix * 8);
        CK(cudaEventRecord(a->frame_ready, ctx->stream)); // (allocation and the points' copy are ordered on the main stream)
        CK(cudaStreamWaitEvent(ctx->copy_stream, a->frame_ready, 0));
        CK(cudaMemcpyAsync(a->vel, d.vel, npix * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(a->frames_ready, ctx->copy_stream));
        field_in_flight = true;
        trace_mark(ctx, "advect_begin: field upload");
    }
    // which of them this run simulates
    const bool shard = d.shard_row_hi > d.shard_row_lo;
    a->sharded = shard;
    a->npts = n_in;
    if (shard && n_in) {
        // count first (one pass over the points), then allocate exactly
        DEV_ALLOC(ctx, &d_cursor, (size_t)(a->height + 1) * 4);
        CK(cudaMemsetAsync(d_cursor, 0, (size_t)(a->height + 1) * 4, ctx->stream));
        row_histogram_kernel<<<std::min<uint32_t>(ceil_div(n_in, 256), (uint32_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(d_in, n_in, a->height, d_cursor);
        CK_LAUNCH(ctx);
        std::vector<uint32_t> rows(a->height);
        CK(cudaMemcpyAsync(rows.data(), d_cursor, (size_t)a->height * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        uint64_t mine = 0;
        for (uint32_t r = d.shard_row_lo; r < std::min(d.shard_row_hi, a->height); ++r) mine += rows[r];
        a->npts = (uint32_t)mine;
    }
    const size_t n = a->npts;
    const size_t nalloc = std::max<size_t>(n + 1, 2); // slack so that 16-byte accesses stay inside
    a->off16 = a->nframes < 65535u; // 0xFFFF is the "never resets" sentinel (stored_offset)
    DEV_ALLOC(ctx, &a->pos, nalloc * 8);
    DEV_ALLOC(ctx, &a->orig, nalloc * 8);
    DEV_ALLOC(ctx, &a->offset, nalloc * 4);
    DEV_ALLOC(ctx, &a->perm, nalloc * 4);

    trace_mark(ctx, "advect_begin: point arrays allocated");
    int rc = choose_cells(a, d.cells);
    if (rc) return rc;
    trace_mark(ctx, "advect_begin: cell type chosen, replay table built");
    {
        // The composite work of a group may run on a side stream next to the advect launch of the next group.
        // With both stages bound by DRAM (round 2: profiles/r2_ncu_steady_summary.txt) that buys nothing in the
        // steady state (C4: 0.1417 ms per frame overlapped, 0.1293 one after the other), so it is off unless asked
        // for; the multi-GPU path, whose stages wait for other GPUs, always overlaps (route_begin).
        const int ov = env_int("CLUMPY_CUDA_OVERLAP", d.overlap > 0 ? 1 : d.overlap < 0 ? 0 : -1);
        a->overlap_opt = ov;
        a->overlap = ov > 0;
        a->keep_pos = env_int("CLUMPY_CUDA_KEEP_POS", 0) == 1;
    }
    // Counter images.  Half cells: three sets of `fuse` images; while the images of group g are composited
    // (which also zeroes those of group g - 1) on the side stream, group g + 1 already counts into the third
    // set.  Other cells: a ring of 2 * fuse images (per-frame composite + clear on the side stream).  Very
    // large images fall back to shorter groups.  desc.fuse / CLUMPY_CUDA_FUSE=<1..8> override.
    a->hist_bytes = a->geom.cells() * cell_bytes_of(a->cell_mode);
    a->hist_stride = round_up(a->hist_bytes, 256);
    {
        const int max_fuse = a->cell_mode == CELL_F16 ? ADVECT_MAX_FUSE : LEGACY_RING / 2;
        int fuse = max_fuse;
        while (fuse > 1 && (size_t)fuse * a->hist_stride > (4ull << 30)) fuse /= 2;
        if (d.fuse > 0) fuse = std::min(d.fuse, fuse);
        a->fuse = std::min(std::max(env_int("CLUMPY_CUDA_FUSE", fuse), 1), max_fuse);
        a->nhist = a->cell_mode == CELL_F16 ? NSETS * a->fuse : std::max(2, 2 * a->fuse);
    }
    DEV_ALLOC(ctx, &a->hist, (size_t)a->nhist * a->hist_stride);
    for (int b = 0; b < a->nhist; ++b)
        a->hist_buf[b] = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(a->hist) + (size_t)b * a->hist_stride);
    CK(cudaMemsetAsync(a->hist, 0, (size_t)a->nhist * a->hist_stride, ctx->stream));
    rc = ensure_image(a, 0);
    if (rc) return rc;
    CK(cudaMemsetAsync(a->imgs[0], 0, npix, ctx->stream)); // advect_points.cc:119: zero-initialised
    a->cur = 0;
    trace_mark(ctx, "advect_begin: allocations, uploads of field and points, cell tables");

    // reset offsets on the device
    const uint32_t* d_offs = d.age_offset;
    if (n_in && !(d.age_offset && d.age_on_device)) {
        if (d.age_offset) CK(cudaMemcpyAsync(d_offs_own, d.age_offset, (size_t)n_in * 4, cudaMemcpyHostToDevice, ctx->stream));
        else CK(cudaStreamWaitEvent(ctx->stream, a->side_done, 0)); // drawn on the side stream (above)
        d_offs = d_offs_own;
    }
    if (n) {
        const uint32_t pgrid = ceil_div(n_in, 256);
        if (shard) {
            uint32_t* cursor = d_cursor + a->height; // (zeroed above, untouched by the histogram)
            if (a->off16) select_rows_kernel<uint16_t><<<pgrid, 256, 0, ctx->stream>>>(d_in, d_offs, n_in, a->nframes, a->height, d.shard_row_lo, d.shard_row_hi, cursor, a->npts, a->pos, a->orig, (uint16_t*)a->offset, a->perm);
            else select_rows_kernel<uint32_t><<<pgrid, 256, 0, ctx->stream>>>(d_in, d_offs, n_in, a->nframes, a->height, d.shard_row_lo, d.shard_row_hi, cursor, a->npts, a->pos, a->orig, (uint32_t*)a->offset, a->perm);
        } else {
            if (a->off16) identity_order_kernel<uint16_t><<<pgrid, 256, 0, ctx->stream>>>(d_in, d_offs, a->npts, a->nframes, a->pos, a->orig, (uint16_t*)a->offset, a->perm);
            else identity_order_kernel<uint32_t><<<pgrid, 256, 0, ctx->stream>>>(d_in, d_offs, a->npts, a->nframes, a->pos, a->orig, (uint32_t*)a->offset, a->perm);
        }
        CK_LAUNCH(ctx);
    }
    rc = sort_points(a, d.regroup);
    if (rc) return rc;
    if (field_in_flight) CK(cudaStreamWaitEvent(ctx->stream, a->frames_ready, 0));
    CK(cudaStreamSynchronize(ctx->stream)); // host buffers of the caller are free again; errors surface here
    trace_mark(ctx, "advect_begin: reset offsets, point records, sort");
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_begin_ex(clumpy_cuda_ctx* ctx, const clumpy_advect_desc* desc, clumpy_advect** out) {
    REQUIRE(out, "out is NULL");
    *out = nullptr;
    REQUIRE(ctx && desc, "null argument");
    REQUIRE(desc->size == sizeof(clumpy_advect_desc), "clumpy_advect_desc.size does not match this library");
    const clumpy_advect_desc& d = *desc;
    REQUIRE(d.vel && (d.pts || d.npts == 0), "null argument");
    REQUIRE(d.width > 0 && d.height > 0, "image dimensions must be positive");
    REQUIRE(d.width < (1u << 30) && d.height < (1u << 30) && (uint64_t)d.width * d.height < (1ull << 32),
            "velocity field too large");
    REQUIRE(d.nframes > 0 && d.nframes <= 0x80000000u, "nframes must be in 1 .. 2^31");
    REQUIRE(d.decay >= 0.0f || d.decay != d.decay, "decay must be non-negative (pass wrapx for the negative-decay mode)");
    REQUIRE(d.shard_row_hi >= d.shard_row_lo, "bad shard rows");
    SpriteRings rings;
    // advect_points.cc:156,175: alpha is the constant 1.0f
    int rc = make_sprite_rings(d.kernel_size, 1.0f, &rings);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    clumpy_advect* a = new (std::nothrow) clumpy_advect();
    if (!a) { set_error("out of host memory"); return CLUMPY_ERR_NOMEM; }
    a->ctx = ctx; a->width = d.width; a->height = d.height; a->nframes = d.nframes;
    a->kernel_size = d.kernel_size; a->wrapx = d.wrapx ? 1 : 0; a->step_size = d.step_size; a->decay = d.decay;
    a->geom = make_hist_geom(d.width, d.height, d.kernel_size);
    // The dense multi-GPU variant splits the counter image into equal row bands (one per rank, whole 16-row
    // tiles): it asks for the padded row count to be a multiple of 16 * world_size.
    const int mult = d.cell_rows_multiple;
    if (mult > 1 && mult <= (1 << 20)) a->geom.rows = (int)round_up((size_t)a->geom.rows, (size_t)mult);
    a->rings = rings;
    rc = advect_begin_impl(a, d);
    if (rc) { advect_release(a); return rc; }
    *out = a;
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_begin(clumpy_cuda_ctx* ctx, const float* pts_host, uint32_t npts,
                                        const float* vel_host, uint32_t width, uint32_t height,
                                        float step_size, int kernel_size, float decay, int wrapx,
                                        uint32_t nframes, const uint32_t* age_offset,
                                        clumpy_advect** out) {
    clumpy_advect_desc d;
    memset(&d, 0, sizeof(d));
    d.size = sizeof(d);
    d.pts = pts_host; d.npts = npts; d.vel = vel_host; d.width = width; d.height = height;
    d.step_size = step_size; d.kernel_size = kernel_size; d.decay = decay; d.wrapx = wrapx; d.nframes = nframes;
    d.age_offset = age_offset;
    return clumpy_cuda_advect_begin_ex(ctx, &d, out);
}

// Home rows of a point list (host or device pointer): rows_host[y] = number of points whose home row is y.
// Multi-GPU callers cut the image into strips of equal point count with it (every rank gets the same
// answer from the same list, so no communication is needed) and pass their strip as desc.shard_row_lo / _hi.
CLUMPY_API int clumpy_cuda_row_histogram(clumpy_cuda_ctx* ctx, const float* pts, uint32_t npts, int pts_on_device,
                                         uint32_t height, uint32_t* rows_host) {
    REQUIRE(ctx && rows_host && (pts || npts == 0) && height > 0, "bad argument");
    CK(cudaSetDevice(ctx->device));
    uint32_t* d_rows = nullptr;
    float2* d_pts = nullptr;
    DEV_ALLOC(ctx, &d_rows, (size_t)height * 4);
    CK(cudaMemsetAsync(d_rows, 0, (size_t)height * 4, ctx->stream));
    const float2* p = reinterpret_cast<const float2*>(pts);
    if (!pts_on_device && npts) {
        DEV_ALLOC(ctx, &d_pts, (size_t)npts * 8);
        CK(cudaMemcpyAsync(d_pts, pts, (size_t)npts * 8, cudaMemcpyHostToDevice, ctx->stream));
        p = d_pts;
    }
    if (npts) {
        row_histogram_kernel<<<std::min<uint32_t>(ceil_div(npts, 256), (uint32_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(p, npts, height, d_rows);
        CK_LAUNCH(ctx);
    }
    CK(cudaMemcpyAsync(rows_host, d_rows, (size_t)height * 4, cudaMemcpyDeviceToHost, ctx->stream));
    dev_release(ctx, d_rows);
    dev_release(ctx, d_pts);
    CK(cudaStreamSynchronize(ctx->stream));
    return CLUMPY_OK;
}

// ---- the advect launch --------------------------------------------------------------------------------------

template <bool WRAP, int MODE, typename OffT>
static void launch_advect_t(clumpy_advect* a, uint32_t phase, int nf, const HistRing& ring) {
    const uint32_t grid = ceil_div(a->npts, ADVECT_THREADS * ADVECT_PTS);
    const int hints = env_int("CLUMPY_CUDA_HINTS", 0); // (A/B switch)
    const GeomDev g{a->geom.width, a->geom.height, a->geom.halo, a->geom.pitch, a->keep_pos ? 1 : 0, hints};
    // (6 CTAs of 40 registers per SM; 8 of 32 was measured and lost: profiles/r2_sweep_single_b.jsonl)
    advect_fused_kernel<WRAP, MODE, OffT><<<grid, ADVECT_THREADS, 0, a->ctx->stream>>>(
        a->pos, a->orig, (const OffT*)a->offset, a->npts, a->vel, a->width, a->height, a->step_size, phase,
        a->nframes, nf, g, ring);
}

template <bool WRAP, int MODE>
static void launch_advect_m(clumpy_advect* a, uint32_t phase, int nf, const HistRing& ring) {
    if (a->off16) launch_advect_t<WRAP, MODE, uint16_t>(a, phase, nf, ring);
    else launch_advect_t<WRAP, MODE, uint32_t>(a, phase, nf, ring);
}

template <bool WRAP>
static void launch_advect_w(clumpy_advect* a, uint32_t phase, bool splat, int nf, const HistRing& ring) {
    if (!splat) launch_advect_m<WRAP, CELL_NONE>(a, phase, nf, ring);
    else if (a->cell_mode == CELL_F16) launch_advect_m<WRAP, CELL_F16>(a, phase, nf, ring);
    else if (a->cell_mode == CELL_X64) launch_advect_m<WRAP, CELL_X64>(a, phase, nf, ring);
    else launch_advect_m<WRAP, CELL_U32>(a, phase, nf, ring);
}

static int launch_advect(clumpy_advect* a, uint32_t phase, bool splat, int nf, const HistRing& ring) {
    if (!a->npts) return CLUMPY_OK;
    if (a->wrapx) launch_advect_w<true>(a, phase, splat, nf, ring);
    else launch_advect_w<false>(a, phase, splat, nf, ring);
    CK_LAUNCH(a->ctx);
    return CLUMPY_OK;
}

// How many frames the next launch may cover: at most `want` and `fuse`, never across a regroup or the
// warm-up / recording boundary (where, with decay 0, frames start to splat).
static uint32_t advect_group_size(const clumpy_advect* a, uint32_t want, int fuse) {
    uint32_t nf = std::min<uint32_t>(std::max<uint32_t>(want, 1u), (uint32_t)fuse);
    nf = std::min<uint32_t>(nf, a->nframes - a->simframe % a->nframes);
    if (a->regroup_every) nf = std::min<uint32_t>(nf, a->regroup_every - std::min(a->frames_since_regroup, a->regroup_every - 1u));
    return std::max<uint32_t>(nf, 1u);
}

// The frames a group produced (recording) are in images slot[]: their rows go to the sink on the copy
// stream, after `ready` (recorded on the stream that wrote the images).
static int enqueue_frame_copies(clumpy_advect* a, cudaEvent_t ready, const int* slot, uint32_t nf, RecordSink* sink) {
    clumpy_cuda_ctx* ctx = a->ctx;
    CK(cudaStreamWaitEvent(ctx->copy_stream, ready, 0));
    for (uint32_t i = 0; i < nf; ++i) {
        sink->slots[i] = slot[i];
        a->copy_pending[slot[i]] = true;
        a->copy_seq[slot[i]] = ++a->copy_counter;
        if (sink->external) continue; // the caller copies and records copy_done itself
        if (sink->nrows)
            CK(cudaMemcpyAsync(sink->dst + (size_t)(sink->done + i) * sink->frame_stride,
                               a->imgs[slot[i]] + (size_t)sink->row0 * a->width, (size_t)sink->nrows * a->width,
                               cudaMemcpyDeviceToHost, ctx->copy_stream));
        CK(cudaEventRecord(a->copy_done[slot[i]], ctx->copy_stream));
    }
    sink->done += nf;
    return CLUMPY_OK;
}

// Where the frames of the next group go: the last one becomes the current framebuffer (in place unless a
// copy is still reading it); when recording, every frame gets an image of its own.
static int choose_frame_images(clumpy_advect* a, cudaStream_t stream, uint32_t nf, bool record, int* slot) {
    uint32_t taken = 0;
    for (uint32_t f = 0; f < nf; ++f) {
        slot[f] = -1;
        if (!record && f + 1 < nf) continue;
        if (!record && !a->copy_pending[a->cur]) { slot[f] = a->cur; continue; }
        int rc = pick_free_image(a, stream, taken, &slot[f]);
        if (rc) return rc;
        taken |= 1u << slot[f];
    }
    return CLUMPY_OK;
}

// Timing hooks of clumpy_cuda_advect_profile_stages: events around the advect launch and around the
// composite work of one group, on the streams they run on.
struct StageEvents { cudaEvent_t adv0, adv1, comp0, comp1; };

// `nf` simulation frames: ONE advect launch on the main stream, then the composite work of the group on
// `side` -- the high-priority aux stream when overlapping, so that it shares the GPU with the advect launch
// of the NEXT group, the main stream otherwise.
static int advect_group(clumpy_advect* a, uint32_t nf, RecordSink* sink = nullptr, StageEvents* ev = nullptr, bool serial = false) {
    clumpy_cuda_ctx* ctx = a->ctx;
    const uint32_t s = a->simframe;
    const bool splat = frame_splats(a);
    const uint32_t phase = s % a->nframes;
    const bool overlap = a->overlap && !serial;
    const cudaStream_t main_stream = ctx->stream;
    const cudaStream_t side = overlap ? ctx->aux_stream : main_stream;
    const bool group_path = a->cell_mode == CELL_F16;
    HistRing ring = {a->hist, a->hist_stride / sizeof(uint32_t), 0, a->nhist};
    const int set = (int)(a->group_index % NSETS);
    if (splat) {
        if (group_path) {
            // this set was zeroed by the composite of two groups ago, whose own set is (set + 1) % 3
            const int z = (set + 1) % NSETS;
            if (a->set_pending[z]) {
                CK(cudaStreamWaitEvent(main_stream, a->set_composited[z], 0));
                a->set_pending[z] = false;
            }
            ring.first = set * a->fuse;
        } else {
            ring.first = a->cellbuf;
            for (uint32_t f = 0; f < nf; ++f) { // these images were last used a ring ago: wait for their clears
                const int b = (a->cellbuf + (int)f) % a->nhist;
                if (a->clear_pending[b]) {
                    CK(cudaStreamWaitEvent(main_stream, a->cleared[b], 0));
                    a->clear_pending[b] = false;
                }
            }
        }
    }
    if (ev) CK(cudaEventRecord(ev->adv0, main_stream));
    int rc = launch_advect(a, phase, splat, (int)nf, ring);
    if (rc) return rc;
    if (ev) CK(cudaEventRecord(ev->adv1, main_stream));
    int slot[ADVECT_MAX_FUSE];
    if (splat) {
        if (overlap) {
            CK(cudaEventRecord(a->advected, main_stream));
            CK(cudaStreamWaitEvent(side, a->advected, 0));
        }
        rc = choose_frame_images(a, side, nf, sink != nullptr, slot);
        if (rc) return rc;
        if (ev) CK(cudaEventRecord(ev->comp0, side));
        if (group_path) {
            GroupCompositeArgs ca;
            ca.width = a->geom.width; ca.height = a->geom.height; ca.pitch = a->geom.pitch; ca.img_pitch = (int)a->width;
            ca.nf = (int)nf;
            ca.img_in = a->imgs[a->cur];
            for (uint32_t f = 0; f < nf; ++f) {
                ca.cells[f] = ring.image((int)f);
                ca.frame_out[f] = slot[f] >= 0 ? a->imgs[slot[f]] : nullptr;
            }
            const int prev = (set + NSETS - 1) % NSETS;
            ca.alone = !overlap; // (beside the next group's advect launch otherwise)
            ca.nzero = a->set_used[prev];
            for (int z = 0; z < ca.nzero; ++z) {
                ca.zero[z] = reinterpret_cast<uint4*>(a->hist_buf[prev * a->fuse + z]);
                ca.zero_n16[z] = (uint32_t)(a->hist_stride / 16);
            }
            rc = launch_composite_group(ctx, side, a->gtab, ca);
            if (rc) return rc;
            a->set_used[prev] = 0;
            a->set_used[set] = (int)nf;
            CK(cudaEventRecord(a->set_composited[set], side));
            a->set_pending[set] = overlap;
            a->group_index++;
        } else {
            for (uint32_t f = 0; f < nf; ++f) {
                const int b = (a->cellbuf + (int)f) % a->nhist;
                uint32_t* const cells = a->hist_buf[b];
                // frames that nobody records are composited in place
                int dst = slot[f];
                if (dst < 0) {
                    dst = a->cur;
                    if (a->copy_pending[dst]) { rc = pick_free_image(a, side, 0, &dst); if (rc) return rc; }
                }
                ctx->stream = side; // the launchers enqueue on ctx->stream
                if (a->cell_mode == CELL_X64)
                    rc = launch_composite_exact(ctx, a->geom, a->rings, cells, a->wrapx, 1, a->decay, a->imgs[a->cur], a->imgs[dst]);
                else
                    rc = launch_composite(ctx, a->geom, a->rings, cells, a->wrapx, 1, a->decay, a->imgs[a->cur], a->imgs[dst]);
                ctx->stream = main_stream;
                if (rc) return rc;
                a->cur = dst;
                CK(cudaMemsetAsync(cells, 0, a->hist_bytes, side));
                if (overlap) {
                    CK(cudaEventRecord(a->cleared[b], side));
                    a->clear_pending[b] = true;
                }
            }
            a->cellbuf = (a->cellbuf + (int)nf) % a->nhist;
        }
        if (ev) CK(cudaEventRecord(ev->comp1, side));
        a->cur = slot[nf - 1];
        if (overlap) a->side_dirty = true;
    } else {
        if (ev) { CK(cudaEventRecord(ev->comp0, side)); CK(cudaEventRecord(ev->comp1, side)); }
        for (uint32_t f = 0; f < nf; ++f) slot[f] = a->cur; // frames that do not splat leave the framebuffer as it is
    }
    if (sink) {
        CK(cudaEventRecord(a->frames_ready, side));
        rc = enqueue_frame_copies(a, a->frames_ready, slot, nf, sink);
        if (rc) return rc;
    }
    a->simframe = s + nf;
    a->frames_since_regroup += nf;
    if (a->regroup_every && a->frames_since_regroup >= a->regroup_every) return regroup_points(a);
    return CLUMPY_OK;
}

// Makes the main stream wait for everything issued on the side stream, so that whatever the caller
// enqueues (or records) on the context's stream next is ordered after the whole frame.
static int advect_join(clumpy_advect* a) {
    if (!a->side_dirty) return CLUMPY_OK;
    clumpy_cuda_ctx* ctx = a->ctx;
    CK(cudaEventRecord(a->side_done, ctx->aux_stream));
    CK(cudaStreamWaitEvent(ctx->stream, a->side_done, 0));
    a->side_dirty = false;
    return CLUMPY_OK;
}

static int advect_steps(clumpy_advect* adv, uint32_t count, RecordSink* sink) {
    for (uint32_t i = 0; i < count;) {
        const uint32_t nf = advect_group_size(adv, count - i, adv->fuse);
        int rc = advect_group(adv, nf, sink);
        if (rc) return rc;
        i += nf;
    }
    return advect_join(adv);
}

CLUMPY_API int clumpy_cuda_advect_step(clumpy_advect* adv, uint32_t count) {
    REQUIRE(adv, "adv is NULL");
    REQUIRE(!adv->route.on, "this run is routed across GPUs: use clumpy_cuda_advect_route_steps");
    CK(cudaSetDevice(adv->ctx->device));
    return advect_steps(adv, count, nullptr);
}

CLUMPY_API int clumpy_cuda_advect_record(clumpy_advect* adv, uint32_t count, uint8_t* dst_host, size_t frame_stride) {
    REQUIRE(adv && dst_host, "null argument");
    REQUIRE(!adv->route.on, "this run is routed across GPUs: use clumpy_cuda_advect_route_record");
    REQUIRE(frame_stride >= (size_t)adv->width * adv->height, "frame_stride is smaller than a frame");
    CK(cudaSetDevice(adv->ctx->device));
    RecordSink sink;
    sink.dst = dst_host; sink.frame_stride = frame_stride; sink.row0 = 0; sink.nrows = adv->height;
    return advect_steps(adv, count, &sink);
}

// The two stages of the pipeline, timed with CUDA events around each launch on the stream it runs on:
// `groups` groups (groups cut short by a regroup or the recording boundary are run but not counted).
// serial = 0: as the pipeline runs them (the composite of one group shares the GPU with the advect launch
// of the next); serial = 1: one stream, each kernel alone on the GPU.  Synchronises.
CLUMPY_API int clumpy_cuda_advect_profile_stages(clumpy_advect* adv, uint32_t groups, int serial, float* advect_ms,
                                                 float* composite_ms, uint32_t* frames_per_launch) {
    REQUIRE(adv && groups > 0 && advect_ms && composite_ms && frames_per_launch, "bad argument");
    REQUIRE(!adv->route.on, "this run is routed across GPUs: use clumpy_cuda_advect_route_profile");
    CK(cudaSetDevice(adv->ctx->device));
    std::vector<StageEvents> ev(groups);
    std::vector<bool> full(groups, false);
    int rc = CLUMPY_OK;
    for (auto& e : ev)
        for (cudaEvent_t* p : {&e.adv0, &e.adv1, &e.comp0, &e.comp1}) {
            *p = nullptr;
            if (rc == CLUMPY_OK && cudaEventCreate(p) != cudaSuccess) { set_error("event creation failed"); rc = CLUMPY_ERR_CUDA; }
        }
    for (uint32_t g = 0; rc == CLUMPY_OK && g < groups; ++g) {
        const uint32_t nf = advect_group_size(adv, (uint32_t)adv->fuse, adv->fuse);
        full[g] = nf == (uint32_t)adv->fuse && frame_splats(adv);
        rc = advect_group(adv, nf, nullptr, &ev[g], serial != 0);
    }
    if (rc == CLUMPY_OK) rc = advect_join(adv);
    if (rc == CLUMPY_OK && cudaStreamSynchronize(adv->ctx->stream) != cudaSuccess) { set_error("profile run failed"); rc = CLUMPY_ERR_CUDA; }
    double ta = 0, tc = 0;
    uint32_t counted = 0;
    for (uint32_t g = 0; rc == CLUMPY_OK && g < groups; ++g) {
        if (!full[g]) continue;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[g].adv0, ev[g].adv1); ta += ms;
        cudaEventElapsedTime(&ms, ev[g].comp0, ev[g].comp1); tc += ms;
        ++counted;
    }
    for (auto& e : ev) for (cudaEvent_t p : {e.adv0, e.adv1, e.comp0, e.comp1}) if (p) cudaEventDestroy(p);
    *advect_ms = counted ? (float)(ta / counted) : 0.f;
    *composite_ms = counted ? (float)(tc / counted) : 0.f;
    *frames_per_launch = (uint32_t)adv->fuse;
    return rc;
}

// ---- one frame split at the exchange point (multi-GPU, dense variant: counters summed by the caller) -------

CLUMPY_API int clumpy_cuda_advect_count(clumpy_advect* adv) {
    REQUIRE(adv, "adv is NULL");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    int rc = advect_join(adv);
    if (rc) return rc;
    REQUIRE(!adv->count_open, "advect_count called twice without advect_finish_frame");
    const bool splat = frame_splats(adv);
    // the private image of this interface is counter image 0: make sure it is clean (clumpy_cuda_advect_step
    // leaves the images of its last group to the next composite)
    for (int s = 0; s < NSETS; ++s) {
        for (int z = 0; z < adv->set_used[s]; ++z) CK(cudaMemsetAsync(adv->hist_buf[s * adv->fuse + z], 0, adv->hist_bytes, ctx->stream));
        adv->set_used[s] = 0;
    }
    for (int b = 0; b < LEGACY_RING; ++b)
        if (adv->clear_pending[b]) { CK(cudaStreamWaitEvent(ctx->stream, adv->cleared[b], 0)); adv->clear_pending[b] = false; }
    const HistRing ring = {adv->hist, adv->hist_stride / sizeof(uint32_t), 0, adv->nhist};
    rc = launch_advect(adv, adv->simframe % adv->nframes, splat, 1, ring);
    if (rc) return rc;
    adv->count_open = true;
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_cells(clumpy_advect* adv, void** cells_dev, size_t* bytes,
                                        uint32_t* pitch_cells, uint32_t* rows, uint32_t* halo,
                                        uint32_t* bytes_per_cell) {
    REQUIRE(adv, "adv is NULL");
    if (cells_dev) *cells_dev = adv->hist_buf[0];
    if (bytes) *bytes = adv->hist_bytes;
    if (pitch_cells) *pitch_cells = (uint32_t)adv->geom.pitch;
    if (rows) *rows = (uint32_t)adv->geom.rows;
    if (halo) *halo = (uint32_t)adv->geom.halo;
    if (bytes_per_cell) *bytes_per_cell = (uint32_t)cell_bytes_of(adv->cell_mode);
    return CLUMPY_OK;
}

// The band as an image of its own: same width and pitch, `nrows` rows, cells given by the caller.
static int composite_rows_impl(clumpy_advect* adv, cudaStream_t stream, const void* cells_dev, uint32_t row0, uint32_t nrows) {
    clumpy_cuda_ctx* ctx = adv->ctx;
    uint8_t* img = adv->imgs[adv->cur] + (size_t)row0 * adv->width;
    if (adv->cell_mode == CELL_F16) {
        GroupCompositeArgs ca;
        ca.width = adv->geom.width; ca.height = (int)nrows; ca.pitch = adv->geom.pitch; ca.img_pitch = (int)adv->width;
        ca.nf = 1; ca.cells[0] = cells_dev; ca.frame_out[0] = img; ca.img_in = img;
        return launch_composite_group(ctx, stream, adv->gtab, ca);
    }
    HistGeom g = adv->geom;
    g.height = (int)nrows;
    const cudaStream_t keep = ctx->stream;
    ctx->stream = stream; // the launchers enqueue on ctx->stream
    int rc;
    if (adv->cell_mode == CELL_X64) rc = launch_composite_exact(ctx, g, adv->rings, cells_dev, adv->wrapx, 1, adv->decay, img, img);
    else rc = launch_composite(ctx, g, adv->rings, reinterpret_cast<const uint32_t*>(cells_dev), adv->wrapx, 1, adv->decay, img, img);
    ctx->stream = keep;
    return rc;
}

CLUMPY_API int clumpy_cuda_advect_composite_rows(clumpy_advect* adv, const void* cells_dev,
                                                 uint32_t row0, uint32_t nrows) {
    REQUIRE(adv && cells_dev, "null argument");
    REQUIRE((uint64_t)row0 + nrows <= adv->height, "row band exceeds the image");
    CK(cudaSetDevice(adv->ctx->device));
    if (!frame_splats(adv) || nrows == 0) return CLUMPY_OK;
    return composite_rows_impl(adv, adv->ctx->stream, cells_dev, row0, nrows);
}

CLUMPY_API int clumpy_cuda_advect_finish_frame(clumpy_advect* adv) {
    REQUIRE(adv, "adv is NULL");
    REQUIRE(adv->count_open, "advect_finish_frame without advect_count");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    if (frame_splats(adv)) CK(cudaMemsetAsync(adv->hist_buf[0], 0, adv->hist_bytes, ctx->stream));
    adv->count_open = false;
    adv->simframe += 1;
    if (adv->regroup_every && ++adv->frames_since_regroup >= adv->regroup_every) return regroup_points(adv);
    return CLUMPY_OK;
}

// ---- multi-GPU, fused exchange: points routed to the band owners through peer memory -------------

CLUMPY_API int clumpy_cuda_ipc_export(clumpy_cuda_ctx* ctx, const void* dptr, void* handle64) {
    REQUIRE(ctx && dptr && handle64, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == CLUMPY_IPC_HANDLE_BYTES, "IPC handle size");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, const_cast<void*>(dptr)));
    memcpy(handle64, &h, sizeof(h));
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_ipc_open(clumpy_cuda_ctx* ctx, const void* handle64, void** dptr) {
    REQUIRE(ctx && dptr && handle64, "null argument");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    CK(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_ipc_close(clumpy_cuda_ctx* ctx, void* dptr) {
    REQUIRE(ctx, "ctx is NULL");
    CK(cudaSetDevice(ctx->device));
    if (dptr) CK(cudaIpcCloseMemHandle(dptr));
    return CLUMPY_OK;
}

// One past the last image row this rank composites: its band's rows, and for the last rank the rest.
static int route_last_row(const clumpy_advect* adv) {
    const auto& rt = adv->route;
    if (rt.rank == rt.world - 1) return (int)adv->height;
    return std::max(0, std::min((int)adv->height, (rt.rank + 1) * rt.chunk - adv->geom.halo));
}

static uint32_t* route_band(const clumpy_advect* adv, int set, int f) {
    return reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(adv->route.band) + (size_t)(set * ROUTE_MAX_FUSE + f) * adv->route.band_bytes);
}

template <bool WRAP, bool SPLAT, bool F16, typename OffT>
static void launch_route_t(clumpy_advect* a, const RouteDev& rd) {
    const uint32_t grid = ceil_div(a->npts, ROUTE_BLOCK); // CTA c owns region c of its segments
    const GeomDev g{a->geom.width, a->geom.height, a->geom.halo, a->geom.pitch, a->keep_pos ? 1 : 0, 0};
    // (6 CTAs of 40 registers per SM.  Measured and dropped: 8 CTAs of 32 registers -- spills, 0.0133 against 0.0125 ms per
    // frame -- and capping the CTAs per SM with unused shared memory to leave room for the side stream's launches.)
    advect_route_kernel<WRAP, SPLAT, F16, OffT><<<grid, ADVECT_THREADS, 0, a->ctx->stream>>>(
        a->pos, a->orig, (const OffT*)a->offset, a->npts, a->vel, a->width, a->height, a->step_size, g, rd);
}

static size_t route_inbox_layout(uint32_t capacity, int world, uint32_t* nregion, size_t* entries_per_frame, size_t* cnt_per_frame,
                                 size_t* cnt_offset, size_t* flags_offset) {
    const uint32_t nr = std::max<uint32_t>(ceil_div(capacity, ROUTE_BLOCK), 1u); // one region per CTA of the largest shard
    const size_t epf = (size_t)world * nr * ROUTE_REGION, cpf = (size_t)world * nr;
    const size_t co = (size_t)ROUTE_PARITIES * ROUTE_MAX_FUSE * epf;
    const size_t fo = round_up(co + (size_t)ROUTE_PARITIES * ROUTE_MAX_FUSE * cpf, 64);
    if (nregion) *nregion = nr;
    if (entries_per_frame) *entries_per_frame = epf;
    if (cnt_per_frame) *cnt_per_frame = cpf;
    if (cnt_offset) *cnt_offset = co;
    if (flags_offset) *flags_offset = fo;
    return round_up((fo + (size_t)ROUTE_MAX_WORLD) * sizeof(uint32_t), 256);
}

CLUMPY_API int clumpy_cuda_advect_route_bytes(uint32_t capacity, int world, size_t* inbox_bytes) {
    REQUIRE(inbox_bytes && world >= 1 && world <= ROUTE_MAX_WORLD, "bad argument");
    *inbox_bytes = route_inbox_layout(capacity, world, nullptr, nullptr, nullptr, nullptr, nullptr);
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_route_begin(clumpy_advect* adv, int rank, int world, uint32_t capacity,
                                              void** inbox_dev, size_t* inbox_bytes) {
    REQUIRE(adv && inbox_dev && inbox_bytes, "null argument");
    REQUIRE(world >= 1 && world <= ROUTE_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world size");
    REQUIRE(adv->cell_mode == CELL_F16 || adv->cell_mode == CELL_U32, "routing needs count-only cells (desc.cells = CLUMPY_CELLS_COUNT)");
    REQUIRE(capacity >= adv->npts, "capacity must be at least the largest shard");
    REQUIRE(!adv->route.on, "routing already set up");
    REQUIRE(adv->simframe == 0, "routing must be set up before the first frame");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    auto& rt = adv->route;
    rt.rank = rank; rt.world = world;
    const size_t ibytes = route_inbox_layout(capacity, world, &rt.nregion, &rt.entries_per_frame, &rt.cnt_per_frame, &rt.cnt_offset, &rt.flags_offset);
    // Bands of `chunk` padded rows, whole 16-row tiles, chosen from the IMAGE height so that for a uniform
    // cloud they coincide with the home-row strips of the shards (4096 rows on 8 ranks: 512 each, not the 528
    // that an even split of the 4098 padded rows would give: with those, rank 7 would start with a fifth of
    // its points in rank 6's band).  The last band also takes whatever is left (the 2 * halo padding rows),
    // which fits in the 32 slack rows of its band buffer.
    rt.chunk = (int)round_up(ceil_div(adv->height, (uint32_t)world), 16);
    REQUIRE(world == 1 || rt.chunk >= 2 * adv->geom.halo, "bands thinner than the sprite are not supported");
    if (adv->overlap_opt < 0) adv->overlap = true; // the receiving half waits for other GPUs: keep it off the main stream
    rt.fuse = std::min(std::max(env_int("CLUMPY_ROUTE_FUSE", adv->cell_mode == CELL_F16 ? ROUTE_MAX_FUSE : 4), 1), ROUTE_MAX_FUSE);
    rt.timeout_ns = (unsigned long long)std::max(1, env_int("CLUMPY_ROUTE_TIMEOUT_MS", 2000)) * 1000000ull;
    // the regroup period is part of the protocol (every rank cuts its groups at the same frames): whole groups
    if (adv->regroup_every) adv->regroup_every = std::max<uint32_t>((uint32_t)rt.fuse, adv->regroup_every / (uint32_t)rt.fuse * (uint32_t)rt.fuse);
    const size_t cell_bytes = cell_bytes_of(adv->cell_mode);
    const size_t band_rows = (size_t)rt.chunk + 2 * adv->geom.halo + 32;
    rt.band_cells = (uint32_t)(band_rows * adv->geom.pitch);
    rt.band_bytes = round_up((size_t)rt.band_cells * cell_bytes, 256);
    trace_mark(ctx, "route_begin: start");
    if (*inbox_dev) { // the caller's arena from an earlier run (its peers still have it mapped)
        REQUIRE(*inbox_bytes >= ibytes, "the inbox arena passed in is too small for this run");
        rt.inbox = static_cast<uint32_t*>(*inbox_dev);
        rt.inbox_owned = false;
    } else {
        CK(cudaMalloc(&rt.inbox, ibytes)); // mapped by the peers: a plain allocation, not the pool
    }
    trace_mark(ctx, "route_begin: inbox allocation");
    DEV_ALLOC(ctx, &rt.band, (size_t)RSETS * ROUTE_MAX_FUSE * rt.band_bytes);
    for (int s = 0; s < RSETS; ++s) CK(cudaEventCreateWithFlags(&rt.composited[s], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&rt.routed, cudaEventDisableTiming));
    CK(cudaHostAlloc(&rt.errors, sizeof(uint32_t), cudaHostAllocMapped));
    *rt.errors = 0u;
    CK(cudaHostGetDevicePointer(&rt.errors_dev, rt.errors, 0));
    DEV_ALLOC(ctx, &rt.ticket, sizeof(uint32_t));
    CK(cudaMemsetAsync(rt.ticket, 0, sizeof(uint32_t), ctx->stream));
    // count words and flags start at zero (the entries need no initialisation)
    CK(cudaMemsetAsync(rt.inbox + rt.cnt_offset, 0, ibytes - rt.cnt_offset * sizeof(uint32_t), ctx->stream));
    CK(cudaMemsetAsync(rt.band, 0, (size_t)RSETS * ROUTE_MAX_FUSE * rt.band_bytes, ctx->stream));
    // Load every kernel of the frame loop NOW.  With lazy module loading the first launch of a kernel
    // can block the host until the device is idle -- and a drain kernel may be sitting on the device
    // waiting for the epoch of a rank that this very host thread has yet to launch (one process
    // driving several ranks).  The drain and the composite run once on the all-zero band, which leaves
    // the (zero) framebuffer as it is; the route kernels are loaded without being run.
    {
        cudaFuncAttributes fa;
#define ROUTE_LOAD2(W, S, F) CK(adv->off16 ? cudaFuncGetAttributes(&fa, advect_route_kernel<W, S, F, uint16_t>) \
                                          : cudaFuncGetAttributes(&fa, advect_route_kernel<W, S, F, uint32_t>))
#define ROUTE_LOAD(W, S) do { if (adv->cell_mode == CELL_F16) ROUTE_LOAD2(W, S, true); else ROUTE_LOAD2(W, S, false); } while (0)
        if (adv->wrapx) { ROUTE_LOAD(true, true); ROUTE_LOAD(true, false); }
        else            { ROUTE_LOAD(false, true); ROUTE_LOAD(false, false); }
#undef ROUTE_LOAD
#undef ROUTE_LOAD2
        CK(cudaFuncGetAttributes(&fa, route_signal_kernel));
        DrainDev dd = {};
        dd.inbox = rt.inbox; dd.cnt = rt.inbox + rt.cnt_offset;
        dd.inbox_frame_stride = rt.entries_per_frame; dd.cnt_frame_stride = rt.cnt_per_frame;
        dd.nregion = rt.nregion; dd.world = world; dd.nf = 1; dd.flags = nullptr; dd.errors = rt.errors_dev;
        dd.band[0] = rt.band; dd.band_cells = rt.band_cells;
        dd.wrap_w = 0; dd.halo = adv->geom.halo; dd.pitch = adv->geom.pitch;
        if (adv->cell_mode == CELL_F16) drain_group_kernel<CELL_F16><<<1, DRAIN_THREADS, 0, ctx->stream>>>(dd);
        else drain_group_kernel<CELL_U32><<<1, DRAIN_THREADS, 0, ctx->stream>>>(dd);
        CK_LAUNCH(ctx);
        const int h = adv->geom.halo, lo = rank * rt.chunk;
        const int y0 = std::max(0, lo - h), y1 = route_last_row(adv);
        if (y1 > y0) {
            const char* cells = reinterpret_cast<const char*>(rt.band) + (size_t)(y0 - (lo - h)) * adv->geom.pitch * cell_bytes;
            int rc = composite_rows_impl(adv, ctx->stream, cells, (uint32_t)y0, (uint32_t)(y1 - y0));
            if (rc) return rc;
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    trace_mark(ctx, "route_begin: band buffers, clears, kernels loaded");
    rt.peer_inbox[rank] = rt.inbox;
    // Profiling aid (ncu must not run a multi-rank job): with CLUMPY_ROUTE_LOOPBACK=1 one process plays
    // rank `rank` of `world` alone.  The caller maps its own inbox as every peer's; the fence then
    // publishes all `world` epochs into its own flags.  Timing and memory traffic are about those of a
    // real rank, the FRAMES ARE NOT (entries for other bands pile up in this rank's inbox).
    rt.loopback = env_int("CLUMPY_ROUTE_LOOPBACK", 0) == 1;
    rt.on = true;
    *inbox_dev = rt.inbox;
    if (rt.inbox_owned) *inbox_bytes = ibytes;
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_route_peers(clumpy_advect* adv, void* const* peer_inbox) {
    REQUIRE(adv && peer_inbox && adv->route.on, "routing is not set up");
    for (int r = 0; r < adv->route.world; ++r)
        adv->route.peer_inbox[r] = (r == adv->route.rank) ? adv->route.inbox : static_cast<uint32_t*>(peer_inbox[r]);
    for (int r = 0; r < adv->route.world; ++r) REQUIRE(adv->route.peer_inbox[r], "missing peer inbox");
    return CLUMPY_OK;
}

struct RouteEvents { cudaEvent_t r0, r1, d0, d1, c1; bool keep_overlap = false; };

// One group of `nf` frames of the routed run has a SENDING half and a RECEIVING half:
//
//   route_front   main stream: advect_route_kernel -- Euler step + reset of this rank's points through the group;
//                 own-band points are counted into band buffer (set, f), the others' cells are stored into their
//                 owners' inboxes (parity); at its end the group's epoch goes to every peer
//   route_back    side stream: drain_group_kernel (waits for every peer's epoch, counts the inbox into the band
//                 buffers), then the group composite of this rank's image rows, which also zeroes the band buffers
//                 of the group before
//
// What may be reused when (G = this group; "its" = this rank's):
//   band set G % 3 was read by composite(G - 3) and zeroed by composite(G - 2): the route launch waits for
//     its composite(G - 2);
//   inbox parity G % 4 was last written for group G - 4 and read by the PEERS' drain(G - 4).  Its composite(G - 2)
//     is behind its drain(G - 2), which saw every peer's epoch G - 2, i.e. every peer's route(G - 2) had
//     run, which that peer launched behind ITS composite(G - 4) and therefore behind its drain(G - 4).  So the
//     same wait covers the inbox: four parities are enough and no peer is ever waited for on the host.
//
// Measured and dropped in round 2 (2 ranks on a quarter of C4 = the per-rank load of 8 GPUs; CLUMPY_ROUTE_TIMELINE
// prints where the launches sit): more band sets and parities, so that route(G) no longer waits for composite(G - 2)
// -- 0.0193 ms per frame, the same; the drain of group G riding on route(G + 2) behind a one-CTA wait launch, with
// the composite alone on the side stream -- 0.0205 (20 frames: 0.0197 against 0.0174).  The three launches of a
// group cost 95 + 40 + 57 us alone and about 150 us together however they are interleaved: the SMs are shared,
// not idle.
struct RouteBack { uint32_t nf = 0; int set = 0, parity = 0; uint32_t epoch = 0; };

static DrainDev route_drain_desc(const clumpy_advect* adv, const RouteBack& b) {
    const auto& rt = adv->route;
    DrainDev dd = {};
    dd.inbox = rt.inbox + (size_t)b.parity * ROUTE_MAX_FUSE * rt.entries_per_frame;
    dd.cnt = rt.inbox + rt.cnt_offset + (size_t)b.parity * ROUTE_MAX_FUSE * rt.cnt_per_frame;
    dd.inbox_frame_stride = rt.entries_per_frame; dd.cnt_frame_stride = rt.cnt_per_frame;
    dd.nregion = rt.nregion; dd.world = rt.world; dd.nf = (int)b.nf;
    dd.flags = rt.inbox + rt.flags_offset;
    dd.epoch = b.epoch; dd.errors = rt.errors_dev; dd.timeout_ns = rt.timeout_ns;
    for (uint32_t f = 0; f < b.nf; ++f) dd.band[f] = route_band(adv, b.set, (int)f);
    dd.band_cells = rt.band_cells;
    dd.wrap_w = adv->wrapx ? (int)adv->width : 0; dd.halo = adv->geom.halo; dd.pitch = adv->geom.pitch;
    return dd;
}

// The sending half.  *back describes this group's receiving half (nf == 0: the group does not splat, nothing to receive).
static int route_front(clumpy_advect* adv, uint32_t nf, RouteEvents* ev, RouteBack* back) {
    clumpy_cuda_ctx* ctx = adv->ctx;
    auto& rt = adv->route;
    const bool splat = frame_splats(adv);
    const cudaStream_t main_stream = ctx->stream;
    const int set = (int)(rt.group_index % RSETS), parity = (int)(rt.group_index % ROUTE_PARITIES);
    if (splat) {
        const int z = (set + 1) % RSETS; // the set of group G - 2, whose composite zeroed the set of group G - 3: this one
        if (rt.composited_pending[z]) {
            CK(cudaStreamWaitEvent(main_stream, rt.composited[z], 0));
            rt.composited_pending[z] = false;
        }
    }
    RouteDev rd = {};
    rd.me = rt.rank; rd.world = rt.world; rd.chunk = rt.chunk; rd.halo = adv->geom.halo; rd.pitch = adv->geom.pitch;
    rd.nregion = rt.nregion;
    rd.nframes = adv->nframes; rd.phase0 = adv->simframe % adv->nframes;
    rd.nf = (int)nf;
    rd.inbox_off = (size_t)parity * ROUTE_MAX_FUSE * rt.entries_per_frame;
    rd.inbox_frame_stride = rt.entries_per_frame;
    rd.cnt_off = (size_t)parity * ROUTE_MAX_FUSE * rt.cnt_per_frame;
    rd.cnt_frame_stride = rt.cnt_per_frame;
    for (uint32_t f = 0; f < nf; ++f) rd.band[f] = route_band(adv, set, (int)f);
    rd.band_f16 = adv->cell_mode == CELL_F16 ? 1 : 0;
    {
        const int my_lo = rt.rank * rt.chunk, gh = adv->geom.halo;
        rd.interior = (uint32_t)std::max(rt.chunk - 2 * rd.halo, 0);
        rd.in_row0 = my_lo + rd.halo - gh;
        rd.in_bias = (gh - my_lo + rd.halo) * rd.pitch + gh;
    }
    rd.epoch = splat ? rt.epoch + 1u : rt.epoch;
    rd.loopback = rt.loopback ? 1 : 0;
    rd.ticket = rt.ticket;
    for (int r = 0; r < rt.world; ++r) {
        rd.inbox[r] = rt.peer_inbox[r];
        rd.cnt[r] = rt.peer_inbox[r] + rt.cnt_offset;
        rd.flags[r] = rt.peer_inbox[r] + rt.flags_offset;
    }
    if (ev) CK(cudaEventRecord(ev->r0, main_stream));
    if (adv->npts) {
#define ROUTE_LAUNCH(W, S) (adv->cell_mode == CELL_F16 ? (adv->off16 ? launch_route_t<W, S, true, uint16_t>(adv, rd) : launch_route_t<W, S, true, uint32_t>(adv, rd)) \
                                                       : (adv->off16 ? launch_route_t<W, S, false, uint16_t>(adv, rd) : launch_route_t<W, S, false, uint32_t>(adv, rd)))
        if (adv->wrapx) { if (splat) ROUTE_LAUNCH(true, true); else ROUTE_LAUNCH(true, false); }
        else            { if (splat) ROUTE_LAUNCH(false, true); else ROUTE_LAUNCH(false, false); }
#undef ROUTE_LAUNCH
        CK_LAUNCH(ctx);
    } else if (splat) { // no route kernel to carry the fence
        route_signal_kernel<<<1, 32, 0, main_stream>>>(rd);
        CK_LAUNCH(ctx);
    }
    if (ev) CK(cudaEventRecord(ev->r1, main_stream));
    *back = RouteBack();
    if (splat) {
        rt.epoch += 1u;
        back->nf = nf; back->set = set; back->parity = parity; back->epoch = rt.epoch;
        rt.group_index++;
    }
    adv->simframe += nf;
    adv->frames_since_regroup += nf;
    if (adv->regroup_every && adv->frames_since_regroup >= adv->regroup_every) return regroup_points(adv);
    return CLUMPY_OK;
}

// The receiving half of group `b`.
static int route_back(clumpy_advect* adv, const RouteBack& b, bool overlap, RecordSink* sink, RouteEvents* ev) {
    clumpy_cuda_ctx* ctx = adv->ctx;
    auto& rt = adv->route;
    const cudaStream_t main_stream = ctx->stream;
    const cudaStream_t side = overlap ? ctx->aux_stream : main_stream;
    const uint32_t nf = b.nf;
    const int set = b.set;
    int slot[ADVECT_MAX_FUSE];
    if (overlap) {
        // the receiving half starts once this rank's own route kernel is done (by then its own epoch is
        // out; a drain that started earlier would only sit on SM slots waiting for it)
        CK(cudaEventRecord(rt.routed, main_stream));
        CK(cudaStreamWaitEvent(side, rt.routed, 0));
    }
    if (ev) CK(cudaEventRecord(ev->d0, side));
    {
        const DrainDev dd = route_drain_desc(adv, b);
        // 16 (frame, region) pairs per warp; the CTAs wait (one thread each) for the peers' epochs before they work
        constexpr uint32_t warps = DRAIN_THREADS / 32;
        const uint32_t dgrid = std::min<uint32_t>((uint32_t)ctx->sm_count * 16u, std::max<uint32_t>(1u, ceil_div((uint32_t)(nf * rt.cnt_per_frame), 16u * warps)));
        if (adv->cell_mode == CELL_F16) drain_group_kernel<CELL_F16><<<dgrid, DRAIN_THREADS, 0, side>>>(dd);
        else drain_group_kernel<CELL_U32><<<dgrid, DRAIN_THREADS, 0, side>>>(dd);
        CK_LAUNCH(ctx);
    }
    if (ev) CK(cudaEventRecord(ev->d1, side));
    // image rows this rank owns: those whose window of padded rows lies in chunk + halos
    const int h = adv->geom.halo, lo = rt.rank * rt.chunk;
    const int y0 = std::max(0, lo - h), y1 = route_last_row(adv);
    const size_t cell_bytes = cell_bytes_of(adv->cell_mode);
    const size_t row_off = (size_t)(y0 - (lo - h)) * adv->geom.pitch * cell_bytes;
    const int prev = (set + RSETS - 1) % RSETS;
    int rc = choose_frame_images(adv, side, nf, sink != nullptr, slot);
    if (rc) return rc;
    if (adv->cell_mode == CELL_F16) {
        GroupCompositeArgs ca;
        ca.width = adv->geom.width; ca.height = std::max(0, y1 - y0); ca.pitch = adv->geom.pitch; ca.img_pitch = (int)adv->width;
        ca.nf = (int)nf;
        ca.img_in = adv->imgs[adv->cur] + (size_t)y0 * adv->width;
        for (uint32_t f = 0; f < nf; ++f) {
            ca.cells[f] = reinterpret_cast<const char*>(route_band(adv, set, (int)f)) + row_off;
            ca.frame_out[f] = slot[f] >= 0 ? adv->imgs[slot[f]] + (size_t)y0 * adv->width : nullptr;
        }
        ca.alone = !overlap; // (beside the next group's route launch otherwise)
        ca.nzero = rt.set_used[prev];
        for (int zf = 0; zf < ca.nzero; ++zf) {
            ca.zero[zf] = reinterpret_cast<uint4*>(route_band(adv, prev, zf));
            ca.zero_n16[zf] = (uint32_t)(rt.band_bytes / 16);
        }
        if (ca.height <= 0) { // a rank without image rows still has band buffers to zero
            for (int zf = 0; zf < ca.nzero; ++zf) CK(cudaMemsetAsync(ca.zero[zf], 0, rt.band_bytes, side));
        } else {
            rc = launch_composite_group(ctx, side, adv->gtab, ca);
            if (rc) return rc;
        }
    } else {
        for (int zf = 0; zf < rt.set_used[prev]; ++zf) CK(cudaMemsetAsync(route_band(adv, prev, zf), 0, rt.band_bytes, side));
        for (uint32_t f = 0; f < nf && y1 > y0; ++f) {
            const int dst = slot[f] >= 0 ? slot[f] : adv->cur;
            HistGeom g = adv->geom;
            g.height = y1 - y0;
            ctx->stream = side;
            rc = launch_composite(ctx, g, adv->rings, reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(route_band(adv, set, (int)f)) + row_off),
                                  adv->wrapx, 1, adv->decay, adv->imgs[adv->cur] + (size_t)y0 * adv->width, adv->imgs[dst] + (size_t)y0 * adv->width);
            ctx->stream = main_stream;
            if (rc) return rc;
            adv->cur = dst;
        }
    }
    rt.set_used[prev] = 0;
    rt.set_used[set] = (int)nf;
    adv->cur = slot[nf - 1];
    CK(cudaEventRecord(rt.composited[set], side));
    rt.composited_pending[set] = overlap;
    if (ev) CK(cudaEventRecord(ev->c1, side));
    if (overlap) adv->side_dirty = true;
    if (sink) return enqueue_frame_copies(adv, rt.composited[set], slot, nf, sink);
    return CLUMPY_OK;
}

static int route_steps_impl(clumpy_advect* adv, uint32_t count, RecordSink* sink, RouteEvents* ev) {
    REQUIRE(adv && adv->route.on, "routing is not set up");
    REQUIRE(!adv->count_open, "advect_route_steps inside an open frame");
    int rc = route_check(adv);
    if (rc) return rc;
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    auto& rt = adv->route;
    for (int r = 0; r < rt.world; ++r) REQUIRE(rt.peer_inbox[r], "peer inboxes are not mapped (route_peers)");
    const bool keep_overlap = ev && ev->keep_overlap;
    const bool overlap = adv->overlap && (!ev || keep_overlap);
    for (uint32_t i = 0; i < count;) {
        // a group never crosses a regroup or the warm-up / recording boundary (every rank cuts its groups at
        // the same frames: the group structure is part of the protocol)
        const uint32_t nf = advect_group_size(adv, count - i, rt.fuse);
        RouteBack mine;
        rc = route_front(adv, nf, ev, &mine);
        if (rc) return rc;
        if (mine.nf) {
            rc = route_back(adv, mine, overlap, sink, ev);
            if (rc) return rc;
        } else {
            // a group that does not splat leaves the framebuffer as it is
            if (ev) for (cudaEvent_t e : {ev->d0, ev->d1, ev->c1}) CK(cudaEventRecord(e, ctx->stream));
            if (sink) {
                int slot[ADVECT_MAX_FUSE];
                for (uint32_t f = 0; f < nf; ++f) slot[f] = adv->cur;
                CK(cudaEventRecord(adv->frames_ready, overlap ? ctx->aux_stream : ctx->stream));
                rc = enqueue_frame_copies(adv, adv->frames_ready, slot, nf, sink);
                if (rc) return rc;
            }
        }
        i += nf;
        if (ev) break; // profiling: one group per call
    }
    if (keep_overlap) return CLUMPY_OK; // (the timeline joins once, at its end)
    return advect_join(adv); // whatever the caller enqueues next sees the finished framebuffer
}

CLUMPY_API int clumpy_cuda_advect_route_steps(clumpy_advect* adv, uint32_t count) {
    return route_steps_impl(adv, count, nullptr, nullptr);
}

CLUMPY_API int clumpy_cuda_advect_route_record(clumpy_advect* adv, uint32_t count, uint8_t* dst_host, size_t frame_stride) {
    REQUIRE(adv && dst_host && adv->route.on, "routing is not set up");
    const int h = adv->geom.halo, lo = adv->route.rank * adv->route.chunk;
    const int y0 = std::max(0, lo - h), y1 = std::max(y0, route_last_row(adv));
    REQUIRE(frame_stride >= (size_t)(y1 - y0) * adv->width, "frame_stride is smaller than this rank's band");
    RecordSink sink;
    sink.dst = dst_host; sink.frame_stride = frame_stride; sink.row0 = (uint32_t)y0; sink.nrows = (uint32_t)(y1 - y0);
    return route_steps_impl(adv, count, &sink, nullptr);
}

// `groups` groups of the routed run, launched one by one with CUDA events around every kernel: mean device
// time per GROUP (ms) of the route kernel, the drain kernel (which includes the wait for the slowest peer)
// and the band composite, and the frames per group.  Every rank must call it with the same arguments.
// Synchronises.
CLUMPY_API int clumpy_cuda_advect_route_profile(clumpy_advect* adv, uint32_t groups, float* route_ms,
                                                float* drain_ms, float* composite_ms, uint32_t* frames_per_group) {
    REQUIRE(adv && groups > 0 && adv->route.on, "bad argument");
    CK(cudaSetDevice(adv->ctx->device));
    std::vector<RouteEvents> ev(groups);
    int rc = CLUMPY_OK;
    for (auto& e : ev)
        for (cudaEvent_t* p : {&e.r0, &e.r1, &e.d0, &e.d1, &e.c1}) {
            *p = nullptr;
            if (rc == CLUMPY_OK && cudaEventCreate(p) != cudaSuccess) { set_error("event creation failed"); rc = CLUMPY_ERR_CUDA; }
        }
    // CLUMPY_ROUTE_TIMELINE=1 (development aid): leave the two streams as they run in production and print where
    // every launch sat; the three means then overlap each other and do not add up to the step
    const bool timeline = env_int("CLUMPY_ROUTE_TIMELINE", 0) == 1;
    for (auto& e : ev) e.keep_overlap = timeline;
    std::vector<uint32_t> nfs(groups, 0);
    for (uint32_t g = 0; rc == CLUMPY_OK && g < groups; ++g) {
        nfs[g] = frame_splats(adv) ? advect_group_size(adv, (uint32_t)adv->route.fuse, adv->route.fuse) : 0u; // 0: not counted
        rc = route_steps_impl(adv, (uint32_t)adv->route.fuse, nullptr, &ev[g]);
    }
    if (rc == CLUMPY_OK && timeline) rc = advect_join(adv);
    if (rc == CLUMPY_OK && cudaStreamSynchronize(adv->ctx->stream) != cudaSuccess) { set_error("profile run failed"); rc = CLUMPY_ERR_CUDA; }
    double t[3] = {0, 0, 0};
    uint32_t counted = 0;
    for (uint32_t g = 0; rc == CLUMPY_OK && g < groups; ++g) {
        if (nfs[g] != (uint32_t)adv->route.fuse) continue;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[g].r0, ev[g].r1); t[0] += ms;
        cudaEventElapsedTime(&ms, ev[g].d0, ev[g].d1); t[1] += ms;
        cudaEventElapsedTime(&ms, ev[g].d1, ev[g].c1); t[2] += ms;
        ++counted;
    }
    if (rc == CLUMPY_OK && timeline) {
        // development aid: where every launch of the last groups sat on this rank's clock (us since the first event)
        for (uint32_t g = groups > 6 ? groups - 6 : 0; g < groups; ++g) {
            float at[5] = {0, 0, 0, 0, 0};
            const cudaEvent_t es[5] = {ev[g].r0, ev[g].r1, ev[g].d0, ev[g].d1, ev[g].c1};
            for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&at[i], ev[0].r0, es[i]);
            fprintf(stderr, "[timeline rank %d group %u nf %u] route %.0f..%.0f  drain %.0f..%.0f  composite ..%.0f us\n", adv->route.rank, g,
                    nfs[g], at[0] * 1e3, at[1] * 1e3, at[2] * 1e3, at[3] * 1e3, at[4] * 1e3);
        }
    }
    for (auto& e : ev) for (cudaEvent_t p : {e.r0, e.r1, e.d0, e.d1, e.c1}) if (p) cudaEventDestroy(p);
    if (route_ms) *route_ms = counted ? (float)(t[0] / counted) : 0.f;
    if (drain_ms) *drain_ms = counted ? (float)(t[1] / counted) : 0.f;
    if (composite_ms) *composite_ms = counted ? (float)(t[2] / counted) : 0.f;
    if (frames_per_group) *frames_per_group = (uint32_t)adv->route.fuse;
    if (rc == CLUMPY_OK) rc = route_check(adv);
    return rc;
}

// Number of fence time-outs so far (a peer that never published its epoch).  Synchronises.
CLUMPY_API int clumpy_cuda_advect_route_errors(clumpy_advect* adv, uint32_t* errors) {
    REQUIRE(adv && errors && adv->route.on, "routing is not set up");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    for (cudaStream_t s : {ctx->stream, ctx->aux_stream}) CK(cudaStreamSynchronize(s));
    *errors = *reinterpret_cast<volatile uint32_t*>(adv->route.errors);
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_owned_rows(clumpy_advect* adv, uint32_t* row0, uint32_t* nrows) {
    REQUIRE(adv && row0 && nrows && adv->route.on, "routing is not set up");
    const int h = adv->geom.halo, lo = adv->route.rank * adv->route.chunk;
    const int y0 = std::max(0, lo - h), y1 = route_last_row(adv);
    *row0 = (uint32_t)y0;
    *nrows = (uint32_t)std::max(0, y1 - y0);
    return CLUMPY_OK;
}

// ---- reading results ------------------------------------------------------------------------------------------

CLUMPY_API int clumpy_cuda_advect_image_dev(clumpy_advect* adv, void** img_dev) {
    REQUIRE(adv && img_dev, "null argument");
    *img_dev = adv->imgs[adv->cur];
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_simframe(clumpy_advect* adv, uint32_t* simframe) {
    REQUIRE(adv && simframe, "null argument");
    *simframe = adv->simframe;
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_npts(clumpy_advect* adv, uint32_t* npts) {
    REQUIRE(adv && npts, "null argument");
    *npts = adv->npts;
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_frame_async(clumpy_advect* adv, uint8_t* dst_host) {
    REQUIRE(adv && dst_host, "null argument");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    const int b = adv->cur;
    CK(cudaEventRecord(adv->frame_ready, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, adv->frame_ready, 0));
    CK(cudaMemcpyAsync(dst_host, adv->imgs[b], (size_t)adv->width * adv->height, cudaMemcpyDeviceToHost, ctx->copy_stream));
    CK(cudaEventRecord(adv->copy_done[b], ctx->copy_stream));
    adv->copy_pending[b] = true;
    adv->copy_seq[b] = ++adv->copy_counter;
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_wait_frames(clumpy_advect* adv) {
    REQUIRE(adv, "adv is NULL");
    CK(cudaSetDevice(adv->ctx->device));
    CK(cudaStreamSynchronize(adv->ctx->copy_stream));
    for (int i = 0; i < MAX_IMAGES; ++i) adv->copy_pending[i] = false;
    return route_check(adv);
}

CLUMPY_API int clumpy_cuda_advect_read_points(clumpy_advect* adv, float* out_host) {
    REQUIRE(adv && (out_host || adv->npts == 0), "null argument");
    REQUIRE(!adv->sharded, "this run holds a shard of the caller's points: use clumpy_cuda_advect_read_points_indexed");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    if (adv->npts == 0) return CLUMPY_OK;
    float2* tmp = nullptr;
    DEV_ALLOC(ctx, &tmp, (size_t)adv->npts * 8);
    unpermute_kernel<<<ceil_div(adv->npts, 256), 256, 0, ctx->stream>>>(adv->pos, adv->perm, adv->npts, tmp);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, tmp, (size_t)adv->npts * 8, cudaMemcpyDeviceToHost, ctx->stream);
    dev_release(ctx, tmp);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { set_error("read_points failed: %s", cudaGetErrorString(e)); return CLUMPY_ERR_CUDA; }
    return CLUMPY_OK;
}

// Positions of this run's points in storage order together with the index each had in the caller's list
// (shards of a multi-GPU run: the caller scatters pos_host[i] to its place index_host[i]).
CLUMPY_API int clumpy_cuda_advect_read_points_indexed(clumpy_advect* adv, float* pos_host, uint32_t* index_host) {
    REQUIRE(adv && ((pos_host && index_host) || adv->npts == 0), "null argument");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    if (adv->npts == 0) return CLUMPY_OK;
    CK(cudaMemcpyAsync(pos_host, adv->pos, (size_t)adv->npts * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(index_host, adv->perm, (size_t)adv->npts * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_read_image(clumpy_advect* adv, uint8_t* out_host) {
    REQUIRE(adv && out_host, "null argument");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out_host, adv->imgs[adv->cur], (size_t)adv->width * adv->height, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return route_check(adv);
}

CLUMPY_API int clumpy_cuda_advect_end(clumpy_advect* adv) {
    if (!adv) return CLUMPY_OK;
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
    trace_mark(ctx, "advect_points: start");
    clumpy_advect* a = nullptr;
    int rc = clumpy_cuda_advect_begin(ctx, pts_host, npts, vel_host, width, height, step_size,
                                      kernel_size, decay, wrapx, nframes, age_offset, &a);
    if (rc) return rc;
    const size_t npix = (size_t)width * height;
    // Recorded frames go from the framebuffer ring to a small ring of pinned buffers (kept by the context
    // between commands) on the copy stream, and from there to the caller, in order, while the GPU simulates.
    constexpr int NPIN = 3;
    uint8_t* pinned[NPIN] = {};
    cudaEvent_t landed[NPIN] = {};
    auto cleanup = [&]() {
        clumpy_cuda_advect_end(a);
        for (int b = 0; b < NPIN; ++b) { ctx_pinned_put(ctx, pinned[b], npix); if (landed[b]) cudaEventDestroy(landed[b]); }
    };
    for (int b = 0; b < NPIN; ++b) {
        pinned[b] = static_cast<uint8_t*>(ctx_pinned_get(ctx, npix));
        if (!pinned[b]) { cleanup(); return CLUMPY_ERR_NOMEM; }
        if (cudaEventCreateWithFlags(&landed[b], cudaEventDisableTiming) != cudaSuccess) { cleanup(); set_error("event creation failed"); return CLUMPY_ERR_CUDA; }
    }
    trace_mark(ctx, "advect_points: begin done, pinned frame buffers");

    rc = clumpy_cuda_advect_step(a, nframes); // warm-up frames (advect_points.cc:139-158)
    trace_mark(ctx, "advect_points: warm-up frames");
    uint32_t issued = 0, delivered = 0; // recorded frames whose copy was enqueued / handed to the caller
    auto deliver_one = [&]() -> int {
        const int b = (int)(delivered % NPIN);
        if (cudaEventSynchronize(landed[b]) != cudaSuccess) { set_error("frame copy failed"); return CLUMPY_ERR_CUDA; }
        if (on_frame(user, delivered, pinned[b]) != 0) { set_error("frame callback aborted the run"); return CLUMPY_ERR_STATE; }
        ++delivered;
        return CLUMPY_OK;
    };
    RecordSink sink;
    sink.external = true; // advect_group reports the images; the copies are issued here, one pinned buffer at a time
    while (rc == CLUMPY_OK && issued < nframes) {
        // one group of frames, every frame in an image of its own ...
        const uint32_t nf = advect_group_size(a, nframes - issued, a->fuse);
        rc = advect_group(a, nf, &sink);
        // ... each copied to the next pinned buffer as soon as the caller has consumed what was in it
        for (uint32_t f = 0; rc == CLUMPY_OK && f < nf; ++f) {
            while (rc == CLUMPY_OK && issued - delivered >= (uint32_t)NPIN) rc = deliver_one();
            if (rc) break;
            const int b = (int)(issued % NPIN), img = sink.slots[f];
            cudaError_t e = cudaMemcpyAsync(pinned[b], a->imgs[img], npix, cudaMemcpyDeviceToHost, ctx->copy_stream);
            if (e == cudaSuccess) e = cudaEventRecord(landed[b], ctx->copy_stream);
            if (e == cudaSuccess) e = cudaEventRecord(a->copy_done[img], ctx->copy_stream); // the image is free again
            if (e != cudaSuccess) { set_error("frame copy failed: %s", cudaGetErrorString(e)); rc = CLUMPY_ERR_CUDA; break; }
            ++issued;
        }
    }
    if (rc == CLUMPY_OK) rc = advect_join(a);
    while (rc == CLUMPY_OK && delivered < issued) rc = deliver_one();
    trace_mark(ctx, "advect_points: recorded frames + copies + callbacks");
    cleanup();
    trace_mark(nullptr, "advect_points: release");
    return rc;
}
