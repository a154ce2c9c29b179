/*
 * clumpy_cuda.h -- C ABI of libclumpy_cuda.so, the B200 (sm_100a) implementation of clumpy's
 * point-advection hot path:  generate_simplex -> curl_2d -> advect_points (+ splat_points).
 *
 * This is the drop-in boundary.  A clumpy command's `exec` (clumpy_command.hh:8-29) keeps its
 * argument parsing, validation messages and .npy I/O on the host and calls ONE of the entry
 * points below instead of its CPU loop.  Each entry point names the reference loop it replaces
 * (paths relative to the clumpy source root).  See INTEGRATION.md for the binding a maintainer
 * would add on the reference side.
 *
 * Conventions
 *   - plain C types only; the caller owns every host pointer, the library owns device memory;
 *   - every function returns CLUMPY_OK (0) or a CLUMPY_ERR_* code and never throws;
 *     clumpy_cuda_last_error() returns the message of the calling thread's last failure;
 *   - there is NO CPU fallback: without a usable CUDA device clumpy_cuda_create fails;
 *   - `*_dev` variants take DEVICE pointers (memory resident in HBM, e.g. from
 *     clumpy_cuda_dev_alloc) and enqueue work on the context's stream without synchronising
 *     unless a host result (min/max) is requested;
 *   - images are row-major (H, W[, 2]); points are (N, 2) float32 = (x, y) pairs;
 *   - host buffers may be pageable or pinned (clumpy_cuda_host_alloc); pinned ones make the
 *     copies asynchronous and run at PCIe speed.
 */
#ifndef CLUMPY_CUDA_H
#define CLUMPY_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLUMPY_OK 0
#define CLUMPY_ERR_INVALID 1 /* bad argument (even kernel size, null pointer, zero dims ...) */
#define CLUMPY_ERR_CUDA 2    /* a CUDA runtime call or kernel failed                        */
#define CLUMPY_ERR_NOMEM 3   /* host or device allocation failed                            */
#define CLUMPY_ERR_STATE 4   /* call made in the wrong order (e.g. step after the last frame) */

#define CLUMPY_MAX_KERNEL_SIZE 31 /* largest odd splat kernel the device path accepts */

typedef struct clumpy_cuda_ctx clumpy_cuda_ctx; /* one device + its streams and scratch */
typedef struct clumpy_advect clumpy_advect;     /* device-resident state of one advect_points run */

/* ---- library / context ---------------------------------------------------------------- */

const char* clumpy_cuda_last_error(void);
int clumpy_cuda_device_count(int* count);

/* `stream` is a cudaStream_t to run on (e.g. the caller's current stream so that the caller's
 * CUDA events bracket the work), or NULL for a stream owned by the context. */
int clumpy_cuda_create(int device, void* stream, clumpy_cuda_ctx** out);
void clumpy_cuda_destroy(clumpy_cuda_ctx* ctx);
int clumpy_cuda_sync(clumpy_cuda_ctx* ctx);

void* clumpy_cuda_host_alloc(size_t bytes); /* pinned host memory, NULL on failure */
void clumpy_cuda_host_free(void* p);
int clumpy_cuda_dev_alloc(clumpy_cuda_ctx* ctx, size_t bytes, void** dptr);
int clumpy_cuda_dev_free(clumpy_cuda_ctx* ctx, void* dptr);
int clumpy_cuda_memcpy_h2d(clumpy_cuda_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int clumpy_cuda_memcpy_d2h(clumpy_cuda_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* Device-side stopwatch on the context's stream (CUDA events). */
int clumpy_cuda_timer_start(clumpy_cuda_ctx* ctx);
int clumpy_cuda_timer_stop(clumpy_cuda_ctx* ctx, float* elapsed_ms); /* synchronises */

/* Number of kernels this library has launched on `ctx` since creation. */
int clumpy_cuda_launch_count(clumpy_cuda_ctx* ctx, uint64_t* launches);

/* Measurement aid: throughput of the reductions the splat is made of.  Issues `nupdates`
 * fire-and-forget reductions per launch -- op 0: red.global.add.noftz.f16x2 on half cells, op 1:
 * red.global.add.u32 -- on a scratch region of `region_bytes`; pattern 0: uniform random cells, 1: all on
 * `hot_cells` cells, 2: update i near cell i * ncells / nupdates (a cloud stored in tile order).  Returns the
 * mean device time of one launch over `iters` launches (CUDA events, clears excluded).  Synchronises. */
int clumpy_cuda_atomic_peak(clumpy_cuda_ctx* ctx, int op, int pattern, size_t region_bytes,
                            uint64_t nupdates, uint32_t hot_cells, int iters, float* ms_per_launch);

/* ---- generate_simplex ------------------------------------------------------------------ */

/* Host helper: the 256-entry permutation of commands/generate_simplex.cc:241-280. */
int clumpy_cuda_simplex_perm(int64_t seed, int16_t* perm256);

/* Replaces the per-pixel loop of commands/generate_simplex.cc:59-82 (which calls
 * open_simplex_noise2, :298-417).  out_host is (height, width) float32.  minval/maxval receive
 * the "Noise range" the command prints (:78-82); either may be NULL. */
int clumpy_cuda_generate_simplex(clumpy_cuda_ctx* ctx, uint32_t width, uint32_t height,
                                 double amplitude, double frequency, int64_t seed,
                                 float* out_host, float* minval, float* maxval);

/* Row band [row0, row0+nrows) of the (height, width) image into out_dev (nrows x width floats).
 * Bands are independent, which is how the field is sharded across GPUs. */
int clumpy_cuda_generate_simplex_dev(clumpy_cuda_ctx* ctx, uint32_t width, uint32_t height,
                                     uint32_t row0, uint32_t nrows, double amplitude,
                                     double frequency, int64_t seed, float* out_dev,
                                     float* minval, float* maxval);

/* Several octaves in one pass: the sum, in FP32 and in the order given, of the fields generate_simplex produces
 * for (amplitudes[o], frequencies[o], seeds[o]) -- what the reference's users compute by running the command once
 * per octave and adding the files with numpy (README.md:28-36, extras/test.py:18-21).  Bit-identical to adding this
 * library's single-octave outputs; 1 <= noctaves <= 8.  Row bands as above. */
int clumpy_cuda_generate_simplex_octaves_dev(clumpy_cuda_ctx* ctx, uint32_t width, uint32_t height,
                                             uint32_t row0, uint32_t nrows, uint32_t noctaves,
                                             const double* amplitudes, const double* frequencies,
                                             const int64_t* seeds, float* out_dev, float* minval, float* maxval);

/* ---- curl_2d --------------------------------------------------------------------------- */

/* Replaces the loop of commands/curl_2d.cc:57-71.  src_host (height, width) float32 ->
 * dst_host (height, width, 2) float32 with texel = (dpdy, dpdx).  minval/maxval follow the
 * reference's printed "Curl range" including its quirk (:64-65); either may be NULL. */
int clumpy_cuda_curl_2d(clumpy_cuda_ctx* ctx, const float* src_host, uint32_t width,
                        uint32_t height, float* dst_host, float* minval, float* maxval);

/* A band of `nrows` rows starting at src_dev.  next_row_dev points at the row below the band
 * (the one-row halo owned by the next band / next GPU) or is NULL when the band ends at the
 * bottom edge of the image, where the reference clamps (:60). */
int clumpy_cuda_curl_2d_dev(clumpy_cuda_ctx* ctx, const float* src_dev, uint32_t width,
                            uint32_t nrows, const float* next_row_dev, float* dst_dev,
                            float* minval, float* maxval);

/* ---- pendulum_phase, cull_points (the producers either side of the path) ------ */

/* Replaces the loop of commands/pendulum_phase.cc:54-66: out_host (height, width, 2) float32 with
 * texel(row, col) = (omega, -friction * omega - sin(theta)), theta = sx * (col / W - 0.5),
 * omega = sy * (row / H - 0.5), sx = (float)(scale_x * M_PI / 0.5), sy = scale_y.  Bit-identical to
 * the reference: the W values sinf(theta) and the H values omega are evaluated on the host exactly as
 * the reference evaluates them per texel (same libm), the kernel combines them. */
int clumpy_cuda_pendulum_phase(clumpy_cuda_ctx* ctx, uint32_t width, uint32_t height, float friction,
                               float scale_x, float scale_y, float* out_host);

/* Rows [row0, row0 + nrows) of the same field into device memory (rows are independent: the
 * multi-GPU / fused-pipeline entry point). */
int clumpy_cuda_pendulum_phase_dev(clumpy_cuda_ctx* ctx, uint32_t width, uint32_t height,
                                   uint32_t row0, uint32_t nrows, float friction, float scale_x,
                                   float scale_y, float* out_dev);

/* Replaces cull_points() (commands/cull_points.cc:97-111): copies to out_host, in their original
 * order, the points whose nearest mask texel is positive (coordinates converted float -> uint32_t as
 * the reference does; outside the (height, width) float32 mask reads 0).  out_host holds up to npts
 * points; *nkept receives how many were kept ("Removed {npts - nkept} points."). */
int clumpy_cuda_cull_points(clumpy_cuda_ctx* ctx, const float* pts_host, uint32_t npts,
                            const float* mask_host, uint32_t width, uint32_t height,
                            float* out_host, uint32_t* nkept);

/* ---- splat_points ---------------------------------------------------------------------- */

/* Replaces splat_points() (commands/splat_points.cc:108-116): marks the texel under each point
 * with 255.  img_host (height, width) u8 is read, updated and written back. */
int clumpy_cuda_splat_points(clumpy_cuda_ctx* ctx, const float* pts_host, uint32_t npts,
                             uint32_t width, uint32_t height, uint8_t* img_host);

/* Replaces splat_disks() (commands/splat_points.cc:118-161): src-over blend of a k x k smoothstep
 * sprite of opacity `alpha` at every point into the u8 image.  wrapx is the reference's
 * `advect_wrapx` global (:20).  Even kernel sizes return CLUMPY_ERR_INVALID (the reference
 * prints "Kernel size must be an odd integer." and exits 1, :122-125). */
int clumpy_cuda_splat_disks(clumpy_cuda_ctx* ctx, const float* pts_host, uint32_t npts,
                            uint32_t width, uint32_t height, uint8_t* img_host, float alpha,
                            int kernel_size, int wrapx);

/* The same two rasterisers on `ndevices` GPUs of one box from this process (`clumpy splat_points` calls it when
 * CLUMPY_DEVICES lists several devices): alpha == 1 && kernel_size == 1 takes the fast path of splat_points.cc:93-94,
 * anything else is splat_disks.  Points are sharded by index range, every GPU counts its shard into a private
 * accumulator, the GPU that owns a row band sums that band out of all accumulators through peer memory (NVLink),
 * composites it and copies its rows back.  The result equals the one-GPU entry points' (the accumulators are sums). */
int clumpy_cuda_splat_multi(const int* devices, int ndevices, const float* pts_host, uint32_t npts,
                            uint32_t width, uint32_t height, uint8_t* img_host, float alpha,
                            int kernel_size, int wrapx);

/* ---- advect_points --------------------------------------------------------------------- */

/* Host helper: the per-point reset offsets of commands/advect_points.cc:127-135
 * (std::mt19937 seeded with 0 + std::uniform_int_distribution<>(0, nframes-1), libstdc++). */
int clumpy_cuda_age_offsets(uint32_t nframes, uint32_t npts, uint32_t* out);
/* The same draws into DEVICE memory, generated on the device (asynchronous on the context's stream): what
 * clumpy_cuda_advect_begin does when age_offset is NULL. */
int clumpy_cuda_age_offsets_dev(clumpy_cuda_ctx* ctx, uint32_t nframes, uint32_t npts, uint32_t* out_dev);

/* Set-up of commands/advect_points.cc:83-135: uploads the points (npts, 2) and the velocity
 * field (height, width, 2), allocates the u8 framebuffer (zeroed) and the splat accumulator.
 * `decay` is the non-negative decay; `wrapx` != 0 is what the command derives from a negative
 * decay argument (:78-81).  age_offset may be NULL, in which case the offsets of clumpy_cuda_age_offsets
 * are drawn (on the device).  nframes is the number of RECORDED frames; the run has 2*nframes simulation frames. */
int clumpy_cuda_advect_begin(clumpy_cuda_ctx* ctx, const float* pts_host, uint32_t npts,
                             const float* vel_host, uint32_t width, uint32_t height,
                             float step_size, int kernel_size, float decay, int wrapx,
                             uint32_t nframes, const uint32_t* age_offset, clumpy_advect** out);

/* The same set-up with everything spelled out (the device-resident pipeline).  Points, field and
 * offsets may already be on the device; a device field is used IN PLACE (it must outlive the run), which is how
 * `generate_simplex_dev -> curl_2d_dev -> advect` chains without a host round trip.  shard_row_lo < shard_row_hi
 * keeps only the points whose HOME row (the row of their initial position, truncated like the splat) lies in
 * [lo, hi): the multi-GPU shards, selected on the device (see clumpy_cuda_row_histogram).  age_offset values
 * >= nframes mean "never resets" (the reference's own sentinel, advect_points.cc:150).  Options replace the
 * CLUMPY_CUDA_* environment switches of round 1 (which remain as overrides for A/B runs): 0 = automatic. */
#define CLUMPY_CELLS_AUTO 0   /* sparse cloud: exact-order 64-bit cells; else the count-only type            */
#define CLUMPY_CELLS_COUNT 1  /* count-only cells: half floats for k <= 3, 32-bit otherwise (multi-GPU runs) */
#define CLUMPY_CELLS_F16 2
#define CLUMPY_CELLS_U32 3
#define CLUMPY_CELLS_EXACT 4
typedef struct clumpy_advect_desc {
    uint32_t size;                 /* sizeof(clumpy_advect_desc) */
    const float* pts;              /* (npts, 2) float32 */
    uint32_t npts;
    int pts_on_device;
    const float* vel;              /* (height, width, 2) float32 */
    int vel_on_device;
    uint32_t width, height;
    float step_size;
    int kernel_size;
    float decay;                   /* non-negative; negative-decay mode is wrapx = 1 */
    int wrapx;
    uint32_t nframes;
    const uint32_t* age_offset;    /* npts values indexed like pts, or NULL: clumpy_cuda_age_offsets */
    int age_on_device;
    uint32_t shard_row_lo, shard_row_hi;
    int cells;                     /* CLUMPY_CELLS_* */
    int overlap;                   /* composites on a side stream: 0 auto, 1 on, -1 off */
    int fuse;                      /* simulation frames per launch: 0 auto, 1 .. 8 */
    int regroup;                   /* frames between regroups by tile: 0 auto, -1 never */
    int cell_rows_multiple;        /* dense multi-GPU variant: padded counter rows a multiple of this */
} clumpy_advect_desc;
int clumpy_cuda_advect_begin_ex(clumpy_cuda_ctx* ctx, const clumpy_advect_desc* desc, clumpy_advect** out);

/* rows_host[y] = number of points whose home row is y (y < height), counted on the device; pts may be a host
 * or a device pointer.  Multi-GPU callers cut the image into strips of equal point count with it. */
int clumpy_cuda_row_histogram(clumpy_cuda_ctx* ctx, const float* pts, uint32_t npts, int pts_on_device,
                              uint32_t height, uint32_t* rows_host);

/* Advances `count` simulation frames (commands/advect_points.cc:139-176 without the file write):
 * Euler step + reset of every point, u8 decay, splat.  Asynchronous; no host copies.  Frames are run in
 * groups of up to 8: one advect launch (positions stay in registers) + one composite launch per group. */
int clumpy_cuda_advect_step(clumpy_advect* adv, uint32_t count);

/* Like clumpy_cuda_advect_step, and every one of the `count` frames is copied to dst_host + i * frame_stride
 * (height * width u8 each) on the context's copy stream while later frames are simulated.  dst_host should be
 * pinned.  clumpy_cuda_advect_wait_frames waits for the copies. */
int clumpy_cuda_advect_record(clumpy_advect* adv, uint32_t count, uint8_t* dst_host, size_t frame_stride);

/* The two stages of the pipeline, each timed with CUDA events around its launch on the stream it runs on, over
 * `groups` full groups: mean ms per launch of the advect kernel and of the composite work, and the frames one
 * launch covers.  serial = 0: as the pipeline runs them (the composite of one group shares the GPU with the
 * advect launch of the next); serial = 1: each kernel alone on the GPU.  Synchronises.  bench.py's roofline. */
int clumpy_cuda_advect_profile_stages(clumpy_advect* adv, uint32_t groups, int serial, float* advect_ms,
                                      float* composite_ms, uint32_t* frames_per_launch);

/* ---- one simulation frame split at its exchange point (multi-GPU) ------------------------------
 * Points advect independently, so they are sharded across GPUs, each with the whole velocity field.
 * What the shards share is the splat: every GPU counts ITS points into a private counter image (one
 * cell per texel, padded by kernel_size/2 on every side), the images are summed across GPUs (NCCL,
 * by the caller), and each GPU composites the rows it owns.  One frame is then
 *
 *     clumpy_cuda_advect_count          Euler step + reset + count of this GPU's points
 *     clumpy_cuda_advect_cells          where that counter image is (device pointer, geometry)
 *     ... caller sums the images across GPUs (all-reduce, or reduce-scatter by row bands) ...
 *     clumpy_cuda_advect_composite_rows decay + splat replay for image rows [row0, row0+nrows)
 *     clumpy_cuda_advect_finish_frame   clears the private image, advances the frame counter
 *
 * Counter images are row-major with `pitch_cells` cells per row; the cell of texel (x, y) is at
 * [(y + halo) * pitch + (x + halo)].  composite_rows takes a pointer to the padded row `row0` (the
 * cells of image row row0 - halo) of an image with that pitch that holds at least
 * round_up(nrows, 16) + 2*halo rows from there; it updates the framebuffer rows in place. */
int clumpy_cuda_advect_count(clumpy_advect* adv);
int clumpy_cuda_advect_cells(clumpy_advect* adv, void** cells_dev, size_t* bytes,
                             uint32_t* pitch_cells, uint32_t* rows, uint32_t* halo,
                             uint32_t* bytes_per_cell);
int clumpy_cuda_advect_composite_rows(clumpy_advect* adv, const void* cells_dev, uint32_t row0,
                                      uint32_t nrows);
int clumpy_cuda_advect_finish_frame(clumpy_advect* adv);

/* ---- multi-GPU with the exchange fused into the advect kernel -----------------------------------
 * Replaces the loop of commands/advect_points.cc:139-176 when the points of one run are spread over
 * several GPUs (points advect independently, :144-153; the splat, splat_points.cc:118-161, is the one
 * thing they share).  The counter image is distributed by bands of chunk = round_up(ceil(H / world), 16)
 * padded rows (band g on GPU g, the last band also takes the padding rows) and never exists as a
 * whole.  Frames run in GROUPS of up to 8.  Per group and GPU:
 *   route kernel      advects this GPU's points through the group (positions stay in registers), counts the
 *                     ones that land in its own band straight into the frame's band buffer and stores every
 *                     other point's cell index (4 bytes) into the inbox of the GPU that owns the cell's band --
 *                     and of the neighbouring band when the row lies within the sprite's reach of the edge --
 *                     through peer-mapped memory (NVLink / NVSwitch stores from inside the kernel).  No barrier
 *                     and no fence inside the frame loop; at the end a CTA fences its remote stores and takes
 *                     a ticket, and the CTA that takes the last ticket stores the group's EPOCH into a flag
 *                     per source in every peer's memory.
 *   drain kernel      waits for every source's epoch (bounded: CLUMPY_ROUTE_TIMEOUT_MS, default 2000; a time-out
 *                     makes every later call on the run fail with CLUMPY_ERR_STATE), counts the inbox entries
 *                     of the group's frames into the band buffers.
 *   group composite   the band's image rows through the frames of the group (composite_group.cu); also zeroes
 *                     the band buffers of the group before.
 * No host-side collective, no NCCL on the data path: a whole run is enqueued without synchronisation.
 * Best with shards by HOME ROW (desc.shard_row_lo / _hi): most points then never leave their GPU.
 *
 *   set-up:   advect_begin_ex with this rank's shard, desc.cells = CLUMPY_CELLS_COUNT
 *             -> route_begin (allocates this rank's inbox; export it with clumpy_cuda_ipc_export,
 *             exchange the handles, open the peers' with clumpy_cuda_ipc_open -- or, for several
 *             ranks in one process, pass the device pointers themselves) -> route_peers
 *   frames:   route_steps(count) / route_record(count, ...)
 *
 * `capacity` = point count of the largest shard (the same value on every rank).  Every rank must run the same
 * sequence of frame counts (groups are cut at the same frames everywhere).
 * route_begin allocates the inbox when *inbox_dev is NULL on entry and returns it with its size; the run frees it at
 * advect_end.  With *inbox_dev != NULL it uses that memory (*inbox_bytes of it, from clumpy_cuda_dev_alloc or an
 * earlier run's size) and leaves it to the caller: an application that runs the command repeatedly keeps one arena
 * per GPU, and its peers' mappings of it, across runs (clumpy_b200/multigpu.py does). */
#define CLUMPY_IPC_HANDLE_BYTES 64
int clumpy_cuda_ipc_export(clumpy_cuda_ctx* ctx, const void* dptr, void* handle64);
int clumpy_cuda_ipc_open(clumpy_cuda_ctx* ctx, const void* handle64, void** dptr);
int clumpy_cuda_ipc_close(clumpy_cuda_ctx* ctx, void* dptr);
int clumpy_cuda_advect_route_bytes(uint32_t capacity, int world, size_t* inbox_bytes); /* size of an inbox */
int clumpy_cuda_advect_route_begin(clumpy_advect* adv, int rank, int world, uint32_t capacity,
                                   void** inbox_dev, size_t* inbox_bytes);
int clumpy_cuda_advect_route_peers(clumpy_advect* adv, void* const* peer_inbox);
int clumpy_cuda_advect_owned_rows(clumpy_advect* adv, uint32_t* row0, uint32_t* nrows);
int clumpy_cuda_advect_route_steps(clumpy_advect* adv, uint32_t count);
/* route_steps, and this rank's rows (clumpy_cuda_advect_owned_rows) of every one of the `count` frames are
 * copied to dst_host + i * frame_stride on the copy stream (each GPU over its own PCIe link). */
int clumpy_cuda_advect_route_record(clumpy_advect* adv, uint32_t count, uint8_t* dst_host, size_t frame_stride);
/* `groups` groups launched one by one with CUDA events around every kernel: mean device time per GROUP (ms) of
 * the route kernel, the drain kernel (includes the wait for the slowest peer) and the band composite, and the
 * frames per group.  Every rank calls it with the same arguments.  Synchronises. */
int clumpy_cuda_advect_route_profile(clumpy_advect* adv, uint32_t groups, float* route_ms,
                                     float* drain_ms, float* composite_ms, uint32_t* frames_per_group);
/* Fence time-outs seen so far.  Synchronises.  (Any non-zero count also fails route_steps / read_image /
 * wait_frames with CLUMPY_ERR_STATE.) */
int clumpy_cuda_advect_route_errors(clumpy_advect* adv, uint32_t* errors);

/* Device address of the current framebuffer (height * width u8), for callers that gather row
 * bands across GPUs themselves. */
int clumpy_cuda_advect_image_dev(clumpy_advect* adv, void** img_dev);

/* Number of simulation frames done so far; frames [nframes, 2*nframes) are the recorded ones. */
int clumpy_cuda_advect_simframe(clumpy_advect* adv, uint32_t* simframe);
/* Number of points this run simulates (the shard's size when desc.shard_row_* selected one). */
int clumpy_cuda_advect_npts(clumpy_advect* adv, uint32_t* npts);

/* Enqueues a device->host copy of the current framebuffer (height*width u8) on the context's
 * copy stream, ordered after the steps issued so far; later steps do not wait for it unless they
 * would overwrite the buffer being copied.  dst_host should be pinned for the copy to overlap. */
int clumpy_cuda_advect_frame_async(clumpy_advect* adv, uint8_t* dst_host);
int clumpy_cuda_advect_wait_frames(clumpy_advect* adv);

/* Diagnostics for parity tests (the reference CLI never outputs trajectories): current positions
 * in the caller's original point order, and the current framebuffer.  Both synchronise. */
int clumpy_cuda_advect_read_points(clumpy_advect* adv, float* out_host);
int clumpy_cuda_advect_read_image(clumpy_advect* adv, uint8_t* out_host);
/* Shards: positions in storage order + the index each point had in the caller's list (npts of each). */
int clumpy_cuda_advect_read_points_indexed(clumpy_advect* adv, float* pos_host, uint32_t* index_host);

int clumpy_cuda_advect_end(clumpy_advect* adv);

/* The whole command (commands/advect_points.cc:137-179): 2*nframes simulation frames; each of the
 * last nframes framebuffers is copied to pinned host memory on a side stream and handed to
 * `on_frame(user, animframe, pixels)` in order while the GPU keeps simulating.  A non-zero return
 * from the callback aborts the run with CLUMPY_ERR_STATE. */
typedef int (*clumpy_frame_fn)(void* user, uint32_t animframe, const uint8_t* pixels);
int clumpy_cuda_advect_points(clumpy_cuda_ctx* ctx, const float* pts_host, uint32_t npts,
                              const float* vel_host, uint32_t width, uint32_t height,
                              float step_size, int kernel_size, float decay, int wrapx,
                              uint32_t nframes, const uint32_t* age_offset,
                              clumpy_frame_fn on_frame, void* user);

/* The same command on `ndevices` GPUs of one box, driven from this one process (commands/advect_points.cc:137-179
 * behind the same plugin interface: `clumpy advect_points` calls it when CLUMPY_DEVICES lists several devices).
 * One context per device, peer access between all of them, points sharded by home row on the devices, reset
 * offsets drawn once, the routed frame loop above, every GPU copying its row band of each recorded frame over its
 * own PCIe link; the bands are stitched on the host and handed to on_frame in order.  A device may be listed more
 * than once (the ranks then share it). */
int clumpy_cuda_advect_points_multi(const int* devices, int ndevices, const float* pts_host, uint32_t npts,
                                    const float* vel_host, uint32_t width, uint32_t height,
                                    float step_size, int kernel_size, float decay, int wrapx,
                                    uint32_t nframes, const uint32_t* age_offset,
                                    clumpy_frame_fn on_frame, void* user);

#ifdef __cplusplus
}
#endif
#endif /* CLUMPY_CUDA_H */
