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
//   offset  u16/u32  reset phase of each point    read every frame (streaming); u16 if nframes <= 65536
//   perm    uint32   original index of each slot  only for read_points and regrouping
//   (a second copy of those four: regroup_points writes there and swaps)
//   vel     float2   (H, W) velocity field        random 8-byte gathers
//   hist    2 counter images (see CellMode)       one atomic per point, read + cleared by the composite
//   img[2]  uint8    framebuffer (ping-pong only while a frame copy is in flight)
//
// Points are kept in the order of the 8x4-texel tile they are in (counting sort at set-up, repeated
// every nframes/16 frames: see regroup_points), so a warp's velocity gathers and counter updates stay
// within a few cache lines.  Results do not depend on the order (integer counts commute; trajectories
// are per point); read_points undoes the permutation.  Sparse clouds stay in the caller's order: their
// cells carry the original point index (CELL_X64) so that the composite pass can keep the reference's
// hit order.
#include "device_utils.cuh"
#include "internal.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

namespace clumpy {

constexpr int ADVECT_THREADS = 256;
constexpr int ADVECT_PTS = 4; // points per thread (independent load chains in flight)

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

// How a point is counted into the splat accumulator.
enum CellMode : int {
    CELL_NONE = -1, // frame without a splat (warm-up with decay == 0)
    CELL_U32 = 0,   // 32-bit counters, one RED.ADD per point (any sprite)
    CELL_U8 = 1,    // byte counters saturating at `cap`, packed 4 per word, CAS loop (dense path)
    CELL_F16 = 2,   // half-float counters, two per word, one RED.ADD.F16x2 per point: a half counts
                    // exactly to 2048 and then stays put, so it saturates by itself and cannot carry
    CELL_X64 = 3,   // 64-bit cells = count (24 bits) + sum of original point indices (40 bits), one
                    // 64-bit RED.ADD per point: keeps the reference's hit order (composite_exact.cu)
};

__device__ __forceinline__ void cell_add_f16(uint32_t* __restrict__ words, long long cell, uint64_t keep) {
    // (1, 0) or (0, 1) as a half2 bit pattern: 0x3C00 is 1.0
    const uint32_t one = (cell & 1) ? 0x3C000000u : 0x00003C00u;
    asm volatile("red.global.add.L2::cache_hint.noftz.f16x2 [%0], %1, %2;"
                 :: "l"(words + (cell >> 1)), "r"(one), "l"(keep) : "memory");
}

// Zeroes a counter image with evict-last stores (cudaMemset would leave the lines with normal
// priority and they would be gone from L2 before the next frame counts into them).
__global__ void __launch_bounds__(256)
clear_cells_kernel(uint4* __restrict__ p, size_t n16) {
    const uint64_t keep = make_keep_policy();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        st_hint_u4(p + i, make_uint4(0u, 0u, 0u, 0u), keep);
}

// Byte counter += 1, saturating at cap.  The new word is computed from the observed old word, so a
// cell can never overflow into its neighbour however many threads hit it at once, and once a cell
// is saturated later points see that with one L2 read and issue no atomic at all.
//
// Warp-aggregated: points are stored sorted by tile, so several lanes of a warp usually target the
// same word (4 neighbouring texels).  Lanes are grouped by word (match.any); the lowest lane of each
// group adds the group's per-byte counts with ONE compare-and-swap, instead of the lanes fighting
// each other through retries.  Must be called by all 32 lanes; `cell` < 0 means "no update".
__device__ __forceinline__ void cell_add_sat_u8(uint32_t* __restrict__ words, long long cell, uint32_t cap) {
    const uint32_t lane = threadIdx.x & 31u;
    const bool valid = cell >= 0;
    const uint32_t word = valid ? (uint32_t)(cell >> 2) : (0xFFFFFFFFu - lane); // unique key when idle
    const uint32_t byte = (uint32_t)cell & 3u;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, word);
    const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, valid && byte == 0u);
    const uint32_t b1 = __ballot_sync(0xFFFFFFFFu, valid && byte == 1u);
    const uint32_t b2 = __ballot_sync(0xFFFFFFFFu, valid && byte == 2u);
    const uint32_t b3 = __ballot_sync(0xFFFFFFFFu, valid && byte == 3u);
    if (!valid || lane != (uint32_t)(__ffs(peers) - 1)) return;
    const uint32_t c0 = __popc(peers & b0), c1 = __popc(peers & b1), c2 = __popc(peers & b2), c3 = __popc(peers & b3);
    uint32_t* w = words + word;
    uint32_t old = __ldcg(w);
    for (;;) {
        const uint32_t n0 = min(cap, (old & 0xFFu) + c0), n1 = min(cap, ((old >> 8) & 0xFFu) + c1);
        const uint32_t n2 = min(cap, ((old >> 16) & 0xFFu) + c2), n3 = min(cap, (old >> 24) + c3);
        const uint32_t want = n0 | (n1 << 8) | (n2 << 16) | (n3 << 24);
        if (want == old) break; // every touched byte is already saturated
        const uint32_t prev = atomicCAS(w, old, want);
        if (prev == old) break;
        old = prev;
    }
}

// Euler step + reset of this thread's ADVECT_PTS points (advect_points.cc:144-153,163-171); returns
// the new positions in p[] and which of them exist in live[].
//
// A CTA owns ADVECT_THREADS * ADVECT_PTS consecutive points; thread t takes t, t + 256, t + 512, ...
// so every warp-wide access covers 32 CONSECUTIVE points: position / offset traffic is perfectly
// coalesced and, the points being stored in tile order, a warp's gather touches only a few lines.
// All loads of a phase are issued before the first use.
template <bool WRAP, typename OffT>
__device__ __forceinline__ void advance_points(float2* __restrict__ pos, const float2* __restrict__ orig,
                                               const OffT* __restrict__ offset, uint32_t npts,
                                               const float2* __restrict__ vel, uint32_t width, uint32_t height,
                                               float step, uint32_t phase, uint32_t base, bool stream_hint,
                                               float2 (&p)[ADVECT_PTS], bool (&live)[ADVECT_PTS]) {
    const uint64_t pol = stream_hint ? make_stream_policy() : make_normal_policy();
    uint32_t off[ADVECT_PTS];
#pragma unroll
    for (int k = 0; k < ADVECT_PTS; ++k) {
        const uint32_t i = base + k * ADVECT_THREADS;
        live[k] = i < npts;
        p[k] = make_float2(0.f, 0.f);
        off[k] = 0xFFFFFFFFu;
        if (live[k]) {
            p[k] = ld_stream_f2(pos + i, pol);
            off[k] = ld_stream_off(offset + i, pol);
        }
    }
    // nearest-texel fetch (advect_points.cc:23-37): all gathers in flight together
    float2 v[ADVECT_PTS];
#pragma unroll
    for (int k = 0; k < ADVECT_PTS; ++k) {
        uint32_t x = f2u32_x86(p[k].x);
        const uint32_t y = f2u32_x86(p[k].y);
        if (WRAP) x = (width + (x % width)) % width;
        v[k] = make_float2(0.f, 0.f);
        if (live[k] && x < width && y < height) v[k] = __ldg(vel + ((size_t)y * width + x));
    }
#pragma unroll
    for (int k = 0; k < ADVECT_PTS; ++k) {
        const uint32_t i = base + k * ADVECT_THREADS;
        // pt += step * v: rounded multiply, then rounded add (:146)
        p[k] = make_float2(__fadd_rn(p[k].x, __fmul_rn(step, v[k].x)), __fadd_rn(p[k].y, __fmul_rn(step, v[k].y)));
        if (off[k] == phase) p[k] = orig[i]; // reset AFTER the move (:147-152,166-170)
        if (live[k]) st_stream_f2(pos + i, p[k], pol);
    }
}

template <bool WRAP, int MODE, typename OffT>
__global__ void __launch_bounds__(ADVECT_THREADS)
advect_kernel(float2* __restrict__ pos, const float2* __restrict__ orig,
              const OffT* __restrict__ offset, uint32_t npts, const float2* __restrict__ vel,
              uint32_t width, uint32_t height, float step, uint32_t phase, GeomDev g,
              uint32_t* __restrict__ hist, uint32_t cap, int keep_cells) {
    const uint32_t base = blockIdx.x * (ADVECT_THREADS * ADVECT_PTS) + threadIdx.x;
    const uint64_t keep = keep_cells ? make_keep_policy() : make_normal_policy();
    float2 p[ADVECT_PTS];
    bool live[ADVECT_PTS]; // no early return: the byte-cell update below is warp-collective
    advance_points<WRAP, OffT>(pos, orig, offset, npts, vel, width, height, step, phase, base, g.keep_pos == 0, p, live);
    if (MODE != CELL_NONE) {
#pragma unroll
        for (int k = 0; k < ADVECT_PTS; ++k) {
            // splat_points.cc:143-144: centre texel by (int32_t) truncation
            const long long at = live[k] ? hist_index<WRAP>(f2i32_x86(p[k].x), f2i32_x86(p[k].y), g) : -1;
            if (MODE == CELL_X64) {
                // exact-order path: the points are in the caller's order, so the slot IS the original index
                const unsigned long long inc = ((unsigned long long)(base + k * ADVECT_THREADS) << 24) | 1ull;
                if (at >= 0) atomicAdd(reinterpret_cast<unsigned long long*>(hist) + at, inc);
            }
            else if (MODE == CELL_U32) { if (at >= 0) atomicAdd(hist + at, 1u); }
            else if (MODE == CELL_F16) { if (at >= 0) cell_add_f16(hist, at, keep); }
            else cell_add_sat_u8(hist, at, cap);
        }
    }
}

// Several frames per launch ("temporal blocking"): the positions are loaded once, advanced through `nf`
// simulation frames in registers -- gather, Euler step, reset, count, each frame counting into the next
// counter image of a ring -- and stored once.  Position and phase traffic (20 of the 28 algorithmic
// bytes per point-step) is paid once per group instead of once per frame; the gathers and the counter
// updates are what is left per frame.  Same arithmetic per frame as advect_kernel, so same results.
template <bool WRAP, int MODE, typename OffT>
__global__ void __launch_bounds__(ADVECT_THREADS)
advect_fused_kernel(float2* __restrict__ pos, const float2* __restrict__ orig,
                    const OffT* __restrict__ offset, uint32_t npts, const float2* __restrict__ vel,
                    uint32_t width, uint32_t height, float step, uint32_t phase0, uint32_t nframes, int nf,
                    GeomDev g, uint32_t* __restrict__ hist0, size_t hist_stride_words, int slot0, int ring,
                    uint32_t cap, int keep_cells) {
    const uint32_t base = blockIdx.x * (ADVECT_THREADS * ADVECT_PTS) + threadIdx.x;
    const uint64_t pol = g.keep_pos == 0 ? make_stream_policy() : make_normal_policy();
    const uint64_t keep = keep_cells ? make_keep_policy() : make_normal_policy();
    float2 p[ADVECT_PTS];
    uint32_t off[ADVECT_PTS];
    bool live[ADVECT_PTS]; // no early return: the byte-cell update below is warp-collective
#pragma unroll
    for (int k = 0; k < ADVECT_PTS; ++k) {
        const uint32_t i = base + k * ADVECT_THREADS;
        live[k] = i < npts;
        p[k] = make_float2(0.f, 0.f);
        off[k] = 0xFFFFFFFFu;
        if (live[k]) {
            p[k] = ld_stream_f2(pos + i, pol);
            off[k] = ld_stream_off(offset + i, pol);
        }
    }
    uint32_t phase = phase0;
    for (int f = 0; f < nf; ++f) {
        float2 v[ADVECT_PTS];
#pragma unroll
        for (int k = 0; k < ADVECT_PTS; ++k) { // nearest-texel fetch (advect_points.cc:23-37)
            uint32_t x = f2u32_x86(p[k].x);
            const uint32_t y = f2u32_x86(p[k].y);
            if (WRAP) x = (width + (x % width)) % width;
            v[k] = make_float2(0.f, 0.f);
            if (live[k] && x < width && y < height) v[k] = __ldg(vel + ((size_t)y * width + x));
        }
#pragma unroll
        for (int k = 0; k < ADVECT_PTS; ++k) {
            // pt += step * v: rounded multiply, then rounded add (:146); reset AFTER the move (:147-152,166-170)
            p[k] = make_float2(__fadd_rn(p[k].x, __fmul_rn(step, v[k].x)), __fadd_rn(p[k].y, __fmul_rn(step, v[k].y)));
            if (off[k] == phase) p[k] = orig[base + k * ADVECT_THREADS];
        }
        phase = phase + 1u == nframes ? 0u : phase + 1u;
        if (MODE != CELL_NONE) {
            uint32_t* hist = hist0 + (size_t)((slot0 + f) % ring) * hist_stride_words;
#pragma unroll
            for (int k = 0; k < ADVECT_PTS; ++k) {
                // splat_points.cc:143-144: centre texel by (int32_t) truncation
                const long long at = live[k] ? hist_index<WRAP>(f2i32_x86(p[k].x), f2i32_x86(p[k].y), g) : -1;
                if (MODE == CELL_X64) {
                    const unsigned long long inc = ((unsigned long long)(base + k * ADVECT_THREADS) << 24) | 1ull;
                    if (at >= 0) atomicAdd(reinterpret_cast<unsigned long long*>(hist) + at, inc);
                }
                else if (MODE == CELL_U32) { if (at >= 0) atomicAdd(hist + at, 1u); }
                else if (MODE == CELL_F16) { if (at >= 0) cell_add_f16(hist, at, keep); }
                else cell_add_sat_u8(hist, at, cap);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < ADVECT_PTS; ++k)
        if (live[k]) st_stream_f2(pos + base + k * ADVECT_THREADS, p[k], pol);
}

// ---- multi-GPU: the splat exchange fused into the advect kernel ------------------------------------
//
// With G GPUs the counter image is distributed by bands of `chunk` padded rows (band d on GPU d) and
// never exists as a whole.  Each GPU advects ITS points and, instead of counting them locally and
// summing dense images afterwards (67 MB per frame at 8192x4096), counts the points that land in its
// own band straight into its band buffer and writes every other point's cell -- one 4-byte index into
// the OWNER's band buffer -- into the owner's inbox through peer-mapped memory (NVLink / NVSwitch
// stores issued from inside the kernel).  A point whose row lies within `halo` rows of a band edge is
// also sent to the neighbouring band, whose composite needs it, so no separate halo exchange is left.
//
// No global slot counters: CTA c of source rank s owns the fixed region [c * ROUTE_BLOCK,
// (c + 1) * ROUTE_BLOCK) of segment s in every destination's inbox (a CTA advects ROUTE_BLOCK points and
// a point sends at most one entry per destination and frame), ranks its entries with a shared-memory
// atomic, and leaves the number it wrote in the destination's count word [s][c].  The receiver reads
// the count words, zeroes the ones it used, and counts the entries into its band.
//
// One launch advances the points by up to ROUTE_MAX_FUSE simulation frames: positions are loaded
// once, stay in registers, and are stored once, and the launch overhead -- which at ~2 M points per
// GPU (C4 on 8 GPUs) is as long as the work of a frame -- is paid once per group.  Every frame of the
// group has its own band buffer, inbox parity, count words and ticket, and its own frame fence: an
// epoch flag per source in the receiver's memory, published by the extra last CTA of the grid as
// soon as all CTAs have finished that frame.
constexpr int ROUTE_MAX_WORLD = 8;
constexpr uint32_t ROUTE_BLOCK = ADVECT_THREADS * ADVECT_PTS; // points per CTA = entry slots per region
constexpr int ROUTE_RING = 8;     // band buffers / inbox parities / tickets in flight per rank
constexpr int ROUTE_AHEAD = 4;    // the drain of frame s zeroes the band buffer of frame s + ROUTE_AHEAD
constexpr int ROUTE_MAX_FUSE = 4; // frames per launch (<= ROUTE_AHEAD)

struct RouteDev {
    int me, world;
    int chunk, halo, pitch;             // band geometry in padded rows / cells
    uint32_t nregion;                   // regions (= CTAs of the largest shard) per source segment
    uint32_t nframes, phase0;           // reset period; phase of the first frame of the group
    int nf;                             // frames in this launch
    int slot0;                          // ring slot of the first frame (frame f uses (slot0 + f) % ROUTE_RING)
    size_t inbox_stride, cnt_stride, band_stride; // uint32 units between consecutive ring slots
    uint32_t* inbox[ROUTE_MAX_WORLD];   // rank r's entries, ring slot 0 (peer-mapped): [ring][world][nregion * ROUTE_BLOCK]
    uint32_t* cnt[ROUTE_MAX_WORLD];     // rank r's count words, ring slot 0 (peer-mapped): [ring][world][nregion]
    uint32_t* flags[ROUTE_MAX_WORLD];   // rank r's epoch flags (peer-mapped): [world]
    uint32_t* band;                     // this rank's band buffers, ring slot 0
    int band_f16;                       // band cells are half floats (else uint32)
    int signal;                         // 1: the grid has an extra CTA that publishes the epochs; 0: the caller fences
    int loopback;                       // profiling aid: every "peer" is this rank itself (see route_begin)
    uint32_t epoch0;                    // epoch of the first frame of the group
    uint32_t* ticket;                   // [ring] CTAs that have finished the frame of that slot
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// flags[me] = epoch in rank d's memory.  Called after every store of the frame has been fenced.
__device__ __forceinline__ void publish_frame(const RouteDev& rt, int d, uint32_t epoch) {
    const int slot = rt.loopback ? d : rt.me; // loopback: play every source of my own inbox
    *reinterpret_cast<volatile uint32_t*>(rt.flags[d] + slot) = epoch;
}

// What a frame of the group writes to.
struct RouteFrame {
    uint32_t* band;      // this rank's band buffer of the frame
    size_t inbox_off;    // offset of the frame's parity inside every inbox / count-word array
    size_t cnt_off;
    uint32_t* cta_n;     // shared: entries this CTA has written per destination this frame
};

// One entry for rank `dst`: the CTA-wide rank comes from a shared-memory atomic, the slot is fixed by
// the CTA's region, the store goes over NVLink.
__device__ __forceinline__ void route_send(const RouteDev& rt, const RouteFrame& fr, int dst, uint32_t entry) {
    const uint32_t r = atomicAdd(fr.cta_n + dst, 1u);
    rt.inbox[dst][fr.inbox_off + ((size_t)rt.me * rt.nregion + blockIdx.x) * ROUTE_BLOCK + r] = entry;
}

// A cell of padded row `r` (relative to band `owner`'s first row) and padded column `ux`: counted
// straight into this rank's band buffer when it is ours, sent to the owner's inbox otherwise; rows
// within the sprite's reach of a band edge also go to the neighbouring band, whose composite needs them.
__device__ __forceinline__ void route_cell(const RouteDev& rt, const RouteFrame& fr, int owner, int r, uint32_t ux) {
    const uint32_t entry = (uint32_t)((r + rt.halo) * rt.pitch) + ux; // band rows: [halo above | chunk | halo below]
    if (owner == rt.me) {
        if (rt.band_f16) cell_add_f16(fr.band, (long long)entry, make_normal_policy());
        else atomicAdd(fr.band + entry, 1u);
    } else {
        route_send(rt, fr, owner, entry);
    }
    if (r < rt.halo && owner > 0) route_send(rt, fr, owner - 1, (uint32_t)((r + rt.chunk + rt.halo) * rt.pitch) + ux);
    if (r >= rt.chunk - rt.halo && owner < rt.world - 1) route_send(rt, fr, owner + 1, (uint32_t)((r - rt.chunk + rt.halo) * rt.pitch) + ux);
}

// The extra last CTA of a fenced launch: for every frame of the group, wait until all the other CTAs
// have finished it and publish its epoch to every peer (acquire = ticket read + fence here; release =
// fence + ticket on the senders' side).
__device__ __forceinline__ void route_publisher(const RouteDev& rt) {
    if (threadIdx.x >= 32) return;
    for (int f = 0; f < rt.nf; ++f) {
        if (threadIdx.x == 0) {
            uint32_t* t = rt.ticket + (rt.slot0 + f) % ROUTE_RING;
            while (*reinterpret_cast<volatile uint32_t*>(t) != gridDim.x - 1u) __nanosleep(40);
            *t = 0u;
            __threadfence_system();
        }
        __syncwarp();
        if ((int)threadIdx.x < rt.world) publish_frame(rt, (int)threadIdx.x, rt.epoch0 + (uint32_t)f);
    }
}

template <bool WRAP, bool SPLAT, typename OffT, int MINB>
__global__ void __launch_bounds__(ADVECT_THREADS, MINB)
advect_route_kernel(float2* __restrict__ pos, const float2* __restrict__ orig,
                    const OffT* __restrict__ offset, uint32_t npts, const float2* __restrict__ vel,
                    uint32_t width, uint32_t height, float step, GeomDev g, RouteDev rt) {
    __shared__ uint32_t cta_n[ROUTE_MAX_FUSE][ROUTE_MAX_WORLD];
    if (SPLAT && rt.signal && blockIdx.x == gridDim.x - 1u) { route_publisher(rt); return; }
    if (SPLAT) {
        if (threadIdx.x < ROUTE_MAX_FUSE * ROUTE_MAX_WORLD) (&cta_n[0][0])[threadIdx.x] = 0u;
        __syncthreads(); // here, where every warp arrives at once, not after the loads
    }
    const uint64_t pol = g.keep_pos == 0 ? make_stream_policy() : make_normal_policy();
    const uint32_t base = blockIdx.x * ROUTE_BLOCK + threadIdx.x;
    const int my_lo = rt.me * rt.chunk; // first padded row of this rank's band
    // positions and reset phases: loaded once for the whole group of frames
    float2 p[ADVECT_PTS];
    uint32_t off[ADVECT_PTS];
    bool live[ADVECT_PTS];
#pragma unroll
    for (int k = 0; k < ADVECT_PTS; ++k) {
        const uint32_t i = base + k * ADVECT_THREADS;
        live[k] = i < npts;
        p[k] = make_float2(0.f, 0.f);
        off[k] = 0xFFFFFFFFu;
        if (live[k]) {
            p[k] = ld_stream_f2(pos + i, pol);
            off[k] = ld_stream_off(offset + i, pol);
        }
    }
    uint32_t phase = rt.phase0;
    for (int f = 0; f < rt.nf; ++f) {
        // Euler step + reset, exactly as advance_points does it (advect_points.cc:144-153,163-171)
        float2 v[ADVECT_PTS];
#pragma unroll
        for (int k = 0; k < ADVECT_PTS; ++k) { // nearest-texel fetch (advect_points.cc:23-37)
            uint32_t x = f2u32_x86(p[k].x);
            const uint32_t y = f2u32_x86(p[k].y);
            if (WRAP) x = (width + (x % width)) % width;
            v[k] = make_float2(0.f, 0.f);
            if (live[k] && x < width && y < height) v[k] = __ldg(vel + ((size_t)y * width + x));
        }
#pragma unroll
        for (int k = 0; k < ADVECT_PTS; ++k) {
            // pt += step * v: rounded multiply, then rounded add (:146)
            p[k] = make_float2(__fadd_rn(p[k].x, __fmul_rn(step, v[k].x)), __fadd_rn(p[k].y, __fmul_rn(step, v[k].y)));
            if (off[k] == phase) p[k] = orig[base + k * ADVECT_THREADS]; // reset AFTER the move (:147-152,166-170)
        }
        phase = phase + 1u == rt.nframes ? 0u : phase + 1u;
        if (!SPLAT) continue;

        const int slot = (rt.slot0 + f) % ROUTE_RING;
        RouteFrame fr;
        fr.band = rt.band + (size_t)slot * rt.band_stride;
        fr.inbox_off = (size_t)slot * rt.inbox_stride;
        fr.cnt_off = (size_t)slot * rt.cnt_stride;
        fr.cta_n = cta_n[f];
#pragma unroll
        for (int k = 0; k < ADVECT_PTS; ++k) {
            // splat_points.cc:143-144: centre texel by (int32_t) truncation; same tests as hist_index
            const int32_t cx = f2i32_x86(p[k].x), cy = f2i32_x86(p[k].y);
            const uint32_t uy = (uint32_t)cy + (uint32_t)g.halo;
            uint32_t ux;
            bool inside = live[k] && uy < (uint32_t)(g.height + 2 * g.halo);
            if (WRAP) {
                int32_t m = cx % g.width;
                if (m < 0) m += g.width;
                ux = (uint32_t)(m + g.halo);
            } else {
                ux = (uint32_t)cx + (uint32_t)g.halo;
                inside = inside && ux < (uint32_t)(g.width + 2 * g.halo);
            }
            if (!inside) continue;
            const int r = (int)uy - my_lo;
            if ((uint32_t)r < (uint32_t)rt.chunk) {
                route_cell(rt, fr, rt.me, r, ux);
            } else {
                const int owner = min((int)(uy / (uint32_t)rt.chunk), rt.world - 1);
                route_cell(rt, fr, owner, (int)uy - owner * rt.chunk, ux);
            }
        }
        __syncthreads(); // all entries of this CTA for this frame are written, cta_n[f] is final
        if (threadIdx.x < 32) {
            // Warp 0 leaves the non-zero counts with their destinations (the receivers zero them again)
            // and takes part in the frame fence while the other warps go on with the next frame.  A CTA
            // that stored into an inbox fences those stores at system scope (bar.sync / syncwarp order
            // them before thread 0's cumulative fence.sys; CTAs whose points all landed in this rank's own
            // band skip the fence, which costs microseconds) and then adds itself to the frame's ticket
            // with a fire-and-forget reduction.
            const uint32_t n = (int)threadIdx.x < rt.world ? fr.cta_n[threadIdx.x] : 0u;
            if (n) *reinterpret_cast<volatile uint32_t*>(rt.cnt[threadIdx.x] + fr.cnt_off + (size_t)rt.me * rt.nregion + blockIdx.x) = n;
            const bool cta_sent = __any_sync(0xFFFFFFFFu, n != 0u);
            if (rt.signal && threadIdx.x == 0) {
                if (cta_sent) __threadfence_system();
                asm volatile("red.global.add.u32 [%0], 1;" :: "l"(rt.ticket + slot) : "memory");
            }
        }
    }
#pragma unroll
    for (int k = 0; k < ADVECT_PTS; ++k)
        if (live[k]) st_stream_f2(pos + base + k * ADVECT_THREADS, p[k], pol);
}

// The same fence for a rank that has no points this run (no route kernel to carry it).
__global__ void route_signal_kernel(RouteDev rt) {
    for (int f = 0; f < rt.nf; ++f)
        if ((int)threadIdx.x < rt.world) publish_frame(rt, (int)threadIdx.x, rt.epoch0 + (uint32_t)f);
}

// Counts the entries the ranks sent into this rank's band buffer: one warp per (source, region).
// With `flags`, first waits until rank s has published this frame's epoch (bounded: on a timeout it
// records the error and skips the source, so a dead peer cannot hang the GPU).  Count words are zeroed
// as they are used: senders only ever write non-zero ones.  CELLS: CELL_F16 or CELL_U32.
template <int CELLS>
__global__ void __launch_bounds__(256)
drain_inbox_kernel(const uint32_t* __restrict__ inbox, uint32_t* cnt, uint32_t nregion,
                   const uint32_t* flags, uint32_t epoch, uint32_t* __restrict__ errors,
                   uint32_t* __restrict__ band, uint32_t band_cells,
                   uint4* __restrict__ other_band, uint32_t other_n16) {
    const int s = blockIdx.y;
    // The band buffer is multi-buffered by splat frame: while this frame's entries go into `band`,
    // the buffer an earlier frame composited from is zeroed for the next one (useful work to do
    // before waiting for the peers).
    for (uint32_t i = (blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x; i < other_n16;
         i += gridDim.x * gridDim.y * blockDim.x)
        other_band[i] = make_uint4(0u, 0u, 0u, 0u);
    if (flags) {
        __shared__ int arrived;
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            int ok = 1;
            // epochs only grow; compare as signed distance so that wrap-around stays harmless.  The
            // acquire load orders the reads of the entries (by the whole CTA, after the barrier) behind it.
            while ((int32_t)(ld_acquire_sys(flags + s) - epoch) < 0) {
                __nanosleep(100);
                if (clock64() - t0 > 4000000000LL) { ok = 0; break; } // ~2 s
            }
            if (!ok) atomicAdd(errors, 1u);
            arrived = ok;
        }
        __syncthreads();
        if (!arrived) return;
    }
    // Regions of source s are dealt round-robin to the warps (neighbouring regions are busy together:
    // the CTAs near a band edge).  A warp first loads the count words of up to 32 of its regions in one
    // go (most are zero: far ranks send nothing), then works through the non-empty ones.
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarp = gridDim.x * (blockDim.x >> 5);
    const uint64_t keep = make_normal_policy();
    for (uint32_t k0 = 0; warp + k0 * nwarp < nregion; k0 += 32u) {
        const uint32_t c = warp + (k0 + lane) * nwarp;
        uint32_t* word = cnt + (size_t)s * nregion + c;
        const uint32_t mine = (c < nregion) ? min(__ldcg(word), ROUTE_BLOCK) : 0u; // written by a peer: read at L2
        if (mine) *word = 0u;
        uint32_t todo = __ballot_sync(0xFFFFFFFFu, mine != 0u);
        while (todo) {
            const int r = __ffs(todo) - 1;
            todo &= todo - 1u;
            const uint32_t n = __shfl_sync(0xFFFFFFFFu, mine, r);
            const uint32_t* seg = inbox + ((size_t)s * nregion + warp + (k0 + (uint32_t)r) * nwarp) * ROUTE_BLOCK;
            for (uint32_t i0 = 0; i0 < n; i0 += 256u) { // 8 loads in flight per lane: the loop is latency-bound
                uint32_t e[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t i = i0 + (uint32_t)j * 32u + lane;
                    e[j] = i < n ? __ldcg(seg + i) : 0xFFFFFFFFu;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (e[j] >= band_cells) continue;
                    if (CELLS == CELL_F16) cell_add_f16(band, (long long)e[j], keep);
                    else atomicAdd(band + e[j], 1u);
                }
            }
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

// cursors[] holds the exclusive scan; each point claims the next slot of the tile it is in NOW and
// takes its whole record (position, home, reset phase, original index) along.
template <typename OffT>
__global__ void __launch_bounds__(256)
regroup_scatter_kernel(const float2* __restrict__ pos, const float2* __restrict__ orig,
                       const OffT* __restrict__ offset, const uint32_t* __restrict__ perm, uint32_t npts,
                       SortGrid sg, uint32_t* __restrict__ cursors, float2* __restrict__ pos_out,
                       float2* __restrict__ orig_out, OffT* __restrict__ offset_out,
                       uint32_t* __restrict__ perm_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const float2 p = pos[i];
    const uint32_t slot = atomicAdd(cursors + home_tile(p, sg), 1u);
    pos_out[slot] = p;
    orig_out[slot] = orig[i];
    offset_out[slot] = offset[i];
    perm_out[slot] = perm[i];
}

template <typename OffT>
__global__ void __launch_bounds__(256)
identity_order_kernel(const float2* __restrict__ pts, const uint32_t* __restrict__ offs, uint32_t npts,
                      float2* __restrict__ pos, float2* __restrict__ orig, OffT* __restrict__ offset,
                      uint32_t* __restrict__ perm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const float2 p = pts[i];
    pos[i] = p; orig[i] = p; offset[i] = (OffT)offs[i]; perm[i] = i;
}

__global__ void __launch_bounds__(256)
unpermute_kernel(const float2* __restrict__ pos, const uint32_t* __restrict__ perm, uint32_t npts,
                 float2* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npts) out[perm[i]] = pos[i];
}

} // namespace clumpy

using namespace clumpy;

constexpr int HIST_RING = 8; // counter images in flight on the single-GPU path (see advect_group)

struct clumpy_advect {
    clumpy_cuda_ctx* ctx = nullptr;
    uint32_t npts = 0, width = 0, height = 0, nframes = 0, simframe = 0;
    int kernel_size = 1, wrapx = 0;
    float step_size = 0.f, decay = 0.f;
    HistGeom geom;
    SpriteRings rings;
    float2 *pos = nullptr, *orig = nullptr, *vel = nullptr;
    void* offset = nullptr;  // uint16_t[npts] when nframes <= 65536 (off16), else uint32_t[npts]
    bool off16 = false;
    uint32_t *perm = nullptr;
    // second copy of the per-point arrays (regroup writes there, then the roles swap) + sort scratch
    float2 *pos_alt = nullptr, *orig_alt = nullptr;
    void* offset_alt = nullptr;
    uint32_t *perm_alt = nullptr, *sort_counts = nullptr, *sort_sums = nullptr;
    SortGrid sort_grid = {};
    uint32_t regroup_every = 0, frames_since_regroup = 0; // 0: never regroup
    uint32_t* hist = nullptr; // the allocation: two counter images back to back (see CellMode)
    uint32_t* hist_buf[HIST_RING] = {}; // frame s counts into hist_buf[cellbuf]
    int hist_ring = 2;        // counter images in the ring (HIST_RING, or 2 when they are too big)
    int fuse = 1;             // simulation frames per advect launch (<= hist_ring / 2)
    size_t hist_stride = 0;   // bytes between consecutive counter images
    int cellbuf = 0;
    int cell_mode = CELL_U32;
    size_t hist_bytes = 0;    // bytes of ONE counter image
    // overlap: composite + clear of frame s on ctx->aux_stream while advect of frame s+1 runs on
    // ctx->stream.  advected[b]: counts of buffer b complete; cleared[b]: buffer b zeroed again.
    bool overlap = false, side_dirty = false;
    bool keep_cells = false;  // half-float counters are accessed with L2 evict-last hints
    bool keep_pos = false;    // positions / phases are accessed with the normal L2 policy instead of
                              // evict-first (when the whole working set of a rank fits in L2)
    bool count_open = false;  // advect_count was called, advect_finish_frame not yet
    // multi-GPU routing (clumpy_cuda_advect_route_*): this rank's inbox, band buffer and peers
    struct Route {
        bool on = false;
        int rank = 0, world = 1, chunk = 0;
        int fuse = 1;                // frames per route launch (<= ROUTE_MAX_FUSE)
        int minb = 6;                // route kernel variant: 6 CTAs per SM (40 registers; measured 2 % faster) or 8 (32)
        uint32_t nregion = 0;        // regions of ROUTE_BLOCK entry slots per source segment
        uint32_t* inbox = nullptr;   // [ROUTE_RING][world sources][nregion * ROUTE_BLOCK] cell indices, then
                                     // [ROUTE_RING][world sources][nregion] count words, then [world] epoch flags
        size_t entries_per_slot = 0, cnt_offset = 0, flags_offset = 0; // in uint32 units from `inbox`
        uint32_t* counts = nullptr;  // [world] zeros: what clumpy_cuda_advect_route hands out for the caller's fence
        uint32_t* band = nullptr;    // ROUTE_RING x [halo | chunk | halo | slack] rows of counter cells: frame s
                                     // counts into slot(s); its drain zeroes the buffer of frame s + ROUTE_AHEAD
        int slot = 0;                // ring slot of the current splat frame
        cudaEvent_t routed = nullptr, drained[ROUTE_RING] = {}, composited[ROUTE_RING] = {}, side2_done = nullptr;
        bool drain_pending[ROUTE_RING] = {}, composite_pending[ROUTE_RING] = {}, side2_dirty = false;
        uint32_t* ticket = nullptr;  // [ROUTE_RING] CTAs of the route kernel that have finished the slot's frame
        size_t band_bytes = 0;       // bytes of ONE band buffer (a multiple of 256)
        uint32_t band_cells = 0;
        uint32_t* peer_inbox[ROUTE_MAX_WORLD] = {nullptr};
        uint32_t* errors = nullptr;  // fence time-outs seen by this rank's drain kernels
        uint32_t epoch_base = 0;
        bool loopback = false;       // CLUMPY_ROUTE_LOOPBACK=1: see route_begin
    } route;
    bool clear_pending[HIST_RING] = {};
    cudaEvent_t advected[2] = {nullptr, nullptr}, cleared[HIST_RING] = {}, side_done = nullptr;
    DenseTables dense;        // replay tables of the byte-counter path
    bool l2_window = false;   // an L2 persisting window for the counters is set on the stream
    uint8_t* img[2] = {nullptr, nullptr};
    int cur = 0;                                  // which img holds the current framebuffer
    bool copy_pending[2] = {false, false};        // a frame D2H was issued from img[b]
    cudaEvent_t copy_done[2] = {nullptr, nullptr};
    cudaEvent_t frame_ready = nullptr;
};

static void advect_release(clumpy_advect* a) {
    if (!a) return;
    if (a->l2_window) { // drop the window and give the persisting lines back
        cudaStreamAttrValue attr = {};
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(a->ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaCtxResetPersistingL2Cache();
        cudaGetLastError();
    }
    cudaFree(a->pos); cudaFree(a->orig); cudaFree(a->vel); cudaFree(a->offset); cudaFree(a->perm);
    cudaFree(a->pos_alt); cudaFree(a->orig_alt); cudaFree(a->offset_alt); cudaFree(a->perm_alt);
    cudaFree(a->sort_counts); cudaFree(a->sort_sums);
    cudaFree(a->route.inbox); cudaFree(a->route.counts); cudaFree(a->route.band); cudaFree(a->route.errors);
    cudaFree(a->route.ticket);
    for (int b = 0; b < ROUTE_RING; ++b) {
        if (a->route.drained[b]) cudaEventDestroy(a->route.drained[b]);
        if (a->route.composited[b]) cudaEventDestroy(a->route.composited[b]);
    }
    if (a->route.side2_done) cudaEventDestroy(a->route.side2_done);
    if (a->route.routed) cudaEventDestroy(a->route.routed);
    cudaFree(a->hist); cudaFree(a->img[0]); cudaFree(a->img[1]);
    free_dense_tables(&a->dense);
    for (int b = 0; b < 2; ++b) {
        if (a->copy_done[b]) cudaEventDestroy(a->copy_done[b]);
        if (a->advected[b]) cudaEventDestroy(a->advected[b]);
    }
    for (int b = 0; b < HIST_RING; ++b) if (a->cleared[b]) cudaEventDestroy(a->cleared[b]);
    if (a->frame_ready) cudaEventDestroy(a->frame_ready);
    if (a->side_done) cudaEventDestroy(a->side_done);
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
    const uint32_t pgrid = ceil_div(n, 256), nblocks = ceil_div(sg.ntiles, SCAN_BLOCK);
    CK(cudaMemsetAsync(a->sort_counts, 0, (size_t)sg.ntiles * 4, ctx->stream));
    sort_count_kernel<<<pgrid, 256, 0, ctx->stream>>>(a->pos, n, sg, a->sort_counts);
    CK_LAUNCH(ctx);
    scan_local_kernel<<<nblocks, SCAN_BLOCK, 0, ctx->stream>>>(a->sort_counts, sg.ntiles, a->sort_sums);
    CK_LAUNCH(ctx);
    scan_sums_kernel<<<1, SCAN_BLOCK, 0, ctx->stream>>>(a->sort_sums, nblocks);
    CK_LAUNCH(ctx);
    scan_add_kernel<<<nblocks, SCAN_BLOCK, 0, ctx->stream>>>(a->sort_counts, sg.ntiles, a->sort_sums);
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

// Fills the per-point arrays in the caller's order, sets up the sort grid and does the first regroup.
static int sort_points(clumpy_advect* a, const float2* d_in, const uint32_t* d_offs) {
    clumpy_cuda_ctx* ctx = a->ctx;
    const uint32_t n = a->npts;
    if (n == 0) return CLUMPY_OK;
    const uint32_t pgrid = ceil_div(n, 256);
    if (a->off16) identity_order_kernel<uint16_t><<<pgrid, 256, 0, ctx->stream>>>(d_in, d_offs, n, a->pos, a->orig, (uint16_t*)a->offset, a->perm);
    else identity_order_kernel<uint32_t><<<pgrid, 256, 0, ctx->stream>>>(d_in, d_offs, n, a->pos, a->orig, (uint32_t*)a->offset, a->perm);
    CK_LAUNCH(ctx);
    const char* no_sort = getenv("CLUMPY_CUDA_NO_SORT");
    if (no_sort && no_sort[0] == '1') return CLUMPY_OK;
    if (a->cell_mode == CELL_X64) return CLUMPY_OK; // exact-order path: slot == original index

    SortGrid sg;
    sg.width = a->width; sg.height = a->height;
    // 8 x 4 texels per tile (measured best on B200 among 8x4, 16x2, 16x4, 32x1, 16x1).
    // CLUMPY_CUDA_SORT_TILE="sx,sy" (log2) overrides (A/B runs).
    sg.shift_x = 3; sg.shift_y = 2;
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
    a->sort_grid = sg;
    const size_t nalloc = std::max<size_t>((size_t)n + 1, 2);
    CK(cudaMalloc(&a->sort_counts, (size_t)sg.ntiles * 4));
    CK(cudaMalloc(&a->sort_sums, (size_t)ceil_div(sg.ntiles, SCAN_BLOCK) * 4));
    CK(cudaMalloc(&a->pos_alt, nalloc * 8));
    CK(cudaMalloc(&a->orig_alt, nalloc * 8));
    CK(cudaMalloc(&a->offset_alt, nalloc * 4));
    CK(cudaMalloc(&a->perm_alt, nalloc * 4));
    // Regroup period: by then about 1/16 of the points have been reset away from their group.
    // Small clouds do not need it.  CLUMPY_CUDA_REGROUP=<frames> overrides, 0 disables.
    a->regroup_every = n >= (1u << 16) ? std::max<uint32_t>(8u, a->nframes / 16u) : 0u;
    if (const char* r = getenv("CLUMPY_CUDA_REGROUP")) a->regroup_every = (uint32_t)atoi(r);
    return regroup_points(a);
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
    // Counter cells: bytes with replay tables when the sprite allows it (every advect_points sprite
    // up to k = 7 does), 32-bit counters otherwise.  CLUMPY_CUDA_CELLS=u32 forces the latter (A/B runs).
    a->off16 = a->nframes <= 65536u;
    a->cell_mode = CELL_U32;
    //   auto (default)  sparse cloud (8 * npts <= pixels): exact-order 64-bit cells, points kept in
    //                   the caller's order; otherwise half-float cells + replay tables, points
    //                   regrouped by tile; 32-bit cells when the sprite has no tables (k > 7)
    //   CLUMPY_CUDA_CELLS = exact | f16 | u8 | u32 forces one (A/B runs, tests)
    {
        const char* force = getenv("CLUMPY_CUDA_CELLS");
        std::string want = force ? force : "auto";
        const bool can_exact = exact_mode_supported(a->kernel_size, a->npts);
        // Dense clouds: table replay goes ring by ring, which is within 1 LSB when every ring
        // contracts strongly (k <= 3: opacities 0.65 / 0.5 / 0.35).  Larger sprites have a 0.028 ring;
        // for them the 32-bit path with its proportionally interleaved replay is the accurate one.
        if (want == "auto")
            want = (can_exact && 8ull * a->npts <= (uint64_t)npix) ? "exact" : (a->kernel_size <= 3 ? "f16" : "u32");
        if (want == "exact" && can_exact) {
            a->cell_mode = CELL_X64;
        } else if (want != "u32") {
            int rc = build_dense_tables(ctx, a->rings, a->decay, &a->dense);
            if (rc) return rc;
            if (a->dense.valid) a->cell_mode = (want == "u8") ? CELL_U8 : CELL_F16;
        }
    }
    bool composite_cap_from_env = false;
    a->hist_bytes = a->cell_mode == CELL_U8 ? a->geom.cells()
                  : a->cell_mode == CELL_F16 ? a->geom.cells() * 2
                  : a->cell_mode == CELL_X64 ? a->geom.cells() * 8 : a->geom.bytes();
    // Two counter images: while frame s is composited and its image cleared on the side stream,
    // frame s+1 already counts into the other one.  CLUMPY_CUDA_OVERLAP=0 runs everything on one
    // stream (A/B runs); small clouds gain nothing from the overlap and skip it.
    {
        const char* ov = getenv("CLUMPY_CUDA_OVERLAP");
        a->overlap = ov ? (ov[0] != '0') : (a->npts >= (1u << 18));
        const char* kc = getenv("CLUMPY_CUDA_KEEP_CELLS");
        // L2 evict-last hints on the counters: measured to make no difference on B200 (0.2132 vs
        // 0.2173 ms per step), so off unless asked for.
        a->keep_cells = a->cell_mode == CELL_F16 && (kc ? kc[0] != '0' : false);
        ctx->keep_cells = a->keep_cells ? 1 : 0;
        if (const char* kp = getenv("CLUMPY_CUDA_KEEP_POS")) a->keep_pos = kp[0] == '1';
        if (const char* cap = getenv("CLUMPY_CUDA_COMPOSITE_CTAS")) ctx->composite_ctas_per_sm = atoi(cap);
        else ctx->composite_ctas_per_sm = a->overlap ? 4 : 0;
        composite_cap_from_env = getenv("CLUMPY_CUDA_COMPOSITE_CTAS") != nullptr;
    }
    // A ring of counter images: one advect launch advances `fuse` frames, each counting into the next
    // image of the ring, while the images of the previous group are still being composited and cleared
    // on the side stream (so the ring holds two groups).  Very large images (> 512 MB each) fall back to
    // two images and one frame per launch.  CLUMPY_CUDA_FUSE=<1..4> overrides (1 = the one-frame kernel).
    const size_t image_stride = round_up(a->hist_bytes, 256);
    a->hist_stride = image_stride;
    a->hist_ring = image_stride <= (512ull << 20) ? HIST_RING : 2;
    a->fuse = a->hist_ring / 2;
    if (const char* fz = getenv("CLUMPY_CUDA_FUSE")) a->fuse = std::min(std::max(atoi(fz), 1), a->hist_ring / 2);
    if (a->fuse == 1) a->hist_ring = 2;
    // With several frames per advect launch the composites are the longer stage of the pipeline: give
    // them 5 CTAs per SM instead of 4 (measured on C4: step 0.199 -> 0.177 ms, advect launch 0.417 ->
    // 0.467 ms per 4 frames; 6 CTAs: 0.175 / 0.518, 8 CTAs: 0.179 / 0.609).
    if (a->overlap && a->fuse > 1 && !composite_cap_from_env) ctx->composite_ctas_per_sm = 5;
    CK(cudaMalloc(&a->hist, (size_t)a->hist_ring * image_stride));
    for (int b = 0; b < a->hist_ring; ++b)
        a->hist_buf[b] = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(a->hist) + (size_t)b * image_stride);
    CK(cudaMalloc(&a->img[0], npix));
    CK(cudaMalloc(&a->img[1], npix));
    for (int b = 0; b < 2; ++b) {
        CK(cudaEventCreateWithFlags(&a->copy_done[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&a->advected[b], cudaEventDisableTiming));
    }
    for (int b = 0; b < HIST_RING; ++b) CK(cudaEventCreateWithFlags(&a->cleared[b], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&a->frame_ready, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&a->side_done, cudaEventDisableTiming));
    CK(cudaMemsetAsync(a->hist, 0, (size_t)a->hist_ring * image_stride, ctx->stream));
    // Optional (CLUMPY_CUDA_PERSIST=1): an L2 persisting window over the counters, which are touched
    // three times per frame while ~0.6 GB of streaming traffic passes by.  Measured on B200 it is a
    // wash for one 64 MB image (0.165 vs 0.178 ms advect) and a 2.3x LOSS for the two images of the
    // overlapped schedule, because the carve-out is taken away from the velocity field's share of L2;
    // so it is off by default.
    {
        const char* on = getenv("CLUMPY_CUDA_PERSIST");
        cudaDeviceProp prop;
        if ((a->cell_mode == CELL_U8 || a->cell_mode == CELL_F16) && (on && on[0] == '1') &&
            cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
            const size_t window = std::min<size_t>((size_t)a->hist_ring * image_stride, (size_t)prop.accessPolicyMaxWindowSize);
            const size_t want = std::min<size_t>(window, (size_t)prop.persistingL2CacheMaxSize);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaStreamAttrValue attr = {};
            attr.accessPolicyWindow.base_ptr = a->hist;
            attr.accessPolicyWindow.num_bytes = window;
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)window);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
            cudaStreamSetAttribute(ctx->aux_stream, cudaStreamAttributeAccessPolicyWindow, &attr);
            cudaGetLastError(); // best effort: a refusal only costs bandwidth
            a->l2_window = true;
        }
    }
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
    // Multi-GPU callers split the counter image into equal row bands (one per rank, whole 16-row
    // tiles): they ask for the padded row count to be a multiple of 16 * world_size.
    if (const char* m = getenv("CLUMPY_CUDA_CELL_ROWS_MULTIPLE")) {
        const int mult = atoi(m);
        if (mult > 1 && mult <= (1 << 20)) a->geom.rows = (int)round_up((size_t)a->geom.rows, (size_t)mult);
    }
    a->rings = rings;
    rc = advect_begin_impl(a, pts_host, vel_host, age_offset);
    if (rc) { advect_release(a); return rc; }
    *out = a;
    return CLUMPY_OK;
}

template <bool WRAP, int MODE, typename OffT>
static void launch_advect_t(clumpy_advect* a, uint32_t phase, int nf) {
    const uint32_t grid = ceil_div(a->npts, ADVECT_THREADS * ADVECT_PTS);
    const GeomDev g{a->geom.width, a->geom.height, a->geom.halo, a->geom.pitch, a->keep_pos ? 1 : 0};
    if (nf == 1)
        advect_kernel<WRAP, MODE, OffT><<<grid, ADVECT_THREADS, 0, a->ctx->stream>>>(
            a->pos, a->orig, (const OffT*)a->offset, a->npts, a->vel, a->width, a->height, a->step_size, phase, g,
            a->hist_buf[a->cellbuf], a->dense.cell_cap, a->keep_cells ? 1 : 0);
    else
        advect_fused_kernel<WRAP, MODE, OffT><<<grid, ADVECT_THREADS, 0, a->ctx->stream>>>(
            a->pos, a->orig, (const OffT*)a->offset, a->npts, a->vel, a->width, a->height, a->step_size, phase,
            a->nframes, nf, g, a->hist, a->hist_stride / sizeof(uint32_t), a->cellbuf, a->hist_ring,
            a->dense.cell_cap, a->keep_cells ? 1 : 0);
}

template <bool WRAP, int MODE>
static void launch_advect_m(clumpy_advect* a, uint32_t phase, int nf) {
    if (a->off16) launch_advect_t<WRAP, MODE, uint16_t>(a, phase, nf);
    else launch_advect_t<WRAP, MODE, uint32_t>(a, phase, nf);
}

template <bool WRAP>
static void launch_advect(clumpy_advect* a, uint32_t phase, bool splat, int nf = 1) {
    if (!splat) launch_advect_m<WRAP, CELL_NONE>(a, phase, nf);
    else if (a->cell_mode == CELL_U8) launch_advect_m<WRAP, CELL_U8>(a, phase, nf);
    else if (a->cell_mode == CELL_F16) launch_advect_m<WRAP, CELL_F16>(a, phase, nf);
    else if (a->cell_mode == CELL_X64) launch_advect_m<WRAP, CELL_X64>(a, phase, nf);
    else launch_advect_m<WRAP, CELL_U32>(a, phase, nf);
}

// How many frames the next advect launch may cover: at most `want` and a->fuse, never across a
// regroup or the warm-up / recording boundary (where, with decay 0, frames start to splat).
static uint32_t advect_group_size(const clumpy_advect* a, uint32_t want) {
    uint32_t nf = std::min<uint32_t>(std::max<uint32_t>(want, 1u), (uint32_t)a->fuse);
    nf = std::min<uint32_t>(nf, a->nframes - a->simframe % a->nframes);
    if (a->regroup_every) nf = std::min<uint32_t>(nf, a->regroup_every - std::min(a->frames_since_regroup, a->regroup_every - 1u));
    return std::max<uint32_t>(nf, 1u);
}

// `nf` simulation frames: ONE advect launch on the main stream (frame f of the group counts into
// counter image (cellbuf + f) % ring), then per frame a composite and a clear of its counter image on
// `side` -- the high-priority aux stream when overlapping, so that they share the SMs with the advect
// launch of the NEXT group (the advect kernel is DRAM-bound, the composite issue-bound), the main
// stream otherwise and always while profiling.
// ev (optional, nf = 1): 4 events recorded before the advect kernel, after it, after the composite kernel
// and after the counter clear, for clumpy_cuda_advect_profile.  kev (optional): 2 events around the
// advect launch alone, in the pipeline as it runs (clumpy_cuda_advect_profile_fused).
static int advect_group(clumpy_advect* a, uint32_t nf, cudaEvent_t* ev = nullptr, cudaEvent_t* kev = nullptr) {
    clumpy_cuda_ctx* ctx = a->ctx;
    const uint32_t s = a->simframe;
    const bool warm = s < a->nframes;
    const bool splat = !(warm && a->decay == 0.0f); // advect_points.cc:154
    const uint32_t phase = s % a->nframes;
    const bool overlap = a->overlap && !ev;
    const cudaStream_t main_stream = ctx->stream;
    const cudaStream_t side = overlap ? ctx->aux_stream : main_stream;
    const int ring = a->hist_ring, b0 = a->cellbuf;
    if (splat)
        for (uint32_t f = 0; f < nf; ++f) { // these images were last used ring frames ago: wait for their clears
            const int b = (b0 + (int)f) % ring;
            if (a->clear_pending[b]) {
                CK(cudaStreamWaitEvent(main_stream, a->cleared[b], 0));
                a->clear_pending[b] = false;
            }
        }
    if (ev) CK(cudaEventRecord(ev[0], main_stream));
    if (kev) CK(cudaEventRecord(kev[0], main_stream));
    if (a->npts) {
        if (a->wrapx) launch_advect<true>(a, phase, splat, (int)nf);
        else launch_advect<false>(a, phase, splat, (int)nf);
        CK_LAUNCH(ctx);
    }
    if (ev) CK(cudaEventRecord(ev[1], main_stream));
    if (kev) CK(cudaEventRecord(kev[1], main_stream));
    if (splat) {
        if (overlap) {
            CK(cudaEventRecord(a->advected[0], main_stream));
            CK(cudaStreamWaitEvent(side, a->advected[0], 0));
        }
        for (uint32_t f = 0; f < nf; ++f) {
            const int b = (b0 + (int)f) % ring;
            uint32_t* const cells = a->hist_buf[b];
            // write in place unless a frame copy is still reading the current buffer
            int dst = a->cur;
            if (a->copy_pending[a->cur]) {
                dst = a->cur ^ 1;
                if (a->copy_pending[dst]) {
                    CK(cudaStreamWaitEvent(side, a->copy_done[dst], 0));
                    a->copy_pending[dst] = false;
                }
            }
            ctx->stream = side; // the launchers enqueue on ctx->stream
            int rc;
            if (a->cell_mode == CELL_U8)
                rc = launch_composite_dense(ctx, a->geom, a->dense, reinterpret_cast<const uint8_t*>(cells), a->wrapx,
                                            a->img[a->cur], a->img[dst]);
            else if (a->cell_mode == CELL_F16)
                rc = launch_composite_half(ctx, a->geom, a->dense, cells, a->wrapx, a->img[a->cur], a->img[dst]);
            else if (a->cell_mode == CELL_X64)
                rc = launch_composite_exact(ctx, a->geom, a->rings, cells, a->wrapx, 1, a->decay, a->img[a->cur], a->img[dst]);
            else
                rc = launch_composite(ctx, a->geom, a->rings, cells, a->wrapx, 1, a->decay, a->img[a->cur], a->img[dst]);
            ctx->stream = main_stream;
            if (rc) return rc;
            a->cur = dst;
            if (ev) CK(cudaEventRecord(ev[2], main_stream));
            if (a->keep_cells && a->hist_bytes % 16 == 0) {
                const size_t n16 = a->hist_bytes / 16;
                const uint32_t grid = (uint32_t)std::min<size_t>((n16 + 255) / 256, (size_t)ctx->sm_count * 8);
                clear_cells_kernel<<<grid, 256, 0, side>>>(reinterpret_cast<uint4*>(cells), n16);
                CK_LAUNCH(ctx);
            } else {
                CK(cudaMemsetAsync(cells, 0, a->hist_bytes, side));
            }
            if (overlap) {
                CK(cudaEventRecord(a->cleared[b], side));
                a->clear_pending[b] = true;
                a->side_dirty = true;
            }
        }
        // One-stream runs have cleared every image again by the time the next launch starts; the ring
        // only matters (and only advances) when composites lag behind on the side stream or a group
        // spans several images.
        if (overlap || nf > 1) a->cellbuf = (b0 + (int)nf) % ring;
    } else if (ev) {
        CK(cudaEventRecord(ev[2], main_stream));
    }
    if (ev) CK(cudaEventRecord(ev[3], main_stream));
    a->simframe = s + nf;
    a->frames_since_regroup += nf;
    if (a->regroup_every && a->frames_since_regroup >= a->regroup_every) return regroup_points(a);
    return CLUMPY_OK;
}

// Makes the main stream wait for everything issued on the side stream, so that whatever the caller
// enqueues (or records) on the context's stream next is ordered after the whole frame.
static int advect_join(clumpy_advect* a) {
    if (a->route.side2_dirty) {
        CK(cudaEventRecord(a->route.side2_done, a->ctx->aux2_stream));
        CK(cudaStreamWaitEvent(a->ctx->stream, a->route.side2_done, 0));
        a->route.side2_dirty = false;
    }
    if (!a->side_dirty) return CLUMPY_OK;
    CK(cudaEventRecord(a->side_done, a->ctx->aux_stream));
    CK(cudaStreamWaitEvent(a->ctx->stream, a->side_done, 0));
    a->side_dirty = false;
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_step(clumpy_advect* adv, uint32_t count) {
    REQUIRE(adv, "adv is NULL");
    CK(cudaSetDevice(adv->ctx->device));
    for (uint32_t i = 0; i < count;) {
        const uint32_t nf = advect_group_size(adv, count - i);
        int rc = advect_group(adv, nf);
        if (rc) return rc;
        i += nf;
    }
    return advect_join(adv);
}

CLUMPY_API int clumpy_cuda_advect_profile(clumpy_advect* adv, uint32_t count, float* advect_ms,
                                          float* composite_ms, float* clear_ms) {
    REQUIRE(adv && count > 0, "bad argument");
    CK(cudaSetDevice(adv->ctx->device));
    std::vector<cudaEvent_t> ev(4 * (size_t)count, nullptr);
    int rc = CLUMPY_OK;
    for (auto& e : ev)
        if (cudaEventCreate(&e) != cudaSuccess) { set_error("event creation failed"); rc = CLUMPY_ERR_CUDA; break; }
    for (uint32_t i = 0; rc == CLUMPY_OK && i < count; ++i) rc = advect_group(adv, 1, &ev[4 * (size_t)i]);
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

// The advect launch as the pipeline runs it -- several frames per launch, composites of the previous
// group sharing the GPU -- timed with CUDA events around the launch alone.  Runs `groups` full groups
// (groups cut short by a regroup or the recording boundary are run but not counted).  Synchronises.
CLUMPY_API int clumpy_cuda_advect_profile_fused(clumpy_advect* adv, uint32_t groups, float* launch_ms,
                                                uint32_t* frames_per_launch) {
    REQUIRE(adv && groups > 0 && launch_ms && frames_per_launch, "bad argument");
    CK(cudaSetDevice(adv->ctx->device));
    std::vector<cudaEvent_t> ev(2 * (size_t)groups, nullptr);
    std::vector<bool> full(groups, false);
    int rc = CLUMPY_OK;
    for (auto& e : ev)
        if (cudaEventCreate(&e) != cudaSuccess) { set_error("event creation failed"); rc = CLUMPY_ERR_CUDA; break; }
    for (uint32_t g = 0; rc == CLUMPY_OK && g < groups; ++g) {
        const uint32_t nf = advect_group_size(adv, (uint32_t)adv->fuse);
        full[g] = nf == (uint32_t)adv->fuse;
        rc = advect_group(adv, nf, nullptr, &ev[2 * (size_t)g]);
    }
    if (rc == CLUMPY_OK) rc = advect_join(adv);
    if (rc == CLUMPY_OK && cudaStreamSynchronize(adv->ctx->stream) != cudaSuccess) { set_error("profile run failed"); rc = CLUMPY_ERR_CUDA; }
    double total = 0;
    uint32_t counted = 0;
    for (uint32_t g = 0; rc == CLUMPY_OK && g < groups; ++g) {
        if (!full[g]) continue;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[2 * (size_t)g], ev[2 * (size_t)g + 1]);
        total += ms;
        ++counted;
    }
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    *launch_ms = counted ? (float)(total / counted) : 0.f;
    *frames_per_launch = (uint32_t)adv->fuse;
    return rc;
}

// ---- one frame split at the exchange point (multi-GPU: points sharded, counters summed) ----------

static bool frame_splats(const clumpy_advect* a) {
    return !(a->simframe < a->nframes && a->decay == 0.0f); // advect_points.cc:154
}

CLUMPY_API int clumpy_cuda_advect_count(clumpy_advect* adv) {
    REQUIRE(adv, "adv is NULL");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    int rc = advect_join(adv);
    if (rc) return rc;
    REQUIRE(!adv->count_open, "advect_count called twice without advect_finish_frame");
    const bool splat = frame_splats(adv);
    if (splat && adv->clear_pending[adv->cellbuf]) {
        CK(cudaStreamWaitEvent(ctx->stream, adv->cleared[adv->cellbuf], 0));
        adv->clear_pending[adv->cellbuf] = false;
    }
    if (adv->npts) {
        const uint32_t phase = adv->simframe % adv->nframes;
        if (adv->wrapx) launch_advect<true>(adv, phase, splat);
        else launch_advect<false>(adv, phase, splat);
        CK_LAUNCH(ctx);
    }
    adv->count_open = true;
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_cells(clumpy_advect* adv, void** cells_dev, size_t* bytes,
                                        uint32_t* pitch_cells, uint32_t* rows, uint32_t* halo,
                                        uint32_t* bytes_per_cell) {
    REQUIRE(adv, "adv is NULL");
    if (cells_dev) *cells_dev = adv->hist_buf[adv->cellbuf];
    if (bytes) *bytes = adv->hist_bytes;
    if (pitch_cells) *pitch_cells = (uint32_t)adv->geom.pitch;
    if (rows) *rows = (uint32_t)adv->geom.rows;
    if (halo) *halo = (uint32_t)adv->geom.halo;
    if (bytes_per_cell)
        *bytes_per_cell = adv->cell_mode == CELL_U8 ? 1u : adv->cell_mode == CELL_F16 ? 2u : adv->cell_mode == CELL_X64 ? 8u : 4u;
    return CLUMPY_OK;
}

static int composite_rows_impl(clumpy_advect* adv, const void* cells_dev, uint32_t row0, uint32_t nrows);

CLUMPY_API int clumpy_cuda_advect_composite_rows(clumpy_advect* adv, const void* cells_dev,
                                                 uint32_t row0, uint32_t nrows) {
    REQUIRE(adv && cells_dev, "null argument");
    REQUIRE((uint64_t)row0 + nrows <= adv->height, "row band exceeds the image");
    CK(cudaSetDevice(adv->ctx->device));
    if (!frame_splats(adv) || nrows == 0) return CLUMPY_OK;
    return composite_rows_impl(adv, cells_dev, row0, nrows);
}

static int composite_rows_impl(clumpy_advect* adv, const void* cells_dev, uint32_t row0, uint32_t nrows) {
    clumpy_cuda_ctx* ctx = adv->ctx;
    // the band as an image of its own: same width and pitch, `nrows` rows, cells given by the caller
    HistGeom g = adv->geom;
    g.height = (int)nrows;
    uint8_t* img = adv->img[adv->cur] + (size_t)row0 * adv->width;
    if (adv->cell_mode == CELL_U8)
        return launch_composite_dense(ctx, g, adv->dense, reinterpret_cast<const uint8_t*>(cells_dev), adv->wrapx, img, img);
    if (adv->cell_mode == CELL_F16) return launch_composite_half(ctx, g, adv->dense, cells_dev, adv->wrapx, img, img);
    if (adv->cell_mode == CELL_X64) return launch_composite_exact(ctx, g, adv->rings, cells_dev, adv->wrapx, 1, adv->decay, img, img);
    return launch_composite(ctx, g, adv->rings, reinterpret_cast<const uint32_t*>(cells_dev), adv->wrapx, 1, adv->decay, img, img);
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

CLUMPY_API int clumpy_cuda_advect_route_begin(clumpy_advect* adv, int rank, int world, uint32_t capacity,
                                              void** inbox_dev, size_t* inbox_bytes) {
    REQUIRE(adv && inbox_dev && inbox_bytes, "null argument");
    REQUIRE(world >= 1 && world <= ROUTE_MAX_WORLD && rank >= 0 && rank < world, "bad rank / world size");
    REQUIRE(adv->cell_mode == CELL_F16 || adv->cell_mode == CELL_U32, "routing needs count-only cells (f16 or u32)");
    REQUIRE(capacity >= adv->npts, "capacity must be at least the largest shard");
    REQUIRE(!adv->route.on, "routing already set up");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    auto& rt = adv->route;
    rt.rank = rank; rt.world = world;
    rt.nregion = std::max<uint32_t>(ceil_div(capacity, ROUTE_BLOCK), 1u); // one region per CTA of the largest shard
    // Bands of `chunk` padded rows, whole 16-row composite tiles, chosen from the IMAGE height so that
    // for a uniform cloud they coincide with the home-row strips of the shards (4096 rows on 8 ranks: 512
    // each, not the 528 that an even split of the 4098 padded rows would give: with those, rank 7 would
    // start with a fifth of its points in rank 6's band).  The last band also takes whatever is left
    // (the 2 * halo padding rows), which fits in the 32 slack rows of its band buffer.
    rt.chunk = (int)round_up(ceil_div(adv->height, (uint32_t)world), 16);
    REQUIRE(world == 1 || rt.chunk >= 2 * adv->geom.halo, "bands thinner than the sprite are not supported");
    rt.fuse = ROUTE_MAX_FUSE;
    if (const char* fz = getenv("CLUMPY_ROUTE_FUSE")) rt.fuse = std::min(std::max(atoi(fz), 1), ROUTE_MAX_FUSE);
    if (const char* mb = getenv("CLUMPY_ROUTE_MINB")) rt.minb = atoi(mb) == 8 ? 8 : 6;
    const size_t cell_bytes = adv->cell_mode == CELL_F16 ? 2 : 4;
    const size_t band_rows = (size_t)rt.chunk + 2 * adv->geom.halo + 32;
    rt.band_cells = (uint32_t)(band_rows * adv->geom.pitch);
    rt.band_bytes = round_up((size_t)rt.band_cells * cell_bytes, 256);
    rt.entries_per_slot = (size_t)world * rt.nregion * ROUTE_BLOCK;
    rt.cnt_offset = ROUTE_RING * rt.entries_per_slot;
    rt.flags_offset = round_up(rt.cnt_offset + (size_t)ROUTE_RING * world * rt.nregion, 64);
    const size_t ibytes = round_up((rt.flags_offset + (size_t)world) * sizeof(uint32_t), 256);
    CK(cudaMalloc(&rt.inbox, ibytes));
    CK(cudaMalloc(&rt.counts, ROUTE_MAX_WORLD * sizeof(uint32_t)));
    CK(cudaMalloc(&rt.band, ROUTE_RING * rt.band_bytes));
    for (int b = 0; b < ROUTE_RING; ++b) {
        CK(cudaEventCreateWithFlags(&rt.drained[b], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&rt.composited[b], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&rt.side2_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&rt.routed, cudaEventDisableTiming));
    CK(cudaMalloc(&rt.errors, sizeof(uint32_t)));
    CK(cudaMalloc(&rt.ticket, ROUTE_RING * sizeof(uint32_t)));
    CK(cudaMemsetAsync(rt.ticket, 0, ROUTE_RING * sizeof(uint32_t), ctx->stream));
    // count words and flags start at zero (the entries need no initialisation)
    CK(cudaMemsetAsync(rt.inbox + rt.cnt_offset, 0, ibytes - rt.cnt_offset * sizeof(uint32_t), ctx->stream));
    CK(cudaMemsetAsync(rt.errors, 0, sizeof(uint32_t), ctx->stream));
    CK(cudaMemsetAsync(rt.counts, 0, ROUTE_MAX_WORLD * sizeof(uint32_t), ctx->stream));
    CK(cudaMemsetAsync(rt.band, 0, ROUTE_RING * rt.band_bytes, ctx->stream));
    // Load every kernel of the frame loop NOW.  With lazy module loading the first launch of a kernel
    // can block the host until the device is idle -- and a drain kernel may be sitting on the device
    // waiting for the epoch of a rank that this very host thread has yet to launch (one process
    // driving several ranks).  The drain and the composite run once on the all-zero band, which leaves
    // the (zero) framebuffer as it is; the route kernels are loaded without being run.
    {
        cudaFuncAttributes fa;
#define ROUTE_LOAD(W, S) CK(adv->off16 ? (rt.minb == 8 ? cudaFuncGetAttributes(&fa, advect_route_kernel<W, S, uint16_t, 8>) \
                                                        : cudaFuncGetAttributes(&fa, advect_route_kernel<W, S, uint16_t, 6>)) \
                                       : (rt.minb == 8 ? cudaFuncGetAttributes(&fa, advect_route_kernel<W, S, uint32_t, 8>) \
                                                        : cudaFuncGetAttributes(&fa, advect_route_kernel<W, S, uint32_t, 6>)))
        if (adv->wrapx) { ROUTE_LOAD(true, true); ROUTE_LOAD(true, false); }
        else            { ROUTE_LOAD(false, true); ROUTE_LOAD(false, false); }
#undef ROUTE_LOAD
        CK(cudaFuncGetAttributes(&fa, route_signal_kernel));
        uint32_t* cnt0 = rt.inbox + rt.cnt_offset;
        uint4* band1 = reinterpret_cast<uint4*>(reinterpret_cast<char*>(rt.band) + rt.band_bytes);
        const dim3 grid(1, world);
        if (adv->cell_mode == CELL_F16)
            drain_inbox_kernel<CELL_F16><<<grid, 256, 0, ctx->stream>>>(rt.inbox, cnt0, rt.nregion, nullptr, 0u, rt.errors,
                                                                         rt.band, rt.band_cells, band1, (uint32_t)(rt.band_bytes / 16));
        else
            drain_inbox_kernel<CELL_U32><<<grid, 256, 0, ctx->stream>>>(rt.inbox, cnt0, rt.nregion, nullptr, 0u, rt.errors,
                                                                         rt.band, rt.band_cells, band1, (uint32_t)(rt.band_bytes / 16));
        CK_LAUNCH(ctx);
        const int h = adv->geom.halo, lo = rank * rt.chunk;
        const int y0 = std::max(0, lo - h), y1 = route_last_row(adv);
        if (y1 > y0) {
            const char* cells = reinterpret_cast<const char*>(rt.band) + (size_t)(y0 - (lo - h)) * adv->geom.pitch * cell_bytes;
            int rc = composite_rows_impl(adv, cells, (uint32_t)y0, (uint32_t)(y1 - y0));
            if (rc) return rc;
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    // The texels a rank gathers lie in and around its own band of the field (home-row sharding): a
    // few tens of MB that every frame reads again, while ~40 MB of positions stream past.  An L2
    // persisting window over them (on the main stream, where only route kernels run) measured no
    // difference on B200 (2 ranks, 1/4 of C4 each: 0.0411 vs 0.0408 ms per frame), so it is off unless
    // CLUMPY_ROUTE_PERSIST_FIELD=1 (A/B runs).
    {
        const char* pf = getenv("CLUMPY_ROUTE_PERSIST_FIELD");
        cudaDeviceProp prop;
        if (world > 1 && (pf && pf[0] == '1') && cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess &&
            prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
            const int h = adv->geom.halo, margin = std::max(16, rt.chunk / 4);
            const long long r0 = std::max<long long>(0, (long long)rank * rt.chunk - h - margin);
            const long long r1 = std::min<long long>(adv->height, (long long)(rank + 1) * rt.chunk - h + margin);
            if (r1 > r0) {
                const size_t row_bytes = (size_t)adv->width * sizeof(float2);
                size_t window = std::min<size_t>((size_t)(r1 - r0) * row_bytes, (size_t)prop.accessPolicyMaxWindowSize);
                const size_t carve = std::min<size_t>(window, (size_t)prop.persistingL2CacheMaxSize);
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
                cudaStreamAttrValue attr = {};
                attr.accessPolicyWindow.base_ptr = reinterpret_cast<char*>(adv->vel) + (size_t)r0 * row_bytes;
                attr.accessPolicyWindow.num_bytes = window;
                attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)window);
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
                cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
                cudaGetLastError(); // best effort: a refusal only costs bandwidth
                adv->l2_window = true;
            }
        }
    }
    rt.peer_inbox[rank] = rt.inbox;
    // Profiling aid (ncu must not run a multi-rank job): with CLUMPY_ROUTE_LOOPBACK=1 one process plays
    // rank `rank` of `world` alone.  The caller maps its own inbox as every peer's; the fence then
    // publishes all `world` epochs into its own flags.  Timing and memory traffic are about those of a
    // real rank, the FRAMES ARE NOT (entries for other bands pile up in this rank's inbox).
    if (const char* lb = getenv("CLUMPY_ROUTE_LOOPBACK")) rt.loopback = lb[0] == '1';
    rt.on = true;
    *inbox_dev = rt.inbox;
    *inbox_bytes = ibytes;
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_route_peers(clumpy_advect* adv, void* const* peer_inbox) {
    REQUIRE(adv && peer_inbox && adv->route.on, "routing is not set up");
    for (int r = 0; r < adv->route.world; ++r)
        adv->route.peer_inbox[r] = (r == adv->route.rank) ? adv->route.inbox : static_cast<uint32_t*>(peer_inbox[r]);
    for (int r = 0; r < adv->route.world; ++r) REQUIRE(adv->route.peer_inbox[r], "missing peer inbox");
    return CLUMPY_OK;
}

template <bool WRAP, bool SPLAT, typename OffT>
static void launch_route_t(clumpy_advect* a, const RouteDev& rd) {
    // CTA c owns region c of its segments; a fenced launch has one extra CTA, the publisher
    const uint32_t grid = ceil_div(a->npts, ROUTE_BLOCK) + ((SPLAT && rd.signal) ? 1u : 0u);
    const GeomDev g{a->geom.width, a->geom.height, a->geom.halo, a->geom.pitch, a->keep_pos ? 1 : 0};
    if (a->route.minb == 8)
        advect_route_kernel<WRAP, SPLAT, OffT, 8><<<grid, ADVECT_THREADS, 0, a->ctx->stream>>>(
            a->pos, a->orig, (const OffT*)a->offset, a->npts, a->vel, a->width, a->height, a->step_size, g, rd);
    else
        advect_route_kernel<WRAP, SPLAT, OffT, 6><<<grid, ADVECT_THREADS, 0, a->ctx->stream>>>(
            a->pos, a->orig, (const OffT*)a->offset, a->npts, a->vel, a->width, a->height, a->step_size, g, rd);
}

// One route launch covering `nf` consecutive frames, starting at the current one.  fenced: the
// kernel itself publishes every frame's epoch to the peers (route_steps); otherwise the caller fences
// with a collective of its own (nf = 1).
static int route_launch(clumpy_advect* adv, bool fenced, int nf, uint32_t epoch0) {
    clumpy_cuda_ctx* ctx = adv->ctx;
    auto& rt = adv->route;
    const bool splat = frame_splats(adv);
    const int slot0 = rt.slot; // ring slots (band buffer, inbox parity, count words, ticket) cycle per splat frame
    // Frame s counts into band buffer slot(s), which the drain of frame s - ROUTE_AHEAD zeroed on its
    // side stream; that drain also saw every peer's epoch of its frame, so every peer had by then
    // launched -- after ITS drain of ROUTE_AHEAD frames earlier -- the kernel covering that frame, and
    // so is done reading the inbox slot this frame writes (ROUTE_RING = 2 * ROUTE_AHEAD frames back).
    // Drains complete in order, so waiting for the one the LAST frame of the group needs covers them all.
    if (splat) {
        const int need = (slot0 + nf - 1 + ROUTE_RING - ROUTE_AHEAD) % ROUTE_RING;
        if (rt.drain_pending[need]) {
            CK(cudaStreamWaitEvent(ctx->stream, rt.drained[need], 0));
            rt.drain_pending[need] = false;
        }
    }
    RouteDev rd = {};
    rd.me = rt.rank; rd.world = rt.world; rd.chunk = rt.chunk; rd.halo = adv->geom.halo; rd.pitch = adv->geom.pitch;
    rd.nregion = rt.nregion;
    rd.nframes = adv->nframes; rd.phase0 = adv->simframe % adv->nframes;
    rd.nf = nf; rd.slot0 = slot0;
    rd.inbox_stride = rt.entries_per_slot;
    rd.cnt_stride = (size_t)rt.world * rt.nregion;
    rd.band_stride = rt.band_bytes / sizeof(uint32_t);
    rd.band = rt.band;
    rd.band_f16 = adv->cell_mode == CELL_F16 ? 1 : 0;
    rd.signal = fenced ? 1 : 0;
    rd.epoch0 = epoch0;
    rd.loopback = rt.loopback ? 1 : 0;
    rd.ticket = rt.ticket;
    for (int r = 0; r < rt.world; ++r) {
        rd.inbox[r] = rt.peer_inbox[r];
        rd.cnt[r] = rt.peer_inbox[r] + rt.cnt_offset;
        rd.flags[r] = rt.peer_inbox[r] + rt.flags_offset;
    }
    if (adv->npts) {
#define ROUTE_LAUNCH(W, S) (adv->off16 ? launch_route_t<W, S, uint16_t>(adv, rd) : launch_route_t<W, S, uint32_t>(adv, rd))
        if (adv->wrapx) { if (splat) ROUTE_LAUNCH(true, true); else ROUTE_LAUNCH(true, false); }
        else            { if (splat) ROUTE_LAUNCH(false, true); else ROUTE_LAUNCH(false, false); }
#undef ROUTE_LAUNCH
        CK_LAUNCH(ctx);
    } else if (fenced && splat) { // no route kernel to carry the fence
        route_signal_kernel<<<1, 32, 0, ctx->stream>>>(rd);
        CK_LAUNCH(ctx);
    }
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_route(clumpy_advect* adv, void** counts_dev) {
    REQUIRE(adv && adv->route.on, "routing is not set up");
    REQUIRE(!adv->count_open, "advect_route called twice without advect_drain");
    CK(cudaSetDevice(adv->ctx->device));
    int rc = route_launch(adv, false, 1, 0u);
    if (rc) return rc;
    adv->count_open = true;
    if (counts_dev) *counts_dev = adv->route.counts;
    return CLUMPY_OK;
}

// The receiving half of the current frame: drain this rank's inbox into the frame's band buffer and
// composite the rows this rank owns.  fenced: the drain waits for every source's epoch flag on the
// device, and (for big clouds) drains run on one side stream and composites on another, ordered by
// the flags and by events only: the main stream carries nothing but route kernels, the three stages of
// consecutive frames overlap, and a rank may run several frames ahead of its slowest peer.
static int route_receive(clumpy_advect* adv, bool fenced, uint32_t epoch, cudaEvent_t* ev = nullptr) {
    clumpy_cuda_ctx* ctx = adv->ctx;
    auto& rt = adv->route;
    const int h = adv->geom.halo;
    if (frame_splats(adv)) {
        const int cur = rt.slot, ahead = (cur + ROUTE_AHEAD) % ROUTE_RING;
        const uint32_t* inbox = rt.inbox + (size_t)cur * rt.entries_per_slot;
        uint32_t* cnt = rt.inbox + rt.cnt_offset + (size_t)cur * rt.world * rt.nregion;
        const uint32_t* flags = fenced ? rt.inbox + rt.flags_offset : nullptr;
        // Band buffers: frame s counts into `cur`; its drain zeroes `ahead` for frame s + ROUTE_AHEAD
        // (last read by the composite of frame s - ROUTE_AHEAD).
        uint32_t* band = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(rt.band) + (size_t)cur * rt.band_bytes);
        uint4* other = reinterpret_cast<uint4*>(reinterpret_cast<char*>(rt.band) + (size_t)ahead * rt.band_bytes);
        const uint32_t other_n16 = (uint32_t)(rt.band_bytes / 16);
        const bool overlap = adv->overlap && fenced && !ev;
        const cudaStream_t main_stream = ctx->stream;
        const cudaStream_t side = overlap ? ctx->aux_stream : main_stream;   // drains
        const cudaStream_t side2 = overlap ? ctx->aux2_stream : main_stream; // composites
        if (overlap && rt.composite_pending[ahead]) {
            CK(cudaStreamWaitEvent(side, rt.composited[ahead], 0));
            rt.composite_pending[ahead] = false;
        }
        // A small grid: these CTAs may sit waiting for a peer's epoch while route kernels run on the
        // main stream, and must leave them most of the SMs (256 CTAs of 1184 resident slots).
        const dim3 grid(std::max<uint32_t>(1u, 256u / (uint32_t)rt.world), rt.world);
        if (adv->cell_mode == CELL_F16)
            drain_inbox_kernel<CELL_F16><<<grid, 256, 0, side>>>(inbox, cnt, rt.nregion, flags, epoch, rt.errors,
                                                                  band, rt.band_cells, other, other_n16);
        else
            drain_inbox_kernel<CELL_U32><<<grid, 256, 0, side>>>(inbox, cnt, rt.nregion, flags, epoch, rt.errors,
                                                                  band, rt.band_cells, other, other_n16);
        CK_LAUNCH(ctx);
        if (overlap) {
            CK(cudaEventRecord(rt.drained[cur], side));
            rt.drain_pending[cur] = true;
            adv->side_dirty = true;
        }
        if (ev) CK(cudaEventRecord(ev[2], main_stream));
        // image rows this rank owns: those whose window of padded rows lies in chunk + halos
        const int lo = rt.rank * rt.chunk;
        const int y0 = std::max(0, lo - h), y1 = route_last_row(adv);
        if (y1 > y0) {
            if (overlap) CK(cudaStreamWaitEvent(side2, rt.drained[cur], 0));
            const size_t cell_bytes = adv->cell_mode == CELL_F16 ? 2 : 4;
            const char* cells = reinterpret_cast<const char*>(band) + (size_t)(y0 - (lo - h)) * adv->geom.pitch * cell_bytes;
            ctx->stream = side2; // the launchers enqueue on ctx->stream
            int rc = clumpy_cuda_advect_composite_rows(adv, cells, (uint32_t)y0, (uint32_t)(y1 - y0));
            ctx->stream = main_stream;
            if (rc) return rc;
            if (overlap) {
                CK(cudaEventRecord(rt.composited[cur], side2));
                rt.composite_pending[cur] = true;
                rt.side2_dirty = true;
            }
        }
        rt.slot = (cur + 1) % ROUTE_RING;
    } else if (ev) {
        CK(cudaEventRecord(ev[2], ctx->stream));
    }
    if (ev) CK(cudaEventRecord(ev[3], ctx->stream));
    adv->simframe += 1;
    ++adv->frames_since_regroup;
    return CLUMPY_OK;
}

// The caller has fenced the frame itself (any collective on the context's stream that every rank
// enters after its clumpy_cuda_advect_route); `counts_all_dev` is ignored and may be NULL.
CLUMPY_API int clumpy_cuda_advect_drain(clumpy_advect* adv, const uint32_t* counts_all_dev) {
    (void)counts_all_dev;
    REQUIRE(adv && adv->route.on, "routing is not set up");
    REQUIRE(adv->count_open, "advect_drain without advect_route");
    CK(cudaSetDevice(adv->ctx->device));
    int rc = route_receive(adv, false, 0u);
    adv->count_open = false;
    if (rc) return rc;
    if (adv->regroup_every && adv->frames_since_regroup >= adv->regroup_every) {
        rc = regroup_points(adv);
        if (rc) return rc;
    }
    return advect_join(adv);
}

// `count` whole frames with the fence on the device.  Frames are launched in groups of up to
// route.fuse: ONE route kernel advances the points through the group on the main stream (its extra
// last CTA stores each frame's epoch into every peer's flags as soon as all CTAs have finished that
// frame); per frame, a drain (zeroes the band buffer of the frame ROUTE_AHEAD frames on, waits for
// every peer's epoch, counts the inbox) and a composite run on the two side streams.  No host-side
// collective, no synchronisation: the caller can enqueue thousands of frames in one call.
static int route_steps_impl(clumpy_advect* adv, uint32_t count, cudaEvent_t* ev) {
    REQUIRE(adv && adv->route.on, "routing is not set up");
    REQUIRE(!adv->count_open, "advect_route_steps inside an open frame");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    auto& rt = adv->route;
    for (int r = 0; r < rt.world; ++r) REQUIRE(rt.peer_inbox[r], "peer inboxes are not mapped (route_peers)");
    for (uint32_t i = 0; i < count;) {
        // a group never crosses a regroup or the warm-up / recording boundary (where, with decay 0,
        // frames start to splat) and is timed frame by frame when profiling
        uint32_t nf = ev ? 1u : std::min<uint32_t>((uint32_t)rt.fuse, count - i);
        nf = std::min<uint32_t>(nf, adv->nframes - adv->simframe % adv->nframes);
        if (adv->regroup_every) nf = std::min<uint32_t>(nf, std::max<uint32_t>(1u, adv->regroup_every - std::min(adv->frames_since_regroup, adv->regroup_every - 1u)));
        const uint32_t epoch0 = rt.epoch_base + adv->simframe + 1u;
        cudaEvent_t* e = ev ? ev + 4 * (size_t)i : nullptr;
        if (e) CK(cudaEventRecord(e[0], ctx->stream));
        int rc = route_launch(adv, true, (int)nf, epoch0);
        if (rc) return rc;
        if (e) CK(cudaEventRecord(e[1], ctx->stream));
        // The drains of the group start once its route kernel is done (by then this rank's own epochs
        // are out; a drain that started earlier would only sit on SM slots waiting for them).
        if (adv->overlap && !ev && frame_splats(adv)) {
            CK(cudaEventRecord(rt.routed, ctx->stream));
            CK(cudaStreamWaitEvent(ctx->aux_stream, rt.routed, 0));
        }
        for (uint32_t f = 0; f < nf; ++f) {
            rc = route_receive(adv, true, epoch0 + f, e);
            if (rc) return rc;
        }
        if (adv->regroup_every && adv->frames_since_regroup >= adv->regroup_every) {
            rc = regroup_points(adv);
            if (rc) return rc;
        }
        i += nf;
    }
    return advect_join(adv); // whatever the caller enqueues next sees the finished framebuffer
}

CLUMPY_API int clumpy_cuda_advect_route_steps(clumpy_advect* adv, uint32_t count) {
    return route_steps_impl(adv, count, nullptr);
}

// route_steps with CUDA events around every kernel: mean device time per frame (ms) of the route
// kernel, the drain kernel (which includes the wait for the slowest peer) and the band composite.
// Frames are launched one by one here.  Every rank must call it with the same count.  Synchronises.
CLUMPY_API int clumpy_cuda_advect_route_profile(clumpy_advect* adv, uint32_t count, float* route_ms,
                                                float* drain_ms, float* composite_ms) {
    REQUIRE(adv && count > 0, "bad argument");
    CK(cudaSetDevice(adv->ctx->device));
    std::vector<cudaEvent_t> ev(4 * (size_t)count, nullptr);
    int rc = CLUMPY_OK;
    for (auto& e : ev)
        if (cudaEventCreate(&e) != cudaSuccess) { set_error("event creation failed"); rc = CLUMPY_ERR_CUDA; break; }
    if (rc == CLUMPY_OK) rc = route_steps_impl(adv, count, ev.data());
    double t[3] = {0, 0, 0};
    if (rc == CLUMPY_OK && cudaStreamSynchronize(adv->ctx->stream) != cudaSuccess) { set_error("profile run failed"); rc = CLUMPY_ERR_CUDA; }
    for (uint32_t i = 0; rc == CLUMPY_OK && i < count; ++i)
        for (int k = 0; k < 3; ++k) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[4 * (size_t)i + k], ev[4 * (size_t)i + k + 1]);
            t[k] += ms;
        }
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (route_ms) *route_ms = (float)(t[0] / count);
    if (drain_ms) *drain_ms = (float)(t[1] / count);
    if (composite_ms) *composite_ms = (float)(t[2] / count);
    return rc;
}

// Number of fence time-outs so far (a peer that never published its epoch).  Synchronises.
CLUMPY_API int clumpy_cuda_advect_route_errors(clumpy_advect* adv, uint32_t* errors) {
    REQUIRE(adv && errors && adv->route.on, "routing is not set up");
    CK(cudaSetDevice(adv->ctx->device));
    CK(cudaMemcpyAsync(errors, adv->route.errors, 4, cudaMemcpyDeviceToHost, adv->ctx->stream));
    CK(cudaStreamSynchronize(adv->ctx->stream));
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

CLUMPY_API int clumpy_cuda_advect_image_dev(clumpy_advect* adv, void** img_dev) {
    REQUIRE(adv && img_dev, "null argument");
    *img_dev = adv->img[adv->cur];
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_advect_finish_frame(clumpy_advect* adv) {
    REQUIRE(adv, "adv is NULL");
    REQUIRE(adv->count_open, "advect_finish_frame without advect_count");
    clumpy_cuda_ctx* ctx = adv->ctx;
    CK(cudaSetDevice(ctx->device));
    if (frame_splats(adv)) CK(cudaMemsetAsync(adv->hist_buf[adv->cellbuf], 0, adv->hist_bytes, ctx->stream));
    adv->count_open = false;
    adv->simframe += 1;
    if (adv->regroup_every && ++adv->frames_since_regroup >= adv->regroup_every) return regroup_points(adv);
    return CLUMPY_OK;
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
