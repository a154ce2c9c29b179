// advect_kernels.cuh -- device side of clumpy `advect_points` for sm_100a (host side: advect.cu).
//
// Replaces the per-point loop of commands/advect_points.cc:144-153,163-171 and the hit loop of
// splat_points.cc:142-160.  Facts the kernels rely on:
//   * velocity lookup is NEAREST texel (advect_points.cc:27-36), index conversion via int64;
//   * pt += step * v is a rounded multiply then a rounded add (the reference binary has no FMA);
//   * the age / age_offset bookkeeping (:124-135,147-152,166-170) is equivalent to
//         "point i returns to its original position at the end of simulation frame s
//          iff s % nframes == age_offset_i",
//     so no per-point age is stored;
//   * trajectories never depend on the framebuffer.
#pragma once

#include "device_utils.cuh"
#include "internal.h"

#include <cuda_fp16.h>

namespace clumpy {

constexpr int ADVECT_THREADS = 256;
constexpr int ADVECT_PTS = 4;      // points per thread (independent load chains in flight)
constexpr int ADVECT_MAX_FUSE = 8; // simulation frames one launch can advance (= GC_MAX_FRAMES)

// How a point is counted into the splat accumulator.
enum CellMode : int {
    CELL_NONE = -1, // frame without a splat (warm-up with decay == 0)
    CELL_U32 = 0,   // 32-bit counters, one RED.ADD per point (any sprite)
    CELL_F16 = 2,   // half-float counters, two per word, one RED.ADD.F16x2 per point: a half counts
                    // exactly to 2048 and then stays put, so it saturates by itself and cannot carry
    CELL_X64 = 3,   // 64-bit cells = count (24 bits) + sum of original point indices (40 bits), one
                    // 64-bit RED.ADD per point: keeps the reference's hit order (composite_exact.cu)
};

__device__ __forceinline__ void cell_add_f16(uint32_t* __restrict__ words, long long cell) {
    // (1, 0) or (0, 1) as a half2 bit pattern: 0x3C00 is 1.0
    const uint32_t one = (cell & 1) ? 0x3C000000u : 0x00003C00u;
    asm volatile("red.global.add.noftz.f16x2 [%0], %1;" :: "l"(words + (cell >> 1)), "r"(one) : "memory");
}
// (Warp-aggregated counts -- the lanes of a warp that hit the same cell sending ONE reduction, found with
// match.any -- were measured in round 2 and lose: 0.0273 vs 0.0165 ms per frame for the route kernel at 2.1 M points,
// profiles/r2_sweep_multi_2gpu_quarter.txt.  At 0.5 points per texel almost every lane has a cell of its own.)
// Points OUTSIDE the image whose sprite still reaches it: a point that has left the image gets no velocity any more
// (advect_points.cc:27-32) and sits still, just outside, until its reset -- at C4 some 60 of them per halo cell in
// the steady state, every one reducing into the same cell every frame (one cell takes 0.65 G reductions per second).
// The regroup stores such points in the order of their ring cell (home_tile), so the lanes of a warp share a few
// cells: they find each other with match.any and send ONE reduction per cell.  Only lanes outside the image come
// here; for the others (almost every lane its own cell) the match costs more than it saves (measured).
__device__ __noinline__ void cell_add_f16_outside(uint32_t* __restrict__ words, long long cell) {
    const uint32_t same = __match_any_sync(__activemask(), cell);
    if ((uint32_t)(__ffs(same) - 1) != (threadIdx.x & 31u)) return;
    const uint32_t h = (uint32_t)__half_as_ushort(__uint2half_rn((uint32_t)__popc(same))); // <= 32: exact
    asm volatile("red.global.add.noftz.f16x2 [%0], %1;" :: "l"(words + (cell >> 1)), "r"((cell & 1) ? (h << 16) : h) : "memory");
}
__device__ __forceinline__ void cell_add_f16(uint32_t* __restrict__ words, long long cell, uint64_t pol) {
    const uint32_t one = (cell & 1) ? 0x3C000000u : 0x00003C00u;
    asm volatile("red.global.add.L2::cache_hint.noftz.f16x2 [%0], %1, %2;" :: "l"(words + (cell >> 1)), "r"(one), "l"(pol) : "memory");
}

// Counts a point whose truncated position is (cx, cy) (splat_points.cc:143-144) into `hist`.
// Half cells in x-wrap mode: the group composite (composite_group.cu) reads its halo columns as they are, so a
// point within `halo` columns of the left / right edge is also counted into the halo column on the other
// side -- the cell its wrapped taps would be read from (splat_points.cc:149-151).
template <bool WRAP, int MODE>
__device__ __forceinline__ void count_point(uint32_t* __restrict__ hist, int32_t cx, int32_t cy, const GeomDev& g,
                                            uint32_t slot, uint64_t red_pol) {
    const long long at = hist_index<WRAP>(cx, cy, g);
    if (at < 0) return;
    if (MODE == CELL_X64) {
        // exact-order path: the points are in the caller's order, so the slot IS the original index
        atomicAdd(reinterpret_cast<unsigned long long*>(hist) + at, ((unsigned long long)slot << 24) | 1ull);
    } else if (MODE == CELL_U32) {
        atomicAdd(hist + at, 1u);
    } else if (MODE == CELL_F16) {
        if (!WRAP && ((uint32_t)cx >= (uint32_t)g.width || (uint32_t)cy >= (uint32_t)g.height)) { cell_add_f16_outside(hist, at); return; }
        cell_add_f16(hist, at, red_pol);
        if (WRAP && g.halo > 0) {
            int32_t m = cx % g.width;
            if (m < 0) m += g.width;
            if (m >= g.width - g.halo) cell_add_f16(hist, at - g.width);
            if (m < g.halo) cell_add_f16(hist, at + g.width);
        }
    }
}

// Up to ADVECT_MAX_FUSE simulation frames per launch ("temporal blocking"): the positions are loaded once,
// advanced through `nf` frames in registers -- gather, Euler step, reset, count, each frame counting into
// its own counter image -- and stored once.  Position and phase traffic (20 of the 28 algorithmic bytes per
// point-step) is paid once per group instead of once per frame; the gathers and the counter updates are what
// is left per frame.
//
// A CTA owns ADVECT_THREADS * ADVECT_PTS consecutive points; thread t takes t, t + 256, t + 512, ... so every
// warp-wide access covers 32 CONSECUTIVE points: position / offset traffic is perfectly coalesced and, the
// points being stored in tile order, a warp's gather touches only a few lines.
struct HistRing { // frame f of the group counts into image (first + f) % count
    uint32_t* base;
    size_t stride_words;
    int first, count;
    __host__ __device__ uint32_t* image(int f) const { return base + (size_t)((first + f) % count) * stride_words; }
};

template <bool WRAP, int MODE, typename OffT>
__global__ void __launch_bounds__(ADVECT_THREADS, 6)
advect_fused_kernel(float2* __restrict__ pos, const float2* __restrict__ orig,
                    const OffT* __restrict__ offset, uint32_t npts, const float2* __restrict__ vel,
                    uint32_t width, uint32_t height, float step, uint32_t phase0, uint32_t nframes, int nf,
                    GeomDev g, HistRing ring) {
    const uint32_t base = blockIdx.x * (ADVECT_THREADS * ADVECT_PTS) + threadIdx.x;
    const uint64_t pol = g.keep_pos == 0 ? make_stream_policy() : make_normal_policy();
    const uint64_t vel_pol = (g.hints & 1) ? make_keep_policy() : make_normal_policy();
    const uint64_t red_pol = (g.hints & 2) ? make_stream_policy() : make_normal_policy();
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
    uint32_t phase = phase0;
    for (int f = 0; f < nf; ++f) {
        float2 v[ADVECT_PTS];
#pragma unroll
        for (int k = 0; k < ADVECT_PTS; ++k) { // nearest-texel fetch (advect_points.cc:23-37)
            uint32_t x = f2u32_x86(p[k].x);
            const uint32_t y = f2u32_x86(p[k].y);
            if (WRAP) x = (width + (x % width)) % width;
            v[k] = make_float2(0.f, 0.f);
            if (live[k] && x < width && y < height) v[k] = ld_gather_f2(vel + ((size_t)y * width + x), vel_pol);
        }
#pragma unroll
        for (int k = 0; k < ADVECT_PTS; ++k) {
            // pt += step * v: rounded multiply, then rounded add (:146); reset AFTER the move (:147-152,166-170)
            p[k] = make_float2(__fadd_rn(p[k].x, __fmul_rn(step, v[k].x)), __fadd_rn(p[k].y, __fmul_rn(step, v[k].y)));
            if (off[k] == phase) p[k] = orig[base + k * ADVECT_THREADS];
        }
        phase = phase + 1u == nframes ? 0u : phase + 1u;
        if (MODE != CELL_NONE) {
            uint32_t* hist = ring.image(f);
#pragma unroll
            for (int k = 0; k < ADVECT_PTS; ++k) // splat_points.cc:143-144: centre texel by (int32_t) truncation
                if (live[k]) count_point<WRAP, MODE>(hist, f2i32_x86(p[k].x), f2i32_x86(p[k].y), g, base + k * ADVECT_THREADS, red_pol);
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
// (c + 1) * ROUTE_BLOCK) of segment s in every destination's inbox for every frame of the group (a CTA
// advects ROUTE_BLOCK points and a point sends at most one entry per destination and frame), ranks its
// entries with a shared-memory atomic, and at the END of the launch leaves the numbers it wrote in the
// destination's count words [frame][s][c].
//
// One launch advances the points by a GROUP of up to ROUTE_MAX_FUSE frames; positions stay in registers.
// Inside the frame loop there is no barrier and no fence.  The fence is per GROUP: a CTA that is done
// fences its remote stores at system scope and takes a ticket; the CTA that takes the last ticket stores
// the group's EPOCH into a flag per source in every peer's memory.  The receiver (drain_group_kernel) waits
// for every source's epoch, counts the entries into its band buffers and zeroes the count words it used.
constexpr int ROUTE_MAX_WORLD = 8;
constexpr int ROUTE_PTS = 4;                                  // points per thread of the route kernel
constexpr uint32_t ROUTE_BLOCK = ADVECT_THREADS * ROUTE_PTS;  // points per CTA
constexpr uint32_t ROUTE_REGION = ROUTE_BLOCK;                // entry slots per (CTA, destination, frame)
constexpr int ROUTE_MAX_FUSE = ADVECT_MAX_FUSE;               // frames per launch
constexpr int ROUTE_PARITIES = 4; // inbox slot sets in flight (see route_front in advect.cu for why 4 is enough)
constexpr int ROUTE_SETS = 3;     // band-buffer sets: counted into | composited | being zeroed

struct RouteDev {
    int me, world;
    int chunk, halo, pitch;             // band geometry in padded rows / cells
    uint32_t nregion;                   // regions (= CTAs of the largest shard) per source segment
    uint32_t nframes, phase0;           // reset period; phase of the first frame of the group
    int nf;                             // frames in this launch
    size_t inbox_off;                   // uint32 offset of this group's parity inside every inbox
    size_t inbox_frame_stride;          // ... between consecutive frames of the group
    size_t cnt_off, cnt_frame_stride;   // the same for the count words
    uint32_t* inbox[ROUTE_MAX_WORLD];   // rank r's entries (peer-mapped): [parity][frame][source][region][ROUTE_REGION]
    uint32_t* cnt[ROUTE_MAX_WORLD];     // rank r's count words (peer-mapped): [parity][frame][source][region]
    uint32_t* flags[ROUTE_MAX_WORLD];   // rank r's epoch flags (peer-mapped): [source]
    uint32_t* band[ROUTE_MAX_FUSE];     // this rank's band buffer of frame f
    int band_f16;                       // band cells are half floats (else uint32)
    int loopback;                       // profiling aid: every "peer" is this rank itself (see route_begin)
    uint32_t epoch;                     // epoch of the group
    uint32_t* ticket;                   // CTAs of this launch that are done
    // the common case of a splat (see advect_route_kernel): image rows [in_row0, in_row0 + interior) of this rank's
    // band are out of the sprite's reach of the band edges; cell index there = cy * pitch + cx + in_bias
    int32_t in_row0, in_bias;
    uint32_t interior;
};

// One count into a band buffer.  Half cells in x-wrap mode (wrap_w = image width, else 0): the group
// composite reads its halo columns as they are, so a cell within `halo` columns of the left / right edge is
// also counted into the halo column on the other side (see count_point).
__device__ __forceinline__ void band_add(uint32_t* __restrict__ band, uint32_t entry, int f16, int wrap_w, int halo, int pitch, bool outside = false) {
    if (!f16) { atomicAdd(band + entry, 1u); return; }
    if (outside) { cell_add_f16_outside(band, (long long)entry); return; } // (never in x-wrap mode: see count_point)
    cell_add_f16(band, (long long)entry);
    if (wrap_w && halo > 0) {
        const int m = (int)(entry % (uint32_t)pitch) - halo;
        if (m >= wrap_w - halo) cell_add_f16(band, (long long)entry - wrap_w);
        if (m < halo) cell_add_f16(band, (long long)entry + wrap_w);
    }
}

// One entry for rank `dst`: the CTA-wide rank comes from a shared-memory atomic, the slot is fixed by
// the CTA's region, the store goes over NVLink.  (Collecting a CTA's entries in shared memory and writing them out
// as whole lines at the end of the launch was measured in round 2 and is SLOWER: 0.0198 vs 0.0163 ms per frame at
// 2.1 M points per rank; the 4-byte stores are not what the kernel waits for.)
__device__ __forceinline__ void route_send(const RouteDev& rt, int f, uint32_t* cta_n, int dst, uint32_t entry) {
    const uint32_t r = atomicAdd(cta_n + dst, 1u);
    rt.inbox[dst][rt.inbox_off + (size_t)f * rt.inbox_frame_stride + ((size_t)rt.me * rt.nregion + blockIdx.x) * ROUTE_REGION + r] = entry;
}

// A cell of padded row `r` (relative to band `owner`'s first row) and padded column `ux`: counted
// straight into this rank's band buffer when it is ours, sent to the owner's inbox otherwise; rows
// within the sprite's reach of a band edge also go to the neighbouring band, whose composite needs them.
__device__ __forceinline__ void route_cell(const RouteDev& rt, int f, uint32_t* cta_n, int owner, int r, uint32_t ux, int wrap_w, bool outside) {
    const uint32_t entry = (uint32_t)((r + rt.halo) * rt.pitch) + ux; // band rows: [halo above | chunk | halo below]
    if (owner == rt.me) band_add(rt.band[f], entry, rt.band_f16, wrap_w, rt.halo, rt.pitch, outside);
    else route_send(rt, f, cta_n, owner, entry);
    if (r < rt.halo && owner > 0) {
        const uint32_t e = (uint32_t)((r + rt.chunk + rt.halo) * rt.pitch) + ux;
        if (owner - 1 == rt.me) band_add(rt.band[f], e, rt.band_f16, wrap_w, rt.halo, rt.pitch);
        else route_send(rt, f, cta_n, owner - 1, e);
    }
    if (r >= rt.chunk - rt.halo && owner < rt.world - 1) {
        const uint32_t e = (uint32_t)((r - rt.chunk + rt.halo) * rt.pitch) + ux;
        if (owner + 1 == rt.me) band_add(rt.band[f], e, rt.band_f16, wrap_w, rt.halo, rt.pitch);
        else route_send(rt, f, cta_n, owner + 1, e);
    }
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// The receiving half of a group: once every source has published the group's epoch, the entries of every frame
// of the group are counted into that frame's band buffer, one warp per (frame, source, region).  Count words are
// zeroed as they are used: senders only ever write non-zero ones.
struct DrainDev {
    const uint32_t* inbox;      // this group's parity: [frame][source][region][ROUTE_REGION]
    uint32_t* cnt;              // [frame][source][region]
    size_t inbox_frame_stride, cnt_frame_stride;
    uint32_t nregion;
    int world, nf;
    const uint32_t* flags;      // [source]; NULL: do not wait
    uint32_t epoch;
    uint32_t* errors;
    unsigned long long timeout_ns;
    uint32_t* band[ROUTE_MAX_FUSE];
    uint32_t band_cells;
    int wrap_w, halo, pitch;    // x-wrap mode: image width (else 0); band geometry
};

// Warp `warp` of `nwarp`'s share of the group.  The (frame, region) pairs are dealt to the warps ROUND-ROBIN (pair
// p goes to warp p % nwarp): the non-empty regions are the CTAs near a band edge, which are neighbours in region
// order, so a contiguous deal would leave all the work to a few warps.  A warp loads the count words of up to
// 32 of its pairs in one go (most are zero: far CTAs send nothing), then works through the non-empty ones.
template <int CELLS>
__device__ __forceinline__ void drain_share(const DrainDev& d, uint32_t warp, uint32_t nwarp, uint32_t lane) {
    const uint32_t per_frame = (uint32_t)d.world * d.nregion; // the count words of a frame are contiguous
    const uint32_t total = per_frame * (uint32_t)d.nf;
    for (uint32_t k0 = 0; warp + k0 * nwarp < total; k0 += 32u) {
        const uint32_t pair = warp + (k0 + lane) * nwarp;
        const uint32_t pf = pair < total ? pair / per_frame : 0u, pc = pair - pf * per_frame;
        uint32_t* word = d.cnt + (size_t)pf * d.cnt_frame_stride + pc;
        const uint32_t mine = (pair < total) ? min(__ldcg(word), ROUTE_REGION) : 0u; // written by a peer: read at L2
        if (mine) *word = 0u;
        uint32_t todo = __ballot_sync(0xFFFFFFFFu, mine != 0u);
        while (todo) {
            const int r = __ffs(todo) - 1;
            todo &= todo - 1u;
            const uint32_t n = __shfl_sync(0xFFFFFFFFu, mine, r);
            const uint32_t f = __shfl_sync(0xFFFFFFFFu, pf, r), c = __shfl_sync(0xFFFFFFFFu, pc, r);
            const uint32_t* seg = d.inbox + (size_t)f * d.inbox_frame_stride + (size_t)c * ROUTE_REGION;
            uint32_t* band = d.band[f];
            for (uint32_t i0 = 0; i0 < n; i0 += 256u) { // 8 loads in flight per lane: the loop is latency-bound
                uint32_t e[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t i = i0 + (uint32_t)j * 32u + lane;
                    e[j] = i < n ? __ldcg(seg + i) : 0xFFFFFFFFu;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (e[j] >= d.band_cells) continue;
                    band_add(band, e[j], CELLS == CELL_F16, d.wrap_w, d.halo, d.pitch);
                }
            }
        }
    }
}

// The launch: waits for the epochs first (bounded: on a time-out it records the error and does nothing else -- the
// host turns a non-zero error count into a failed call), then counts.
// DRAIN_THREADS x 32 registers = 4096: what an SM has left beside six CTAs of the route kernel (6 x 256 x 40), so the
// drain of one group starts next to the route launch of the next instead of waiting for its CTAs to retire.
constexpr int DRAIN_THREADS = 128;

template <int CELLS>
__global__ void __launch_bounds__(DRAIN_THREADS)
drain_group_kernel(DrainDev d) {
    if (d.flags) {
        __shared__ int arrived;
        if (threadIdx.x == 0) {
            const unsigned long long t0 = globaltimer_ns();
            int ok = 1;
            for (int s = 0; s < d.world && ok; ++s) {
                // epochs only grow; compare as signed distance so that wrap-around stays harmless.  The
                // acquire load orders the reads of the entries (by the whole CTA, after the barrier) behind it.
                while ((int32_t)(ld_acquire_sys(d.flags + s) - d.epoch) < 0) {
                    __nanosleep(100);
                    if (globaltimer_ns() - t0 > d.timeout_ns) { ok = 0; break; }
                }
            }
            if (!ok && blockIdx.x == 0) atomicAdd(d.errors, 1u);
            arrived = ok;
        }
        __syncthreads();
        if (!arrived) return;
    }
    drain_share<CELLS>(d, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), gridDim.x * (blockDim.x >> 5), threadIdx.x & 31u);
}

// The group fence, sender side.  Called by every CTA after a barrier that follows its last remote store:
// thread 0's fence.sys is cumulative over the stores of the whole CTA (they happen-before it through the
// barrier); the ticket orders the CTAs; the last one fences again and publishes.
__device__ __forceinline__ void route_group_fence(const RouteDev& rt, bool cta_sent) {
    if (threadIdx.x != 0) return;
    // a CTA whose points all landed in this rank's own band has nothing to publish: its reductions are consumed
    // by this rank's own later launches (stream order); the fence waits for acknowledgements from the peers and
    // is worth skipping
    if (cta_sent) __threadfence_system();
    const uint32_t done = atomicAdd(rt.ticket, 1u);
    if (done != gridDim.x - 1u) return;
    *rt.ticket = 0u; // for the next launch (launches of one rank are stream-ordered)
    __threadfence_system();
    for (int d = 0; d < rt.world; ++d) {
        const int slot = rt.loopback ? d : rt.me; // loopback: play every source of my own inbox
        *reinterpret_cast<volatile uint32_t*>(rt.flags[d] + slot) = rt.epoch;
    }
}

template <bool WRAP, bool SPLAT, bool F16, typename OffT>
__global__ void __launch_bounds__(ADVECT_THREADS, 6)
advect_route_kernel(float2* __restrict__ pos, const float2* __restrict__ orig,
                    const OffT* __restrict__ offset, uint32_t npts, const float2* __restrict__ vel,
                    uint32_t width, uint32_t height, float step, GeomDev g, RouteDev rt) {
    __shared__ uint32_t cta_n[ROUTE_MAX_FUSE][ROUTE_MAX_WORLD];
    if (SPLAT) {
        if (threadIdx.x < ROUTE_MAX_FUSE * ROUTE_MAX_WORLD) (&cta_n[0][0])[threadIdx.x] = 0u;
        __syncthreads(); // here, where every warp arrives at once, not after the loads
    }
    const uint64_t pol = g.keep_pos == 0 ? make_stream_policy() : make_normal_policy();
    const uint32_t base = blockIdx.x * ROUTE_BLOCK + threadIdx.x;
    const int my_lo = rt.me * rt.chunk; // first padded row of this rank's band
    // positions and reset phases: loaded once for the whole group of frames.  A slot past the end of the shard gets
    // a position that fails every test below by itself (no velocity texel, no cell, never reset), so the frame loop
    // does not ask `live` again.
    float2 p[ROUTE_PTS];
    uint32_t off[ROUTE_PTS];
    bool live[ROUTE_PTS];
#pragma unroll
    for (int k = 0; k < ROUTE_PTS; ++k) {
        const uint32_t i = base + k * ADVECT_THREADS;
        live[k] = i < npts;
        p[k] = make_float2(-1.0e8f, -1.0e8f);
        off[k] = 0xFFFFFFFFu;
        if (live[k]) {
            p[k] = ld_stream_f2(pos + i, pol);
            off[k] = ld_stream_off(offset + i, pol);
        }
    }
    // Each position is truncated once: the integers serve the splat of the frame that produced the position
    // (splat_points.cc:143-144) and the velocity fetch of the next frame (advect_points.cc:23-37).  The two casts
    // of the reference differ from the plain truncation only for |coordinate| >= 2^31 and NaN; `odd` says whether
    // ANY coordinate of this thread is such a value (a sum of magnitudes: NaN and overflow propagate, and a sum of
    // non-negative terms is never below its largest one), and only then are fetch_index / splat_index asked.
    int32_t tx[ROUTE_PTS], ty[ROUTE_PTS];
    bool odd;
    auto truncate = [&]() {
        float mag = 0.f;
#pragma unroll
        for (int k = 0; k < ROUTE_PTS; ++k) {
            tx[k] = __float2int_rz(p[k].x);
            ty[k] = __float2int_rz(p[k].y);
            mag = __fadd_rn(mag, __fadd_rn(fabsf(p[k].x), fabsf(p[k].y)));
        }
        odd = !(mag < 2147483648.0f);
    };
    truncate();
    // The common case of a splat -- the cell lies in this rank's own band, out of the sprite's reach of the band
    // edges, inside the image -- is two compares, one multiply-add and the reduction: with c = (cx, cy) the centre
    // texel, the test is in_row0 <= cy < in_row0 + interior and 0 <= cx < width, and the cell index in the band
    // buffer ((r + halo) * pitch + ux in route_cell's terms) is cy * pitch + cx + in_bias (set by the host).
    const uint32_t padded_w = (uint32_t)(g.width + 2 * g.halo), padded_h = (uint32_t)(g.height + 2 * g.halo);
    const char* const vel_bytes = reinterpret_cast<const char*>(vel);
    const uint32_t vel_pitch = width * (uint32_t)sizeof(float2); // (a row of the field is below 4 GB)
    uint32_t phase = rt.phase0;
    for (int f = 0; f < rt.nf; ++f) {
        uint32_t gx[ROUTE_PTS], gy[ROUTE_PTS];
#pragma unroll
        for (int k = 0; k < ROUTE_PTS; ++k) { gx[k] = (uint32_t)tx[k]; gy[k] = (uint32_t)ty[k]; }
        if (odd) {
#pragma unroll
            for (int k = 0; k < ROUTE_PTS; ++k) { gx[k] = fetch_index(tx[k], p[k].x); gy[k] = fetch_index(ty[k], p[k].y); }
        }
        float2 v[ROUTE_PTS];
#pragma unroll
        for (int k = 0; k < ROUTE_PTS; ++k) { // nearest-texel fetch (advect_points.cc:23-37)
            uint32_t x = gx[k];
            if (WRAP) x = (width + (x % width)) % width;
            v[k] = make_float2(0.f, 0.f);
            if (x < width && gy[k] < height)
                v[k] = __ldg(reinterpret_cast<const float2*>(vel_bytes + (uint64_t)gy[k] * vel_pitch + (uint64_t)x * sizeof(float2)));
        }
#pragma unroll
        for (int k = 0; k < ROUTE_PTS; ++k) {
            // pt += step * v: rounded multiply, then rounded add (:146)
            p[k] = make_float2(__fadd_rn(p[k].x, __fmul_rn(step, v[k].x)), __fadd_rn(p[k].y, __fmul_rn(step, v[k].y)));
        }
        bool reset = false; // one point in nframes per frame: a single branch for the thread's points
#pragma unroll
        for (int k = 0; k < ROUTE_PTS; ++k) reset |= off[k] == phase;
        if (reset) {
#pragma unroll
            for (int k = 0; k < ROUTE_PTS; ++k)
                if (off[k] == phase) p[k] = orig[base + k * ADVECT_THREADS]; // reset AFTER the move (:147-152,166-170)
        }
        truncate();
        phase = phase + 1u == rt.nframes ? 0u : phase + 1u;
        uint32_t* const band = rt.band[f];
#pragma unroll
        for (int k = 0; SPLAT && k < ROUTE_PTS; ++k) {
            if (!WRAP && !odd && (uint32_t)(ty[k] - rt.in_row0) < rt.interior && (uint32_t)tx[k] < (uint32_t)g.width) {
                const uint32_t entry = (uint32_t)(ty[k] * rt.pitch + tx[k] + rt.in_bias);
                if (F16) cell_add_f16(band, (long long)entry);
                else atomicAdd(band + entry, 1u);
                continue;
            }
            // splat_points.cc:143-144: centre texel by (int32_t) truncation; same tests as hist_index
            const int32_t cx = splat_index(tx[k], p[k].x), cy = splat_index(ty[k], p[k].y);
            const uint32_t uy = (uint32_t)cy + (uint32_t)g.halo;
            if (uy >= padded_h) continue;
            uint32_t ux;
            if (WRAP) {
                int32_t m = cx % g.width;
                if (m < 0) m += g.width;
                ux = (uint32_t)(m + g.halo);
            } else {
                ux = (uint32_t)cx + (uint32_t)g.halo;
                if (ux >= padded_w) continue;
            }
            int r = (int)uy - my_lo;
            int owner = rt.me;
            if ((uint32_t)r >= (uint32_t)rt.chunk) {
                owner = min((int)(uy / (uint32_t)rt.chunk), rt.world - 1);
                r = (int)uy - owner * rt.chunk;
            }
            const bool outside = !WRAP && ((uint32_t)cx >= (uint32_t)g.width || (uint32_t)cy >= (uint32_t)g.height);
            route_cell(rt, f, cta_n[f], owner, r, ux, WRAP ? g.width : 0, outside);
        }
    }
#pragma unroll
    for (int k = 0; k < ROUTE_PTS; ++k)
        if (live[k]) st_stream_f2(pos + base + k * ADVECT_THREADS, p[k], pol);
    if (!SPLAT) return;
    __syncthreads(); // every entry of this CTA is written, cta_n is final
    // the non-zero counts go to their destinations (the receivers zero them again)
    uint32_t n = 0u;
    if (threadIdx.x < (uint32_t)(rt.nf * ROUTE_MAX_WORLD)) {
        const int f = threadIdx.x / ROUTE_MAX_WORLD, d = threadIdx.x % ROUTE_MAX_WORLD;
        n = d < rt.world ? cta_n[f][d] : 0u;
        if (n) *reinterpret_cast<volatile uint32_t*>(rt.cnt[d] + rt.cnt_off + (size_t)f * rt.cnt_frame_stride + (size_t)rt.me * rt.nregion + blockIdx.x) = n;
    }
    const bool cta_sent = __syncthreads_or(n != 0u) != 0;
    route_group_fence(rt, cta_sent);
}

// The same fence for a rank that has no points this run (no route kernel to carry it).
__global__ void route_signal_kernel(RouteDev rt) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        __threadfence_system();
        for (int d = 0; d < rt.world; ++d) *reinterpret_cast<volatile uint32_t*>(rt.flags[d] + (rt.loopback ? d : rt.me)) = rt.epoch;
    }
}

// ---- bucket sort by tile ----------------------------------------------------------------------------

struct SortGrid {
    int shift_x, shift_y; // tile = 2^shift_x x 2^shift_y texels
    uint32_t tiles_x, ntiles; // ntiles: the tiles of the image, then the cells of the ring around it
    uint32_t width, height;
    uint32_t nbins;           // tiles + ring cells
};

// The sort key of a point: the tile it is in -- or, for a point OUTSIDE the image, the cell of the one-texel ring
// around the image nearest to it (top row, bottom row, left column, right column), so that the points that sit
// still out there end up ordered by the cell they reduce into (cell_add_f16_outside).
__device__ __forceinline__ uint32_t home_tile(float2 p, const SortGrid& sg) {
    const int32_t x = f2i32_x86(p.x), y = f2i32_x86(p.y);
    const int32_t w = (int32_t)sg.width, h = (int32_t)sg.height;
    if ((uint32_t)x < sg.width && (uint32_t)y < sg.height)
        return (uint32_t)(y >> sg.shift_y) * sg.tiles_x + (uint32_t)(x >> sg.shift_x);
    const int32_t cx = max(-1, min(x, w)), cy = max(-1, min(y, h));
    uint32_t ring;
    if (cy < 0) ring = (uint32_t)(cx + 1);                              // above: columns -1 .. w
    else if (cy >= h) ring = (uint32_t)(w + 2) + (uint32_t)(cx + 1);    // below
    else if (cx < 0) ring = 2u * (uint32_t)(w + 2) + (uint32_t)cy;      // left: rows 0 .. h-1
    else ring = 2u * (uint32_t)(w + 2) + (uint32_t)h + (uint32_t)cy;    // right
    return sg.ntiles + ring;
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

// A reset offset as stored: offsets >= nframes never match a frame's phase (the reference's own "already
// reset" sentinel is age_offset == nframes, advect_points.cc:150) and are stored as all ones -- 16-bit
// storage is only used when nframes < 65535, so the sentinel is never a valid phase.
template <typename OffT>
__device__ __forceinline__ OffT stored_offset(uint32_t off, uint32_t nframes) {
    return off < nframes ? (OffT)off : (OffT)~(OffT)0;
}

// The per-point arrays in the caller's order.
template <typename OffT>
__global__ void __launch_bounds__(256)
identity_order_kernel(const float2* __restrict__ pts, const uint32_t* __restrict__ offs, uint32_t npts, uint32_t nframes,
                      float2* __restrict__ pos, float2* __restrict__ orig, OffT* __restrict__ offset,
                      uint32_t* __restrict__ perm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const float2 p = pts[i];
    pos[i] = p; orig[i] = p; offset[i] = stored_offset<OffT>(offs[i], nframes); perm[i] = i;
}

__global__ void __launch_bounds__(256)
unpermute_kernel(const float2* __restrict__ pos, const uint32_t* __restrict__ perm, uint32_t npts,
                 float2* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npts) out[perm[i]] = pos[i];
}

// ---- multi-GPU shards chosen on the device ------------------------------------------------------------
// Rank r simulates the points whose HOME row lies in [row_lo, row_hi).  Same truncation as the splat
// ((int32_t) cast, splat_points.cc:144); anything outside the image goes to the nearest row.
__device__ __forceinline__ uint32_t home_row(float2 p, uint32_t height) {
    const float y = p.y;
    if (!(y == y)) return 0u;                         // NaN
    if (y >= (float)height) return height - 1u;
    if (y <= 0.0f) return 0u;
    return min((uint32_t)y, height - 1u);
}

__global__ void __launch_bounds__(256)
row_histogram_kernel(const float2* __restrict__ pts, uint32_t npts, uint32_t height, uint32_t* __restrict__ rows) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += gridDim.x * blockDim.x)
        atomicAdd(rows + home_row(pts[i], height), 1u);
}

// Unordered compaction (the order of a shard does not matter: counts commute, trajectories are per point and
// the original index travels with the point): warp-aggregated slot claims.
template <typename OffT>
__global__ void __launch_bounds__(256)
select_rows_kernel(const float2* __restrict__ pts, const uint32_t* __restrict__ offs, uint32_t npts, uint32_t nframes, uint32_t height,
                   uint32_t row_lo, uint32_t row_hi, uint32_t* __restrict__ cursor, uint32_t capacity,
                   float2* __restrict__ pos, float2* __restrict__ orig, OffT* __restrict__ offset,
                   uint32_t* __restrict__ perm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float2 p = make_float2(0.f, 0.f);
    bool take = false;
    if (i < npts) {
        p = pts[i];
        const uint32_t r = home_row(p, height);
        take = r >= row_lo && r < row_hi;
    }
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, take);
    if (!mask) return;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0u;
    if (lane == (uint32_t)(__ffs(mask) - 1)) base = atomicAdd(cursor, (uint32_t)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(mask) - 1);
    if (!take) return;
    const uint32_t slot = base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
    if (slot >= capacity) return;
    pos[slot] = p; orig[slot] = p; offset[slot] = stored_offset<OffT>(offs[i], nframes); perm[slot] = i;
}

} // namespace clumpy
