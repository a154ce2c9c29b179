// composite_group.cu -- the composite pass of advect_points for half-float counter cells, k <= 3:
// u8 decay (advect_points.cc:155,174) + the reference's truncating blend replayed ring by ring
// (splat_points.cc:142-160), for a GROUP of up to 8 consecutive frames in one launch.
//
// What a pixel needs per frame is its old value d and, per sprite ring r (taps of equal opacity, ascending
// distance), the number n_r of points whose ring-r tap lands on it; the new value is
//     f_2^n2( f_1^n1( f_0^n0( decay(d) ) ) ),      f_r(d) = (u8)((1 - a_r) * d + a_r * 255)  (each op rounded)
// f_r is monotone on 0..255, so its iterates are stationary after D_r applications for every start value
// (k = 3: D = 6, 8, 12), and the whole per-frame map is ONE lookup
//     T[(min(n0,D0) * (D1+1) + min(n1,D1)) * (D2+1) + min(n2,D2)][d]
// in a table of 7 * 9 * 13 = 819 rows of 256 bytes (decay folded in) that fills the shared memory of an SM.
//
// Design (B200, round 2; the round-1 kernel staged counter tiles in shared memory and did four dependent
// byte lookups per pixel: 65 instructions per pixel, 12 M bank conflicts per launch, 0.105 ms per frame):
//   * no tile staging.  A lane owns 4 adjacent pixels and walks 8 rows; the counter row it needs is ONE
//     aligned 8-byte load (cells x0-1 .. x0+2 as two half2), the two cells to its right come from the next
//     lane by shuffle (lane 31 loads that word itself), rows above / below are the same registers one row
//     earlier / later.  Every counter row is read once per 8-row strip (+ 2 halo rows), fully coalesced.
//   * ring sums in packed half2 arithmetic (exact below 2048, monotone above -- anything >= D_r selects
//     the last row); the table row index is formed by two HFMA2 and read off the mantissa bits.
//   * ONE table lookup per pixel and frame.  Rows are 260 bytes apart, so rows that differ by one start in
//     neighbouring banks: lanes looking at similar pixel values with different counts spread over the banks.
//   * the table (213 KB for k = 3) arrives by TMA bulk copy (cp.async.bulk + mbarrier) while the CTA
//     zeroes the counter images of the PREVIOUS group -- the separate clear pass of round 1 is gone.
//   * pixels stay in registers across the frames of the group: the framebuffer is read once and written
//     once per group (recorded frames are written out individually).
//   * multi-GPU: the same kernel composites a rank's row band; it first waits (bounded) for every peer's
//     epoch of the group (see advect.cu, "frame fence").
#include "device_utils.cuh"
#include "internal.h"
#include "rings.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

namespace clumpy {

constexpr int GC_THREADS = 512;   // default: 16 warps; one CTA per SM (the k = 3 table takes the shared memory)
constexpr int GC_ROWS = 8;        // image rows per warp tile
constexpr int GC_TILE_W = 128;    // pixels per warp tile row: 32 lanes x 4 pixels
constexpr int GC_ROW_BYTES = 260; // table row stride

// ---- host: the combined replay table ---------------------------------------------------------------

int build_group_table(clumpy_cuda_ctx* ctx, const SpriteRings& rings, float decay, GroupTable* out) {
    *out = GroupTable();
    out->kernel_size = rings.kernel_size;
    if (rings.kernel_size > 3 || rings.nring > 3) return CLUMPY_OK; // not valid: caller uses another cell type
    // per ring: iterates of the truncating blend until stationary
    std::vector<std::vector<uint8_t>> iter(3);
    int dims[3] = {1, 1, 1};
    for (int r = 0; r < 3; ++r) {
        out->depth[r] = 0;
        std::vector<uint8_t>& rows = iter[r];
        rows.resize(256);
        for (int d = 0; d < 256; ++d) rows[d] = (uint8_t)d; // n = 0: identity
        const float a = r < rings.nring ? rings.ring_alpha[r] : 0.0f;
        if (a != 0.0f) { // f_0 is the identity
            int depth = 0;
            for (;;) {
                uint8_t cur[256];
                const uint8_t* prev = rows.data() + (size_t)depth * 256;
                for (int d = 0; d < 256; ++d) cur[d] = (uint8_t)blend_once_host(prev[d], a);
                if (memcmp(cur, prev, 256) == 0) break; // stationary: more hits change nothing
                if (++depth > 255) return CLUMPY_OK;    // never settles: not valid
                rows.insert(rows.end(), cur, cur + 256);
            }
            out->depth[r] = depth;
        }
        dims[r] = out->depth[r] + 1;
    }
    const size_t nrows = (size_t)dims[0] * dims[1] * dims[2];
    if (nrows > 1023) return CLUMPY_OK; // the row index must fit the 10 mantissa bits of a half
    const size_t bytes = round_up(nrows * GC_ROW_BYTES, 16);
    if (bytes > 220 * 1024) return CLUMPY_OK;
    std::vector<uint8_t> table(bytes, 0);
    uint8_t dec[256];
    for (int d = 0; d < 256; ++d) dec[d] = (uint8_t)decay_once_host((uint32_t)d, decay);
    for (int n0 = 0; n0 < dims[0]; ++n0)
        for (int n1 = 0; n1 < dims[1]; ++n1)
            for (int n2 = 0; n2 < dims[2]; ++n2) {
                uint8_t* row = table.data() + ((size_t)(n0 * dims[1] + n1) * dims[2] + n2) * GC_ROW_BYTES;
                for (int d = 0; d < 256; ++d)
                    row[d] = iter[2][(size_t)n2 * 256 + iter[1][(size_t)n1 * 256 + iter[0][(size_t)n0 * 256 + dec[d]]]];
            }
    CK(cudaMalloc(&out->d_table, bytes));
    CK(cudaMemcpyAsync(out->d_table, table.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); // `table` is a local
    out->nrows = (int)nrows;
    out->bytes = bytes;
    out->valid = true;
    return CLUMPY_OK;
}

void free_group_table(GroupTable* t) {
    if (t->d_table) cudaFree(t->d_table);
    t->d_table = nullptr;
    t->valid = false;
}

// ---- device ------------------------------------------------------------------------------------------

__device__ __forceinline__ __half2 as_h2(uint32_t u) { return *reinterpret_cast<const __half2*>(&u); }
__device__ __forceinline__ uint32_t as_u32(__half2 h) { return *reinterpret_cast<const uint32_t*>(&h); }
// (a.high, b.low) as a half2
__device__ __forceinline__ uint32_t hi_lo(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x5432); }

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// The new value of 4 pixels packed in `px` given their packed table row indices (as the low 10 bits of the
// halves of u01 / u23: 0x6400 | index, see below).
__device__ __forceinline__ uint32_t lookup4(const uint8_t* __restrict__ tab, uint32_t px, uint32_t u01, uint32_t u23) {
    const uint32_t i0 = u01 & 0x3FFu, i1 = (u01 >> 16) & 0x3FFu, i2 = u23 & 0x3FFu, i3 = (u23 >> 16) & 0x3FFu;
    const uint32_t v0 = tab[i0 * GC_ROW_BYTES + (px & 0xFFu)];
    const uint32_t v1 = tab[i1 * GC_ROW_BYTES + ((px >> 8) & 0xFFu)];
    const uint32_t v2 = tab[i2 * GC_ROW_BYTES + ((px >> 16) & 0xFFu)];
    const uint32_t v3 = tab[i3 * GC_ROW_BYTES + (px >> 24)];
    return v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
}

struct CellRow { uint32_t w0, w1, e; }; // cells (x0-1, x0), (x0+1, x0+2), (x0+3, x0+4) as half2 bit patterns

// The GC_ROWS + 2 counter rows a lane needs for frame f of its tile, as they are in memory (no shuffles here: the loads
// stay in flight until the words are used).
template <int N>
__device__ __forceinline__ void load_rows(const GroupCompositeArgs& a, int f, int y0, int tx, int lane, uint64_t pol, CellRow (&dst)[N]) {
    const __half* base = static_cast<const __half*>(a.cells[f]) + (size_t)y0 * a.pitch + tx * GC_TILE_W + 4 * lane;
#pragma unroll
    for (int r = 0; r < N; ++r) {
        const __half* p = base + (size_t)r * a.pitch;
        const uint2 v = ld_stream_u2(reinterpret_cast<const uint2*>(p), pol);
        dst[r].w0 = v.x;
        dst[r].w1 = v.y;
        dst[r].e = 0u;
        if (lane == 31) dst[r].e = ld_stream_u1(reinterpret_cast<const uint32_t*>(p + 4), pol);
    }
}

// THREADS: 512 by default (64 registers each: half the register file stays free for the advect kernel that
// shares the SM); 256 (128 registers, no spills) and 1024 exist for A/B runs (CLUMPY_CUDA_COMPOSITE_THREADS).
// PREFETCH (K = 3): the counter rows of frame f + 1 are loaded before frame f is worked through -- 30 more registers per
// thread, which the table's 213 KB make free (one CTA per SM either way).  Without it a warp issues its 10 row
// loads, waits a DRAM latency, computes, and only then asks for the next frame: 4 warps per scheduler cannot
// cover that and the issue slots were 45 % busy (round 2 ncu).
template <int K, int THREADS, bool PREFETCH>
__global__ void __launch_bounds__(THREADS, (THREADS == 1024 || PREFETCH) ? 1 : 2)
composite_group_kernel(GroupCompositeArgs a, const uint8_t* __restrict__ table_g, uint32_t table_bytes) {
    extern __shared__ __align__(128) uint8_t tab[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ int fence_ok;
    const int lane = threadIdx.x & 31;

    // 1. the table: TMA bulk copy, completion on an mbarrier
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(&bar)), "r"(table_bytes) : "memory");
        for (uint32_t off = 0; off < table_bytes; off += 65536u) {
            const uint32_t n = min(65536u, table_bytes - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_addr(tab + off)), "l"(table_g + off), "r"(n), "r"(smem_addr(&bar)) : "memory");
        }
    }

    // 2. (the counter images of the group before this one -- read by the launch before this one on the same stream,
    //    counted into again two groups on -- are zeroed slice by slice inside the tile loop below, where the stores
    //    share the DRAM with the latency-bound reads of the composite; as a phase of its own in front of it the
    //    zeroing of 8 x 67 MB was 90 us of the launch's 315)
    const uint64_t zero_pol = (a.hints & 8) ? make_stream_policy() : make_normal_policy();
    const uint64_t cell_pol = (a.hints & 4) ? make_stream_policy() : make_normal_policy();
    // 3. the wait for the table sits in front of the first lookup (and in front of every way out of the kernel: a CTA
    //    must not leave while a bulk copy into its shared memory is in flight)
    bool table_here = false;
    auto wait_for_table = [&]() {
        if (table_here) return;
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_addr(&bar)) : "memory");
        table_here = true;
    };

    // 4. multi-GPU: every source's entries of this group must have arrived (epoch flags in this rank's
    //    memory, stored by the peers' route kernels after a system-scope fence).  Bounded wait.
    if (a.flags) {
        if (threadIdx.x == 0) {
            int ok = 1;
            const unsigned long long t0 = globaltimer_ns();
            for (int s = 0; s < a.nsrc && ok; ++s) {
                // epochs only grow; compare as signed distance so that wrap-around stays harmless
                while ((int32_t)(ld_acquire_sys_u32(a.flags + s) - a.epoch) < 0) {
                    __nanosleep(200);
                    if (globaltimer_ns() - t0 > a.timeout_ns) { ok = 0; break; }
                }
            }
            if (!ok && blockIdx.x == 0) atomicAdd(a.errors, 1u);
            fence_ok = ok;
        }
        __syncthreads(); // orders every thread's reads of the cells behind thread 0's acquire
        if (!fence_ok) { wait_for_table(); return; } // the host turns the error count into a failed call (advect.cu)
    }

    const __half2 lim0 = __float2half2_rn(a.depth[0]), lim1 = __float2half2_rn(a.depth[1]), lim2 = __float2half2_rn(a.depth[2]);
    const __half2 mul1 = __float2half2_rn(a.depth[1] + 1.0f), mul2 = __float2half2_rn(a.depth[2] + 1.0f);
    const __half2 bias = __float2half2_rn(1024.0f); // 1024 + i has the bit pattern 0x6400 | i for i < 1024

    const int warp = blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5), nwarp = gridDim.x * (THREADS / 32);
    for (int t = warp; t < a.ntiles; t += nwarp) {
        const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
        const int y0 = ty * GC_ROWS, x0 = tx * GC_TILE_W + 4 * lane;
        const bool col_ok = x0 < a.width;
        const bool full = a.vec_ok && (x0 + 3 < a.width);

        uint32_t px[GC_ROWS];
#pragma unroll
        for (int r = 0; r < GC_ROWS; ++r) {
            px[r] = 0u;
            if (y0 + r < a.height && col_ok) {
                const uint8_t* src = a.img_in + (size_t)(y0 + r) * a.img_pitch + x0;
                if (full) px[r] = __ldcg(reinterpret_cast<const uint32_t*>(src));
                else
                    for (int j = 0; j < 4; ++j)
                        if (x0 + j < a.width) px[r] |= (uint32_t)src[j] << (8 * j);
            }
        }

        // this tile's slice of every image to zero: units [zlo, zhi) of zero_units 16-byte units
        const uint32_t zlo = (uint32_t)((unsigned long long)a.zero_units * (uint32_t)t / (uint32_t)a.ntiles);
        const uint32_t zhi = (uint32_t)((unsigned long long)a.zero_units * ((uint32_t)t + 1u) / (uint32_t)a.ntiles);
        CellRow ahead[(K == 3 && PREFETCH) ? GC_ROWS + 2 : 1]; // raw words of the NEXT frame's rows (e: lane 31 only)
        if (K == 3 && PREFETCH) load_rows(a, 0, y0, tx, lane, cell_pol, ahead);
        wait_for_table();
        for (int f = 0; f < a.nf; ++f) {
            for (int z = f; z < a.nzero; z += a.nf) // (a shorter group after a longer one: several images per frame)
                for (uint32_t i = zlo + (uint32_t)lane; i < zhi; i += 32u) st_hint_u4(a.zero[z] + i, make_uint4(0u, 0u, 0u, 0u), zero_pol);
            // first cell of this lane in padded row (y0 + r): image x = x0 - K / 2, image y = y0 + r - K / 2
            const __half* base = static_cast<const __half*>(a.cells[f]) + (size_t)y0 * a.pitch + tx * GC_TILE_W + 4 * lane;
            if (K == 1) {
#pragma unroll
                for (int r = 0; r < GC_ROWS; ++r) {
                    const uint2 v = ld_stream_u2(reinterpret_cast<const uint2*>(base + (size_t)r * a.pitch), cell_pol);
                    const __half2 u01 = __hadd2(__hmin2(as_h2(v.x), lim0), bias);
                    const __half2 u23 = __hadd2(__hmin2(as_h2(v.y), lim0), bias);
                    px[r] = lookup4(tab, px[r], as_u32(u01), as_u32(u23));
                }
            } else {
                CellRow row[GC_ROWS + 2];
                if (PREFETCH) {
#pragma unroll
                    for (int r = 0; r < GC_ROWS + 2; ++r) {
                        row[r].w0 = ahead[r].w0;
                        row[r].w1 = ahead[r].w1;
                        // the next lane's first word = cells (x0+3, x0+4); lane 31 has read it from memory
                        const uint32_t up = __shfl_down_sync(0xFFFFFFFFu, ahead[r].w0, 1);
                        row[r].e = lane == 31 ? ahead[r].e : up;
                    }
                    if (f + 1 < a.nf) load_rows(a, f + 1, y0, tx, lane, cell_pol, ahead);
                } else {
#pragma unroll
                    for (int r = 0; r < GC_ROWS + 2; ++r) {
                        const __half* p = base + (size_t)r * a.pitch;
                        const uint2 v = ld_stream_u2(reinterpret_cast<const uint2*>(p), cell_pol);
                        row[r].w0 = v.x;
                        row[r].w1 = v.y;
                        // the next lane's first word = cells (x0+3, x0+4); lane 31 reads it from memory
                        uint32_t e = 0u;
                        if (lane == 31) e = ld_stream_u1(reinterpret_cast<const uint32_t*>(p + 4), cell_pol);
                        const uint32_t up = __shfl_down_sync(0xFFFFFFFFu, v.x, 1);
                        row[r].e = lane == 31 ? e : up;
                    }
                }
#pragma unroll
                for (int r = 0; r < GC_ROWS; ++r) {
                    const CellRow &up = row[r], &mid = row[r + 1], &dn = row[r + 2];
                    // vertical pair sums of the three columns groups
                    const __half2 v0 = __hadd2(as_h2(up.w0), as_h2(dn.w0));
                    const __half2 v1 = __hadd2(as_h2(up.w1), as_h2(dn.w1));
                    const __half2 ve = __hadd2(as_h2(up.e), as_h2(dn.e));
                    // ring 0: the pixel's own cell; ring 1: left + right + above + below; ring 2: the corners
                    const __half2 c01 = as_h2(hi_lo(mid.w0, mid.w1)), c23 = as_h2(hi_lo(mid.w1, mid.e));
                    const __half2 s01 = __hadd2(__hadd2(as_h2(mid.w0), as_h2(mid.w1)), as_h2(hi_lo(as_u32(v0), as_u32(v1))));
                    const __half2 s23 = __hadd2(__hadd2(as_h2(mid.w1), as_h2(mid.e)), as_h2(hi_lo(as_u32(v1), as_u32(ve))));
                    const __half2 q01 = __hadd2(v0, v1), q23 = __hadd2(v1, ve);
                    // row index (n0 * (D1+1) + n1) * (D2+1) + n2, exact in half arithmetic (< 1024), biased by 1024
                    const __half2 t01 = __hfma2(__hmin2(c01, lim0), mul1, __hmin2(s01, lim1));
                    const __half2 t23 = __hfma2(__hmin2(c23, lim0), mul1, __hmin2(s23, lim1));
                    const __half2 u01 = __hfma2(t01, mul2, __hadd2(__hmin2(q01, lim2), bias));
                    const __half2 u23 = __hfma2(t23, mul2, __hadd2(__hmin2(q23, lim2), bias));
                    px[r] = lookup4(tab, px[r], as_u32(u01), as_u32(u23));
                }
            }
            uint8_t* out = a.frame_out[f];
            if (out && col_ok) {
#pragma unroll
                for (int r = 0; r < GC_ROWS; ++r) {
                    if (y0 + r >= a.height) break;
                    uint8_t* dst = out + (size_t)(y0 + r) * a.img_pitch + x0;
                    if (full) *reinterpret_cast<uint32_t*>(dst) = px[r];
                    else
                        for (int j = 0; j < 4; ++j)
                            if (x0 + j < a.width) dst[j] = (uint8_t)(px[r] >> (8 * j));
                }
            }
        }
    }
    wait_for_table();
}

int launch_composite_group(clumpy_cuda_ctx* ctx, cudaStream_t stream, const GroupTable& t, GroupCompositeArgs& a) {
    if (a.width <= 0 || a.height <= 0 || a.nf <= 0) return CLUMPY_OK;
    REQUIRE(t.valid && a.nf <= GC_MAX_FRAMES && a.frame_out[a.nf - 1], "bad composite group");
    a.tiles_x = (int)ceil_div((uint32_t)a.width, GC_TILE_W);
    a.ntiles = a.tiles_x * (int)ceil_div((uint32_t)a.height, GC_ROWS);
    for (int r = 0; r < 3; ++r) a.depth[r] = (float)t.depth[r];
    // the kernel zeroes the same slice of every image per tile: the images must have one size (they do: one set)
    a.zero_units = a.nzero > 0 ? a.zero_n16[0] : 0u;
    for (int z = 1; z < a.nzero; ++z) REQUIRE(a.zero_n16[z] == a.zero_units, "counter images of one set differ in size");
    bool vec = a.img_pitch % 4 == 0 && (uintptr_t)a.img_in % 4 == 0;
    for (int f = 0; f < a.nf; ++f) vec = vec && (uintptr_t)a.frame_out[f] % 4 == 0;
    a.vec_ok = vec ? 1 : 0;
    const int hints_env = [] { const char* v = getenv("CLUMPY_CUDA_HINTS"); return v ? atoi(v) : 0; }();
    a.hints = hints_env;
    const int threads_env = [] { const char* v = getenv("CLUMPY_CUDA_COMPOSITE_THREADS"); return v ? atoi(v) : 0; }();
    const int threads = threads_env == 256 || threads_env == 1024 ? threads_env : GC_THREADS;
    int grid = std::min(ctx->sm_count, (int)ceil_div((uint32_t)a.ntiles, (uint32_t)threads / 32));
    if (a.nzero > 0) grid = ctx->sm_count;
    using Kern = void (*)(GroupCompositeArgs, const uint8_t*, uint32_t);
    // The look-ahead variant takes all 128 registers of its 512 threads.  Beside a route launch on the other stream
    // (multi-GPU) that costs more than it gives -- no route CTA fits next to it: 0.0198 against 0.0183 ms per frame
    // on two ranks with a quarter of C4 -- so it is for launches that have the GPU to themselves.
    // CLUMPY_CUDA_COMPOSITE_PREFETCH=0|1 forces it off / on (A/B); 1024 threads leave 64 registers each: no room.
    const int prefetch_env = [] { const char* v = getenv("CLUMPY_CUDA_COMPOSITE_PREFETCH"); return v ? atoi(v) : -1; }();
    const bool prefetch = (prefetch_env < 0 ? a.alone : prefetch_env != 0) && t.kernel_size == 3 && threads == 512;
    const Kern kern = t.kernel_size == 1 ? (threads == 256 ? composite_group_kernel<1, 256, false> : threads == 1024 ? composite_group_kernel<1, 1024, false> : composite_group_kernel<1, 512, false>)
                    : prefetch           ? composite_group_kernel<3, 512, true>
                                         : (threads == 256 ? composite_group_kernel<3, 256, false> : threads == 1024 ? composite_group_kernel<3, 1024, false> : composite_group_kernel<3, 512, false>);
    // the opt-in to more than 48 KB of dynamic shared memory is per device and function
    static thread_local std::vector<std::pair<int, Kern>> opted;
    if (std::find(opted.begin(), opted.end(), std::make_pair(ctx->device, kern)) == opted.end()) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
        opted.emplace_back(ctx->device, kern);
    }
    kern<<<grid, threads, t.bytes, stream>>>(a, t.d_table, (uint32_t)t.bytes);
    CK_LAUNCH(ctx);
    return CLUMPY_OK;
}

} // namespace clumpy
