// device_utils.cuh -- small device helpers shared by the kernels.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace clumpy {

// float -> uint32_t exactly as the reference binary does it (gcc/x86-64 compiles the implicit
// conversion at advect_points.cc:35 and splat_points.cc:110-111 to cvttss2si r64 + low half):
// truncate toward zero through int64; values that do not fit (and NaN) give 0.
__device__ __forceinline__ uint32_t f2u32_x86(float f) {
    if (fabsf(f) < 2147483648.0f) return (uint32_t)__float2int_rz(f); // common case, one F2I
    if (!(fabsf(f) < 9223372036854775808.0f)) return 0u;
    return (uint32_t)(unsigned long long)__float2ll_rz(f);
}

// (int32_t) cast of splat_points.cc:143-144: cvttss2si r32, "indefinite" INT32_MIN when the value
// does not fit or is NaN.
__device__ __forceinline__ int32_t f2i32_x86(float f) {
    return (fabsf(f) < 2147483648.0f) ? __float2int_rz(f) : (int32_t)0x80000000;
}

// Both conversions of one coordinate from ONE F2I: i = __float2int_rz(f) is kept, and for f inside (-2^31, 2^31)
// the uint32 the reference's fetch sees (cvttss2si r64, low half) is (uint32_t)i and the int32 its splat sees is i;
// outside (or NaN), f2u32_x86 / f2i32_x86 give the answers.  The route kernel, which is issue-bound at per-rank
// sizes, truncates a position once and uses it for the splat of this frame and the gather of the next.
__device__ __forceinline__ uint32_t fetch_index(int32_t i, float f) { return fabsf(f) < 2147483648.0f ? (uint32_t)i : f2u32_x86(f); }
__device__ __forceinline__ int32_t splat_index(int32_t i, float f) { return fabsf(f) < 2147483648.0f ? i : (int32_t)0x80000000; }

// float -> uint8_t store of advect_points.cc:155 / splat_points.cc:155: cvttss2si r32, low byte.
__device__ __forceinline__ uint32_t f2u8_x86(float f) {
    return (uint32_t)f2i32_x86(f) & 0xFFu;
}

__device__ __forceinline__ uint32_t encode_ordered(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Block-wide min/max, then one atomic pair per block.  `lo`/`hi` follow std::min/std::max
// semantics of the callers (values that lose every `<` comparison never get here as NaN).
// has_lo / has_hi tell whether the thread saw any candidate.
__device__ __forceinline__ void block_minmax_commit(float lo, bool has_lo, float hi, bool has_hi,
                                                    uint32_t* mm) {
    uint32_t elo = has_lo ? encode_ordered(lo) : 0xFFFFFFFFu;
    uint32_t ehi = has_hi ? encode_ordered(hi) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        elo = min(elo, __shfl_xor_sync(0xFFFFFFFFu, elo, o));
        ehi = max(ehi, __shfl_xor_sync(0xFFFFFFFFu, ehi, o));
    }
    __shared__ uint32_t s_lo[32], s_hi[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x * blockDim.y + 31) >> 5;
    if (lane == 0) { s_lo[warp] = elo; s_hi[warp] = ehi; }
    __syncthreads();
    if (warp == 0) {
        elo = lane < nwarp ? s_lo[lane] : 0xFFFFFFFFu;
        ehi = lane < nwarp ? s_hi[lane] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            elo = min(elo, __shfl_xor_sync(0xFFFFFFFFu, elo, o));
            ehi = max(ehi, __shfl_xor_sync(0xFFFFFFFFu, ehi, o));
        }
        if (lane == 0) {
            if (elo != 0xFFFFFFFFu) atomicMin(mm, elo);
            if (ehi != 0u) atomicMax(mm + 1, ehi);
        }
    }
}

// Geometry of the padded counter image (see splat.cu) as the kernels see it.
struct GeomDev {
    int width, height, halo, pitch;
    int keep_pos; // advect kernels: 1 = positions keep the normal L2 policy (default 0: evict-first stream)
    int hints;    // L2 eviction priorities (CLUMPY_CUDA_HINTS, A/B): bit 0 = velocity gathers evict-last,
                  // bit 1 = counter reductions evict-first
};

// Index of the counter for a point whose truncated position is (cx, cy), or -1 when none of its
// sprite taps can land inside the image.
template <bool WRAP>
__device__ __forceinline__ long long hist_index(int32_t cx, int32_t cy, const GeomDev& g) {
    const uint32_t uy = (uint32_t)cy + (uint32_t)g.halo;
    if (uy >= (uint32_t)(g.height + 2 * g.halo)) return -1;
    uint32_t ux;
    if (WRAP) {
        int32_t m = cx % g.width; // splat_points.cc:150: (W + x0 % W) % W is the true modulus
        if (m < 0) m += g.width;
        ux = (uint32_t)(m + g.halo);
    } else {
        ux = (uint32_t)cx + (uint32_t)g.halo;
        if (ux >= (uint32_t)(g.width + 2 * g.halo)) return -1;
    }
    return (long long)uy * g.pitch + ux;
}

// Streaming (read-once / write-once) accesses: mark them evict-first in L2 and keep them out of L1
// so that they do not displace what should stay cached (velocity texels, splat counters, the
// framebuffer).  The policy is a register operand (createpolicy), valid for any vector width.
__device__ __forceinline__ uint64_t make_stream_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// The opposite hint for the splat counters: touched three times per frame (count, composite, clear)
// with ~0.6 GB of streaming traffic in between, so ask L2 to evict them last.
__device__ __forceinline__ uint64_t make_keep_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t make_normal_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float2 ld_gather_f2(const float2* p, uint64_t pol) { // read-only data, through L1
    float2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_hint_u32(const uint32_t* p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_hint_u4(uint4* p, uint4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 ld_stream_f4(const float4* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float2 ld_stream_f2(const float2* p, uint64_t pol) {
    float2 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;"
                 : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint2 ld_stream_u2(const uint2* p, uint64_t pol) {
    uint2 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
                 : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream_u1(const uint32_t* p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream_off(const uint32_t* p, uint64_t pol) { return ld_stream_u1(p, pol); }
__device__ __forceinline__ uint32_t ld_stream_off(const uint16_t* p, uint64_t pol) {
    uint16_t v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_stream_f2(float2* p, float2 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

} // namespace clumpy
