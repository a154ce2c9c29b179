// rings.cuh -- sprite rings shared by the composite kernels, and the reference's per-hit arithmetic
// as host+device functions (the host copies build the replay tables of the dense path).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace clumpy {

// Compile-time map tap (i, j) -> rank of its squared distance among the distinct distances of a
// K x K sprite (the same ranking make_sprite_rings uses on the host), and how many taps share it.
template <int K>
struct RingTable {
    int8_t ring[K * K];
    int8_t mult[16];
    int nring;
    int max_mult;
    constexpr RingTable() : ring{}, mult{}, nring(0), max_mult(0) {
        constexpr int H = K / 2;
        int n = 0;
        for (int d2 = 0; d2 <= 2 * H * H; ++d2) {
            int found = 0;
            for (int j = 0; j < K; ++j)
                for (int i = 0; i < K; ++i)
                    if ((i - H) * (i - H) + (j - H) * (j - H) == d2) { ring[i + j * K] = (int8_t)n; ++found; }
            if (found) {
                mult[n] = (int8_t)found;
                if (found > max_mult) max_mult = found;
                ++n;
            }
        }
        nring = n;
    }
};

// float -> u8 store as the reference binary does it: cvttss2si r32 (INT32_MIN when the value does
// not fit or is NaN), low byte.
__host__ __device__ inline uint32_t to_u8_x86(float f) {
    const float af = f < 0.0f ? -f : f;
    const int32_t i = (af < 2147483648.0f) ? (int32_t)f : (int32_t)0x80000000;
    return (uint32_t)i & 0xFFu;
}

// splat_points.cc:153-155 with opacity a: (uint8_t)((1 - a) * dst + a * 255), every operation rounded
// on its own.  Host copy: plain x86-64 code has no FMA to contract into; device callers use the
// *_rn intrinsics (see splat.cu).
inline uint32_t blend_once_host(uint32_t d, float a) {
    volatile float one_minus_a = 1.0f - a;
    volatile float left = one_minus_a * (float)d;
    volatile float right = a * 255.0f;
    volatile float sum = left + right;
    return to_u8_x86(sum);
}

// advect_points.cc:155,174
inline uint32_t decay_once_host(uint32_t d, float decay) {
    volatile float v = (float)d * decay;
    return to_u8_x86(v);
}

} // namespace clumpy
