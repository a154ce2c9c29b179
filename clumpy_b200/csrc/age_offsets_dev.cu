// age_offsets_dev.cu -- the per-point reset offsets of commands/advect_points.cc:127-135 drawn ON THE DEVICE:
//     std::mt19937 g; g.seed(0); std::uniform_int_distribution<> d(0, nframes - 1); offset[i] = d(g);
// bit for bit (same restatement as age_offsets.cc, which stays as the host entry point and as this kernel's
// checker in tests/test_gpu_parity.py).
//
// Why here: on the host the 16.7 M draws of C4 take 10 ms and 67 MB of PCIe, which a single-GPU command hides
// behind its uploads but a multi-GPU rank -- whose points are already on the device -- waits for (11 of the
// 27 ms of a 20-frame command on 2 GPUs, profiles/r2_e2e_multi_trace.txt).
//
// mt19937 is one sequential recurrence, x[n + 624] = f(x[n], x[n + 1], x[n + 397]): at any time the next 227
// words depend only on words that exist, so ONE CTA advances the generator 227 (227, 170) words per barrier with
// the state in shared memory, tempers each word, maps it to a draw (libstdc++: Lemire's multiply-shift,
// bits/uniform_int_dist.h) and stores it.  A raw word is REJECTED with probability (2^32 mod range) / 2^32 (four
// times in 16.7 M draws for range 1000) and then every later draw moves up by one: a block of 624 words that holds
// a rejected one is compacted by thread 0 (rare), all other blocks store straight to their places.
#include "internal.h"

namespace clumpy {

namespace {

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b, uint32_t m) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
    return m ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908B0DFu);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    return y;
}

constexpr int MT_N = 624, MT_M = 397, MT_LAG = MT_N - MT_M; // 227

__global__ void __launch_bounds__(256)
age_offsets_kernel(uint32_t range, uint32_t n, uint32_t* __restrict__ out) {
    __shared__ uint32_t st[2][MT_N];
    __shared__ uint32_t raw[MT_N];
    __shared__ uint32_t kept;
    const uint32_t t = threadIdx.x;
    if (t == 0) { // g.seed(0)
        uint32_t v = 0u;
        st[0][0] = v;
        for (uint32_t i = 1; i < MT_N; ++i) { v = 1812433253u * (v ^ (v >> 30)) + i; st[0][i] = v; }
    }
    __syncthreads();
    const uint32_t threshold = (0u - range) % range;
    uint32_t produced = 0;
    int cur = 0;
    while (produced < n) {
        const uint32_t* __restrict__ c = st[cur];
        uint32_t* __restrict__ x = st[cur ^ 1];
        bool rejected = false;
        auto emit = [&](uint32_t k, uint32_t v) {
            x[k] = v;
            const uint32_t y = mt_temper(v);
            raw[k] = y;
            const unsigned long long prod = (unsigned long long)y * range;
            rejected |= (uint32_t)prod < threshold;
            const uint32_t i = produced + k; // its place when no word of this block is rejected
            if (i < n) out[i] = (uint32_t)(prod >> 32);
        };
        if (t < MT_LAG) emit(t, mt_twist(c[t], c[t + 1], c[t + MT_M]));
        __syncthreads();
        if (t < MT_LAG) { const uint32_t k = t + MT_LAG; emit(k, mt_twist(c[k], c[k + 1], x[k - MT_LAG])); }
        __syncthreads();
        if (t < MT_N - 1 - 2 * MT_LAG) { const uint32_t k = t + 2 * MT_LAG; emit(k, mt_twist(c[k], c[k + 1], x[k - MT_LAG])); }
        else if (t == MT_N - 1 - 2 * MT_LAG) emit(MT_N - 1, mt_twist(c[MT_N - 1], x[0], x[MT_M - 1]));
        if (__syncthreads_or(rejected)) {
            if (t == 0) {
                uint32_t j = 0;
                for (uint32_t k = 0; k < MT_N; ++k) {
                    const unsigned long long prod = (unsigned long long)raw[k] * range;
                    if ((uint32_t)prod < threshold) continue;
                    if (produced + j < n) out[produced + j] = (uint32_t)(prod >> 32);
                    ++j;
                }
                kept = j;
            }
            __syncthreads();
            produced += kept;
            __syncthreads(); // (kept is rewritten by the next block that needs it)
        } else {
            produced += MT_N;
        }
        cur ^= 1;
    }
}

} // namespace

// `n` draws from [0, nframes) into out_dev, asynchronously on `stream`.
int launch_age_offsets(clumpy_cuda_ctx* ctx, cudaStream_t stream, uint32_t nframes, uint32_t n, uint32_t* out_dev) {
    if (n == 0) return CLUMPY_OK;
    age_offsets_kernel<<<1, 256, 0, stream>>>(nframes, n, out_dev);
    CK_LAUNCH(ctx);
    return CLUMPY_OK;
}

} // namespace clumpy

using namespace clumpy;

// Device-side twin of clumpy_cuda_age_offsets: `npts` draws into device memory, on the context's stream.
CLUMPY_API int clumpy_cuda_age_offsets_dev(clumpy_cuda_ctx* ctx, uint32_t nframes, uint32_t npts, uint32_t* out_dev) {
    REQUIRE(ctx && (out_dev || npts == 0), "null argument");
    REQUIRE(nframes > 0 && nframes <= 0x80000000u, "nframes must be in 1 .. 2^31");
    CK(cudaSetDevice(ctx->device));
    return launch_age_offsets(ctx, ctx->stream, nframes, npts, out_dev);
}
