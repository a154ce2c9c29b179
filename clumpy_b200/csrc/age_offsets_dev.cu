// age_offsets_dev.cu -- the per-point reset offsets of commands/advect_points.cc:127-135 drawn ON THE DEVICE:
//     std::mt19937 g; g.seed(0); std::uniform_int_distribution<> d(0, nframes - 1); offset[i] = d(g);
// bit for bit (same restatement as age_offsets.cc, which stays as the host entry point and as this file's checker
// in tests/test_gpu_parity.py).
//
// Why here: on the host the 16.7 M draws of C4 take 10 ms and 67 MB of PCIe, which a single-GPU command hides
// behind its uploads but a multi-GPU rank -- whose points are already on the device -- waits for (11 of the
// 27 ms of a 20-frame command on 2 GPUs, round 2).
//
// mt19937 is one sequential recurrence, x[n + 624] = f(x[n], x[n + 1], x[n + 397]): at any time the next 227
// words depend only on words that exist, so a CTA advances the generator 227 (227, 170) words per barrier with the
// state in shared memory -- 0.47 us per 624 words, 12.7 ms for C4 if ONE CTA did it all.  But the seed is the
// constant 0, so the state at any fixed distance into the sequence is a constant too: the host computes a table of
// 64 such states once per process (age_offsets.cc, mt_checkpoints.h) and CTA s generates stretch s of the sequence.
//
// A word is mapped to a draw by libstdc++'s rule (Lemire's multiply-shift, bits/uniform_int_dist.h): it is
// REJECTED with probability (2^32 mod range) / 2^32 -- four times in 16.7 M draws for range 1000, a quarter of the
// words for range 3 << 29 -- and every later draw then moves up by one place.  So there are two launches: the first
// counts the rejected words of every stretch, the second generates the stretches again and stores the accepted
// draws of stretch s behind those of the stretches before it.  The LAST stretch has no end: its CTA goes on until
// draw n - 1 is out, however many words that takes (beyond 33.5 M draws, or with many rejections, it is the
// sequential part; nframes is the length of an animation, so in practice the first case only).
#include "internal.h"
#include "mt_checkpoints.h"

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
constexpr int MT_THREADS = 256;

// WRITE = false: rejected[s] = number of rejected words among the MT_CKPT_BLOCKS * 624 words of stretch s
//               (grid = the stretches that have a successor).
// WRITE = true:  the accepted draws of stretch s go to out[first ...), first = s * stretch - (rejected before s);
//               the last CTA of the grid runs until out[n - 1] is written.
template <bool WRITE>
__global__ void __launch_bounds__(MT_THREADS)
age_offsets_kernel(const uint32_t* __restrict__ checkpoints, uint32_t range, uint32_t n, uint32_t* __restrict__ rejected,
                   uint32_t* __restrict__ out) {
    __shared__ uint32_t st[2][MT_N];
    __shared__ uint32_t raw[MT_N];
    __shared__ uint32_t kept;
    const uint32_t t = threadIdx.x, s = blockIdx.x;
    for (uint32_t i = t; i < MT_N; i += MT_THREADS) st[0][i] = checkpoints[(size_t)s * MT_N + i];
    const uint32_t threshold = (0u - range) % range;
    const bool open_ended = WRITE && s == gridDim.x - 1u;
    unsigned long long produced = (unsigned long long)s * MT_CKPT_BLOCKS * MT_N; // index of the next draw
    if (WRITE)
        for (uint32_t i = 0; i < s; ++i) produced -= rejected[i];
    uint32_t nrejected = 0; // (thread 0's count, WRITE = false)
    __syncthreads();
    int cur = 0;
    for (uint32_t block = 0; open_ended ? produced < n : block < (uint32_t)MT_CKPT_BLOCKS; ++block) {
        const uint32_t* __restrict__ c = st[cur];
        uint32_t* __restrict__ x = st[cur ^ 1];
        bool reject = false;
        uint32_t draw[3] = {0u, 0u, 0u};
        auto emit = [&](int phase, uint32_t k, uint32_t v) {
            x[k] = v;
            const uint32_t y = mt_temper(v);
            raw[k] = y;
            const unsigned long long prod = (unsigned long long)y * range;
            reject |= (uint32_t)prod < threshold;
            draw[phase] = (uint32_t)(prod >> 32);
        };
        if (t < MT_LAG) emit(0, t, mt_twist(c[t], c[t + 1], c[t + MT_M]));
        __syncthreads();
        if (t < MT_LAG) { const uint32_t k = t + MT_LAG; emit(1, k, mt_twist(c[k], c[k + 1], x[k - MT_LAG])); }
        __syncthreads();
        if (t < MT_N - 1 - 2 * MT_LAG) { const uint32_t k = t + 2 * MT_LAG; emit(2, k, mt_twist(c[k], c[k + 1], x[k - MT_LAG])); }
        else if (t == MT_N - 1 - 2 * MT_LAG) emit(2, MT_N - 1, mt_twist(c[MT_N - 1], x[0], x[MT_M - 1]));
        if (!__syncthreads_or(reject)) {
            // the common case: word k of the block is draw `produced + k`
            if (WRITE) {
                if (t < MT_LAG) {
                    if (produced + t < n) out[produced + t] = draw[0];
                    if (produced + t + MT_LAG < n) out[produced + t + MT_LAG] = draw[1];
                }
                if (t <= MT_N - 1 - 2 * MT_LAG) {
                    const uint32_t k = t == MT_N - 1 - 2 * MT_LAG ? MT_N - 1 : t + 2 * MT_LAG;
                    if (produced + k < n) out[produced + k] = draw[2];
                }
            }
            produced += MT_N;
        } else {
            // a rejected word in this block: thread 0 goes through it word by word
            if (t == 0) {
                uint32_t j = 0;
                for (uint32_t k = 0; k < MT_N; ++k) {
                    const unsigned long long prod = (unsigned long long)raw[k] * range;
                    if ((uint32_t)prod < threshold) continue;
                    if (WRITE && produced + j < n) out[produced + j] = (uint32_t)(prod >> 32);
                    ++j;
                }
                kept = j;
                nrejected += MT_N - j;
            }
            __syncthreads();
            produced += kept;
            __syncthreads(); // (kept is rewritten by the next block that needs it)
        }
        cur ^= 1;
    }
    if (!WRITE && t == 0) rejected[s] = nrejected;
}

} // namespace

// `n` draws from [0, nframes) into out_dev, asynchronously on `stream`.
int launch_age_offsets(clumpy_cuda_ctx* ctx, cudaStream_t stream, uint32_t nframes, uint32_t n, uint32_t* out_dev) {
    if (n == 0) return CLUMPY_OK;
    if (!ctx->mt_checkpoints_dev) {
        const uint32_t* table = mt_checkpoints(); // (computed in the background since the first context was created)
        const size_t bytes = (size_t)MT_CKPT_COUNT * MT_N * sizeof(uint32_t);
        CK(cudaMalloc(&ctx->mt_checkpoints_dev, bytes + MT_CKPT_COUNT * sizeof(uint32_t)));
        CK(cudaMemcpy(ctx->mt_checkpoints_dev, table, bytes, cudaMemcpyHostToDevice));
    }
    const uint32_t* table_dev = ctx->mt_checkpoints_dev;
    uint32_t* rejected = ctx->mt_checkpoints_dev + (size_t)MT_CKPT_COUNT * MT_N;
    const unsigned long long stretch = (unsigned long long)MT_CKPT_BLOCKS * MT_N;
    const uint32_t stretches = (uint32_t)std::min<unsigned long long>((n + stretch - 1) / stretch, MT_CKPT_COUNT);
    if (stretches > 1) {
        age_offsets_kernel<false><<<stretches - 1, MT_THREADS, 0, stream>>>(table_dev, nframes, n, rejected, nullptr);
        CK_LAUNCH(ctx);
    }
    age_offsets_kernel<true><<<stretches, MT_THREADS, 0, stream>>>(table_dev, nframes, n, rejected, out_dev);
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
