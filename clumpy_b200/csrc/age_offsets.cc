// age_offsets.cc -- the per-point reset offsets of commands/advect_points.cc:127-135:
//     std::mt19937 g; g.seed(0); std::uniform_int_distribution<> d(0, nframes - 1); offset[i] = d(g);
// bit for bit, without going through <random> one draw at a time (79 ms for the 16.7 M points of the C4
// workload on the GPU box's host -- 40 % of the whole command's wall clock in round 1).
//
// The arithmetic is fixed by the standard (mt19937) and by libstdc++'s distribution, which for a 32-bit
// generator is Lemire's multiply-shift with rejection (bits/uniform_int_dist.h, _S_nd):
//     prod = uint64(g()) * range;  low = uint32(prod);
//     if (low < range) { threshold = -range % range; while (low < threshold) { prod = uint64(g()) * range; low = uint32(prod); } }
//     return prod >> 32;
// A rejection consumes an extra generator output and so shifts every later draw; it happens with probability
// (2^32 mod range) / 2^32 per draw (about 4 times in 16.7 M draws for range 1000).  The fast path turns a whole
// block of 624 generator outputs into draws with branch-free, vectorisable loops and only notes whether ANY of
// them had low < range; if so that block is redone draw by draw.  tests/test_abi.py compares the result
// with <random> itself for ranges from 1 to 2^31 - 1.
#include <algorithm>
#include <cstdint>
#include <future>
#include <mutex>
#include <vector>

#include "mt_checkpoints.h"

namespace {

struct Mt19937 {
    uint32_t state[624];
    uint32_t out[624]; // tempered outputs of the current block
    int pos = 624;

    void seed(uint32_t v) {
        state[0] = v;
        for (int i = 1; i < 624; ++i) state[i] = 1812433253u * (state[i - 1] ^ (state[i - 1] >> 30)) + (uint32_t)i;
        pos = 624;
    }
    static inline uint32_t twist(uint32_t a, uint32_t b, uint32_t m) {
        const uint32_t y = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
        return m ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908B0DFu);
    }
    // The next 624 words.  Within each of the two loops an element depends only on elements at least 227
    // positions away, so they vectorise.
    __attribute__((target_clones("avx2", "default"))) void refill() {
        uint32_t* __restrict__ p = state;
#pragma GCC ivdep
        for (int k = 0; k < 227; ++k) p[k] = twist(p[k], p[k + 1], p[k + 397]);
#pragma GCC ivdep
        for (int k = 227; k < 623; ++k) p[k] = twist(p[k], p[k + 1], p[k - 227]);
        p[623] = twist(p[623], p[0], p[396]);
        for (int k = 0; k < 624; ++k) {
            uint32_t y = p[k];
            y ^= y >> 11;
            y ^= (y << 7) & 0x9D2C5680u;
            y ^= (y << 15) & 0xEFC60000u;
            y ^= y >> 18;
            out[k] = y;
        }
        pos = 0;
    }
    // the state alone (no outputs): what the checkpoint table needs
    __attribute__((target_clones("avx2", "default"))) void advance_state() {
        uint32_t* __restrict__ p = state;
#pragma GCC ivdep
        for (int k = 0; k < 227; ++k) p[k] = twist(p[k], p[k + 1], p[k + 397]);
#pragma GCC ivdep
        for (int k = 227; k < 623; ++k) p[k] = twist(p[k], p[k + 1], p[k - 227]);
        p[623] = twist(p[623], p[0], p[396]);
    }
    inline uint32_t next() {
        if (pos == 624) refill();
        return out[pos++];
    }
};

} // namespace

namespace clumpy {

void age_offsets_host(uint32_t nframes, uint32_t npts, uint32_t* dst) {
    Mt19937 g;
    g.seed(0);
    const uint32_t range = nframes; // urange + 1 with urange = nframes - 1
    const uint32_t threshold = (0u - range) % range;
    uint32_t i = 0;
    while (i < npts) {
        if (g.pos == 624) g.refill();
        const uint32_t n = std::min<uint32_t>(624u - (uint32_t)g.pos, npts - i);
        const uint32_t* __restrict__ r = g.out + g.pos;
        uint32_t* __restrict__ o = dst + i;
        uint32_t risky = 0;
        for (uint32_t k = 0; k < n; ++k) {
            const uint64_t prod = (uint64_t)r[k] * range;
            o[k] = (uint32_t)(prod >> 32);
            risky |= (uint32_t)((uint32_t)prod < range);
        }
        if (!risky) { g.pos += (int)n; i += n; continue; }
        // some draw of this block may be rejected: redo the block one draw at a time
        uint32_t done = 0;
        while (done < n) {
            uint64_t prod = (uint64_t)g.next() * range;
            uint32_t low = (uint32_t)prod;
            if (low < range)
                while (low < threshold) { prod = (uint64_t)g.next() * range; low = (uint32_t)prod; }
            o[done++] = (uint32_t)(prod >> 32);
        }
        i += done;
    }
}

// ---- generator states at fixed distances, for the device-side draws (age_offsets_dev.cu) --------------------------
// The reference always seeds with 0 (advect_points.cc:129), so the generator's word sequence is one fixed sequence
// and its state after MT_CKPT_BLOCKS * s refills is a constant.  The table of those states lets MT_CKPT_COUNT CTAs
// generate disjoint stretches of the sequence at once.  It is computed once per process on a background thread
// (started when the first context is created; ~10 ms of one core) instead of being shipped as 160 KB of literals.
namespace {
std::once_flag ckpt_once;
std::shared_future<std::vector<uint32_t>> ckpt_future;

std::vector<uint32_t> compute_checkpoints() {
    std::vector<uint32_t> table((size_t)MT_CKPT_COUNT * 624);
    Mt19937 g;
    g.seed(0);
    for (int s = 0; s < MT_CKPT_COUNT; ++s) {
        std::copy(g.state, g.state + 624, table.begin() + (size_t)s * 624);
        if (s + 1 < MT_CKPT_COUNT)
            for (int b = 0; b < MT_CKPT_BLOCKS; ++b) g.advance_state();
    }
    return table;
}
} // namespace

void mt_checkpoints_prefetch() {
    std::call_once(ckpt_once, [] { ckpt_future = std::async(std::launch::async, compute_checkpoints).share(); });
}

const uint32_t* mt_checkpoints() {
    mt_checkpoints_prefetch();
    const std::shared_future<std::vector<uint32_t>> mine = ckpt_future; // (each caller waits through its own copy)
    return mine.get().data(); // the vector lives in the shared state, i.e. as long as ckpt_future does
}

} // namespace clumpy
