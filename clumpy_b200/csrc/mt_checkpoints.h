// mt_checkpoints.h -- shared between age_offsets.cc (host: computes the table) and age_offsets_dev.cu (device: uses it).
#pragma once
#include <cstdint>

namespace clumpy {

// checkpoint s = the 624 state words of std::mt19937 seeded with 0 BEFORE refill number s * MT_CKPT_BLOCKS, i.e. the
// state from which words s * MT_CKPT_BLOCKS * 624 ... of the generator's output come next
constexpr int MT_CKPT_COUNT = 64;
constexpr int MT_CKPT_BLOCKS = 840; // 524 160 words per stretch; 64 stretches cover 33.5 M words

void mt_checkpoints_prefetch();      // starts the background computation (idempotent)
const uint32_t* mt_checkpoints();    // MT_CKPT_COUNT x 624 words; waits for the computation if it is still running

} // namespace clumpy
