// common.cuh -- shared structures of the build / query kernels and the C ABI glue.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/boomphf_cuda.h"
#include "hash.cuh"

namespace bphf {

// ---- HBM layout -----------------------------------------------------------------------------
// One "arena" of u64 words holds the bit vectors of ALL levels back to back; level l starts at
// word_off[l], a multiple of 8 words (64 B), and is zero padded up to the next multiple of 8.
// Rank entry r of the arena (one per 8 words / 512 bits) therefore covers arena words [8r, 8r+8),
// and the reference's per-level rank arrays (src/lib.rs:277-297: running popcount across all
// levels, one entry per 8 words) are exactly consecutive slices of ONE exclusive prefix sum over
// the arena's per-block popcounts -- level l's ranks are arena_ranks[word_off[l]/8 ...].
// During construction a second arena with the same offsets holds the collision bit vectors.
constexpr uint32_t LEVEL_ALIGN_WORDS = 8;

struct LevelDesc {
    uint64_t bits;      // BitVector.bits = capacity()  (src/bitvector.rs:359)
    uint64_t n_words;   // u64s(bits)                   (src/bitvector.rs:426-428)
    uint64_t word_off;  // first arena word of this level (multiple of 8)
    uint64_t k_in;      // keys entering this level (build only)
    uint64_t sp0;       // seed_xor_p0(seed index) used to BUILD this level
    uint64_t sp0_query; // seed_xor_p0(level index): what Mphf::hash uses (src/lib.rs:327)
};

enum BuildStatus : uint32_t {
    ST_RUNNING = 0,
    ST_DONE = 1,
    ST_ERR_MAX_ITERS = 2,  // level count exceeded MAX_ITERS (src/lib.rs:265-268, 398-401)
    ST_NEED_WORDS = 3,     // arena too small for the next level: host grows it and resumes
    ST_NEED_LIST = 4,      // redo list buffer too small for the next level: host grows it and resumes
    ST_ERR_INTERNAL = 5,
};

// device-resident build state; the level loop runs without host round trips (kernels for all
// expected levels are enqueued up front and read their sizes from here)
struct BuildState {
    uint32_t status;
    uint32_t n_levels;    // levels completed so far
    uint32_t tail_level;  // first level owned by the single-CTA tail kernel (0xffffffff: none yet)
    uint32_t fin_ticket;  // "last block done" ticket of the finalize kernel
    uint64_t redo_count;  // append cursor of the redo list being written by the current pass
    uint64_t uniq_count;  // popcount(a & ~collide) of the level being finalized
    uint64_t words_used;  // arena words handed out so far (multiple of 8)
    uint64_t words_cap;   // arena capacity in words (excluding the tail reserve)
    uint64_t list_cap[2]; // capacity (keys) of redo buffers: level l's key list lives in buf[l & 1]
    uint64_t seed0;       // starting_seed.unwrap_or(0)
    double gamma;
    uint64_t need;        // how much ST_NEED_* asks for
    uint64_t err_info;
    LevelDesc lv[BPHF_MAX_LEVELS + 1];
};

// tail kernel: one CTA finishes all remaining levels out of shared memory once a level has
// k_in <= TAIL_MAX_KEYS and bits <= TAIL_MAX_BITS
constexpr uint32_t TAIL_MAX_KEYS = 4096;
constexpr uint32_t TAIL_MAX_BITS = 65536;                    // 8 KB per bit array
constexpr uint32_t TAIL_MAX_WORDS = TAIL_MAX_BITS / 64;      // 1024
// arena words reserved so the tail can never run out: every remaining level needs <= this many
constexpr uint64_t TAIL_RESERVE_WORDS = (uint64_t)(BPHF_MAX_LEVELS + 1) * (TAIL_MAX_WORDS + LEVEL_ALIGN_WORDS);

__host__ __device__ __forceinline__ uint64_t round_up8(uint64_t w) { return (w + 7) & ~7ULL; }

// fill in lv[level] for k keys; returns false when the arena cannot hold it (need -> words required)
__host__ __device__ inline void make_level(BuildState* st, uint32_t level, uint64_t k) {
    LevelDesc& d = st->lv[level];
    d.bits = level_size(st->gamma, k);
    d.n_words = (d.bits + 63) / 64;
    d.word_off = st->words_used;
    d.k_in = k;
    d.sp0 = seed_xor_p0(st->seed0 + level);
    d.sp0_query = seed_xor_p0(level);
}

}  // namespace bphf
