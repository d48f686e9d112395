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

constexpr int SHARD_MAX = 8;  // shards (GPUs) of a sharded build

struct LevelDesc {
    uint64_t bits;      // BitVector.bits = capacity()  (src/bitvector.rs:359)
    uint64_t n_words;   // u64s(bits)                   (src/bitvector.rs:426-428)
    uint64_t word_off;  // first arena word of this level (multiple of 8)
    uint64_t k_in;      // keys entering this level, over ALL shards (build only): sizes the level
    uint64_t k_loc;     // keys of THIS shard entering the level (== k_in for an unsharded build)
    uint64_t sp0;       // seed_xor_p0(seed index) used to BUILD this level
    uint64_t sp0_query; // seed_xor_p0(level index): what Mphf::hash uses (src/lib.rs:327)
    // bucketed construction (build only): the level's bit range is cut into b_count buckets of b_range bits
    // (multiple of 512); the keys entering the level are stored grouped by bucket, bucket b at
    // buf[level & 1] + b * b_cap, so one CTA owns a bucket's slice of (a, collide) in shared memory
    uint64_t b_shift;   // log2(b_range): bucket = idx >> b_shift
    uint32_t b_range;   // 0 = plain level (one contiguous key list, L2-atomic kernels); else a power of two
    uint32_t b_count;
    uint32_t b_cap;     // key capacity of each bucket region
    uint32_t pad_;
    // sharded build, exchange ("owner computes") level: shard g owns bits [g * x_slice_bits, (g + 1) * x_slice_bits) of
    // the level and marks every shard's keys that fall there (xchg_kernels.cuh); 0 = not an exchange level
    uint64_t x_slice_bits;
};

enum BuildStatus : uint32_t {
    ST_RUNNING = 0,
    ST_DONE = 1,
    ST_ERR_MAX_ITERS = 2,  // level count exceeded MAX_ITERS (src/lib.rs:265-268, 398-401)
    ST_NEED_WORDS = 3,     // arena too small for the next level: host grows it and resumes
    ST_NEED_LIST = 4,      // redo list buffer too small for the next level: host grows it and resumes
    ST_ERR_INTERNAL = 5,
    ST_BUCKET_OVERFLOW = 6,  // a bucket region or a kernel's shared memory was too small: host rebuilds unbucketed
    ST_ERR_COMM = 7,         // sharded build: a peer did not arrive at a barrier in time, or reported an error
    ST_ERR_LIST_OVERFLOW = 8,  // sharded build: this shard's redo list outgrew its buffer (no host round trip there)
};

// device-resident build state; the level loop runs without host round trips (kernels for all
// expected levels are enqueued up front and read their sizes from here)
// Layout: the words that thousands of warps update or poll at the same time sit in 128-byte lines of their own.  They
// used to share the first line of the struct: the CTAs waiting at the cascade's grid barrier poll bar_gen, and every
// update of the append cursor by the CTAs still working queued behind those polls in the same L2 slice
// (profiles/r02_summary.md).
struct alignas(128) BuildState {
    uint32_t status;
    uint32_t n_levels;    // levels completed so far
    uint32_t tail_level;  // first level owned by the single-CTA tail kernel (0xffffffff: none yet)
    uint32_t pad0_;
    alignas(128) uint64_t redo_count;  // append cursor of the redo list being written by the current pass
    alignas(128) uint64_t uniq_count;  // popcount(a & ~collide) of the level being finalized
    uint32_t fin_ticket;               // "last block done" ticket of the finalize kernel
    alignas(128) uint32_t bar_count;   // grid barrier of the persistent cascade kernel: arrivals ...
    alignas(128) uint32_t bar_gen;     // ... and the generation the waiting CTAs poll
    alignas(128) uint64_t words_used;  // arena words handed out so far (multiple of 8)
    uint64_t words_cap;   // arena capacity in words (excluding the tail reserve)
    uint64_t list_cap[2]; // capacity (keys) of redo buffers: level l's key list lives in buf[l & 1]
    uint64_t seed0;       // starting_seed.unwrap_or(0)
    double gamma;
    uint64_t need;        // how much ST_NEED_* asks for
    uint64_t err_info;
    uint64_t bucket_min_keys;  // levels with fewer keys are built by the plain (L2-atomic) kernels
    uint32_t bucket_slots;     // CTAs that run concurrently (SM count x 2): bucket counts are multiples of it
    uint32_t bucket_enable;
    uint64_t bucket_keys_seen; // sum of bucket fills seen by the mark kernel (consistency check)
    uint32_t shard_rank;       // sharded build: this shard's index ...
    uint32_t shard_world;      // ... of shard_world shards (1 = unsharded)
    uint32_t gather_level;     // sharded: level at which all remaining keys are collected on every shard (host plan)
    uint32_t list_ready_level; // level whose complete key list already sits in buf[level & 1] (after the collect)
    uint64_t peer_count[SHARD_MAX];  // sharded: every shard's key count, as carried by the last list-sum barrier
    // Mphf::from_chunked_iterator_parallel (src/lib.rs:539-667): levels use the caller's gamma until fewer than
    // min(50 M, chunk_thr) keys are left; the levels after that one are built by new_parallel(1.7, ..) (:654)
    uint64_t chunk_thr;        // (0.01f32 * n as f32) as u64; only read when chunk_mode != 0
    uint64_t chunk_max_iters;  // max_iters of the chunked levels (chunk_has_max != 0)
    uint32_t chunk_mode, chunk_has_max;
    uint32_t switch_level;     // first level built with gamma 1.7 (NO_LEVEL: not reached yet)
    // striped key lists (build_kernels.cuh, pass_body): n_regions regions of reg_cap key slots in each list buffer,
    // reg_cnt[(l & 1) * n_regions + r] keys of level l's list in region r
    uint32_t n_regions;
    uint64_t reg_cap;
    uint32_t* reg_cnt;
    // sharded build: the first xchg_levels levels are exchange levels (host plan, identical on every shard);
    // xreg_cnt[(l & 1) * R * SHARD_MAX + r * world + o] = keys of level l that region r of this shard sent to owner o
    uint32_t xchg_levels;
    uint32_t x_sub_cap;
    uint32_t* xreg_cnt;
    KeyCfg kc;                 // how the keys' Hash impl feeds the hasher (u64 unless built through bphf_build_keys)
    double shard_frac;         // this shard's share of the keys (n_local / n_global; 1 unsharded): sizes bucket regions
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

// ---- sharded build: peer-visible memory of every shard ------------------------------------
// Each shard owns ONE peer-mapped allocation: [words arena | collide arena | tail exchange slot | flags].
// Passed to the kernels by value; pointer p of every array is shard p's copy (own pointers for p == rank).
constexpr uint32_t FLAG_STRIDE = 16;                      // u64 per writer: one 128-byte line each
// sharded build: once at most this many keys are left (over all shards) every shard collects all of them and the
// remaining levels run redundantly on every GPU with the unsharded kernels -- the many small levels would
// otherwise pay two device barriers and a merge each
constexpr uint64_t GATHER_MAX_KEYS = 1ull << 20;
constexpr uint32_t NO_LEVEL = 0xffffffffu;
constexpr uint64_t XKEYS_WORDS = GATHER_MAX_KEYS + 16;    // [0] = count (tail hand-over), [8..] = keys
struct CommDev {
    uint32_t rank, world;
    uint64_t* words[SHARD_MAX];   // level bit vectors (raw a while marking, final a after the merge)
    uint64_t* coll[SHARD_MAX];    // collision bit vectors
    uint64_t* xkeys[SHARD_MAX];   // keys this shard hands to the tail
    uint64_t* flags[SHARD_MAX];   // flags[p][r * FLAG_STRIDE + {0: epoch, 1: payload, 2: status, 3: payload2}] written by shard r
    // exchange levels (xchg_kernels.cuh): [sender s][region r][x_sub_cap] blocks
    uint32_t* xidx[SHARD_MAX];    // xidx[p]: slot indices (relative to p's slice) that the shards sent to owner p
    uint32_t* xcnt[SHARD_MAX];    // xcnt[p][s * R + r]: how many of them block (s, r) holds
    uint32_t* xbits[SHARD_MAX];   // xbits[p]: survivor bits that the owners returned to sender p, [owner o][region r][x_sub_cap / 32]
    uint32_t x_sub_cap;           // entries per (sender, region, owner) block; multiple of 256
    uint32_t x_regions;           // R
    uint64_t barrier_timeout_ns;  // how long shard_barrier_kernel waits for a peer before it reports ST_ERR_COMM
};

// ---- bucket geometry --------------------------------------------------------------------------------
// A bucketed level is cut into b_count ranges of 2^b_shift bits (power of two: bucket = idx >> shift, slot within
// the bucket = idx & (range - 1); a multiple of 512 bits, so a bucket owns whole rank blocks and 64-byte lines).
constexpr uint32_t BUCKET_SHIFT_MAX = 18;    // 2^18 bits: a + collide = 64 KB of shared memory -> 3 CTAs / SM
constexpr uint32_t BUCKET_SHIFT_MIN = 10;
constexpr uint32_t BUCKET_R_MAX = 1u << BUCKET_SHIFT_MAX;
constexpr uint32_t BUCKET_MAX_COUNT = 8192;  // per-tile histogram / base tables live in shared memory
constexpr uint32_t MAX_BUCKETS_ALLOC = 8192; // buckets per parity in the cursor array
// every bucket cursor sits in its own 32-byte sector: the per-tile reservations (one atomicAdd per bucket per
// tile) would otherwise pile onto a handful of L2 lines
constexpr uint32_t CURSOR_STRIDE = 8;

struct BucketGeom {
    uint32_t shift, range, count, cap;
};

// count ~ 2..4 x the concurrent CTA slots so every SM stays busy and the tail of the last wave is short
// frac: share of the level's k keys that this shard holds (bucket regions only hold the shard's own keys)
__host__ __device__ inline bool bucket_geometry(uint64_t bits, uint64_t k, uint32_t slots, uint64_t min_keys,
                                               double frac, BucketGeom* g) {
    if (slots == 0 || k < min_keys || (bits >> 32) != 0) return false;
    const uint64_t target = 2ull * slots;
    uint32_t shift = BUCKET_SHIFT_MAX;
    while (shift > BUCKET_SHIFT_MIN && (bits >> shift) < target) shift--;
    const uint64_t range = 1ull << shift;
    const uint64_t count = (bits + range - 1) >> shift;
    if (count > BUCKET_MAX_COUNT) return false;
    // expected fill k * range / bits, + 6 sigma + slack
    const double mean = (double)k * (frac < 1.0 ? frac * 1.02 : 1.0) * (double)range / (double)bits;
#ifdef __CUDA_ARCH__
    const double sd = sqrt(mean);
#else
    const double sd = __builtin_sqrt(mean);
#endif
    uint64_t cap = (uint64_t)(mean + 6.0 * sd + 96.0);
    cap = (cap + 31) & ~31ULL;
    if (cap * count > 0xffffffffULL) return false;  // destinations are 32-bit key indices
    g->shift = shift;
    g->range = (uint32_t)range;
    g->count = (uint32_t)count;
    g->cap = (uint32_t)cap;
    return true;
}

// fill in lv[level] for k keys; returns false when the arena cannot hold it (need -> words required)
__host__ __device__ inline void make_level(BuildState* st, uint32_t level, uint64_t k) {
    LevelDesc& d = st->lv[level];
    double gamma = st->gamma;
    if (st->chunk_mode) {
        if (level >= st->switch_level) gamma = 1.7;  // the tail of from_chunked_iterator_parallel, src/lib.rs:654
        else if (st->switch_level == NO_LEVEL && k < 50000000ULL && k < st->chunk_thr) st->switch_level = level + 1;
    }
    d.bits = level_size(gamma, k);
    d.n_words = (d.bits + 63) / 64;
    d.word_off = st->words_used;
    d.k_in = k;
    d.k_loc = (st->shard_world > 1 && level > 0) ? 0 : k;  // sharded: known once this shard's filter has run
    d.sp0 = level_spx(st->kc, st->seed0 + level);
    d.sp0_query = level_spx(st->kc, level);
    d.b_shift = 0;
    d.b_range = d.b_count = d.b_cap = 0;
    d.pad_ = 0;
    d.x_slice_bits = 0;
    if (level < st->xchg_levels) {  // (0 for unsharded builds)
        const uint64_t slice_words = round_up8((d.n_words + st->shard_world - 1) / st->shard_world);
        d.x_slice_bits = slice_words * 64;  // < 2^32 by the host's choice of xchg_levels: slots travel as 32-bit offsets
    }
}

}  // namespace bphf
