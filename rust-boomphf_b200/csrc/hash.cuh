// hash.cuh -- seeded hash + range reduction for u64 keys, bit-for-bit what the reference computes.
//
//   fold                 /root/reference/src/lib.rs:55-58
//   hash_with_seed       src/lib.rs:60-65  (wyhash::WyHash::with_seed(1 << (iter+iter)); u64::hash; finish)
//   fastmod / hashmod    src/lib.rs:72-88
//
// The arithmetic of wyhash itself lives in the third-party crate `wyhash` (v1 algorithm,
// crate versions 0.3-0.5).  For a u64 key the Hasher sees ONE 8-byte little-endian write:
//   h   = wymum(seed ^ P0, r(k) ^ P1)          (no 32-byte block; 8-byte tail case)
//   out = wymum(h, 8 ^ P5)                     (finish: total length 8)
// with r(k) = the wyhash-v1 "read64_swapped" of the 8 bytes = rotate-by-32 of k.
#pragma once
#include <stdint.h>

// single switch shared with the oracle (BO_WY_SWAP_8BYTE_TAIL): 1 = swapped 32-bit halves
#ifndef BPHF_WY_SWAP_8BYTE_TAIL
#define BPHF_WY_SWAP_8BYTE_TAIL 1
#endif

namespace bphf {

constexpr uint64_t WY_P0 = 0xa0761d6478bd642fULL;
constexpr uint64_t WY_P1 = 0xe7037ed1a0b428dbULL;
constexpr uint64_t WY_P2 = 0x8ebc6af09c88c6e3ULL;
constexpr uint64_t WY_P3 = 0x589965cc75374cc3ULL;
constexpr uint64_t WY_P4 = 0x1d8e4e27c47d124fULL;
constexpr uint64_t WY_P5 = 0xeb44accab455d165ULL;
constexpr uint64_t WY_FIN = 8ULL ^ WY_P5;  // finish(): total_len ^ P5, total_len = 8

// seed ^ P0 for seed index `si` (iter of the reference): 1 << ((si + si) & 63) -- the release-mode
// wrap of `1 << (iter + iter)` (src/lib.rs:62); debug builds panic for iter >= 32.
__host__ __device__ __forceinline__ uint64_t seed_xor_p0(uint64_t si) {
    return (1ULL << ((si + si) & 63)) ^ WY_P0;
}

__host__ __device__ __forceinline__ uint64_t wymum(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b) ^ (a * b);
#else
    const unsigned __int128 r = (unsigned __int128)a * (unsigned __int128)b;
    return (uint64_t)(r >> 64) ^ (uint64_t)r;
#endif
}

__host__ __device__ __forceinline__ uint64_t wy_read8(uint64_t k) {
#if BPHF_WY_SWAP_8BYTE_TAIL
    return (k << 32) | (k >> 32);
#else
    return k;
#endif
}

// ---- key kinds --------------------------------------------------------------------
// How `T::hash` feeds the hasher, for every T whose byte stream fits one tail word:
//   u64 / i64 / usize / isize : one 8-byte write                       (Hasher::write_u64 -> write(&bytes))
//   u32 / i32, u16 / i16, u8 / i8 : one 4 / 2 / 1-byte write
//   [u8; N], &[u8], Vec<u8> with N <= 8 : write_usize(N) (8 bytes), then one N-byte write  (impl Hash for [T])
// The wyhash Hasher runs the v1 core once per write, carrying its state, and finish() mixes in the total length:
//   single write of w <= 8 bytes :  wymum( wymum(seed ^ P0, tail(key) ^ P1), w ^ P5 )
//   length prefix + N bytes      :  h1 = wymum(seed ^ P0, tail8(N) ^ P1);  wymum( wymum(h1 ^ P0, tail(key) ^ P1), (8+N) ^ P5 )
// so every kind is  wymum( wymum(spx(level), word ^ P1), fin )  with a per-level constant spx, a per-build constant
// fin and the key's tail word (read_rest of the wyhash crate: e.g. 3 bytes -> (rd16 << 8) | byte2, 8 bytes -> the
// halves swapped).  u64 keys are swapped inside the kernels; every other kind is converted to tail words once.
//
// Every other T: Hash (u128, [u8; N > 8], Vec<u8>, String, Vec<String>, tuples, ... -- the remaining quickcheck types of
// the reference, src/lib.rs:851-879) is a STREAM key: the exact sequence of `Hasher::write` calls T::hash makes,
// recorded by the caller as byte segments (bytes + segment ends + segments per key).  The kernels then carry the key's
// INDEX in their key lists and hash64(StreamCfg) walks its segments: one v1 core per write, the state carried from
// write to write, finish(total bytes).  A zero-length write leaves the state alone (the crate's `write` loops over
// `bytes.chunks(..)`, which yields nothing for an empty slice; BPHF_WY_EMPTY_WRITE_IS_NOOP, same switch in the oracle).
// Stream keys run in their own instantiations of the level and lookup kernels (template parameter KC = StreamCfg), so
// the u64 kernels carry neither the branch nor the registers.
enum {
    BPHF_KIND_U64 = 0,
    BPHF_KIND_UINT = 1 /* 1, 2 or 4 bytes */,
    BPHF_KIND_BYTES = 2 /* [u8; N], N <= 8 */,
    BPHF_KIND_STREAM = 3 /* any Hash byte stream; keys are indices into the segment tables below */
};
#ifndef BPHF_WY_EMPTY_WRITE_IS_NOOP
#define BPHF_WY_EMPTY_WRITE_IS_NOOP 1
#endif

struct KeyCfg {
    uint64_t fin;     // total bytes written ^ P5
    uint32_t swap;    // 1: keys are raw u64 values, the kernels apply the 8-byte tail read; 0: keys are tail words
    uint32_t prefix;  // 1: an 8-byte length prefix (usize N) is written before the key bytes
    uint32_t width;   // key bytes, 1..8
    uint32_t kind;
    // BPHF_KIND_STREAM only (device pointers, valid for the duration of one build / one lookup batch)
    const uint8_t* s_bytes;      // all write segments back to back
    const uint64_t* s_seg_ends;  // end offset (exclusive) of segment j in s_bytes; segment j starts at s_seg_ends[j-1]
    const uint64_t* s_key_segs;  // key i owns segments [s_key_segs[i], s_key_segs[i+1])
};
// the same configuration, typed so that overload resolution picks the stream hash (kernels take KC as a template type)
struct StreamCfg : KeyCfg {
    __host__ __device__ StreamCfg() {}
    __host__ __device__ StreamCfg(const KeyCfg& c) : KeyCfg(c) {}
};

__host__ __device__ inline KeyCfg make_key_cfg(uint32_t kind, uint32_t width) {
    KeyCfg c;
    c.kind = kind;
    c.width = kind == BPHF_KIND_U64 ? 8u : width;
    c.prefix = kind == BPHF_KIND_BYTES ? 1u : 0u;
    c.swap = kind == BPHF_KIND_U64 ? 1u : 0u;
    c.fin = (uint64_t)(c.width + (c.prefix ? 8u : 0u)) ^ WY_P5;
    c.s_bytes = nullptr;
    c.s_seg_ends = nullptr;
    c.s_key_segs = nullptr;
    return c;
}

// first operand of the key's wymum for seed index `si` (iter of the reference)
__host__ __device__ inline uint64_t level_spx(const KeyCfg& c, uint64_t si) {
    const uint64_t sp0 = seed_xor_p0(si);
    if (!c.prefix) return sp0;
    return wymum(sp0, wy_read8((uint64_t)c.width) ^ WY_P1) ^ WY_P0;  // after write_usize(N)
}

// read_rest of the wyhash crate for a key of `width` bytes held little-endian in the low bytes of k
__host__ __device__ inline uint64_t key_tail_word(uint64_t k, uint32_t width) {
    const uint64_t b0 = k & 0xff, r16 = k & 0xffff, r32 = k & 0xffffffffULL;
    switch (width) {
        case 1: return b0;
        case 2: return r16;
        case 3: return (r16 << 8) | ((k >> 16) & 0xff);
        case 4: return r32;
        case 5: return (r32 << 8) | ((k >> 32) & 0xff);
        case 6: return (r32 << 16) | ((k >> 32) & 0xffff);
        case 7: return (r32 << 24) | (((k >> 32) & 0xffff) << 8) | ((k >> 48) & 0xff);
        default: return wy_read8(k);
    }
}

// hash_with_seed(iter, &key) given spx = level_spx(cfg, iter)
__device__ __forceinline__ uint64_t hash64(const KeyCfg& c, uint64_t spx, uint64_t key) {
    const uint64_t w = c.swap ? wy_read8(key) : key;
    return wymum(wymum(spx, w ^ WY_P1), c.fin);
}

// ---- stream keys: the full v1 core over global-memory bytes ---------------------------------------------
__device__ __forceinline__ uint64_t wy_ld(const uint8_t* p, uint32_t nbytes) {  // little-endian value of <= 8 bytes
    uint64_t v = 0;
    for (uint32_t i = 0; i < nbytes; i++) v |= (uint64_t)__ldg(p + i) << (8 * i);
    return v;
}
__device__ __forceinline__ uint64_t wy_ld8(const uint8_t* p) {
    if (((uintptr_t)p & 7) == 0) return __ldg(reinterpret_cast<const unsigned long long*>(p));
    return wy_ld(p, 8);
}
__device__ __forceinline__ uint64_t wy_ld8_swapped(const uint8_t* p) {  // read64_swapped: unconditional in the tails
    const uint64_t v = wy_ld8(p);
    return (v << 32) | (v >> 32);
}
// one Hasher::write of `len` bytes carrying the state `seed` (wyhash v1 core: 32-byte blocks, then the 1..31-byte tail)
__device__ inline uint64_t wy_core(const uint8_t* p, uint64_t len, uint64_t seed) {
    for (uint64_t i = 0; i + 32 <= len; i += 32, p += 32)
        seed = wymum(seed ^ WY_P0, wymum(wy_ld8(p) ^ WY_P1, wy_ld8(p + 8) ^ WY_P2) ^
                                       wymum(wy_ld8(p + 16) ^ WY_P3, wy_ld8(p + 24) ^ WY_P4));
    seed ^= WY_P0;
    const uint32_t rest = (uint32_t)(len & 31);
    if (rest) {
        const uint32_t q = (rest - 1) >> 3, r = rest - 8 * q;  // r = 1..8 bytes in the last word
        const uint64_t last = key_tail_word(wy_ld(p + 8 * q, r), r);
        if (q == 0) seed = wymum(seed, last ^ WY_P1);
        else if (q == 1) seed = wymum(wy_ld8_swapped(p) ^ seed, last ^ WY_P2);
        else if (q == 2) seed = wymum(wy_ld8_swapped(p) ^ seed, wy_ld8_swapped(p + 8) ^ WY_P2) ^ wymum(seed, last ^ WY_P3);
        else seed = wymum(wy_ld8_swapped(p) ^ seed, wy_ld8_swapped(p + 8) ^ WY_P2) ^
                    wymum(wy_ld8_swapped(p + 16) ^ seed, last ^ WY_P4);
    }
    return seed;
}
// hash_with_seed(iter, &key) for stream key number `idx`; spx = seed ^ P0 (level_spx of a stream cfg)
__device__ __noinline__ uint64_t hash64(const StreamCfg& c, uint64_t spx, uint64_t idx) {
    uint64_t h = spx ^ WY_P0;
    const uint64_t s0 = __ldg(c.s_key_segs + idx), s1 = __ldg(c.s_key_segs + idx + 1);
    const uint64_t start = s0 ? __ldg(c.s_seg_ends + s0 - 1) : 0;
    uint64_t pos = start;
    for (uint64_t sg = s0; sg < s1; sg++) {
        const uint64_t end = __ldg(c.s_seg_ends + sg);
#if BPHF_WY_EMPTY_WRITE_IS_NOOP
        if (end != pos)
#endif
            h = wy_core(c.s_bytes + pos, end - pos, h);
        pos = end;
    }
    return wymum(h, (pos - start) ^ WY_P5);
}

__device__ __forceinline__ uint32_t fold32(uint64_t v) { return (uint32_t)v ^ (uint32_t)(v >> 32); }

// hashmod for a level with n < 2^32 (Lemire fast range on the folded hash)
template <class KC>
__device__ __forceinline__ uint32_t hashmod_small(const KC& c, uint64_t spx, uint64_t key, uint32_t n) {
    return __umulhi(fold32(hash64(c, spx, key)), n);
}
// hashmod for a level with n >= 2^32 (plain modulo of the 64-bit hash)
template <class KC>
__device__ __forceinline__ uint64_t hashmod_big(const KC& c, uint64_t spx, uint64_t key, uint64_t n) {
    return hash64(c, spx, key) % n;
}

// level size: max(255, (gamma * k as f64) as u64)      src/lib.rs:241,256,369,384
// (u64 -> f64 rounds to nearest, product rounds to nearest, -> u64 truncates; gamma*k < 2^63 here)
__host__ __device__ __forceinline__ uint64_t level_size(double gamma, uint64_t k) {
#ifdef __CUDA_ARCH__
    double s = __dmul_rn(gamma, __ull2double_rn(k));
    uint64_t u = (s >= 18446744073709551616.0) ? 0xFFFFFFFFFFFFFFFFULL : __double2ull_rz(s);
#else
    volatile double s = gamma * (double)k;
    uint64_t u = !(s > 0.0) ? 0 : (s >= 18446744073709551616.0 ? 0xFFFFFFFFFFFFFFFFULL : (uint64_t)s);
#endif
    return u < 255 ? 255 : u;
}

}  // namespace bphf
