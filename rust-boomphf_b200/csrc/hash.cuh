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
enum { BPHF_KIND_U64 = 0, BPHF_KIND_UINT = 1 /* 1, 2 or 4 bytes */, BPHF_KIND_BYTES = 2 /* [u8; N], N <= 8 */ };

struct KeyCfg {
    uint64_t fin;     // total bytes written ^ P5
    uint32_t swap;    // 1: keys are raw u64 values, the kernels apply the 8-byte tail read; 0: keys are tail words
    uint32_t prefix;  // 1: an 8-byte length prefix (usize N) is written before the key bytes
    uint32_t width;   // key bytes, 1..8
    uint32_t kind;
};

__host__ __device__ inline KeyCfg make_key_cfg(uint32_t kind, uint32_t width) {
    KeyCfg c;
    c.kind = kind;
    c.width = kind == BPHF_KIND_U64 ? 8u : width;
    c.prefix = kind == BPHF_KIND_BYTES ? 1u : 0u;
    c.swap = kind == BPHF_KIND_U64 ? 1u : 0u;
    c.fin = (uint64_t)(c.width + (c.prefix ? 8u : 0u)) ^ WY_P5;
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

__device__ __forceinline__ uint32_t fold32(uint64_t v) { return (uint32_t)v ^ (uint32_t)(v >> 32); }

// hashmod for a level with n < 2^32 (Lemire fast range on the folded hash)
__device__ __forceinline__ uint32_t hashmod_small(const KeyCfg& c, uint64_t spx, uint64_t key, uint32_t n) {
    return __umulhi(fold32(hash64(c, spx, key)), n);
}
// hashmod for a level with n >= 2^32 (plain modulo of the 64-bit hash)
__device__ __forceinline__ uint64_t hashmod_big(const KeyCfg& c, uint64_t spx, uint64_t key, uint64_t n) {
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
