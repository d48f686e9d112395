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

__device__ __forceinline__ uint64_t wymum(uint64_t a, uint64_t b) { return __umul64hi(a, b) ^ (a * b); }

__device__ __forceinline__ uint64_t wy_read8(uint64_t k) {
#if BPHF_WY_SWAP_8BYTE_TAIL
    return (k << 32) | (k >> 32);
#else
    return k;
#endif
}

// hash_with_seed(iter, &key) given sp0 = seed_xor_p0(iter)
__device__ __forceinline__ uint64_t hash64(uint64_t sp0, uint64_t key) {
    return wymum(wymum(sp0, wy_read8(key) ^ WY_P1), WY_FIN);
}

__device__ __forceinline__ uint32_t fold32(uint64_t v) { return (uint32_t)v ^ (uint32_t)(v >> 32); }

// hashmod for a level with n < 2^32 (Lemire fast range on the folded hash)
__device__ __forceinline__ uint32_t hashmod_small(uint64_t sp0, uint64_t key, uint32_t n) {
    return __umulhi(fold32(hash64(sp0, key)), n);
}
// hashmod for a level with n >= 2^32 (plain modulo of the 64-bit hash)
__device__ __forceinline__ uint64_t hashmod_big(uint64_t sp0, uint64_t key, uint64_t n) {
    return hash64(sp0, key) % n;
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
