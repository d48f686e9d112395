// query_kernels.cuh -- batched Mphf::hash / try_hash and the BoomHashMap helpers.
//
// Reference functions replaced (file:line into /root/reference):
//   Mphf::get_rank   src/lib.rs:299-318
//   Mphf::hash       src/lib.rs:324-335      Mphf::try_hash  src/lib.rs:340-355
//   BoomHashMap::create_map  src/hashmap.rs:27-40   BoomHashMap::get  src/hashmap.rs:49-66
#pragma once
#include "common.cuh"

namespace bphf {

// per-level lookup record (device copy of the level table)
struct QLevel {
    uint64_t bits;     // BitVector.capacity()
    uint64_t bit_off;  // word_off * 64 (multiple of 512)
    uint64_t sp0;      // seed_xor_p0(level index)
};

constexpr int QUERY_THREADS = 256;

// walk the levels (src/lib.rs:325-332), first level whose bit is set wins; then get_rank:
//   ranks[idx/512] + popcount of the words of the same 512-bit block below idx/64
//   + popcount of the bits of the final word strictly below idx (skipped when idx % 64 == 0)
// With level offsets that are multiples of 512 bits, "idx/512 within the level" and "arena bit / 512"
// select the same rank entry and the same words.
__device__ __forceinline__ uint64_t lookup_one(const QLevel* __restrict__ lv, uint32_t n_levels,
                                               const uint64_t* __restrict__ words,
                                               const uint64_t* __restrict__ ranks, uint64_t key) {
    for (uint32_t l = 0; l < n_levels; l++) {
        const uint64_t bits = lv[l].bits;
        const uint64_t idx = (bits >> 32) ? hashmod_big(lv[l].sp0, key, bits)
                                          : (uint64_t)hashmod_small(lv[l].sp0, key, (uint32_t)bits);
        const uint64_t g = lv[l].bit_off + idx;
        const uint64_t wi = g >> 6;
        const uint64_t w = __ldg(words + wi);
        const uint32_t sh = (uint32_t)g & 63;
        if ((w >> sh) & 1ULL) {
            uint64_t rank = __ldg(ranks + (g >> 9));
            for (uint64_t j = wi & ~7ULL; j < wi; j++) rank += __popcll(__ldg(words + j));
            rank += __popcll(w & ((1ULL << sh) - 1ULL));
            return rank;
        }
    }
    return BPHF_NONE;
}

__global__ void __launch_bounds__(QUERY_THREADS)
    query_kernel(const QLevel* __restrict__ lv_g, uint32_t n_levels, const uint64_t* __restrict__ words,
                 const uint64_t* __restrict__ ranks, const uint64_t* __restrict__ keys, uint64_t n,
                 uint64_t* __restrict__ out) {
    __shared__ QLevel lv[BPHF_MAX_LEVELS + 1];
    for (uint32_t i = threadIdx.x; i < n_levels; i += QUERY_THREADS) lv[i] = lv_g[i];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * QUERY_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * QUERY_THREADS + threadIdx.x; i < n; i += stride) {
        out[i] = lookup_one(lv, n_levels, words, ranks, keys[i]);
    }
}

// create_map as an out-of-place scatter: keys_out[hash(k)] = k, values_out[hash(k)] = v.
// bad counts keys whose hash is None or >= n (a key outside the construction set)
__global__ void __launch_bounds__(QUERY_THREADS)
    map_permute_kernel(const QLevel* __restrict__ lv_g, uint32_t n_levels, const uint64_t* __restrict__ words,
                       const uint64_t* __restrict__ ranks, const uint64_t* __restrict__ keys,
                       const uint64_t* __restrict__ values, uint64_t n, uint64_t* __restrict__ keys_out,
                       uint64_t* __restrict__ values_out, unsigned long long* __restrict__ bad) {
    __shared__ QLevel lv[BPHF_MAX_LEVELS + 1];
    for (uint32_t i = threadIdx.x; i < n_levels; i += QUERY_THREADS) lv[i] = lv_g[i];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * QUERY_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * QUERY_THREADS + threadIdx.x; i < n; i += stride) {
        const uint64_t k = keys[i];
        const uint64_t pos = lookup_one(lv, n_levels, words, ranks, k);
        if (pos >= n) {
            atomicAdd(bad, 1ULL);
            continue;
        }
        keys_out[pos] = k;
        if (values) values_out[pos] = values[i];
    }
}

// BoomHashMap::get over a batch
__global__ void __launch_bounds__(QUERY_THREADS)
    map_get_kernel(const QLevel* __restrict__ lv_g, uint32_t n_levels, const uint64_t* __restrict__ words,
                   const uint64_t* __restrict__ ranks, const uint64_t* __restrict__ map_keys,
                   const uint64_t* __restrict__ map_values, uint64_t n_map, const uint64_t* __restrict__ queries,
                   uint64_t nq, uint64_t* __restrict__ values_out, uint8_t* __restrict__ found_out) {
    __shared__ QLevel lv[BPHF_MAX_LEVELS + 1];
    for (uint32_t i = threadIdx.x; i < n_levels; i += QUERY_THREADS) lv[i] = lv_g[i];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * QUERY_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * QUERY_THREADS + threadIdx.x; i < nq; i += stride) {
        const uint64_t q = queries[i];
        const uint64_t pos = lookup_one(lv, n_levels, words, ranks, q);
        const bool found = pos < n_map && __ldg(map_keys + pos) == q;
        values_out[i] = found ? (map_values ? __ldg(map_values + pos) : pos) : 0;
        found_out[i] = found ? 1 : 0;
    }
}

// ---- utilities ----------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64_fin(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

__global__ void keygen_kernel(uint64_t* __restrict__ out, uint64_t start, uint64_t n, uint64_t seed, int kind) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const uint64_t i = start + j;
        uint64_t k;
        if (kind == 0) {
            k = splitmix64_fin(seed + i * 0x9E3779B97F4A7C15ULL);
        } else if (kind == 1) {
            const uint64_t M = (1ULL << 62) - 1;
            uint64_t z = (i + seed) & M;
            z = (z * 0x9E3779B97F4A7C15ULL) & M;
            z ^= z >> 31;
            z = (z * 0xbf58476d1ce4e5b9ULL) & M;
            z ^= z >> 29;
            k = z;
        } else {
            k = 2 * i;
        }
        out[j] = k;
    }
}

// MPHF property check: every value < n and no value repeats (bitmap of n bits, pre-zeroed)
__global__ void check_perm_kernel(const uint64_t* __restrict__ vals, uint64_t n, uint32_t* __restrict__ bitmap,
                                  unsigned long long* __restrict__ bad) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t v = vals[i];
        if (v >= n) {
            local++;
            continue;
        }
        const uint32_t m = 1u << (v & 31);
        if (atomicOr(bitmap + (v >> 5), m) & m) local++;
    }
    if (local) atomicAdd(bad, local);
}

__global__ void checksum_kernel(const uint64_t* __restrict__ vals, uint64_t n, unsigned long long* __restrict__ acc) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t s = 0, x = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t h = splitmix64_fin(vals[i] + 0x9E3779B97F4A7C15ULL);
        s += h;
        x ^= h;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, d);
        x ^= __shfl_xor_sync(0xffffffffu, x, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(acc, (unsigned long long)s);
        atomicXor(acc + 1, (unsigned long long)x);
    }
}

}  // namespace bphf
