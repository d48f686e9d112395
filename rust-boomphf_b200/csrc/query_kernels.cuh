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
template <class KC>
__device__ __forceinline__ uint64_t lookup_one(const KC& kc, const QLevel* __restrict__ lv, uint32_t n_levels,
                                               const uint64_t* __restrict__ words,
                                               const uint64_t* __restrict__ ranks, uint64_t key) {
    for (uint32_t l = 0; l < n_levels; l++) {
        const uint64_t bits = lv[l].bits;
        const uint64_t idx = (bits >> 32) ? hashmod_big(kc, lv[l].sp0, key, bits)
                                          : (uint64_t)hashmod_small(kc, lv[l].sp0, key, (uint32_t)bits);
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

// ---------------------------------------------------------------------------------------------
// Packed lookup structure.  Random 32-byte-sector accesses are what bounds a lookup on B200 (one sector per clock
// and SM through the L1 tag stage, measured): the reference layout costs a word probe per visited level plus, on
// the hit, a rank entry and up to 7 more words (src/lib.rs:299-318) -- 6.5 sectors per key at gamma = 1.7.
// Here every level is re-packed into 16-byte granules {96 bits of the level, u32 rank of the granule's first bit
// relative to the level} so that ONE 16-byte load both tests the bit and yields everything get_rank needs:
//   rank = level_base + granule.rank + popc(full words below) + popc(partial word)
// which is the same number the reference computes (same bits, same running popcount).  Built on the device
// right after compute_ranks; used when every level has fewer than 2^32 bits (else the generic kernel runs).
// ---------------------------------------------------------------------------------------------
struct QLevelP {
    uint64_t sp0;        // seed_xor_p0(level index)
    uint64_t base_rank;  // ranks[first block of the level]: set bits of all previous levels
    uint64_t gran_off;   // first granule of the level in the packed array
    uint32_t bits;       // BitVector.capacity() (< 2^32)
    uint32_t n_gran;     // ceil(bits / 96)
};
constexpr uint32_t GRAN_BITS = 96;

__global__ void set_base_rank_kernel(QLevelP* __restrict__ lvp, const QLevel* __restrict__ lv, uint32_t n_levels,
                                     const uint64_t* __restrict__ ranks) {
    for (uint32_t l = threadIdx.x; l < n_levels; l += blockDim.x) lvp[l].base_rank = ranks[lv[l].bit_off / 512];
}

// one thread per granule: 96-bit window of the level's words + popcount of the level below it
__global__ void __launch_bounds__(256)
    pack_query_kernel(const QLevelP* __restrict__ lvp, const QLevel* __restrict__ lv, uint32_t n_levels,
                      const uint64_t* __restrict__ words, const uint64_t* __restrict__ ranks, uint64_t total_gran,
                      uint4* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_gran; g += stride) {
        // level of this granule (few levels: linear search from the top is cheap, the big levels come first)
        uint32_t l = 0;
        while (l + 1 < n_levels && g >= lvp[l + 1].gran_off) l++;
        const uint64_t gi = g - lvp[l].gran_off;
        const uint64_t s = gi * GRAN_BITS;                 // first bit of the granule within the level
        const uint64_t w0 = lv[l].bit_off / 64;            // first arena word of the level (multiple of 8)
        const uint64_t n_words = (lv[l].bits + 63) / 64;
        auto word = [&](uint64_t i) -> uint64_t { return i < n_words ? words[w0 + i] : 0ULL; };
        const uint64_t wi = s >> 6;
        const uint32_t sh = (uint32_t)s & 63;
        const uint64_t a = word(wi), b = word(wi + 1), c = word(wi + 2);
        uint64_t lo = sh ? ((a >> sh) | (b << (64 - sh))) : a;
        uint64_t hi = sh ? ((b >> sh) | (c << (64 - sh))) : b;
        // rank of bit s relative to the level: block rank + words of the block below wi + bits of word wi below sh
        uint64_t r = ranks[(w0 + wi) / 8] - lvp[l].base_rank;
        for (uint64_t j = wi & ~7ULL; j < wi; j++) r += __popcll(word(j));
        if (sh) r += __popcll(a & ((1ULL << sh) - 1ULL));
        // bits at or beyond the level's capacity are zero already (the build never sets them)
        out[g] = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)r);
    }
}

template <class KC>
__device__ __forceinline__ uint64_t lookup_packed(const KC& kc, const QLevelP* __restrict__ lv, uint32_t n_levels,
                                                  const uint4* __restrict__ q, uint64_t key) {
    for (uint32_t l = 0; l < n_levels; l++) {
        const uint32_t idx = hashmod_small(kc, lv[l].sp0, key, lv[l].bits);
        const uint32_t g = __umulhi(idx >> 5, 0x55555556u);  // (idx / 32) / 3, exact below 2^31
        const uint32_t bit = idx - g * GRAN_BITS;            // 0..95
        const uint4 v = __ldg(q + lv[l].gran_off + g);
        const uint32_t w = bit >> 5;
        const uint32_t word = w == 0 ? v.x : (w == 1 ? v.y : v.z);
        const uint32_t sh = bit & 31;
        if ((word >> sh) & 1u) {
            uint32_t below = __popc(word & ((1u << sh) - 1u));
            if (w > 0) below += __popc(v.x);
            if (w > 1) below += __popc(v.y);
            return lv[l].base_rank + v.w + below;
        }
    }
    return BPHF_NONE;
}

// one probe of the packed layout: does `key` own a bit in level `lv`?  rank is valid on a hit
struct PackedProbe {
    uint32_t g, bit;
};
template <class KC>
__device__ __forceinline__ PackedProbe packed_slot(const KC& kc, const QLevelP& lv, uint64_t key) {
    const uint32_t idx = hashmod_small(kc, lv.sp0, key, lv.bits);
    PackedProbe p;
    p.g = __umulhi(idx >> 5, 0x55555556u);
    p.bit = idx - p.g * GRAN_BITS;
    return p;
}
__device__ __forceinline__ bool packed_hit(const QLevelP& lv, const uint4 v, uint32_t bit, uint64_t* rank) {
    const uint32_t w = bit >> 5;
    const uint32_t word = w == 0 ? v.x : (w == 1 ? v.y : v.z);
    const uint32_t sh = bit & 31;
    uint32_t below = __popc(word & ((1u << sh) - 1u));
    if (w > 0) below += __popc(v.x);
    if (w > 1) below += __popc(v.y);
    *rank = lv.base_rank + v.w + below;
    return (word >> sh) & 1u;
}

// handle-side view of either layout
struct QView {
    const QLevel* lv;
    const QLevelP* lvp;  // non-null: packed layout available
    const uint64_t* words;
    const uint64_t* ranks;
    const uint4* packed;
    uint32_t n_levels;
    KeyCfg kc;
};

// level table in shared memory + the lookup of either layout
template <bool PACKED>
struct LevelTab;
template <>
struct LevelTab<false> {
    QLevel lv[BPHF_MAX_LEVELS + 1];
    __device__ __forceinline__ void load(const QView& v) {
        for (uint32_t i = threadIdx.x; i < v.n_levels; i += blockDim.x) lv[i] = v.lv[i];
    }
    template <class KC>
    __device__ __forceinline__ uint64_t lookup(const QView& v, const KC& kc, uint64_t key) const {
        return lookup_one(kc, lv, v.n_levels, v.words, v.ranks, key);
    }
};
template <>
struct LevelTab<true> {
    QLevelP lv[BPHF_MAX_LEVELS + 1];
    __device__ __forceinline__ void load(const QView& v) {
        for (uint32_t i = threadIdx.x; i < v.n_levels; i += blockDim.x) lv[i] = v.lvp[i];
    }
    template <class KC>
    __device__ __forceinline__ uint64_t lookup(const QView& v, const KC& kc, uint64_t key) const {
        return lookup_packed(kc, lv, v.n_levels, v.packed, key);
    }
};

// Mphf::hash / try_hash over a batch (src/lib.rs:324-355)
// KC = StreamCfg: `keys` may be null, the keys are then the stream indices 0..n-1 themselves
template <bool PACKED, class KC = KeyCfg>
__global__ void __launch_bounds__(QUERY_THREADS)
    query_kernel(QView v, const uint64_t* __restrict__ keys, uint64_t n, uint64_t* __restrict__ out) {
    __shared__ LevelTab<PACKED> tab;
    tab.load(v);
    __syncthreads();
    const KC kc(v.kc);
    const uint64_t stride = (uint64_t)gridDim.x * QUERY_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * QUERY_THREADS + threadIdx.x; i < n; i += stride)
        out[i] = tab.lookup(v, kc, keys ? keys[i] : i);
}

// ---------------------------------------------------------------------------------------------
// query_sync_kernel: the level walk of Mphf::hash (src/lib.rs:325-332) run level-synchronously per WARP.
// A thread-per-key walk leaves ~11 of 32 lanes active (a warp lasts as long as its deepest key, and the hash is
// re-evaluated per level).  Here a warp takes a tile of QS_WTILE keys, probes level 0 for all of them (independent
// loads), and the keys that missed are compacted into a per-warp index queue in shared memory; round l probes
// level l for the queue with full lanes.  The last few stragglers walk the remaining levels one per lane.
// No block barrier anywhere: warps drift apart and hide each other's latency.  Results are staged in shared
// memory and written coalesced.
// ---------------------------------------------------------------------------------------------
constexpr int QS_THREADS = 128;
constexpr int QS_WARPS = QS_THREADS / 32;
constexpr int QS_KPL = 8;                    // keys per lane per tile
constexpr int QS_WTILE = 32 * QS_KPL;        // 256 keys per warp tile: 2 KB of results + 2 x 512 B of queues
// (larger tiles amortise the short rounds better but starve L1: with ~220 KB of shared memory per SM the random
// 16-byte loads have no room to stay in flight and the kernel runs 2.4x slower -- measured)
constexpr int QS_TILE = QS_WTILE;            // (host: tiles are per warp)
constexpr uint32_t QS_STRAGGLERS = 8;        // at most this many keys left: one lane each walks the rest

// append lane's entry to the warp's queue; returns the new count (uniform)
__device__ __forceinline__ uint32_t wqueue_append(bool miss, uint32_t i, uint16_t* __restrict__ q, uint32_t count) {
    const uint32_t m = __ballot_sync(0xffffffffu, miss);
    const int lane = threadIdx.x & 31;
    if (miss) q[count + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
    return count + __popc(m);
}

// the 256-key tiles of a dense key array
struct QSDenseTiles {
    uint64_t n;
    __device__ __forceinline__ uint64_t count() const { return (n + QS_WTILE - 1) / QS_WTILE; }
    __device__ __forceinline__ void locate(uint64_t tile, uint64_t* base, uint32_t* t_n) const {
        *base = tile * QS_WTILE;
        *t_n = (uint32_t)min((uint64_t)QS_WTILE, n - *base);
    }
};
// SUB = false: Mphf::hash of the keys from level 0, ranks (u64, BPHF_NONE) to out.
// SUB = true: the last stage of the partitioned lookup (qpart_kernels.cuh): the keys reached level `level0` unresolved,
// the walk starts there and the ranks go to out32 (0xffffffff = None: the caller made sure every rank is smaller).
// Tiles: where tile t of the key list starts (the same offset applies to the output) and how many keys it holds.
template <bool SUB, class Tiles>
__device__ __forceinline__ void query_sync_body(const QView& v, Tiles tiles, const uint64_t* __restrict__ keys,
                                                uint64_t* __restrict__ out, uint32_t* __restrict__ out32, uint32_t level0) {
    __shared__ QLevelP lv[BPHF_MAX_LEVELS + 1];
    __shared__ uint64_t res_all[QS_WARPS][QS_WTILE];      // key while unresolved, rank (or NONE) once resolved
    __shared__ uint16_t queue_all[QS_WARPS][2][QS_WTILE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t i = tid; i < v.n_levels; i += QS_THREADS) lv[i] = v.lvp[i];
    __syncthreads();
    uint64_t* __restrict__ res = res_all[warp];
    const uint4* __restrict__ q = v.packed;
    const KeyCfg kc = v.kc;
    const uint32_t n_levels = v.n_levels;
    const uint32_t l0 = SUB ? level0 : 0u;  // first level of the walk
    const uint64_t n_tiles = tiles.count();
    const uint64_t w_stride = (uint64_t)gridDim.x * QS_WARPS;
    for (uint64_t tile = (uint64_t)blockIdx.x * QS_WARPS + warp; tile < n_tiles; tile += w_stride) {
        uint64_t base;
        uint32_t t_n;
        tiles.locate(tile, &base, &t_n);
        uint32_t n_q = 0;
        // ---- round 0: every key of the tile against level 0 (batches of 4 independent probes per lane)
#pragma unroll
        for (int h = 0; h < QS_KPL; h += 4) {
            uint64_t key[4];
            PackedProbe pp[4];
            uint4 g[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t i = (h + j) * 32 + lane;
                key[j] = i < t_n ? keys[base + i] : 0;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                pp[j] = packed_slot(kc, lv[l0], key[j]);
                g[j] = __ldg(q + lv[l0].gran_off + pp[j].g);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t i = (h + j) * 32 + lane;
                uint64_t rank;
                const bool hit = packed_hit(lv[l0], g[j], pp[j].bit, &rank);
                res[i] = hit ? rank : key[j];
                n_q = wqueue_append(!hit && i < t_n, i, queue_all[warp][1], n_q);
            }
        }
        __syncwarp();
        // ---- round r: the queue against level r, two entries per lane in flight
        uint32_t r = l0 + 1;
        for (; r < n_levels && n_q > QS_STRAGGLERS; r++) {
            const uint16_t* __restrict__ qin = queue_all[warp][(r - l0) & 1];  // (round 0 filled queue 1)
            uint16_t* __restrict__ qout = queue_all[warp][(r - l0 + 1) & 1];
            const QLevelP L = lv[r];
            uint32_t n_out = 0;
            for (uint32_t t0 = 0; t0 < n_q; t0 += 64) {
                uint32_t i[2];
                uint64_t key[2];
                PackedProbe pp[2];
                uint4 g[2];
                bool live[2];
                const int nb = (t0 + 32 < n_q) ? 2 : 1;  // uniform: skip the second batch when it is empty
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    if (j < nb) {
                        const uint32_t t = t0 + j * 32 + lane;
                        live[j] = t < n_q;
                        i[j] = live[j] ? qin[t] : 0;
                        key[j] = res[i[j]];
                        pp[j] = packed_slot(kc, L, key[j]);
                        g[j] = __ldg(q + L.gran_off + pp[j].g);
                    }
                }
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    if (j >= nb) break;
                    uint64_t rank;
                    const bool hit = packed_hit(L, g[j], pp[j].bit, &rank);
                    if (live[j] && hit) res[i[j]] = rank;
                    n_out = wqueue_append(live[j] && !hit, i[j], qout, n_out);
                }
            }
            n_q = n_out;
            __syncwarp();
        }
        // ---- stragglers: one lane per key walks the remaining levels; None when no level claims the key.
        // The round loop also ends when the levels run out (r == n_levels) with ANY number of keys still queued
        // (absent keys against a shallow MPHF): every queued entry is drained, not just the first 32.
        for (uint32_t t = lane; t < n_q; t += 32) {
            const uint32_t i = queue_all[warp][(r - l0) & 1][t];
            const uint64_t key = res[i];
            uint64_t rank = BPHF_NONE;
            for (uint32_t l = r; l < n_levels; l++) {
                const PackedProbe pp = packed_slot(kc, lv[l], key);
                const uint4 g = __ldg(q + lv[l].gran_off + pp.g);
                uint64_t rk;
                if (packed_hit(lv[l], g, pp.bit, &rk)) {
                    rank = rk;
                    break;
                }
            }
            res[i] = rank;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < QS_KPL; j++) {
            const uint32_t i = j * 32 + lane;
            if (i < t_n) {
                if (SUB) __stcs(out32 + base + i, (uint32_t)res[i]);  // BPHF_NONE truncates to 0xffffffff
                else out[base + i] = res[i];
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(QS_THREADS, 8)
    query_sync_kernel(QView v, const uint64_t* __restrict__ keys, uint64_t n, uint64_t* __restrict__ out) {
    query_sync_body<false>(v, QSDenseTiles{n}, keys, out, nullptr, 0);
}

// create_map as an out-of-place scatter: keys_out[hash(k)] = k, values_out[hash(k)] = v.
// bad counts keys whose hash is None or >= n (a key outside the construction set)
template <bool PACKED>
__global__ void __launch_bounds__(QUERY_THREADS)
    map_permute_kernel(QView v, const uint64_t* __restrict__ keys, const uint64_t* __restrict__ values, uint64_t n,
                       uint64_t* __restrict__ keys_out, uint64_t* __restrict__ values_out,
                       unsigned long long* __restrict__ bad) {
    __shared__ LevelTab<PACKED> tab;
    tab.load(v);
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * QUERY_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * QUERY_THREADS + threadIdx.x; i < n; i += stride) {
        const uint64_t k = keys[i];
        const uint64_t pos = tab.lookup(v, v.kc, k);
        if (pos >= n) {
            atomicAdd(bad, 1ULL);
            continue;
        }
        keys_out[pos] = k;
        if (values) values_out[pos] = values[i];
    }
}

// BoomHashMap::get over a batch
template <bool PACKED>
__global__ void __launch_bounds__(QUERY_THREADS)
    map_get_kernel(QView v, const uint64_t* __restrict__ map_keys, const uint64_t* __restrict__ map_values,
                   uint64_t n_map, const uint64_t* __restrict__ queries, uint64_t nq, uint64_t* __restrict__ values_out,
                   uint8_t* __restrict__ found_out) {
    __shared__ LevelTab<PACKED> tab;
    tab.load(v);
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * QUERY_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * QUERY_THREADS + threadIdx.x; i < nq; i += stride) {
        const uint64_t q = queries[i];
        const uint64_t pos = tab.lookup(v, v.kc, q);
        const bool found = pos < n_map && __ldg(map_keys + pos) == q;
        values_out[i] = found ? (map_values ? __ldg(map_values + pos) : pos) : 0;
        found_out[i] = found ? 1 : 0;
    }
}

// second half of create_map / get once the ranks are known (first half: the batched lookup kernel)
__global__ void __launch_bounds__(256)
    map_scatter_kernel(const uint64_t* __restrict__ pos, const uint64_t* __restrict__ keys, const uint64_t* __restrict__ values,
                       uint64_t n, uint64_t n_total, uint64_t* __restrict__ keys_out, uint64_t* __restrict__ values_out,
                       unsigned long long* __restrict__ bad) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t p = pos[i];
        if (p >= n_total) {
            atomicAdd(bad, 1ULL);
            continue;
        }
        keys_out[p] = keys[i];
        if (values) values_out[p] = values[i];
    }
}
__global__ void __launch_bounds__(256)
    map_gather_kernel(const uint64_t* __restrict__ pos, const uint64_t* __restrict__ map_keys,
                      const uint64_t* __restrict__ map_values, uint64_t n_map, const uint64_t* __restrict__ queries,
                      uint64_t nq, uint64_t* __restrict__ values_out, uint8_t* __restrict__ found_out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) {
        const uint64_t p = pos[i];
        const bool found = p < n_map && __ldg(map_keys + p) == queries[i];
        values_out[i] = found ? (map_values ? __ldg(map_values + p) : p) : 0;
        found_out[i] = found ? 1 : 0;
    }
}

// stream keys (hash.cuh): the level-0 key list is the index of every key; and a structural check of the caller's
// segment tables before any kernel dereferences them (bad[0] = number of violations)
__global__ void __launch_bounds__(256) iota_kernel(uint64_t* __restrict__ out, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = i;
}
__global__ void __launch_bounds__(256)
    stream_check_kernel(const uint64_t* __restrict__ seg_ends, uint64_t n_segs, const uint64_t* __restrict__ key_segs,
                        uint64_t n_keys, uint64_t n_bytes, uint32_t* __restrict__ bad) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = 0;
    for (uint64_t j = t; j < n_segs; j += stride) {
        const uint64_t e = seg_ends[j];
        if (e > n_bytes || (j && seg_ends[j - 1] > e)) v++;
    }
    for (uint64_t i = t; i < n_keys; i += stride)
        if (key_segs[i] > key_segs[i + 1] || key_segs[i + 1] > n_segs) v++;
    if (t == 0 && (key_segs[0] != 0 || key_segs[n_keys] != n_segs)) v++;
    if (v) atomicAdd(bad, v);
}

// packed keys of `width` bytes (u32 / u16 / u8 values, [u8; N]) -> tail words, hash.cuh: key_tail_word
__global__ void key_words_kernel(const uint8_t* __restrict__ in, uint32_t width, uint64_t n, uint64_t* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t k = 0;
        if (width == 4) {
            k = reinterpret_cast<const uint32_t*>(in)[i];
        } else {
            for (uint32_t b = 0; b < width; b++) k |= (uint64_t)in[i * width + b] << (8 * b);
        }
        out[i] = key_tail_word(k, width);
    }
}

// ---- utilities ----------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64_fin(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

// "L2-atomic ceiling" probe (north_star): n_ops random 4-byte operations on an array of n_words words, 8 in flight per
// thread -- kind 0: fire-and-forget OR (REDG), 1: returning fetch_or (ATOMG), 2: ld.cg probe.  bench.py times it in
// process to quote the build kernels against what this very GPU sustains (tools/microbench_atomics.cu is the long form).
__global__ void __launch_bounds__(256) l2_probe_kernel(uint32_t* __restrict__ arr, uint64_t n_words, uint64_t n_ops, int kind,
                                                       uint32_t* __restrict__ sink) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 8;
    uint32_t acc = 0;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x * 8 + threadIdx.x; base < n_ops; base += stride) {
        uint64_t w[8];
        uint32_t m[8], old[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint64_t h = splitmix64_fin((base + (uint64_t)j * blockDim.x) * 0x9E3779B97F4A7C15ULL + 12345);
            w[j] = (uint64_t)(((h >> 32) * n_words) >> 32);
            m[j] = 1u << (h & 31);
            old[j] = 0;
        }
        if (kind == 0) {
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("red.global.or.b32 [%0], %1;" ::"l"(arr + w[j]), "r"(m[j]) : "memory");
        } else if (kind == 1) {
#pragma unroll
            for (int j = 0; j < 8; j++) old[j] = atomicOr(arr + w[j], m[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) old[j] = __ldcg(arr + w[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) acc += old[j];
    }
    if (acc == 0x12345678u) *sink = acc;
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
