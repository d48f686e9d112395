// bucket_kernels.cuh -- shared-memory ("bucketed") construction of the large levels.
//
// Measured on B200 (tools/microbench_atomics.cu, tools/microbench_smem.cu): random 4-byte read-modify-write on
// an L2-resident bit array tops out at ~130 G op/s (every op moves a 32-byte sector across the crossbar, twice
// when the old value is returned), the same find_collisions step on a SHARED-MEMORY bit array runs at ~750 G
// keys/s.  So the large levels are built like this:
//
//   scatter(l) : the keys entering level l are grouped by the bit range their level-l slot falls in
//                ("bucket" = 2^b_shift bits).  scatter(0) streams the input; scatter(l>=1) streams level l-1's
//                buckets, keeps the keys whose level-(l-1) slot collided (Context::filter, src/lib.rs:443-452;
//                collide slice of the source bucket staged in shared memory) and groups the survivors by their
//                level-l bucket.  Tile-local binning in shared memory, one global cursor atomicAdd per
//                (tile, bucket), keys staged in bucket order so the stores are runs.
//   mark(l)    : one CTA per bucket: (a, collide) slice of the bucket lives in shared memory, the bucket's keys
//                are streamed once (Context::find_collisions, src/lib.rs:429-434, with ATOMS), then
//                a & ~collide (the net effect of every a.remove), collide, the per-512-bit popcounts
//                (compute_ranks, src/lib.rs:277-297) and the unique count are written out with coalesced
//                stores.  The last CTA derives k(l+1) and the next level's geometry on the device.
//
// All HBM traffic is coalesced streams: 16 B/key for scatter(0), 8 B/key for mark, 8 B/key + 8 B/survivor for
// scatter(l>=1).  Results are identical to the reference (order-free).
#pragma once
#include "build_kernels.cuh"

namespace bphf {

constexpr int BK_THREADS = 512;
constexpr int BK_KPT = 8;
constexpr int BK_TILE = BK_THREADS * BK_KPT;  // 4096 keys per tile
constexpr uint32_t BK_POS_BITS = 13;          // position of a key within (tile, bucket): < 4096

// destination of a scatter: level l's bucket regions (or one contiguous list when level l is plain).
// Bucket b owns keys[b * cap .. b * cap + cursor[b]).
struct ScatterDst {
    KeyCfg kc;
    uint64_t sp0;
    uint32_t bits, count, cap, shift;  // count == 1 for a plain level: one "bucket" = the contiguous key list
    uint64_t* keys;
    uint32_t* cursor;
};

__device__ __forceinline__ ScatterDst make_dst(const BuildState* st, uint32_t level, uint64_t* keys, uint32_t* cursor) {
    const LevelDesc* d = &st->lv[level];
    ScatterDst o;
    o.kc = st->kc;
    o.sp0 = d->sp0;
    o.bits = (uint32_t)d->bits;
    o.keys = keys;
    o.cursor = cursor;
    if (d->b_range) {
        o.shift = (uint32_t)d->b_shift;
        o.count = d->b_count;
        o.cap = d->b_cap;
    } else {
        o.shift = 0;
        o.count = 1;
        // one contiguous list: bounded by the list buffer (an unsharded build sized it for exactly k(level))
        const uint64_t lc = st->list_cap[level & 1];
        o.cap = lc > 0xffffffffULL ? 0xffffffffu : (uint32_t)lc;
    }
    return o;
}

// shared-memory carve-up of the scatter stage (dynamic shared memory)
struct ScatterSmem {
    uint64_t* stage_key;  // [BK_TILE] keys in bucket order
    uint32_t* stage_dst;  // [BK_TILE] destination key index of every staged key
    uint32_t* hist;       // [count] keys of this tile per bucket (zero between tiles)
    uint32_t* sbase;      // [count] exclusive scan of hist (position in the stage)
    uint32_t* delta;      // [count] destination index - stage position, per bucket
    uint32_t* warp_tot;   // [BK_THREADS / 32]
    uint32_t* ovf;        // [1] a bucket region overflowed in this CTA
};
__host__ __device__ inline size_t scatter_smem_bytes(uint32_t count) {
    return (size_t)BK_TILE * 12 + (size_t)((count + 3) & ~3u) * 12 + (BK_THREADS / 32) * 4 + 64;
}
__device__ __forceinline__ ScatterSmem carve_scatter(unsigned char* p, uint32_t count) {
    ScatterSmem s;
    s.stage_key = reinterpret_cast<uint64_t*>(p);
    p += (size_t)BK_TILE * 8;
    s.stage_dst = reinterpret_cast<uint32_t*>(p);
    p += (size_t)BK_TILE * 4;
    const size_t padded = (size_t)((count + 3) & ~3u);  // tables are read / written two entries at a time
    s.hist = reinterpret_cast<uint32_t*>(p);
    p += padded * 4;
    s.sbase = reinterpret_cast<uint32_t*>(p);
    p += padded * 4;
    s.delta = reinterpret_cast<uint32_t*>(p);
    p += padded * 4;
    s.warp_tot = reinterpret_cast<uint32_t*>(p);
    p += (BK_THREADS / 32) * 4;
    s.ovf = reinterpret_cast<uint32_t*>(p);
    return s;
}

// One tile: every thread brings up to BK_KPT keys (valid mask; FULL = all of them, known at compile time); they
// are hashed for the destination level, binned by destination bucket and written to the bucket regions as runs.
// Must be called by all BK_THREADS threads.  hist[] must be zero on entry and is zero again on exit.  Four block
// barriers per tile; the stage is protected from the next tile by that tile's first barrier.
template <bool FULL>
__device__ __forceinline__ void scatter_tile(BuildState* st, const ScatterDst& dst, const ScatterSmem& sm,
                                             const uint64_t (&key)[BK_KPT], uint32_t valid) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t bp[BK_KPT];  // bucket << 13 | position within (tile, bucket)
    if (dst.count > 1) {
#pragma unroll
        for (int j = 0; j < BK_KPT; j++) {
            bp[j] = 0;
            if (FULL || (valid & (1u << j))) {
                const uint32_t b = hashmod_small(dst.kc, dst.sp0, key[j], dst.bits) >> dst.shift;
                bp[j] = (b << BK_POS_BITS) | atomicAdd(&sm.hist[b], 1u);
            }
        }
    } else {
        // plain destination list: one bucket, position by warp ballot (all lanes would hit the same counter)
#pragma unroll
        for (int j = 0; j < BK_KPT; j++) {
            const bool v = FULL || ((valid >> j) & 1u);
            const uint32_t m = __ballot_sync(0xffffffffu, v);
            uint32_t base = 0;
            if (lane == 0 && m) base = atomicAdd(&sm.hist[0], (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            bp[j] = base + __popc(m & ((1u << lane) - 1u));
        }
    }
    __syncthreads();
    uint32_t total;
    if (dst.count <= 2 * BK_THREADS) {
        // common case: two buckets per thread; the global reservations are issued before the scan so that their
        // round trip overlaps it
        const uint32_t b0 = 2 * tid;
        uint2 h = make_uint2(0, 0);
        if (b0 < dst.count) {
            h = *reinterpret_cast<const uint2*>(sm.hist + b0);  // hist is padded to an even count
            if (b0 + 1 >= dst.count) h.y = 0;
            if (h.x | h.y) *reinterpret_cast<uint2*>(sm.hist + b0) = make_uint2(0, 0);
        }
        uint32_t g0 = 0, g1 = 0;
        if (h.x) g0 = atomicAdd(dst.cursor + (size_t)b0 * CURSOR_STRIDE, h.x);
        if (h.y) g1 = atomicAdd(dst.cursor + (size_t)(b0 + 1) * CURSOR_STRIDE, h.y);
        const uint32_t sum = h.x + h.y;
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) sm.warp_tot[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
        total = 0;
        const uint4* wt = reinterpret_cast<const uint4*>(sm.warp_tot);
#pragma unroll
        for (int q4 = 0; q4 < BK_THREADS / 128; q4++) {
            const uint4 t = wt[q4];
            const uint32_t e[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (q4 * 4 + k < warp) woff += e[k];
                total += e[k];
            }
        }
        if (sum) {
            const uint32_t run = woff + incl - sum;
            const uint32_t region = dst.count > 1 ? dst.cap : 0u;
            if ((h.x && (uint64_t)g0 + h.x > dst.cap) || (h.y && (uint64_t)g1 + h.y > dst.cap)) {
                st->status = ST_BUCKET_OVERFLOW;  // region too small: host rebuilds unbucketed
                *sm.ovf = 1;
            }
            *reinterpret_cast<uint2*>(sm.sbase + b0) = make_uint2(run, run + h.x);
            *reinterpret_cast<uint2*>(sm.delta + b0) = make_uint2(b0 * region + g0 - run, (b0 + 1) * region + g1 - (run + h.x));
        }
    } else {
        // many buckets: each thread owns a contiguous chunk
        const uint32_t per = (dst.count + BK_THREADS - 1) / BK_THREADS;
        const uint32_t b0 = min(dst.count, (uint32_t)tid * per), b1 = min(dst.count, b0 + per);
        uint32_t sum = 0;
        for (uint32_t b = b0; b < b1; b++) sum += sm.hist[b];
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) sm.warp_tot[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
        total = 0;
#pragma unroll
        for (int w = 0; w < BK_THREADS / 32; w++) {
            const uint32_t t = sm.warp_tot[w];
            if (w < warp) woff += t;
            total += t;
        }
        uint32_t run = woff + incl - sum;
        for (uint32_t b = b0; b < b1; b++) {
            const uint32_t h = sm.hist[b];
            if (h) {
                const uint32_t g = atomicAdd(dst.cursor + (size_t)b * CURSOR_STRIDE, h);
                if ((uint64_t)g + h > dst.cap) {
                    st->status = ST_BUCKET_OVERFLOW;
                    *sm.ovf = 1;
                }
                sm.sbase[b] = run;
                sm.delta[b] = b * dst.cap + g - run;
                sm.hist[b] = 0;
                run += h;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < BK_KPT; j++) {
        if (FULL || (valid & (1u << j))) {
            const uint32_t b = dst.count > 1 ? (bp[j] >> BK_POS_BITS) : 0u;
            const uint32_t s = sm.sbase[b] + (dst.count > 1 ? (bp[j] & ((1u << BK_POS_BITS) - 1u)) : bp[j]);
            sm.stage_key[s] = key[j];
            sm.stage_dst[s] = sm.delta[b] + s;
        }
    }
    __syncthreads();
    if (*sm.ovf == 0) {
        if (FULL) {
#pragma unroll
            for (int j = 0; j < BK_KPT; j++) {
                const uint32_t i = j * BK_THREADS + tid;
                dst.keys[sm.stage_dst[i]] = sm.stage_key[i];
            }
        } else {
            for (uint32_t i = tid; i < total; i += BK_THREADS) dst.keys[sm.stage_dst[i]] = sm.stage_key[i];
        }
    }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// -------------------------------------------------------------------------------------------------------
// scatter0_kernel: input keys [key_begin, key_end) -> level-0 buckets
// -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BK_THREADS, 3)
    scatter0_kernel(BuildState* __restrict__ st, const uint64_t* __restrict__ keys, uint64_t key_begin,
                    uint64_t key_end, uint64_t* __restrict__ buf0, uint32_t* __restrict__ cursor0, uint32_t smem_bytes) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    if (st->status != ST_RUNNING || st->n_levels != 0 || st->lv[0].b_range == 0) return;
    const ScatterDst dst = make_dst(st, 0, buf0, cursor0);
    if (scatter_smem_bytes(dst.count) > smem_bytes) {
        if (threadIdx.x == 0) st->status = ST_BUCKET_OVERFLOW;
        return;
    }
    const ScatterSmem sm = carve_scatter(dyn_smem, dst.count);
    for (uint32_t b = threadIdx.x; b < dst.count + 1; b += BK_THREADS) sm.hist[b] = 0;
    if (threadIdx.x == 0) *sm.ovf = 0;
    __syncthreads();
    const uint64_t n = key_end - key_begin;
    const uint64_t n_tiles = (n + BK_TILE - 1) / BK_TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t base = key_begin + tile * BK_TILE + threadIdx.x;
        uint64_t key[BK_KPT];
        // the tile this CTA takes next: pull it into L2 while this one is binned (one 128-byte line per 16 lanes)
        if ((threadIdx.x & 15) == 0 && tile + gridDim.x < n_tiles) {
            const uint64_t nb = key_begin + (tile + gridDim.x) * BK_TILE + threadIdx.x;
#pragma unroll
            for (int j = 0; j < BK_KPT; j++)
                if (nb + (uint64_t)j * BK_THREADS < key_end) prefetch_l2(keys + nb + (uint64_t)j * BK_THREADS);
        }
        if (key_begin + (tile + 1) * BK_TILE <= key_end) {
#pragma unroll
            for (int j = 0; j < BK_KPT; j++) key[j] = ld_stream_u64(keys + base + (uint64_t)j * BK_THREADS);
            scatter_tile<true>(st, dst, sm, key, 0xffu);
        } else {
            uint32_t valid = 0;
#pragma unroll
            for (int j = 0; j < BK_KPT; j++) {
                const uint64_t i = base + (uint64_t)j * BK_THREADS;
                key[j] = 0;
                if (i < key_end) {
                    key[j] = ld_stream_u64(keys + i);
                    valid |= 1u << j;
                }
            }
            scatter_tile<false>(st, dst, sm, key, valid);
        }
    }
}

// -------------------------------------------------------------------------------------------------------
// scatter_kernel(level >= 1): filter level-1's buckets against their collide slice, survivors -> level's buckets.
// One CTA per source bucket (grid-stride over buckets).  The filter streams sub-tiles of the source bucket and
// appends the surviving keys to a pending buffer in shared memory (the same buffer the partition later stages
// into); the partition runs whenever another sub-tile might not fit, i.e. on (nearly) full tiles.  The pending
// keys carry over from one source bucket to the next -- they all go to the same destination level.
// -------------------------------------------------------------------------------------------------------
constexpr int SRC_KPT = 4;
constexpr int SRC_TILE = BK_THREADS * SRC_KPT;  // 2048 source keys per filter step

__device__ __forceinline__ void flush_pending(BuildState* st, const ScatterDst& dst, const ScatterSmem& sm,
                                              uint32_t* pend_cnt) {
    // all threads: *pend_cnt is stable (read after a barrier); the pending keys sit at stage_key[0 .. n_p)
    const uint32_t n_p = *pend_cnt;
    uint64_t key[BK_KPT];
    if (n_p == BK_TILE) {
#pragma unroll
        for (int j = 0; j < BK_KPT; j++) key[j] = sm.stage_key[j * BK_THREADS + threadIdx.x];
        scatter_tile<true>(st, dst, sm, key, 0xffu);
    } else {
        uint32_t valid = 0;
#pragma unroll
        for (int j = 0; j < BK_KPT; j++) {
            const uint32_t i = j * BK_THREADS + threadIdx.x;
            key[j] = 0;
            if (i < n_p) {
                key[j] = sm.stage_key[i];
                valid |= 1u << j;
            }
        }
        scatter_tile<false>(st, dst, sm, key, valid);
    }
    if (threadIdx.x == 0) *pend_cnt = 0;
    __syncthreads();  // the stage has been written out; the buffer takes pending keys again
}

__global__ void __launch_bounds__(BK_THREADS, 2)
    scatter_kernel(BuildState* __restrict__ st, uint32_t level, uint64_t* __restrict__ buf0,
                   uint64_t* __restrict__ buf1, const uint32_t* __restrict__ c32, uint32_t* __restrict__ cursors,
                   uint32_t smem_bytes) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ uint32_t pend_cnt;
    if (st->status != ST_RUNNING || st->n_levels != level || st->lv[level - 1].b_range == 0) return;
    const LevelDesc* pv = &st->lv[level - 1];
    const uint32_t p_range = pv->b_range, p_count = pv->b_count, p_cap = pv->b_cap, p_bits = (uint32_t)pv->bits;
    const uint64_t p_sp0 = pv->sp0;
    const uint64_t p_off32 = pv->word_off * 2;
    const uint32_t p_words32 = (uint32_t)(round_up8(pv->n_words) * 2);  // u32 words of the (padded) level
    const uint64_t* __restrict__ src = ((level - 1) & 1) ? buf1 : buf0;
    const uint32_t* __restrict__ src_cursor = cursors + (size_t)((level - 1) & 1) * MAX_BUCKETS_ALLOC * CURSOR_STRIDE;
    const ScatterDst dst = make_dst(st, level, (level & 1) ? buf1 : buf0,
                                    cursors + (size_t)(level & 1) * MAX_BUCKETS_ALLOC * CURSOR_STRIDE);
    const size_t c_bytes = (size_t)p_range / 8;
    if (c_bytes + scatter_smem_bytes(dst.count) > smem_bytes) {
        if (threadIdx.x == 0) st->status = ST_BUCKET_OVERFLOW;
        return;
    }
    uint32_t* c_sm = reinterpret_cast<uint32_t*>(dyn_smem);
    const ScatterSmem sm = carve_scatter(dyn_smem + c_bytes, dst.count);
    for (uint32_t b = threadIdx.x; b < dst.count + 1; b += BK_THREADS) sm.hist[b] = 0;
    if (threadIdx.x == 0) {
        *sm.ovf = 0;
        pend_cnt = 0;
    }
    const uint32_t p_mask = p_range - 1;
    const int lane = threadIdx.x & 31;

    for (uint32_t bucket = blockIdx.x; bucket < p_count; bucket += gridDim.x) {
        __syncthreads();
        // collide slice of this bucket -> shared memory (coalesced)
        const uint32_t w_first = bucket * (p_range / 32);
        for (uint32_t w = threadIdx.x; w < p_range / 32; w += BK_THREADS)
            c_sm[w] = (w_first + w < p_words32) ? c32[p_off32 + w_first + w] : 0;
        __syncthreads();
        const uint32_t fill = min(src_cursor[(size_t)bucket * CURSOR_STRIDE], p_cap);
        const uint64_t* __restrict__ bsrc = src + (uint64_t)bucket * p_cap;
        for (uint32_t t0 = 0; t0 < fill; t0 += SRC_TILE) {
            uint64_t key[SRC_KPT];
            bool live[SRC_KPT];
#pragma unroll
            for (int j = 0; j < SRC_KPT; j++) {
                const uint32_t i = t0 + j * BK_THREADS + threadIdx.x;
                live[j] = i < fill;
                key[j] = live[j] ? ld_stream_u64(bsrc + i) : 0;
            }
            // Context::filter (src/lib.rs:443-452): the key is redone iff its slot collided
            bool redo[SRC_KPT];
#pragma unroll
            for (int j = 0; j < SRC_KPT; j++) {
                const uint32_t l = hashmod_small(dst.kc, p_sp0, key[j], p_bits) & p_mask;
                redo[j] = live[j] && ((c_sm[l >> 5] >> (l & 31)) & 1u);
            }
            __syncthreads();  // every thread has taken its decision on the previous step's pend_cnt
#pragma unroll
            for (int j = 0; j < SRC_KPT; j++) {
                const uint32_t m = __ballot_sync(0xffffffffu, redo[j]);
                uint32_t base = 0;
                if (lane == 0 && m) base = atomicAdd(&pend_cnt, (uint32_t)__popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (redo[j]) sm.stage_key[base + __popc(m & ((1u << lane) - 1u))] = key[j];
            }
            __syncthreads();  // appends complete: pend_cnt is the same for everybody
            if (pend_cnt > BK_TILE - SRC_TILE) flush_pending(st, dst, sm, &pend_cnt);
        }
    }
    __syncthreads();
    if (pend_cnt) flush_pending(st, dst, sm, &pend_cnt);
}

// -------------------------------------------------------------------------------------------------------
// bucket_mark_kernel(level): find_collisions for one bucket in shared memory + finalize + popcounts
// -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mark_smem(uint32_t* a_sm, uint32_t* c_sm, uint32_t idx) {
    const uint32_t m = 1u << (idx & 31);
    const uint32_t old = atomicOr(&a_sm[idx >> 5], m);
    if (old & m) atomicOr(&c_sm[idx >> 5], m);
}

constexpr int MARK_KPT = 8;

__global__ void __launch_bounds__(BK_THREADS)
    bucket_mark_kernel(BuildState* __restrict__ st, uint32_t level, const uint64_t* __restrict__ buf0,
                       const uint64_t* __restrict__ buf1, uint32_t* __restrict__ a32, uint32_t* __restrict__ c32,
                       uint32_t* __restrict__ blockpop, uint32_t* __restrict__ cursors, uint32_t smem_bytes) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ uint64_t warp_u[BK_THREADS / 32];
    __shared__ bool is_last;
    if (st->status != ST_RUNNING || st->n_levels != level) return;
    const LevelDesc* cu = &st->lv[level];
    if (cu->b_range == 0) return;
    const uint32_t range = cu->b_range, count = cu->b_count, cap = cu->b_cap, bits = (uint32_t)cu->bits;
    const KeyCfg kc = st->kc;
    const uint64_t sp0 = cu->sp0;
    const uint64_t off32 = cu->word_off * 2;
    const uint32_t words32 = (uint32_t)(round_up8(cu->n_words) * 2);
    const uint32_t rw = range / 32;  // u32 words per bucket (multiple of 16)
    if ((size_t)rw * 8 > smem_bytes) {
        if (threadIdx.x == 0) st->status = ST_BUCKET_OVERFLOW;
        return;
    }
    uint32_t* a_sm = reinterpret_cast<uint32_t*>(dyn_smem);
    uint32_t* c_sm = a_sm + rw;
    const uint64_t* __restrict__ src = (level & 1) ? buf1 : buf0;
    const uint32_t* __restrict__ cursor = cursors + (size_t)(level & 1) * MAX_BUCKETS_ALLOC * CURSOR_STRIDE;
    const int lane = threadIdx.x & 31;
    const uint32_t mask = range - 1;
    // sharded build: write the RAW pair (a = cnt >= 1, collide = cnt >= 2 of this shard's keys); the OR-collide
    // merge and finalize_kernel follow
    const bool raw = st->shard_world > 1;
    uint64_t uniq = 0, seen = 0;

    for (uint32_t bucket = blockIdx.x; bucket < count; bucket += gridDim.x) {
        __syncthreads();
        {
            uint4* z = reinterpret_cast<uint4*>(a_sm);
            for (uint32_t w = threadIdx.x; w < rw / 2; w += BK_THREADS) z[w] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        const uint32_t fill = min(cursor[(size_t)bucket * CURSOR_STRIDE], cap);
        if (threadIdx.x == 0) seen += fill;
        const uint64_t* __restrict__ bsrc = src + (uint64_t)bucket * cap;
        for (uint32_t t0 = 0; t0 < fill; t0 += BK_THREADS * MARK_KPT) {
            uint64_t key[MARK_KPT];
            if (t0 + BK_THREADS * MARK_KPT <= fill) {
#pragma unroll
                for (int j = 0; j < MARK_KPT; j++) key[j] = ld_stream_u64(bsrc + t0 + j * BK_THREADS + threadIdx.x);
#pragma unroll
                for (int j = 0; j < MARK_KPT; j++) mark_smem(a_sm, c_sm, hashmod_small(kc, sp0, key[j], bits) & mask);
            } else {
#pragma unroll
                for (int j = 0; j < MARK_KPT; j++) {
                    const uint32_t i = t0 + j * BK_THREADS + threadIdx.x;
                    key[j] = (i < fill) ? ld_stream_u64(bsrc + i) : 0;
                }
#pragma unroll
                for (int j = 0; j < MARK_KPT; j++) {
                    const uint32_t i = t0 + j * BK_THREADS + threadIdx.x;
                    if (i < fill) mark_smem(a_sm, c_sm, hashmod_small(kc, sp0, key[j], bits) & mask);
                }
            }
        }
        __syncthreads();
        // a &= ~collide; write a and collide slices; popcount per 512-bit block (16 consecutive u32 = 16 lanes)
        const uint32_t w_first = bucket * rw;
        for (uint32_t w0 = 0; w0 < rw; w0 += BK_THREADS) {
            const uint32_t w = w0 + threadIdx.x;  // rw is a multiple of 16, BK_THREADS of 32
            uint32_t pc = 0;
            const bool in = (w < rw) && (w_first + w < words32);
            if (in) {
                const uint32_t c = c_sm[w];
                const uint32_t a = raw ? a_sm[w] : (a_sm[w] & ~c);
                a32[off32 + w_first + w] = a;
                c32[off32 + w_first + w] = c;
                pc = __popc(a);
            }
            uniq += pc;
            pc += __shfl_xor_sync(0xffffffffu, pc, 1);
            pc += __shfl_xor_sync(0xffffffffu, pc, 2);
            pc += __shfl_xor_sync(0xffffffffu, pc, 4);
            pc += __shfl_xor_sync(0xffffffffu, pc, 8);
            if (in && !raw && (lane & 15) == 0) blockpop[(off32 + w_first + w) >> 4] = pc;
        }
    }
    // unique count + consistency counter, one atomic per CTA; last CTA finishes the level
    uint64_t u = uniq;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) u += __shfl_xor_sync(0xffffffffu, u, d);
    if (lane == 0) warp_u[threadIdx.x >> 5] = u;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
#pragma unroll
        for (int w = 0; w < BK_THREADS / 32; w++) t += warp_u[w];
        if (t) atomicAdd((unsigned long long*)&st->uniq_count, (unsigned long long)t);
        if (seen) atomicAdd((unsigned long long*)&st->bucket_keys_seen, (unsigned long long)seen);
        __threadfence();
        const uint32_t ticket = atomicAdd(&st->fin_ticket, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        if (threadIdx.x == 0) {
            __threadfence();
            const uint64_t uq = atomicAdd((unsigned long long*)&st->uniq_count, 0ULL);
            const uint64_t sn = atomicAdd((unsigned long long*)&st->bucket_keys_seen, 0ULL);
            const uint64_t k_in = st->lv[level].k_in;
            st->bucket_keys_seen = 0;
            if (raw) {
                // hand over to barrier + merge + finalize_kernel: this shard's key count of the level, counters reset
                st->redo_count = sn;
                st->uniq_count = 0;
                st->fin_ticket = 0;
            } else if (st->status == ST_RUNNING) {
                if (uq > k_in || sn != k_in) {
                    st->status = ST_ERR_INTERNAL;
                    st->err_info = ((uint64_t)level << 48) | (sn & 0xffffffffffULL) | (1ULL << 45);
                } else {
                    finish_level(st, level, k_in - uq);
                }
            }
        }
        __syncthreads();
        // cursors of the next level (it lives in the other parity) start at zero
        uint32_t* nc = cursors + (size_t)((level + 1) & 1) * MAX_BUCKETS_ALLOC * CURSOR_STRIDE;
        for (uint32_t b = threadIdx.x; b < MAX_BUCKETS_ALLOC; b += BK_THREADS) nc[(size_t)b * CURSOR_STRIDE] = 0;
    }
}

}  // namespace bphf
