// build_kernels.cuh -- sm_100a kernels of the MPHF level cascade.
//
// Reference functions replaced (file:line into /root/reference):
//   Context::find_collisions  src/lib.rs:429-434   -> mark()           (pass_kernel, tail_kernel)
//   Context::filter           src/lib.rs:443-452   -> collide test + stream compaction (pass_kernel)
//                                                     and a &= ~collide (finalize_kernel)
//   Mphf::new_parallel loop   src/lib.rs:363-408   -> device-resident BuildState, no host round trip
//   Mphf::compute_ranks       src/lib.rs:277-297   -> finalize_kernel block popcounts + scan kernels
//
// Schedule (differs from the reference on purpose, result identical -- SURVEY.md 3.1):
//   pass(l)     : stream the keys that entered level l-1, keep those whose level-(l-1) slot collided,
//                 append them to level l's key list AND mark them into level l's (a, collide) arrays
//                 in the same kernel.  pass(0) only marks.
//   finalize(l) : a &= ~collide word-wise (the net effect of every a.remove(idx)), per-512-bit-block
//                 popcounts, total unique count u_l; the last CTA derives k_{l+1} = k_l - u_l and the
//                 next level's size/offset/seed on the device.
//   tail        : once k <= TAIL_MAX_KEYS one CTA finishes every remaining level in shared memory.
#pragma once
#include "common.cuh"

namespace bphf {

constexpr int PASS_THREADS = 256;
constexpr int PASS_KPT = 8;                          // keys per thread per tile
constexpr int PASS_TILE = PASS_THREADS * PASS_KPT;   // 2048 keys, 16 KB staging
constexpr int FIN_THREADS = 256;
constexpr int TAIL_THREADS = 1024;
constexpr uint32_t NO_TAIL = 0xffffffffu;

// streaming 8-byte load, cached in L2 only (keys are read once per pass; and the persistent cascade kernel reads
// lists another CTA wrote earlier in the same launch, which a non-coherent L1 path must not serve)
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// fire-and-forget OR (REDG): the collide update never needs the old word.  Spelled in PTX because ptxas keeps the
// returning ATOMG form inside the persistent cascade kernel (fences follow), and REDG is ~1.5x cheaper at the L2
__device__ __forceinline__ void red_or_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- TMA (bulk async copy) + mbarrier helpers: key tiles are prefetched global -> shared by the copy engine, two
// tiles ahead of the tile being processed, so the key stream costs neither LSU slots nor exposed DRAM latency
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// level bookkeeping shared by the finalize kernel (last CTA), the tail kernel and the host
// (resume after growing a buffer).  Level `level` gets k keys; returns false (status set) when a
// buffer must grow first.
__host__ __device__ inline bool try_make_level(BuildState* st, uint32_t level, uint64_t k) {
    make_level(st, level, k);
    LevelDesc& d = st->lv[level];
    const bool tail_ok = (k <= TAIL_MAX_KEYS) && (d.bits <= TAIL_MAX_BITS);
    if (tail_ok) {
        // after a bucketed level the scatter kernel hands the tail its keys as a plain list in buf[level & 1]
        if (level >= 1 && st->lv[level - 1].b_range != 0 && k > st->list_cap[level & 1]) {
            st->status = ST_NEED_LIST;
            st->need = k;
            return false;
        }
        if (st->tail_level == NO_TAIL) st->tail_level = level;
        // arena space for tail levels comes out of the TAIL_RESERVE_WORDS slack; lists live in smem
        st->words_used = d.word_off + round_up8(d.n_words);
        return true;
    }
    const uint64_t need_words = d.word_off + round_up8(d.n_words);
    if (need_words > st->words_cap) {
        st->status = ST_NEED_WORDS;
        st->need = need_words;
        return false;
    }
    // bucketed (shared-memory) construction for a prefix of large levels, plain L2-atomic kernels after that
    BucketGeom g;
    const bool may_bucket = st->bucket_enable && (level == 0 || st->lv[level - 1].b_range != 0);
    if (may_bucket && bucket_geometry(d.bits, k, st->bucket_slots, st->bucket_min_keys, st->shard_world > 1 ? st->shard_frac : 1.0, &g)) {
        const uint64_t need_keys = (uint64_t)g.count * g.cap;
        if (need_keys > st->list_cap[level & 1]) {
            st->status = ST_NEED_LIST;
            st->need = need_keys;
            return false;
        }
        d.b_shift = g.shift;
        d.b_range = g.range;
        d.b_count = g.count;
        d.b_cap = g.cap;
    } else if (level >= 1 && st->shard_world <= 1 && k > st->list_cap[level & 1]) {
        st->status = ST_NEED_LIST;
        st->need = k;
        return false;
    }
    st->words_used = need_words;
    return true;
}

// called once level `level` is complete (a final, popcounts written) with k_next keys left over
__host__ __device__ inline void finish_level(BuildState* st, uint32_t level, uint64_t k_next) {
    st->n_levels = level + 1;
    st->redo_count = 0;
    st->uniq_count = 0;
    st->fin_ticket = 0;
    if (st->chunk_mode) {
        // from_chunked_iterator_parallel: `iter > max_iters` is tested at the top of the loop (src/lib.rs:568-573),
        // BEFORE the `keys_remaining == 0` break (:580-582), so it fires for iter = level + 1 even when nothing is left;
        // the loop is left right after the buffering level (:645-648, next == switch_level) without another test, and
        // the tail is a new_parallel of its own whose counter starts at the switch (:654) and fires like the check below
        const uint32_t next = level + 1;
        const bool tail_phase = st->switch_level != NO_LEVEL && next >= st->switch_level;
        bool fail = next >= BPHF_MAX_LEVELS;
        if (!tail_phase) fail = fail || (st->chunk_has_max && (uint64_t)next > st->chunk_max_iters);
        else fail = fail || (next - st->switch_level > BPHF_MAX_ITERS);
        if (fail) {
            st->status = ST_ERR_MAX_ITERS;
            st->err_info = k_next;
            return;
        }
    } else if (level >= 1 && level + 1 > BPHF_MAX_ITERS) {
        // reference: the `iter > MAX_ITERS` check sits inside `while !redo_keys.is_empty()` after
        // `iter += 1`, i.e. it fires for level index >= 100 even when nothing is left to redo
        st->status = ST_ERR_MAX_ITERS;
        st->err_info = k_next;
        return;
    }
    if (k_next == 0) {
        st->status = ST_DONE;
        return;
    }
    try_make_level(st, level + 1, k_next);
}

// find_collisions (src/lib.rs:429-434) on the 32-bit view of the little-endian u64 words:
// bit i of the level = bit (i & 31) of u32 word (i >> 5).  a: fetch_or returning the old word;
// collide: fire-and-forget OR when the slot was already taken.  The reference's
// `!collide.contains(idx)` pre-check only skips redundant work; the final (a, collide) is the same.
__device__ __forceinline__ void mark_bit(uint32_t* __restrict__ a32, uint32_t* __restrict__ c32, uint64_t w,
                                         uint32_t m) {
    const uint32_t old = atomicOr(a32 + w, m);
    if (old & m) red_or_u32(c32 + w, m);
}

template <bool BIG, class KC>
__device__ __forceinline__ uint64_t level_index(const KC& kc, uint64_t sp0, uint64_t key, uint64_t bits) {
    if (BIG) return hashmod_big(kc, sp0, key, bits);
    return hashmod_small(kc, sp0, key, (uint32_t)bits);
}

// ---------------------------------------------------------------------------------------------
// mark_list_kernel: find_collisions only, over a contiguous key list.  level 0: the input keys
// [key_begin, key_end); level >= 1 (first plain level after the bucketed prefix): the list in buf[level & 1]
// ---------------------------------------------------------------------------------------------
template <bool MAYBE_BIG, class KC = KeyCfg>
__global__ void __launch_bounds__(PASS_THREADS)
    mark_list_kernel(BuildState* __restrict__ st, uint32_t level, const uint64_t* __restrict__ user_keys,
                     uint64_t key_begin, uint64_t key_end, const uint64_t* __restrict__ buf0,
                     const uint64_t* __restrict__ buf1, uint32_t* __restrict__ a32, uint32_t* __restrict__ c32,
                     const uint32_t* __restrict__ list_cursor) {
    if (st->status != ST_RUNNING || st->n_levels != level || level >= st->tail_level) return;
    if (st->lv[level].b_range != 0) return;                         // bucketed level: bucket_mark_kernel
    if (st->lv[level].x_slice_bits != 0) return;                    // exchange level of a sharded build: xchg_kernels.cuh
    const bool list_ready = st->list_ready_level == level;          // collected list of a sharded build
    if (level >= 1 && st->lv[level - 1].b_range == 0 && !list_ready) return;  // plain / exchange predecessor: pass_kernel marks
    const uint64_t* __restrict__ keys = user_keys;
    if (level >= 1) {
        keys = (level & 1) ? buf1 : buf0;
        key_begin = 0;
        key_end = st->lv[level].k_loc;
        if (st->shard_world > 1 && !list_ready) {
            // sharded: this shard's part of the list is what its scatter step appended (count in the list cursor);
            // publish it as the level's local key count for the merge barrier and for finalize
            key_end = list_cursor[0];
            if (blockIdx.x == 0 && threadIdx.x == 0) st->redo_count = key_end;
        }
    }
    const KC kc(st->kc);
    const uint64_t bits = st->lv[level].bits;
    const uint64_t sp0 = st->lv[level].sp0;
    const uint64_t off32 = st->lv[level].word_off * 2;
    const bool big = MAYBE_BIG && (bits >> 32) != 0;
    const uint64_t n = key_end - key_begin;
    const uint64_t n_tiles = (n + PASS_TILE - 1) / PASS_TILE;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t base = key_begin + tile * PASS_TILE + threadIdx.x;
        const bool full = key_begin + (tile + 1) * PASS_TILE <= key_end;
        uint64_t key[PASS_KPT];
#pragma unroll
        for (int j = 0; j < PASS_KPT; j++) {
            const uint64_t i = base + (uint64_t)j * PASS_THREADS;
            key[j] = (full || i < key_end) ? ld_stream_u64(keys + i) : 0;
        }
        // all PASS_KPT fetch_or's are issued back to back (independent), the collide updates follow:
        // the a-word round trips overlap instead of serialising on `if (old & m)`
        uint64_t w[PASS_KPT];
        uint32_t m[PASS_KPT], old[PASS_KPT];
#pragma unroll
        for (int j = 0; j < PASS_KPT; j++) {
            const uint64_t idx = big ? level_index<true>(kc, sp0, key[j], bits) : level_index<false>(kc, sp0, key[j], bits);
            w[j] = off32 + (idx >> 5);
            m[j] = 1u << (idx & 31);
        }
        if (full) {
#pragma unroll
            for (int j = 0; j < PASS_KPT; j++) old[j] = atomicOr(a32 + w[j], m[j]);
        } else {
#pragma unroll
            for (int j = 0; j < PASS_KPT; j++) {
                const uint64_t i = base + (uint64_t)j * PASS_THREADS;
                old[j] = 0;
                if (i < key_end) old[j] = atomicOr(a32 + w[j], m[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < PASS_KPT; j++)
            if (old[j] & m[j]) red_or_u32(c32 + w[j], m[j]);
    }
}

// ---------------------------------------------------------------------------------------------
// pass_kernel: level >= 1.  filter(level-1) + compaction + mark(level) fused.
// ---------------------------------------------------------------------------------------------
// Striped key lists.  The list of level l (the keys that collided at level l-1) is NOT one dense array: it is cut into
// st->n_regions "regions" of st->reg_cap key slots each (buf[l & 1] + r * reg_cap), with reg_cnt[l & 1][r] keys in
// region r.  Region r of level l holds exactly the survivors of region r of level l-1 (level 1: of the input tiles
// r, r + R, r + 2R, ...), so a region never receives more keys than its predecessor held: reg_cap = the tiles of the
// first source per region x 256 can never overflow, and whoever filters region r appends to it with a private cursor.
// Why: with one dense list every 256 survivors cost a reservation on ONE global cursor -- 174 k same-address atomics
// for level 1 of a 1e8-key build -- and those alone held the pass at 0.94 ms against 0.69 ms without them
// (tools/microbench_pass.cu, profiles/r02_summary.md); ordering does not matter (only the redo SET feeds the next
// level).  Lists that other kernels produce stay dense (the input, scatter_kernel's output after the
// bucketed levels, the collected list of a sharded build): list_is_striped() tells which kind level l has.
//
// Schedule: every WARP runs on its own -- a region at a time, 256-key tiles through a two-deep TMA ring of its own
// (cp.async.bulk + mbarrier), all collide probes of a tile issued before anything is waited for, ballot compaction
// into a per-warp ring in shared memory, and whenever 256 survivors are staged: coalesced append to the warp's region
// plus the fetch_or's of level l for them, whose results are consumed one flush later (by then they have arrived).
// No block barrier, no shared cursor: the warps of an SM drift apart, so the probes of some are in flight together
// with the atomics of others.
// GATHER: only filter + compact, into the DENSE array `gather_dst` (a shard's exchange slot, read by its peers as one
// array: reservations on st->redo_count, at most 2^20 keys); level `level` is marked later from the collected list.
#ifndef BPHF_PASS_CTAS
#define BPHF_PASS_CTAS 2  // CTAs per SM of the level kernels: 64 KB of shared memory each, the rest of the array stays L1
#endif
constexpr int WP_TILE = 32 * PASS_KPT;        // 256 keys per warp tile (2 KB)
constexpr int WP_RING = 2 * WP_TILE;          // survivors staged per warp (power of two)
constexpr int WP_WARPS = PASS_THREADS / 32;
constexpr size_t PASS_SMEM_BYTES = (size_t)WP_WARPS * (WP_RING + 2 * WP_TILE) * 8 + WP_WARPS * 2 * 8;

template <bool BIG> struct SlotIdx { typedef uint32_t type; };
template <> struct SlotIdx<true> { typedef uint64_t type; };

// level `level`'s key list is striped iff a filter pass wrote it (see above)
__device__ __forceinline__ bool list_is_striped(const BuildState* st, uint32_t level) {
    return level >= 1 && __ldcg(&st->lv[level - 1].b_range) == 0 && __ldcg(&st->list_ready_level) != level;
}

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// XSRC: the previous level was an exchange level of a sharded build (xchg_kernels.cuh): its keys sit in this shard's
// list grouped by (region, owner) and the owners' answers -- one survivor bit per key, in the same order -- in `xbits`;
// the filter is then a sequential read instead of a random probe.
template <bool MAYBE_BIG, bool GATHER = false, class KC = KeyCfg, bool XSRC = false>
__device__ __forceinline__ void pass_body(BuildState* __restrict__ st, uint32_t level, const uint64_t* __restrict__ user_keys,
                                          uint64_t* __restrict__ buf0, uint64_t* __restrict__ buf1,
                                          uint32_t* __restrict__ a32, uint32_t* __restrict__ c32,
                                          uint64_t* __restrict__ gather_dst = nullptr,
                                          const uint64_t* __restrict__ src_chunk = nullptr, uint64_t chunk_keys = 0,
                                          const uint32_t* __restrict__ xbits = nullptr) {
    typedef typename SlotIdx<MAYBE_BIG>::type slot_t;
    extern __shared__ __align__(128) unsigned char pass_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* __restrict__ ring = reinterpret_cast<uint64_t*>(pass_smem) + warp * WP_RING;
    uint64_t* __restrict__ stage = reinterpret_cast<uint64_t*>(pass_smem) + WP_WARPS * WP_RING + warp * 2 * WP_TILE;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(pass_smem) + WP_WARPS * (WP_RING + 2 * WP_TILE) + warp * 2;

    const KC kc(st->kc);
    const LevelDesc* pv = &st->lv[level - 1];
    const LevelDesc* cu = &st->lv[level];
    // (descriptors and lists may have been written by another CTA earlier in the same launch: read through L2)
    const uint64_t p_bits = __ldcg(&pv->bits), p_sp0 = __ldcg(&pv->sp0), p_off32 = __ldcg(&pv->word_off) * 2;
    const uint64_t c_bits = __ldcg(&cu->bits), c_sp0 = __ldcg(&cu->sp0), c_off32 = __ldcg(&cu->word_off) * 2;
    const bool p_big = MAYBE_BIG && (p_bits >> 32) != 0;
    const bool c_big = MAYBE_BIG && (c_bits >> 32) != 0;
    const uint32_t R = __ldcg(&st->n_regions);
    const uint64_t reg_cap = __ldcg(&st->reg_cap);
    uint32_t* __restrict__ reg_cnt = st->reg_cnt;   // [2][R]
    // ---- source: one chunk of the input (streamed ingest), a dense list (the input, a scatter / collect output) or the
    // striped list the previous filter pass wrote
    const bool src_striped = XSRC || (!src_chunk && list_is_striped(st, level - 1));
    const uint32_t x_world = XSRC ? __ldcg(&st->shard_world) : 1u;
    const uint64_t x_sub_cap = XSRC ? (uint64_t)__ldcg(&st->x_sub_cap) : 0;
    const uint32_t* __restrict__ x_cnt = st->xreg_cnt + (size_t)((level - 1) & 1) * R * SHARD_MAX;
    const uint64_t* __restrict__ src =
        src_chunk ? src_chunk : (level == 1 && !XSRC) ? user_keys : (((level - 1) & 1) ? buf1 : buf0);
    const uint64_t k_dense = src_chunk ? chunk_keys : __ldcg(&pv->k_loc);   // dense sources only
    const uint64_t dense_tiles = (k_dense + WP_TILE - 1) / WP_TILE;
    const uint32_t* __restrict__ src_cnt = reg_cnt + (size_t)((level - 1) & 1) * R;
    uint64_t* __restrict__ dst = GATHER ? gather_dst : ((level & 1) ? buf1 : buf0);
    uint32_t* __restrict__ dst_cnt = reg_cnt + (size_t)(level & 1) * R;
    const bool append = src_chunk != nullptr;  // streamed ingest: the chunks of one level accumulate in the regions
    if (!GATHER && !src_striped && !append) {
        // a dense source decides what a region may receive: uniform check, before anything is written
        const uint64_t per = (dense_tiles + R - 1) / R;
        if (per * WP_TILE > reg_cap) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                st->status = ST_NEED_LIST;
                st->need = (per * WP_TILE + WP_TILE) * R;
            }
            return;
        }
    }
    uint32_t* __restrict__ a_lv = a32 + c_off32;        // level l's occupancy words
    uint32_t* __restrict__ c_lv = c32 + c_off32;        // level l's collision words
    const uint32_t* __restrict__ c_pv = c32 + p_off32;  // level l-1's collision words (final)
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool src_aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;

    uint32_t head = 0, n_st = 0;   // warp-uniform: first staged survivor, number of staged survivors
    slot_t aidx[PASS_KPT];         // slots of the fetch_or's in flight ...
    uint32_t old[PASS_KPT];        // ... and the words they return
    uint32_t amask = 0;            // which of them exist
#pragma unroll
    for (int j = 0; j < PASS_KPT; j++) {
        aidx[j] = 0;
        old[j] = 0;
    }
    // find_collisions (src/lib.rs:429-434), second half: the slot was taken already -> it collides
    auto settle = [&]() {
#pragma unroll
        for (int j = 0; j < PASS_KPT; j++) {
            const uint32_t m = 1u << ((uint32_t)aidx[j] & 31);
            if (((amask >> j) & 1u) && (old[j] & m)) red_or_u32(c_lv + (aidx[j] >> 5), m);
        }
        amask = 0;
    };
    // rayon filter_map().collect() (src/lib.rs:374-377): `cnt` staged keys go to level l's list (coalesced), and the
    // first half of find_collisions of level l runs for them while they are on chip
    uint64_t out_pos = 0;          // append position (region base + keys written; GATHER: set per flush)
    auto flush = [&](uint32_t cnt) {
        settle();
        bool ok = true;
        if (GATHER) {
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd((unsigned long long*)&st->redo_count, (unsigned long long)cnt);
            out_pos = __shfl_sync(0xffffffffu, t, 0);
            if (out_pos + cnt > GATHER_MAX_KEYS) {
                if (lane == 0) st->status = ST_ERR_LIST_OVERFLOW;
                ok = false;
            }
        }
        if (ok) {
#pragma unroll
            for (int j = 0; j < PASS_KPT; j++) {
                const uint32_t i = lane + 32 * j;
                if (i < cnt) {
                    const uint64_t kk = ring[(head + i) & (WP_RING - 1)];
                    dst[out_pos + i] = kk;
                    if (GATHER) continue;
                    const uint64_t idx = c_big ? level_index<true>(kc, c_sp0, kk, c_bits) : level_index<false>(kc, c_sp0, kk, c_bits);
                    aidx[j] = (slot_t)idx;
                    amask |= 1u << j;
                }
            }
            if (!GATHER) {
#pragma unroll
                for (int j = 0; j < PASS_KPT; j++)
                    if ((amask >> j) & 1u) old[j] = atomicOr(a_lv + (aidx[j] >> 5), 1u << ((uint32_t)aidx[j] & 31));
            }
        }
        out_pos += cnt;
        head = (head + cnt) & (WP_RING - 1);
        n_st -= cnt;
        __syncwarp();  // the ring entries just read may be overwritten by the next tile's survivors
    };

    // ---- TMA key ring of this warp: tile j of the current region lands in stage[j & 1]
    if (lane == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
        fence_proxy_async();
        fence_proxy_async_global();  // lists written with ordinary stores (earlier in this launch) are read by bulk copies
    }
    __syncwarp();
    uint32_t phase = 0;        // bit b: parity of mbar[b]'s next completion
    uint64_t total = 0;        // keys this warp appended (all its regions)

    const uint64_t n_gw = (uint64_t)gridDim.x * WP_WARPS;
    for (uint64_t r = (uint64_t)blockIdx.x * WP_WARPS + warp; r < R; r += n_gw) {
        const uint64_t written0 = (append && !GATHER) ? (uint64_t)__ldcg(dst_cnt + r) : 0;
        if (!GATHER) out_pos = r * reg_cap + written0;
        const uint64_t region_first = out_pos;
        for (uint32_t seg = 0; seg < x_world; seg++) {   // XSRC: one block per owner; otherwise one segment
        // tiles of region r: (pointer, valid keys) of tile j
        uint64_t n_rt;          // number of tiles
        uint64_t cnt_src = 0;   // striped: keys in the source region
        const uint64_t x_blk = ((uint64_t)r * x_world + seg) * x_sub_cap;                          // XSRC: key block (r, owner)
        const uint32_t* __restrict__ xb = XSRC ? xbits + (((uint64_t)seg * R + r) * x_sub_cap) / 32 : nullptr;  // its survivor bits
        if (XSRC) {
            cnt_src = __ldcg(x_cnt + r * x_world + seg);
            n_rt = (cnt_src + WP_TILE - 1) / WP_TILE;
        } else if (src_striped) {
            cnt_src = __ldcg(src_cnt + r);
            n_rt = (cnt_src + WP_TILE - 1) / WP_TILE;
        } else {
            n_rt = r < dense_tiles ? (dense_tiles - r + R - 1) / R : 0;
        }
        auto tile_ptr = [&](uint64_t j) -> const uint64_t* {
            return XSRC ? src + x_blk + j * WP_TILE : src_striped ? src + r * reg_cap + j * WP_TILE : src + (r + j * R) * WP_TILE;
        };
        auto tile_valid = [&](uint64_t j) -> uint32_t {
            const uint64_t left = src_striped ? cnt_src - j * WP_TILE : k_dense - (r + j * R) * WP_TILE;
            return left < (uint64_t)WP_TILE ? (uint32_t)left : (uint32_t)WP_TILE;
        };
        // a tile comes through the copy engine when it is whole (a striped region may be over-read by one key: the
        // slot exists) and 16-byte aligned; otherwise with plain loads
        auto tile_tma = [&](uint64_t j) -> bool {
            return src_striped ? true : (src_aligned && tile_valid(j) == (uint32_t)WP_TILE);
        };
        auto issue = [&](uint64_t j) {
            if (j < n_rt && tile_tma(j) && lane == 0) {
                const uint32_t bytes = ((tile_valid(j) + 1u) & ~1u) * 8u;
                mbar_expect_tx(&mbar[j & 1], bytes);
                bulk_load(stage + (j & 1) * WP_TILE, tile_ptr(j), bytes, &mbar[j & 1]);
            }
        };
        issue(0);
        issue(1);
        for (uint64_t j = 0; j < n_rt; j++) {
            const int b = (int)(j & 1);
            const uint32_t valid = tile_valid(j);
            uint64_t key[PASS_KPT];
            if (tile_tma(j)) {
                mbar_wait(&mbar[b], (phase >> b) & 1u);
                phase ^= 1u << b;
                // (slots past the region's count hold whatever was there: a stream key is an INDEX and must be a valid one)
#pragma unroll
                for (int q = 0; q < PASS_KPT; q++) key[q] = (lane + 32u * q < valid) ? stage[b * WP_TILE + lane + 32 * q] : 0;
                __syncwarp();
                if (lane == 0) fence_proxy_async();
            } else {
                const uint64_t* __restrict__ tp = tile_ptr(j);
#pragma unroll
                for (int q = 0; q < PASS_KPT; q++) key[q] = (lane + 32u * q < valid) ? ld_stream_u64(tp + lane + 32 * q) : 0;
            }
            issue(j + 2);  // stage b is free again
            // Context::filter (src/lib.rs:443-452): does this key's level-(l-1) slot collide?  All probes of the tile
            // are issued before anything else is waited for
            uint32_t cw[PASS_KPT];
            uint64_t shs = 0;  // 8 x 5 bits: position of each key's bit in its probe word
#pragma unroll
            for (int q = 0; q < PASS_KPT; q++) {
                if (XSRC) {
                    // the owner's answers for keys 32q .. 32q+31 of this tile: bit `lane` is this key's
                    uint32_t w;
                    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(w) : "l"(xb + j * PASS_KPT + q));
                    cw[q] = (lane + 32u * q < valid) ? w : 0u;
                    shs |= (uint64_t)lane << (5 * q);
                } else {
                    const uint64_t idx = p_big ? level_index<true>(kc, p_sp0, key[q], p_bits) : level_index<false>(kc, p_sp0, key[q], p_bits);
                    cw[q] = (lane + 32u * q < valid) ? __ldcg(c_pv + (idx >> 5)) : 0u;
                    shs |= (uint64_t)((uint32_t)idx & 31) << (5 * q);
                }
            }
            if (n_st >= WP_TILE) flush(WP_TILE);   // the previous tiles' survivors, while the probes are in flight
#pragma unroll
            for (int q = 0; q < PASS_KPT; q++) {
                const bool redo = (cw[q] >> ((uint32_t)(shs >> (5 * q)) & 31)) & 1u;
                const uint32_t mk = __ballot_sync(0xffffffffu, redo);
                if (redo) ring[(head + n_st + __popc(mk & lt_mask)) & (WP_RING - 1)] = key[q];
                n_st += __popc(mk);
            }
            __syncwarp();
        }
        }  // segments
        // end of the region: everything staged belongs to it
        if (n_st >= WP_TILE) flush(WP_TILE);
        if (n_st) flush(n_st);
        if (!GATHER) {
            const uint64_t wrote = out_pos - region_first;
            if (lane == 0) dst_cnt[r] = (uint32_t)(written0 + wrote);
            total += wrote;
        }
    }
    settle();
    if (!GATHER && lane == 0 && total) atomicAdd((unsigned long long*)&st->redo_count, (unsigned long long)total);
    fence_proxy_async_global();  // this warp's list stores, before a later bulk copy (async proxy) reads them
    __syncwarp();
    if (lane == 0) {
        mbar_inval(&mbar[0]);
        mbar_inval(&mbar[1]);
    }
}

template <class KC> struct PassOcc { static constexpr int ctas = BPHF_PASS_CTAS; };
template <> struct PassOcc<StreamCfg> { static constexpr int ctas = 1; };  // the stream hash walks the key's bytes: more registers

template <bool MAYBE_BIG, class KC = KeyCfg>
__global__ void __launch_bounds__(PASS_THREADS, PassOcc<KC>::ctas)
    pass_kernel(BuildState* __restrict__ st, uint32_t level, const uint64_t* __restrict__ user_keys,
                uint64_t* __restrict__ buf0, uint64_t* __restrict__ buf1, uint32_t* __restrict__ a32,
                uint32_t* __restrict__ c32) {
    if (st->status != ST_RUNNING || st->n_levels != level || level >= st->tail_level) return;
    if (st->lv[level - 1].b_range != 0) return;  // bucketed predecessor: scatter_kernel filtered it already
    if (st->lv[level - 1].x_slice_bits != 0 || st->lv[level].x_slice_bits != 0) return;  // exchange levels: xchg_kernels.cuh
    pass_body<MAYBE_BIG, false, KC>(st, level, user_keys, buf0, buf1, a32, c32);
}

// first level after the exchange levels of a sharded build: the filter reads the owners' survivor bits
template <bool MAYBE_BIG>
__global__ void __launch_bounds__(PASS_THREADS, BPHF_PASS_CTAS)
    pass_x_kernel(BuildState* __restrict__ st, uint32_t level, uint64_t* __restrict__ buf0, uint64_t* __restrict__ buf1,
                  uint32_t* __restrict__ a32, uint32_t* __restrict__ c32, const uint32_t* __restrict__ xbits) {
    if (st->status != ST_RUNNING || st->n_levels != level || level >= st->tail_level) return;
    if (st->lv[level - 1].x_slice_bits == 0 || st->lv[level].x_slice_bits != 0 || level == st->gather_level) return;
    pass_body<MAYBE_BIG, false, KeyCfg, true>(st, level, nullptr, buf0, buf1, a32, c32, nullptr, nullptr, 0, xbits);
}

// streamed ingest (host key sets that stay out of HBM): filter(0) + compaction + mark(1) over ONE chunk of the level-0
// keys; the survivors of all chunks accumulate in level 1's list, which is where the key set becomes device resident
template <bool MAYBE_BIG>
__global__ void __launch_bounds__(PASS_THREADS, BPHF_PASS_CTAS)
    pass_chunk_kernel(BuildState* __restrict__ st, const uint64_t* __restrict__ chunk, uint64_t chunk_keys,
                      uint64_t* __restrict__ buf0, uint64_t* __restrict__ buf1, uint32_t* __restrict__ a32,
                      uint32_t* __restrict__ c32) {
    if (st->status != ST_RUNNING || st->n_levels != 1 || 1 >= st->tail_level) return;
    if (st->lv[0].b_range != 0) return;
    pass_body<MAYBE_BIG>(st, 1, nullptr, buf0, buf1, a32, c32, nullptr, chunk, chunk_keys);
}

// ---------------------------------------------------------------------------------------------
// finalize: a &= ~collide; block popcounts; unique count; last CTA sets up level+1
// 4 consecutive lanes own one 8-word rank block (16 B per lane).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void finalize_body(BuildState* __restrict__ st, uint32_t level, uint64_t* __restrict__ words,
                                              const uint64_t* __restrict__ coll, uint32_t* __restrict__ blockpop) {
    __shared__ bool is_last;
    const uint64_t word_off = __ldcg(&st->lv[level].word_off);
    const uint64_t n_pairs = round_up8(__ldcg(&st->lv[level].n_words)) / 2;  // multiple of 4
    const int lane = threadIdx.x & 31;
    ulonglong2* __restrict__ a2 = reinterpret_cast<ulonglong2*>(words + word_off);
    const ulonglong2* __restrict__ c2 = reinterpret_cast<const ulonglong2*>(coll + word_off);
    uint32_t* __restrict__ bp = blockpop + word_off / 8;

    uint64_t uniq = 0;
    const uint64_t stride = (uint64_t)gridDim.x * FIN_THREADS;
    for (uint64_t p0 = (uint64_t)blockIdx.x * FIN_THREADS + (threadIdx.x & ~31u); p0 < n_pairs; p0 += stride) {
        const uint64_t p = p0 + lane;
        uint32_t pc = 0;
        if (p < n_pairs) {
            ulonglong2 a = a2[p];
            const ulonglong2 c = c2[p];
            a.x &= ~c.x;
            a.y &= ~c.y;
            a2[p] = a;
            pc = __popcll(a.x) + __popcll(a.y);
        }
        uniq += pc;
        pc += __shfl_xor_sync(0xffffffffu, pc, 1);
        pc += __shfl_xor_sync(0xffffffffu, pc, 2);
        if (p < n_pairs && (lane & 3) == 0) bp[p >> 2] = pc;
    }
    // block reduction of the unique count, one atomic per CTA
    uint64_t u = uniq;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) u += __shfl_xor_sync(0xffffffffu, u, d);
    __shared__ uint64_t warp_u[FIN_THREADS / 32];
    if (lane == 0) warp_u[threadIdx.x >> 5] = u;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
#pragma unroll
        for (int w = 0; w < FIN_THREADS / 32; w++) t += warp_u[w];
        if (t) atomicAdd((unsigned long long*)&st->uniq_count, (unsigned long long)t);
        __threadfence();
        const uint32_t ticket = atomicAdd(&st->fin_ticket, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        const uint64_t uq = atomicAdd((unsigned long long*)&st->uniq_count, 0ULL);
        const uint64_t k_in = __ldcg(&st->lv[level].k_in);
        const uint64_t listed = atomicAdd((unsigned long long*)&st->redo_count, 0ULL);
        const bool sharded = st->shard_world > 1;
        // sharded: this shard's part of the level's key list is what its filter appended (the shard sum was
        // checked against k_in at the merge barrier)
        if (sharded && level >= 1) st->lv[level].k_loc = listed;
        // plain predecessor: pass_kernel appended this level's list through redo_count, which must match k_in
        // (a bucketed predecessor is checked through the bucket cursors instead)
        const bool check_list = !sharded && level >= 1 && __ldcg(&st->lv[level - 1].b_range) == 0;
        if (uq > k_in || (check_list && listed != k_in)) {
            st->status = ST_ERR_INTERNAL;
            st->err_info = ((uint64_t)level << 48) | uq;
        } else {
            finish_level(st, level, k_in - uq);
        }
    }
}

__global__ void __launch_bounds__(FIN_THREADS)
    finalize_kernel(BuildState* __restrict__ st, uint32_t level, uint64_t* __restrict__ words,
                    const uint64_t* __restrict__ coll, uint32_t* __restrict__ blockpop) {
    if (st->status != ST_RUNNING || st->n_levels != level || level >= st->tail_level) return;
    // bucketed levels are finalized by bucket_mark_kernel, except in a sharded build (the merge comes in between)
    if (st->lv[level].b_range != 0 && st->shard_world <= 1) return;
    if (st->lv[level].x_slice_bits != 0) return;  // exchange level: xfinalize_kernel + the XDONE barrier
    finalize_body(st, level, words, coll, blockpop);
}

// ---------------------------------------------------------------------------------------------
// cascade_kernel: ONE persistent (cooperatively launched) kernel runs pass + finalize of every level from
// st->n_levels on, separated by grid-wide barriers, until the tail takes over or the device has to stop
// (buffer growth, error).  Replaces two launches per level: at 1e8 keys the 12 levels after level 0 spend most
// of their time in launch gaps and ramp-up, not in work.  Same level bodies, same result.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all CTAs of the (co-resident) grid; the fence by thread 0 also drops the SM's L1 contents
__device__ __forceinline__ void grid_barrier(BuildState* st) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t gen = ld_volatile_u32(&st->bar_gen);
        const uint32_t arrived = atomicAdd(&st->bar_count, 1u) + 1u;
        if (arrived == gridDim.x) {
            st->bar_count = 0;
            __threadfence();
            atomicAdd(&st->bar_gen, 1u);
        } else {
            while (ld_volatile_u32(&st->bar_gen) == gen) __nanosleep(100);
        }
        __threadfence();
    }
    __syncthreads();
}

// -DBPHF_TRACE: per-CTA timestamps of the cascade phases (measurement builds only; tools/cascade_trace.py)
#ifdef BPHF_TRACE
constexpr int TRACE_LEVELS = 16, TRACE_POINTS = 5, TRACE_CTAS = 1024;
__device__ uint64_t g_cascade_trace[TRACE_LEVELS][TRACE_POINTS][TRACE_CTAS];
__device__ __forceinline__ void trace_point(uint32_t level, int point) {
    if (threadIdx.x == 0 && level < TRACE_LEVELS && blockIdx.x < TRACE_CTAS) {
        uint64_t t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_cascade_trace[level][point][blockIdx.x] = t;
    }
}
#define TRACE_POINT(level, point) trace_point(level, point)
#else
#define TRACE_POINT(level, point)
#endif

template <bool MAYBE_BIG>
__global__ void __launch_bounds__(PASS_THREADS, BPHF_PASS_CTAS)
    cascade_kernel(BuildState* __restrict__ st, const uint64_t* __restrict__ user_keys, uint64_t* __restrict__ buf0,
                   uint64_t* __restrict__ buf1, uint64_t* __restrict__ words, uint64_t* __restrict__ coll,
                   uint32_t* __restrict__ blockpop) {
    __shared__ uint32_t ctl[4];
    for (;;) {
        if (threadIdx.x == 0) {
            ctl[0] = ld_volatile_u32(&st->status);
            ctl[1] = ld_volatile_u32(&st->n_levels);
            ctl[2] = ld_volatile_u32(&st->tail_level);
        }
        __syncthreads();
        const uint32_t status = ctl[0], level = ctl[1], tail = ctl[2];
        __syncthreads();
        // every CTA reads the same values here: they only change before the second barrier of an iteration
        if (status != ST_RUNNING || level == 0 || level >= tail || level > BPHF_MAX_LEVELS) break;
        // only plain levels with a plain predecessor (uniform: written before the previous grid barrier)
        if (__ldcg(&st->lv[level - 1].b_range) != 0 || __ldcg(&st->lv[level].b_range) != 0) break;
        TRACE_POINT(level, 0);
        pass_body<MAYBE_BIG>(st, level, user_keys, buf0, buf1, reinterpret_cast<uint32_t*>(words),
                             reinterpret_cast<uint32_t*>(coll));
        TRACE_POINT(level, 1);
        grid_barrier(st);
        TRACE_POINT(level, 2);
        // the pass may have stopped before writing anything (a list buffer has to grow first): uniform, set before the barrier
        if (threadIdx.x == 0) ctl[3] = ld_volatile_u32(&st->status);
        __syncthreads();
        if (ctl[3] != ST_RUNNING) break;
        finalize_body(st, level, words, coll, blockpop);
        TRACE_POINT(level, 3);
        grid_barrier(st);
        TRACE_POINT(level, 4);
    }
}

// ---------------------------------------------------------------------------------------------
// tail_kernel: one CTA, all remaining levels in shared memory
// ---------------------------------------------------------------------------------------------
struct TailSmem {
    uint64_t keys[2][TAIL_MAX_KEYS];
    uint32_t a[TAIL_MAX_WORDS * 2 + 16];
    uint32_t c[TAIL_MAX_WORDS * 2 + 16];
    uint32_t count;
    uint32_t uniq;
    uint32_t status;
};

template <class KC = KeyCfg>
__global__ void __launch_bounds__(TAIL_THREADS)
    tail_kernel(BuildState* __restrict__ st, const uint64_t* __restrict__ user_keys, const uint64_t* __restrict__ buf0,
                const uint64_t* __restrict__ buf1, uint32_t* __restrict__ words32, const uint32_t* __restrict__ c32,
                uint32_t* __restrict__ blockpop, CommDev cd) {
    extern __shared__ __align__(16) unsigned char tail_smem_raw[];
    TailSmem& s = *reinterpret_cast<TailSmem*>(tail_smem_raw);
    if (st->status != ST_RUNNING || st->tail_level == NO_TAIL || st->n_levels != st->tail_level) return;
    const uint32_t tid = threadIdx.x;
    const KC kc(st->kc);
    uint32_t level = st->tail_level;
    uint32_t k = (uint32_t)st->lv[level].k_in;
    if (tid == 0) s.count = 0;
    __syncthreads();
    // gather the keys entering `level`
    if (st->shard_world > 1) {
        // sharded build: every shard left its part in its exchange slot (tail_gather_kernel + barrier); all
        // shards read all slots over peer memory and finish the cascade redundantly -> identical replicas
        uint32_t base = 0;
        for (uint32_t p = 0; p < cd.world; p++) {
            const volatile uint64_t* xk = cd.xkeys[p];
            const uint32_t cnt = (uint32_t)min(xk[0], (uint64_t)TAIL_MAX_KEYS);
            for (uint32_t i = tid; i < cnt; i += TAIL_THREADS)
                if (base + i < TAIL_MAX_KEYS) s.keys[0][base + i] = xk[8 + i];
            base += cnt;
        }
        if (tid == 0) s.count = base;
    } else if (level == 0) {
        for (uint32_t i = tid; i < k; i += TAIL_THREADS) s.keys[0][i] = user_keys[i];
        if (tid == 0) s.count = k;
    } else if (st->lv[level - 1].b_range != 0) {
        // bucketed predecessor: scatter_kernel already wrote this level's (filtered) key list to buf[level & 1]
        const uint64_t* src = (level & 1) ? buf1 : buf0;
        for (uint32_t i = tid; i < k; i += TAIL_THREADS) s.keys[0][i] = src[i];
        if (tid == 0) s.count = k;
    } else {
        const LevelDesc* pv = &st->lv[level - 1];
        const uint64_t p_bits = pv->bits, p_sp0 = pv->sp0, p_off32 = pv->word_off * 2, k_src = pv->k_in;
        const bool p_big = (p_bits >> 32) != 0;
        const uint64_t* src = (level == 1) ? user_keys : (((level - 1) & 1) ? buf1 : buf0);
        auto take = [&](uint64_t key) {
            const uint64_t idx = p_big ? hashmod_big(kc, p_sp0, key, p_bits) : (uint64_t)hashmod_small(kc, p_sp0, key, (uint32_t)p_bits);
            if ((c32[p_off32 + (idx >> 5)] >> (idx & 31)) & 1u) {
                const uint32_t pos = atomicAdd(&s.count, 1u);
                if (pos < TAIL_MAX_KEYS) s.keys[0][pos] = key;
            }
        };
        if (list_is_striped(st, level - 1)) {
            // the previous filter pass left level-1's keys in regions (pass_body): a few keys per region by now
            const uint32_t R = st->n_regions;
            const uint64_t reg_cap = st->reg_cap;
            const uint32_t* cnt = st->reg_cnt + (size_t)((level - 1) & 1) * R;
            for (uint32_t r = tid; r < R; r += TAIL_THREADS) {
                const uint32_t n_r = cnt[r];
                for (uint32_t i = 0; i < n_r; i++) take(src[(uint64_t)r * reg_cap + i]);
            }
        } else {
            for (uint64_t i = tid; i < k_src; i += TAIL_THREADS) take(src[i]);
        }
    }
    __syncthreads();
    if (s.count != k) {
        if (tid == 0) { st->status = ST_ERR_INTERNAL; st->err_info = ((uint64_t)level << 48) | s.count | (1ULL << 47); }
        return;
    }
    int cur = 0;
    for (;;) {
        const uint32_t bits = (uint32_t)st->lv[level].bits;
        const uint64_t sp0 = st->lv[level].sp0;
        const uint64_t word_off = st->lv[level].word_off;
        const uint32_t n32 = (uint32_t)round_up8(st->lv[level].n_words) * 2;
        for (uint32_t w = tid; w < n32; w += TAIL_THREADS) { s.a[w] = 0; s.c[w] = 0; }
        if (tid == 0) { s.count = 0; s.uniq = 0; }
        __syncthreads();
        const uint64_t* kin = s.keys[cur];
        uint64_t* kout = s.keys[cur ^ 1];
        for (uint32_t i = tid; i < k; i += TAIL_THREADS) {
            const uint32_t idx = hashmod_small(kc, sp0, kin[i], bits);
            const uint32_t m = 1u << (idx & 31);
            const uint32_t old = atomicOr(&s.a[idx >> 5], m);
            if (old & m) atomicOr(&s.c[idx >> 5], m);
        }
        __syncthreads();
        for (uint32_t w = tid; w < n32; w += TAIL_THREADS) {
            const uint32_t a = s.a[w] & ~s.c[w];
            s.a[w] = a;
            words32[word_off * 2 + w] = a;
        }
        __syncthreads();
        for (uint32_t b = tid; b < n32 / 16; b += TAIL_THREADS) {
            uint32_t pc = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) pc += __popc(s.a[b * 16 + j]);
            blockpop[word_off / 8 + b] = pc;
            if (pc) atomicAdd(&s.uniq, pc);
        }
        for (uint32_t i = tid; i < k; i += TAIL_THREADS) {
            const uint64_t key = kin[i];
            const uint32_t idx = hashmod_small(kc, sp0, key, bits);
            if ((s.c[idx >> 5] >> (idx & 31)) & 1u) kout[atomicAdd(&s.count, 1u)] = key;
        }
        __syncthreads();
        const uint32_t k_next = s.count;
        if (tid == 0) {
            if (s.uniq + k_next != k) {
                st->status = ST_ERR_INTERNAL;
                st->err_info = ((uint64_t)level << 48) | k_next | (1ULL << 46);
            } else {
                finish_level(st, level, k_next);
            }
            s.status = st->status;
        }
        __syncthreads();
        if (s.status != ST_RUNNING) return;
        k = k_next;
        level += 1;
        cur ^= 1;
    }
}

// ---------------------------------------------------------------------------------------------
// compute_ranks (src/lib.rs:277-297): exclusive prefix sum over the arena's block popcounts
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER_THREAD = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ uint64_t block_reduce_u64(uint64_t v, uint64_t* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    uint64_t t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sh[w];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(SCAN_THREADS)
    scan_sums_kernel(const uint32_t* __restrict__ blockpop, uint64_t n, uint64_t* __restrict__ chunk_sums) {
    __shared__ uint64_t sh[SCAN_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK;
    uint64_t v = 0;
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; j++) {
        const uint64_t i = base + (uint64_t)j * SCAN_THREADS + threadIdx.x;
        if (i < n) v += blockpop[i];
    }
    const uint64_t t = block_reduce_u64(v, sh);
    if (threadIdx.x == 0) chunk_sums[blockIdx.x] = t;
}

// single CTA: in-place exclusive scan of chunk sums
__global__ void __launch_bounds__(1024) scan_top_kernel(uint64_t* __restrict__ chunk_sums, uint64_t n_chunks) {
    __shared__ uint64_t sh[32];
    __shared__ uint64_t carry_sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n_chunks; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint64_t v = (i < n_chunks) ? chunk_sums[i] : 0;
        uint64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) sh[warp] = incl;
        __syncthreads();
        uint64_t woff = 0, total = 0;
        for (int w = 0; w < 32; w++) {
            const uint64_t t = sh[w];
            if (w < warp) woff += t;
            total += t;
        }
        const uint64_t carry = carry_sh;
        if (i < n_chunks) chunk_sums[i] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_sh = carry + total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
    scan_apply_kernel(const uint32_t* __restrict__ blockpop, uint64_t n, const uint64_t* __restrict__ chunk_off,
                      uint64_t* __restrict__ ranks) {
    __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // thread t owns SCAN_PER_THREAD consecutive entries
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK + (uint64_t)threadIdx.x * SCAN_PER_THREAD;
    uint32_t v[SCAN_PER_THREAD];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; j++) {
        v[j] = (base + j < n) ? blockpop[base + j] : 0;
        sum += v[j];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; w++) woff += warp_tot[w];
    uint64_t run = chunk_off[blockIdx.x] + woff + incl - sum;
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; j++) {
        if (base + j < n) ranks[base + j] = run;
        run += v[j];
    }
}

}  // namespace bphf
