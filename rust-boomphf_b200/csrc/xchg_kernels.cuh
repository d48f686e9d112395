// xchg_kernels.cuh -- sharded construction of the levels that do not fit the L2: "owner computes".
//
// The OR-collide merge (shard_kernels.cuh) has every shard mark its keys into a FULL-WIDTH (a, collide) pair and then
// folds the pairs over NVLink.  That is right while the pair fits the L2; beyond it (8 x 1e8 keys: 2 x 170 MB per GPU
// for level 0) the marks miss to HBM, the keys have to be grouped into thousands of shared-memory buckets first, and
// the per-GPU bytes grow with the number of GPUs -- weak-scaling efficiency 0.37 at 8 GPUs.  For those
// levels the roles are turned around, and the per-GPU working set becomes what a single-GPU build of n/G keys has:
//
//   shard g OWNS bits [g * slice, (g + 1) * slice) of the level; its slice of (a, collide) stays L2 resident.
//   xscatter : every shard hashes its keys for the level, routes each key to the owner of its slot, keeps the key in
//              a local list GROUPED BY OWNER and pushes the slot's offset within the owner's slice (4 bytes) into the
//              owner's receive buffer over NVLink -- same position in both, so no key ever crosses the link.
//   (barrier)
//   xmark    : the owner runs find_collisions (src/lib.rs:429-434) over everything it received: one fetch_or per
//              offset on its slice, collide update for late arrivals -- the single-GPU level-0 kernel, fed offsets.
//   xfinalize: a &= ~collide on the slice (Context::filter's a.remove, src/lib.rs:447), unique count, and the final
//              slice of a is stored into EVERY shard's arena (the all-gather half of the old merge; collide never
//              leaves the owner).
//   xbits    : for every received offset the owner looks up the collide bit (Context::filter, src/lib.rs:443-452) and
//              returns ONE BIT per key to the sender, in arrival order.
//   (barrier, carrying every owner's unique count: k(l+1) = k(l) - sum, derived identically on every shard)
//   The next level's entry step (xscatter again, or pass_body / the collect of the merge path) streams the grouped keys
//   next to the returned bits: the filter costs no random access at all on the sender.
//
// NVLink traffic per level: 4 bytes per key out + 1 bit per key back + the a slices (W * 8 * (G-1)/G in per GPU),
// against 16 W * (G-1)/G in AND out for the fold.  Results are identical (same (cnt >= 1, cnt >= 2) per slot).
//
// Layout: sender s, region r (the warp-private regions of pass_body), owner o form one block of x_sub_cap entries:
//   keys  : sender's buf[l & 1] + ((r * world + o) * x_sub_cap)           (8 B, local)
//   offsets: owner o's xidx    + ((s * R + r) * x_sub_cap)                (4 B, pushed by the sender)
//   counts : owner o's xcnt[s * R + r]  and the sender's xreg_cnt[l & 1][r * world + o]
//   bits  : sender's xbits     + ((o * R + r) * x_sub_cap) / 32           (written by owner o)
// Blocks are private to one warp on the sending side: no cursor is shared, nothing is reserved atomically.
#pragma once
#include "shard_kernels.cuh"

namespace bphf {

constexpr int XS_STAGE = 64;  // keys staged per owner and warp (power of two); flushed 32 at a time

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// ---------------------------------------------------------------------------------------------
// xscatter_kernel: route this shard's keys of an exchange level to the owners of their slots
// ---------------------------------------------------------------------------------------------
// Per warp, in shared memory: for every owner a ring of XS_STAGE (key, offset) pairs + its fill, head and the number
// of entries already written to the owner's block.  A step takes 32 keys (one per lane): __match_any groups the lanes
// by owner, each group appends to its owner's ring, and an owner with >= 32 staged pairs is flushed: 32 keys to the
// local block (256 B), 32 offsets to the owner's receive block over NVLink (128 B).
constexpr size_t XS_SMEM_BYTES = PASS_SMEM_BYTES + (size_t)WP_WARPS * (SHARD_MAX * XS_STAGE * 4 + 3 * SHARD_MAX * 4);

template <bool MAYBE_BIG>
__global__ void __launch_bounds__(PASS_THREADS, BPHF_PASS_CTAS)
    xscatter_kernel(BuildState* __restrict__ st, uint32_t level, const uint64_t* __restrict__ user_keys,
                    uint64_t* __restrict__ buf0, uint64_t* __restrict__ buf1, CommDev cd) {
    if (st->status != ST_RUNNING || st->n_levels != level || level >= st->tail_level) return;
    const LevelDesc* cu = &st->lv[level];
    const uint64_t slice_bits = cu->x_slice_bits;
    if (slice_bits == 0) return;
    extern __shared__ __align__(128) unsigned char pass_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // carve-up as in pass_body (4 KB key staging, two 2 KB key tiles, two mbarriers) + offsets and ring state
    uint64_t* __restrict__ stg_key = reinterpret_cast<uint64_t*>(pass_smem) + warp * WP_RING;
    uint64_t* __restrict__ stage = reinterpret_cast<uint64_t*>(pass_smem) + WP_WARPS * WP_RING + warp * 2 * WP_TILE;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(pass_smem) + WP_WARPS * (WP_RING + 2 * WP_TILE) + warp * 2;
    uint32_t* __restrict__ stg_loc = reinterpret_cast<uint32_t*>(pass_smem + PASS_SMEM_BYTES) + warp * (SHARD_MAX * XS_STAGE);
    uint32_t* __restrict__ ring_st =
        reinterpret_cast<uint32_t*>(pass_smem + PASS_SMEM_BYTES) + WP_WARPS * (SHARD_MAX * XS_STAGE) + warp * (3 * SHARD_MAX);
    uint32_t* __restrict__ r_cnt = ring_st;                   // staged pairs per owner
    uint32_t* __restrict__ r_head = ring_st + SHARD_MAX;      // ring head per owner
    uint32_t* __restrict__ r_pos = ring_st + 2 * SHARD_MAX;   // entries written to the (region, owner) block
    static_assert(SHARD_MAX * XS_STAGE <= WP_RING, "per-owner key staging must fit the warp's ring");

    const KeyCfg kc(st->kc);
    const uint64_t c_bits = cu->bits, c_sp0 = cu->sp0;
    const bool c_big = MAYBE_BIG && (c_bits >> 32) != 0;
    const uint32_t R = st->n_regions, world = cd.world, me = cd.rank, sub_cap = st->x_sub_cap;
    // owner of a slot below 2^32: floor(slot / slice) through a reciprocal that never overestimates + a correction
    // (slices are below 2^32 bits by the host's choice of exchange levels)
    const uint32_t slice32 = (uint32_t)min(slice_bits, (uint64_t)0xffffffffu);
    const uint32_t slice_inv = (uint32_t)(0xffffffffull / slice32);
    const bool xsrc = level >= 1;  // exchange levels are a prefix: the predecessor is one, too
    const uint64_t* __restrict__ src = xsrc ? (((level - 1) & 1) ? buf1 : buf0) : user_keys;
    const uint64_t k_dense = st->lv[0].k_loc;
    const uint64_t dense_tiles = (k_dense + WP_TILE - 1) / WP_TILE;
    const uint32_t* __restrict__ src_cnt = st->xreg_cnt + (size_t)((level - 1) & 1) * R * SHARD_MAX;
    const uint32_t* __restrict__ src_bits = cd.xbits[me];
    uint64_t* __restrict__ dst = (level & 1) ? buf1 : buf0;
    uint32_t* __restrict__ dst_cnt = st->xreg_cnt + (size_t)(level & 1) * R * SHARD_MAX;
    const bool src_aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    const uint32_t lt_mask = (1u << lane) - 1u;

    if (lane == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
        fence_proxy_async();
        fence_proxy_async_global();
    }
    __syncwarp();
    uint32_t phase = 0;
    uint64_t total = 0;
    bool overflow = false;

    const uint64_t n_gw = (uint64_t)gridDim.x * WP_WARPS;
    for (uint64_t r = (uint64_t)blockIdx.x * WP_WARPS + warp; r < R; r += n_gw) {
        if (lane < 3 * SHARD_MAX) ring_st[lane] = 0;
        __syncwarp();
        // `cnt` staged pairs of owner o -> block (r, o): the key stays here, its offset in o's slice goes to o
        auto flush = [&](uint32_t o, uint32_t cnt) {
            const uint32_t h = r_head[o], pos = r_pos[o];
            if (pos + cnt > sub_cap) {
                overflow = true;  // more keys for one owner than the block holds (grossly unbalanced shards / hash)
            } else if ((uint32_t)lane < cnt) {
                const uint32_t e = o * XS_STAGE + ((h + lane) & (XS_STAGE - 1));
                dst[((uint64_t)r * world + o) * sub_cap + pos + lane] = stg_key[e];
                cd.xidx[o][((uint64_t)me * R + r) * sub_cap + pos + lane] = stg_loc[e];
            }
            __syncwarp();
            if (lane == 0) {
                r_head[o] = (h + cnt) & (XS_STAGE - 1);
                r_cnt[o] -= cnt;
                r_pos[o] = pos + cnt;
            }
            __syncwarp();
        };
        const uint32_t n_seg = xsrc ? world : 1u;
        for (uint32_t seg = 0; seg < n_seg; seg++) {
            // tiles of this segment: the shard's share of the input (level 0), or block (r, seg) of the previous level
            uint64_t n_rt, cnt_src = 0;
            if (xsrc) {
                cnt_src = ld_cg_u32(src_cnt + r * world + seg);
                n_rt = (cnt_src + WP_TILE - 1) / WP_TILE;
            } else {
                n_rt = r < dense_tiles ? (dense_tiles - r + R - 1) / R : 0;
            }
            const uint64_t blk = ((uint64_t)r * world + seg) * sub_cap;                                // xsrc: key block
            const uint32_t* __restrict__ bits = src_bits + (((uint64_t)seg * R + r) * sub_cap) / 32;  // xsrc: its survivor bits
            auto tile_ptr = [&](uint64_t j) -> const uint64_t* {
                return xsrc ? src + blk + j * WP_TILE : src + (r + j * R) * WP_TILE;
            };
            auto tile_valid = [&](uint64_t j) -> uint32_t {
                const uint64_t left = xsrc ? cnt_src - j * WP_TILE : k_dense - (r + j * R) * WP_TILE;
                return left < (uint64_t)WP_TILE ? (uint32_t)left : (uint32_t)WP_TILE;
            };
            auto tile_tma = [&](uint64_t j) -> bool { return xsrc ? true : (src_aligned && tile_valid(j) == (uint32_t)WP_TILE); };
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
                // Context::filter of the previous level: the owner's answers, one bit per key in arrival order
                uint32_t bw[PASS_KPT];
#pragma unroll
                for (int q = 0; q < PASS_KPT; q++) bw[q] = xsrc ? ld_cg_u32(bits + j * PASS_KPT + q) : 0xffffffffu;
                uint64_t key[PASS_KPT];
                if (tile_tma(j)) {
                    mbar_wait(&mbar[b], (phase >> b) & 1u);
                    phase ^= 1u << b;
#pragma unroll
                    for (int q = 0; q < PASS_KPT; q++) key[q] = stage[b * WP_TILE + lane + 32 * q];
                    __syncwarp();
                    if (lane == 0) fence_proxy_async();
                } else {
                    const uint64_t* __restrict__ tp = tile_ptr(j);
#pragma unroll
                    for (int q = 0; q < PASS_KPT; q++) key[q] = (lane + 32u * q < valid) ? ld_stream_u64(tp + lane + 32 * q) : 0;
                }
                issue(j + 2);  // this tile's buffer is free again
#pragma unroll
                for (int q = 0; q < PASS_KPT; q++) {
                    const bool live = (lane + 32u * q < valid) && ((bw[q] >> lane) & 1u);
                    uint32_t own = 0xffu, loc = 0;
                    if (live) {
                        if (c_big) {
                            const uint64_t slot = level_index<true>(kc, c_sp0, key[q], c_bits);
                            own = (uint32_t)(slot / slice_bits);
                            loc = (uint32_t)(slot - (uint64_t)own * slice_bits);
                        } else {
                            const uint32_t slot = (uint32_t)level_index<false>(kc, c_sp0, key[q], c_bits);
                            own = __umulhi(slot, slice_inv);
                            while (slot - own * slice32 >= slice32) own++;  // at most two steps
                            loc = slot - own * slice32;
                        }
                    }
                    const uint32_t grp = __match_any_sync(0xffffffffu, own);
                    if (live) {
                        const uint32_t e = own * XS_STAGE + ((r_head[own] + r_cnt[own] + __popc(grp & lt_mask)) & (XS_STAGE - 1));
                        stg_key[e] = key[q];
                        stg_loc[e] = loc;
                    }
                    __syncwarp();
                    if (live && (grp & lt_mask) == 0) r_cnt[own] += __popc(grp);  // the group's first lane
                    __syncwarp();
                    uint32_t full = __ballot_sync(0xffffffffu, (uint32_t)lane < world && r_cnt[lane & (SHARD_MAX - 1)] >= 32);
                    while (full) {
                        const uint32_t o = __ffs(full) - 1;
                        full &= full - 1;
                        flush(o, 32);
                    }
                }
            }
        }
        // end of the region: the rest of every owner's ring, then the block counts (here and at the owner)
        for (uint32_t o = 0; o < world; o++) {
            const uint32_t left = r_cnt[o];
            if (left) flush(o, left);
            const uint32_t wrote = r_pos[o];
            if (lane == 0) {
                dst_cnt[r * world + o] = overflow ? 0u : wrote;
                cd.xcnt[o][(size_t)me * R + r] = overflow ? 0u : wrote;
            }
            total += wrote;
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (overflow) st->status = ST_ERR_LIST_OVERFLOW;
        if (total) atomicAdd((unsigned long long*)&st->redo_count, (unsigned long long)total);
    }
    fence_proxy_async_global();
    __threadfence_system();  // offsets and counts must be performed at the owners before the barrier kernel signals
    __syncwarp();
    if (lane == 0) {
        mbar_inval(&mbar[0]);
        mbar_inval(&mbar[1]);
    }
}

// (sender, region) blocks of the offsets this owner received, dealt to the warps round robin
template <class F>
__device__ __forceinline__ void for_each_received_tile(const CommDev& cd, uint32_t R, uint32_t sub_cap, F&& body) {
    const int lane = threadIdx.x & 31;
    const uint64_t n_blocks = (uint64_t)cd.world * R;
    const uint64_t n_gw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t* __restrict__ idx = cd.xidx[cd.rank];
    const uint32_t* __restrict__ cnt = cd.xcnt[cd.rank];
    for (uint64_t blk = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); blk < n_blocks; blk += n_gw) {
        const uint32_t n = ld_cg_u32(cnt + blk);
        const uint32_t* __restrict__ base = idx + blk * sub_cap;
        for (uint32_t i0 = 0; i0 < n; i0 += WP_TILE) {
            uint32_t loc[PASS_KPT];
            uint32_t live = 0;
#pragma unroll
            for (int q = 0; q < PASS_KPT; q++) {
                const uint32_t i = i0 + lane + 32 * q;
                loc[q] = 0;
                if (i < n) {
                    loc[q] = ld_cg_u32(base + i);
                    live |= 1u << q;
                }
            }
            body(blk, i0, loc, live);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// xmark_kernel: find_collisions over the offsets received from all shards, on this owner's slice
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PASS_THREADS)
    xmark_kernel(BuildState* __restrict__ st, uint32_t level, CommDev cd, uint32_t* __restrict__ a32, uint32_t* __restrict__ c32) {
    if (st->status != ST_RUNNING || st->n_levels != level || level >= st->tail_level) return;
    const LevelDesc* cu = &st->lv[level];
    const uint64_t slice_bits = cu->x_slice_bits;
    if (slice_bits == 0) return;
    const uint64_t off32 = cu->word_off * 2 + (uint64_t)cd.rank * (slice_bits / 32);
    uint32_t* __restrict__ a_sl = a32 + off32;
    uint32_t* __restrict__ c_sl = c32 + off32;
    for_each_received_tile(cd, st->n_regions, st->x_sub_cap, [&](uint64_t, uint32_t, const uint32_t (&loc)[PASS_KPT], uint32_t live) {
        uint32_t old[PASS_KPT];
#pragma unroll
        for (int q = 0; q < PASS_KPT; q++) {
            old[q] = 0;
            if ((live >> q) & 1u) old[q] = atomicOr(a_sl + (loc[q] >> 5), 1u << (loc[q] & 31));
        }
#pragma unroll
        for (int q = 0; q < PASS_KPT; q++) {
            const uint32_t m = 1u << (loc[q] & 31);
            if (((live >> q) & 1u) && (old[q] & m)) red_or_u32(c_sl + (loc[q] >> 5), m);
        }
    });
}

// ---------------------------------------------------------------------------------------------
// xfinalize_kernel: a &= ~collide on the owner's slice + unique count (local);
// xpush_kernel: the final slice of a -> every other shard's arena.  The push runs on a side stream, next to xbits and the
// levels that follow: nobody needs another owner's slice before the build is over (lookups use the complete replica).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FIN_THREADS)
    xfinalize_kernel(BuildState* __restrict__ st, uint32_t level, CommDev cd) {
    __shared__ uint64_t warp_u[FIN_THREADS / 32];
    if (st->status != ST_RUNNING || st->n_levels != level || level >= st->tail_level) return;
    const LevelDesc* cu = &st->lv[level];
    const uint64_t slice_bits = cu->x_slice_bits;
    if (slice_bits == 0) return;
    const uint64_t pair_off = cu->word_off / 2;
    const uint64_t n_pairs = round_up8(cu->n_words) / 2;
    const uint64_t per = slice_bits / 128;
    const uint64_t p0 = min(n_pairs, per * cd.rank), p1 = min(n_pairs, p0 + per);
    ulonglong2* __restrict__ a2 = reinterpret_cast<ulonglong2*>(cd.words[cd.rank]) + pair_off;
    const ulonglong2* __restrict__ c2 = reinterpret_cast<const ulonglong2*>(cd.coll[cd.rank]) + pair_off;
    uint64_t u = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = p0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < p1; p += stride) {
        ulonglong2 a = a2[p];
        const ulonglong2 c = c2[p];
        a.x &= ~c.x;
        a.y &= ~c.y;
        u += __popcll(a.x) + __popcll(a.y);
        a2[p] = a;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) u += __shfl_xor_sync(0xffffffffu, u, d);
    if ((threadIdx.x & 31) == 0) warp_u[threadIdx.x >> 5] = u;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
#pragma unroll
        for (int w = 0; w < FIN_THREADS / 32; w++) t += warp_u[w];
        if (t) atomicAdd((unsigned long long*)&st->uniq_count, (unsigned long long)t);
    }
}

// (runs concurrently with later kernels of the build stream: it only reads level `level`'s descriptor, which is final)
__global__ void __launch_bounds__(FIN_THREADS)
    xpush_kernel(const BuildState* __restrict__ st, uint32_t level, CommDev cd) {
    const LevelDesc* cu = &st->lv[level];
    const uint64_t slice_bits = cu->x_slice_bits;
    if (slice_bits == 0 || st->n_levels < level) return;
    const uint64_t pair_off = cu->word_off / 2;
    const uint64_t n_pairs = round_up8(cu->n_words) / 2;
    const uint64_t per = slice_bits / 128;
    const uint64_t p0 = min(n_pairs, per * cd.rank), p1 = min(n_pairs, p0 + per);
    const ulonglong2* __restrict__ a2 = reinterpret_cast<const ulonglong2*>(cd.words[cd.rank]) + pair_off;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = p0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < p1; p += stride) {
        const ulonglong2 a = a2[p];
        for (uint32_t r = 0; r < cd.world; r++)
            if (r != cd.rank) reinterpret_cast<ulonglong2*>(cd.words[r])[pair_off + p] = a;
    }
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------------
// xbits_kernel: Context::filter for everything received -- one bit per key back to its sender, in arrival order
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PASS_THREADS)
    xbits_kernel(BuildState* __restrict__ st, uint32_t level, CommDev cd, const uint32_t* __restrict__ c32) {
    if (st->status != ST_RUNNING || st->n_levels != level || level >= st->tail_level) return;
    const LevelDesc* cu = &st->lv[level];
    const uint64_t slice_bits = cu->x_slice_bits;
    if (slice_bits == 0) return;
    const uint32_t* __restrict__ c_sl = c32 + cu->word_off * 2 + (uint64_t)cd.rank * (slice_bits / 32);
    const uint32_t R = st->n_regions, sub_cap = st->x_sub_cap;
    const int lane = threadIdx.x & 31;
    for_each_received_tile(cd, R, sub_cap, [&](uint64_t blk, uint32_t i0, const uint32_t (&loc)[PASS_KPT], uint32_t live) {
        uint32_t cw[PASS_KPT];
#pragma unroll
        for (int q = 0; q < PASS_KPT; q++) cw[q] = ((live >> q) & 1u) ? __ldcg(c_sl + (loc[q] >> 5)) : 0u;
        uint32_t mine = 0;
#pragma unroll
        for (int q = 0; q < PASS_KPT; q++) {
            const uint32_t w = __ballot_sync(0xffffffffu, (cw[q] >> (loc[q] & 31)) & 1u);
            if (lane == q) mine = w;
        }
        // block blk = (sender s, region r) of this owner -> the sender's bits for (owner me, region r)
        const uint32_t s = (uint32_t)(blk / R), r = (uint32_t)(blk % R);
        if (lane < PASS_KPT) cd.xbits[s][(((uint64_t)cd.rank * R + r) * sub_cap + i0) / 32 + lane] = mine;
    });
    __threadfence_system();
}

// ---------------------------------------------------------------------------------------------
// xpop_kernel: per-512-bit popcounts (compute_ranks, src/lib.rs:277-297) of an exchange level, once every owner's
// final slice has arrived (each shard keeps a complete replica and scans its own ranks)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FIN_THREADS)
    xpop_kernel(const BuildState* __restrict__ st, uint32_t level, const uint64_t* __restrict__ words, uint32_t* __restrict__ blockpop) {
    // runs after the level is complete: n_levels has moved on, the descriptor stays
    if (level >= st->n_levels || st->lv[level].x_slice_bits == 0) return;
    if (st->status != ST_RUNNING && st->status != ST_DONE && st->status != ST_NEED_WORDS && st->status != ST_NEED_LIST) return;
    const uint64_t word_off = st->lv[level].word_off;
    const uint64_t n_pairs = round_up8(st->lv[level].n_words) / 2;
    const ulonglong2* __restrict__ a2 = reinterpret_cast<const ulonglong2*>(words + word_off);
    uint32_t* __restrict__ bp = blockpop + word_off / 8;
    const int lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * FIN_THREADS;
    for (uint64_t p0 = (uint64_t)blockIdx.x * FIN_THREADS + (threadIdx.x & ~31u); p0 < n_pairs; p0 += stride) {
        const uint64_t p = p0 + lane;
        uint32_t pc = 0;
        if (p < n_pairs) {
            const ulonglong2 a = a2[p];
            pc = __popcll(a.x) + __popcll(a.y);
        }
        pc += __shfl_xor_sync(0xffffffffu, pc, 1);
        pc += __shfl_xor_sync(0xffffffffu, pc, 2);
        if (p < n_pairs && (lane & 3) == 0) bp[p >> 2] = pc;
    }
}

}  // namespace bphf
