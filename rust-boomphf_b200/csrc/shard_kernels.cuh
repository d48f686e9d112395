// shard_kernels.cuh -- multi-GPU construction: keys sharded across GPUs, level bit vectors merged
// with the OR-collide monoid over NVLink peer memory.
//
// The reference has one shared (a, collide) pair per level that all rayon workers update with fetch_or
// (/root/reference/src/lib.rs:373-377, src/bitvector.rs:253-259).  Here every shard marks ITS keys into its own
// full-width pair (a_g, c_g); the pairs are then folded with
//        (a, c) (+) (a', c') = (a | a', c | c' | (a & a'))            identity (0, 0)
// which is associative and commutative and yields exactly a = (cnt >= 1), collide = (cnt >= 2) of the union of the
// shards -- what the shared-memory fetch_or schedule produces.  NCCL has no such operator (c depends on both a's),
// so the fold is a peer-memory kernel: shard g owns words [g W/G, (g+1) W/G) of the level, LOADS that slice of
// every shard's pair over NVLink (reduce-scatter), folds, and STORES the merged collide and the final
// a & ~collide (the net effect of Context::filter's a.remove, src/lib.rs:447) into every shard's arena
// (all-gather).  After the second barrier every shard holds the complete merged level, filters its own keys
// against it and derives the same k(l+1) from the same popcount: no scalar all-reduce is needed.
//
// Synchronisation between shards is on the device: one flag line per (reader, writer) pair in peer memory,
// release/acquire at system scope, bounded by a timeout so a lost peer ends in an error instead of a hang.
#pragma once
#include "build_kernels.cuh"

namespace bphf {

// A shard that never arrives ends the build with ST_ERR_COMM instead of a hang.  The bound has to cover whatever a peer
// may legitimately still be doing when this shard reaches its first barrier -- the H2D copy of its keys, module load,
// a late host thread -- so it is generous (30 s; BPHF_BARRIER_TIMEOUT_MS overrides it at bphf_comm_create).  A peer
// that FAILS does not cost the timeout: it still signals, with its error status.
constexpr uint64_t BARRIER_TIMEOUT_NS_DEFAULT = 30000000000ULL;

__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// what a barrier carries besides "I am here"
enum BarrierKind : uint32_t {
    BAR_PLAIN = 0,
    BAR_LIST_SUM = 1,  // payload = keys this shard listed for `level`; the shard sum must equal lv[level].k_in
    // end of an exchange level (xchg_kernels.cuh): payload = unique slots of this owner's slice, payload2 = keys this shard
    // sent; every shard derives k(level + 1) = k(level) - sum(payload) and sets up the next level
    BAR_XDONE = 2,
};

// One warp.  Lane p signals shard p (epoch + payload + this shard's status into p's flag line for this shard) and
// then waits for shard p's signal in the own flag array.  Always signals, whatever the local status, so that an
// error on one shard reaches the others as a status instead of a timeout.
__global__ void __launch_bounds__(32)
    shard_barrier_kernel(BuildState* __restrict__ st, CommDev cd, uint64_t epoch, uint32_t kind, uint32_t level) {
    const uint32_t p = threadIdx.x;
    const uint32_t my_status = st->status;
    bool live = my_status == ST_RUNNING && st->n_levels == level && level < st->tail_level;
    const bool x_level = live && st->lv[level].x_slice_bits != 0;
    if (kind == BAR_XDONE) live = x_level;          // the plan enqueues both variants of a level; only one applies
    else if (kind == BAR_LIST_SUM && x_level) live = false;
    uint64_t payload = 0, payload2 = 0;
    if (kind == BAR_LIST_SUM && live) payload = (level == 0) ? st->lv[0].k_loc : st->redo_count;
    if (kind == BAR_XDONE && live) {
        payload = st->uniq_count;
        payload2 = st->redo_count;
    }
    uint64_t peer_payload = 0, peer_payload2 = 0, peer_status = ST_RUNNING;
    bool timeout = false;
    if (p < cd.world) {
        uint64_t* dst = cd.flags[p] + (size_t)cd.rank * FLAG_STRIDE;
        st_relaxed_sys(dst + 1, payload);
        st_relaxed_sys(dst + 3, payload2);
        st_relaxed_sys(dst + 2, (uint64_t)((my_status == ST_RUNNING || my_status == ST_DONE) ? ST_RUNNING : my_status));
        __threadfence_system();
        st_release_sys(dst, epoch);
        const uint64_t* src = cd.flags[cd.rank] + (size_t)p * FLAG_STRIDE;
        const uint64_t t0 = global_timer_ns();
        while (ld_acquire_sys(src) < epoch) {
            if (global_timer_ns() - t0 > cd.barrier_timeout_ns) {
                timeout = true;
                break;
            }
            __nanosleep(200);
        }
        if (!timeout) {
            peer_payload = ld_relaxed_sys(src + 1);
            peer_payload2 = ld_relaxed_sys(src + 3);
            peer_status = ld_relaxed_sys(src + 2);
        }
    }
    if (kind == BAR_LIST_SUM && live && p < cd.world) st->peer_count[p] = peer_payload;
    const bool bad = timeout || peer_status != ST_RUNNING;
    const uint32_t any_bad = __ballot_sync(0xffffffffu, bad);
    uint64_t sum = peer_payload, sum2 = peer_payload2;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        sum2 += __shfl_xor_sync(0xffffffffu, sum2, d);
    }
    if (p == 0 && st->status == ST_RUNNING) {
        if (any_bad) {
            st->status = ST_ERR_COMM;
            st->err_info = ((uint64_t)level << 48) | any_bad;
        } else if (kind == BAR_LIST_SUM && live && sum != st->lv[level].k_in) {
            st->status = ST_ERR_INTERNAL;
            st->err_info = ((uint64_t)level << 48) | (sum & 0xffffffffffULL) | (1ULL << 44);
        } else if (kind == BAR_XDONE && live) {
            const uint64_t k_in = st->lv[level].k_in;
            if (sum2 != k_in || sum > k_in) {
                st->status = ST_ERR_INTERNAL;
                st->err_info = ((uint64_t)level << 48) | (sum2 & 0xffffffffffULL) | (1ULL << 43);
            } else {
                st->lv[level].k_loc = st->redo_count;   // what this shard holds of the level
                finish_level(st, level, k_in - sum);     // identical on every shard: same sums
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// or_collide_merge_kernel: reduce-scatter + all-gather of one level over peer memory (see the header)
// 16 bytes (two words) per lane and array; loads of all shards are issued before the fold.
// ---------------------------------------------------------------------------------------------
constexpr int MERGE_THREADS = 256;

template <int WORLD>
__device__ __forceinline__ void merge_slice(const CommDev& cd, uint64_t pair_off, uint64_t p0, uint64_t p1) {
    const uint64_t stride = (uint64_t)gridDim.x * MERGE_THREADS;
    for (uint64_t p = p0 + (uint64_t)blockIdx.x * MERGE_THREADS + threadIdx.x; p < p1; p += stride) {
        ulonglong2 av[WORLD], cv[WORLD];
#pragma unroll
        for (int r = 0; r < WORLD; r++) {
            av[r] = reinterpret_cast<const ulonglong2*>(cd.words[r])[pair_off + p];
            cv[r] = reinterpret_cast<const ulonglong2*>(cd.coll[r])[pair_off + p];
        }
        ulonglong2 a = av[0], c = cv[0];
#pragma unroll
        for (int r = 1; r < WORLD; r++) {
            c.x |= cv[r].x | (a.x & av[r].x);
            c.y |= cv[r].y | (a.y & av[r].y);
            a.x |= av[r].x;
            a.y |= av[r].y;
        }
        a.x &= ~c.x;
        a.y &= ~c.y;
#pragma unroll
        for (int r = 0; r < WORLD; r++) {
            reinterpret_cast<ulonglong2*>(cd.words[r])[pair_off + p] = a;
            reinterpret_cast<ulonglong2*>(cd.coll[r])[pair_off + p] = c;
        }
    }
}

__global__ void __launch_bounds__(MERGE_THREADS)
    or_collide_merge_kernel(const BuildState* __restrict__ st, uint32_t level, CommDev cd) {
    if (st->status != ST_RUNNING || st->n_levels != level || level >= st->tail_level) return;
    if (st->lv[level].x_slice_bits != 0) return;  // exchange level: nothing to fold, the owners marked their slices
    const uint64_t pair_off = st->lv[level].word_off / 2;
    const uint64_t n_pairs = round_up8(st->lv[level].n_words) / 2;
    const uint64_t per = (n_pairs + cd.world - 1) / cd.world;
    const uint64_t p0 = min(n_pairs, per * cd.rank), p1 = min(n_pairs, p0 + per);
    switch (cd.world) {
        case 2: merge_slice<2>(cd, pair_off, p0, p1); break;
        case 3: merge_slice<3>(cd, pair_off, p0, p1); break;
        case 4: merge_slice<4>(cd, pair_off, p0, p1); break;
        case 5: merge_slice<5>(cd, pair_off, p0, p1); break;
        case 6: merge_slice<6>(cd, pair_off, p0, p1); break;
        case 7: merge_slice<7>(cd, pair_off, p0, p1); break;
        case 8: merge_slice<8>(cd, pair_off, p0, p1); break;
        default: break;
    }
    __threadfence_system();  // the pushes must be performed before the barrier kernel that follows signals
}

// ---------------------------------------------------------------------------------------------
// tail_gather_kernel: this shard's keys entering the tail level -> its exchange slot (one CTA).
// The tail kernel of every shard then reads all slots and finishes the remaining levels redundantly,
// so every shard ends with the identical complete structure.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TAIL_THREADS)
    tail_gather_kernel(BuildState* __restrict__ st, const uint64_t* __restrict__ user_keys,
                       const uint64_t* __restrict__ buf0, const uint64_t* __restrict__ buf1,
                       const uint32_t* __restrict__ c32, uint64_t* __restrict__ xk,
                       const uint32_t* __restrict__ cursors) {
    __shared__ uint32_t count;
    if (st->status != ST_RUNNING || st->tail_level == NO_TAIL || st->n_levels != st->tail_level) return;
    if (st->shard_world <= 1) return;  // all keys were collected on every shard already: the tail reads local lists
    const uint32_t tid = threadIdx.x, level = st->tail_level;
    if (level >= 1 && st->lv[level - 1].x_slice_bits != 0) {
        // the host plan keeps at least two non-exchange levels in front of the tail
        if (tid == 0) { st->status = ST_ERR_INTERNAL; st->err_info = ((uint64_t)level << 48) | (1ULL << 42); }
        return;
    }
    if (tid == 0) count = 0;
    __syncthreads();
    if (level == 0) {
        const uint32_t k = (uint32_t)min(st->lv[0].k_loc, (uint64_t)TAIL_MAX_KEYS);
        for (uint32_t i = tid; i < k; i += TAIL_THREADS) xk[8 + i] = user_keys[i];
        if (tid == 0) count = k;
    } else if (st->lv[level - 1].b_range != 0) {
        // bucketed predecessor: the scatter step already filtered this shard's keys into a plain list
        const uint64_t* src = (level & 1) ? buf1 : buf0;
        const uint32_t k = min(cursors[(size_t)(level & 1) * MAX_BUCKETS_ALLOC * CURSOR_STRIDE], (uint32_t)TAIL_MAX_KEYS);
        for (uint32_t i = tid; i < k; i += TAIL_THREADS) xk[8 + i] = src[i];
        if (tid == 0) count = k;
    } else {
        const LevelDesc* pv = &st->lv[level - 1];
        const uint64_t p_bits = pv->bits, p_sp0 = pv->sp0, p_off32 = pv->word_off * 2, k_src = pv->k_loc;
        const bool p_big = (p_bits >> 32) != 0;
        const uint64_t* src = (level == 1) ? user_keys : (((level - 1) & 1) ? buf1 : buf0);
        auto take = [&](uint64_t key) {
            const uint64_t idx = p_big ? hashmod_big(st->kc, p_sp0, key, p_bits) : (uint64_t)hashmod_small(st->kc, p_sp0, key, (uint32_t)p_bits);
            if ((c32[p_off32 + (idx >> 5)] >> (idx & 31)) & 1u) {
                const uint32_t pos = atomicAdd(&count, 1u);
                if (pos < TAIL_MAX_KEYS) xk[8 + pos] = key;
            }
        };
        if (list_is_striped(st, level - 1)) {
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
    if (tid == 0) {
        xk[0] = count;
        __threadfence_system();
    }
}

// ---------------------------------------------------------------------------------------------
// Collect: once at most GATHER_MAX_KEYS keys are left, every shard filters its keys of level gather_level - 1 into
// its exchange slot (shard_gather_kernel = pass_body without the mark), a list-sum barrier carries the counts,
// and every shard copies all slots into its own list (shard_collect_kernel).  From here on every shard holds the
// whole remaining key set and finishes the cascade with the unsharded kernels, redundantly -> identical replicas.
// ---------------------------------------------------------------------------------------------
template <bool MAYBE_BIG>
__global__ void __launch_bounds__(PASS_THREADS, BPHF_PASS_CTAS)
    shard_gather_kernel(BuildState* __restrict__ st, uint32_t level, const uint64_t* __restrict__ user_keys,
                        uint64_t* __restrict__ buf0, uint64_t* __restrict__ buf1, uint32_t* __restrict__ c32,
                        uint64_t* __restrict__ xk) {
    if (st->status != ST_RUNNING || st->n_levels != level || level != st->gather_level || level >= st->tail_level) return;
    if (st->lv[level - 1].x_slice_bits != 0) return;  // exchange predecessor: shard_gather_x_kernel
    pass_body<MAYBE_BIG, true>(st, level, user_keys, buf0, buf1, nullptr, c32, xk + 8);
}
template <bool MAYBE_BIG>
__global__ void __launch_bounds__(PASS_THREADS, BPHF_PASS_CTAS)
    shard_gather_x_kernel(BuildState* __restrict__ st, uint32_t level, uint64_t* __restrict__ buf0, uint64_t* __restrict__ buf1,
                          uint64_t* __restrict__ xk, const uint32_t* __restrict__ xbits) {
    if (st->status != ST_RUNNING || st->n_levels != level || level != st->gather_level || level >= st->tail_level) return;
    if (st->lv[level - 1].x_slice_bits == 0) return;
    pass_body<MAYBE_BIG, true, KeyCfg, true>(st, level, nullptr, buf0, buf1, nullptr, nullptr, xk + 8, nullptr, 0, xbits);
}

__global__ void __launch_bounds__(256)
    shard_collect_kernel(BuildState* __restrict__ st, uint32_t level, CommDev cd, uint64_t* __restrict__ buf0,
                         uint64_t* __restrict__ buf1) {
    if (st->status != ST_RUNNING || st->n_levels != level || level != st->gather_level || level >= st->tail_level) return;
    uint64_t* __restrict__ dst = (level & 1) ? buf1 : buf0;
    uint64_t off = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint32_t p = 0; p < cd.world; p++) {
        const uint64_t cnt = st->peer_count[p];
        const uint64_t* __restrict__ src = cd.xkeys[p] + 8;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += stride) dst[off + i] = src[i];
        off += cnt;
    }
}

// one thread, after the collect: from now on this shard behaves like an unsharded build of the remaining keys
__global__ void shard_switch_kernel(BuildState* __restrict__ st, uint32_t level) {
    if (st->status != ST_RUNNING || st->n_levels != level || level != st->gather_level || level >= st->tail_level) return;
    uint64_t total = 0;
    for (uint32_t p = 0; p < st->shard_world; p++) total += st->peer_count[p];
    st->lv[level].k_loc = st->lv[level].k_in;
    st->redo_count = total;            // finalize checks it against k(level)
    st->list_ready_level = level;
    st->shard_world = 1;
}

}  // namespace bphf
