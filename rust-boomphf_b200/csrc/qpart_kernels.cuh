// qpart_kernels.cuh -- Mphf::hash / try_hash over a batch (src/lib.rs:324-355) when the lookup structure is larger
// than the L2 (BASELINE configs[3]: 1e9 keys -> 283 MB of level-0 granules, 126 MB of level 1, ...).
//
// The plain kernel (query_sync_kernel) then pays one random DRAM sector per probe of the big levels: 1e9 lookups run at
// the HBM random-access ceiling (~27 G sectors/s), 36 ms, 0.07 of the streaming roofline.  Here the probes of a big level
// are made L2 hits by visiting the level slice by slice:
//
//   stage -1 (the caller's keys)   qfirst           hash level 0, bucket = slot >> shift (<= 15 buckets, each a slice of
//                                                   the level that fits the L2), keys appended to the bucket's lists
//   stage s  (bucket lists, s>=0)  qprobe           bucket by bucket -- the whole grid probes ONE slice at a time --
//                                                   level s is probed; hits store their rank, misses are split again by
//                                                   their level s+1 slot (or go to the walk lists when the remaining
//                                                   levels fit the L2 together)
//   walk stage                     qwalk            query_sync_body<SUB>: remaining levels, L2 resident
//   back up, stage by stage        qrestore         a miss fetches its rank from the slot the split put its key in
//   stage -1                       qrestore<FIRST>  out[i] in the caller's order, coalesced
//
// No index travels with a key: an entry remembers, in the 32-bit word that will hold its rank, the slot of the next
// stage its key went to (and one bit per entry says which of the two the word is); the restore pass of the stage below
// replaces the slot by the rank found there.  The unit of work is 256 consecutive entries, owned by ONE warp from load
// to store (no block barrier anywhere: a warp waiting for its keys or granules never holds up another); positions
// inside the unit's run of a bucket come from shared-memory atomics on 16 per-warp counters.
// Every bucket is QP_SUBS lists with a cursor each (unit u appends to sub-list u % QP_SUBS): one cursor per bucket
// would take 4e6 same-address atomics per 1e9 keys and bucket, and those complete at ~1.4 G/s (profiles/r02_summary.md).
// Every array is read and written in full sectors; bytes per key and stage:
//   first stage   8 (key in) + 4 + 4 (slot out / in) + 4 (rank gathered) + 8 (out)                   = 28
//   stage s       8 + 8 (key out / in) + 4 (word out) + 4 + 4 (merge) + 4 (gathered by the parent)   = 32 (+ 2 bits)
//   walk stage    8 + 8 + 4 + 4                                                                      = 24
// The lists have a fixed capacity (expected fill + slack).  Adversarial batches (one key repeated 1e8 times) overflow a
// list: the split kernels then raise a flag, drop what does not fit, and qfallback_kernel -- launched after every
// batch, a no-op while the flag is clear -- redoes the batch with the plain walk.  No host round trip either way.
#pragma once
#include "build_kernels.cuh"  // TMA bulk copy + mbarrier helpers
#include "query_kernels.cuh"

namespace bphf {

constexpr int QP_THREADS = 256;
constexpr int QP_WARPS = QP_THREADS / 32;
constexpr int QP_KPL = 8;                        // entries per lane
constexpr int QP_UNIT = 32 * QP_KPL;             // 256 entries: one warp's unit of the split (= QS_WTILE)
constexpr uint32_t QP_MAX_BUCKETS = 15;
constexpr uint32_t QP_SUBS = 32;                 // lists (and cursors) per bucket
constexpr uint32_t QP_MAX_LISTS = QP_MAX_BUCKETS * QP_SUBS;
constexpr uint32_t QP_NONE32 = 0xffffffffu;      // rank of "no level claims the key" in the 32-bit result arrays
constexpr uint32_t QP_CURSOR_STRIDE = 32;        // u32 per cursor: one 128-byte line each
constexpr int QP_MAX_STAGES = 4;
#ifndef QP_SPLIT_CTAS
#define QP_SPLIT_CTAS 3   // CTAs per SM of the probing kernel: 80 registers; at 4 (64 registers, spills) it measured 30 % slower
#endif
#ifndef QP_RING
#define QP_RING 1         // 1: the units of the probing kernel arrive through a TMA ring in shared memory; 0: read in place + L1 prefetch (5 % slower)
#endif
#if QP_RING
#define QP_KEY2(p) (*(p))
#else
#define QP_KEY2(p) __ldg(p)
#endif
#ifndef QP_PROBE_BATCH
#define QP_PROBE_BATCH 4  // granule loads of a lane in flight together (8 at 2 CTAs per SM measured no faster)
#endif

// one stage: the keys that reached level `level` unresolved, grouped by the slice of that level their slot falls in;
// list (b, r) = entries [ (b * QP_SUBS + r) * cap, ... + cursor ) of the arrays
struct QPStage {
    uint64_t* keys;
    uint32_t* res;     // per entry: its rank (or QP_NONE32), or -- while its `went` bit is set -- its slot in the next stage
    uint8_t* went;     // [slot / 8]  byte (unit, lane): bit `it` = the lane's entry `it` went on to the next stage
    uint32_t* cursor;  // [list * QP_CURSOR_STRIDE]  fill of each list
    uint32_t n_buckets, shift, cap, level;  // cap: entries per list, a multiple of QP_UNIT
};

// ---- units of a stage, bucket-major: persistent warps advance at the same pace, so at any time the whole grid probes
// the same slice of the level.  pre[l] = units of the lists before l.
__device__ __forceinline__ void qp_prefix(const QPStage& st, uint32_t* pre, uint32_t* fill) {
    const uint32_t nl = st.n_buckets * QP_SUBS;
    if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        uint32_t carry = 0;
        for (uint32_t l0 = 0; l0 < nl; l0 += 32) {
            const uint32_t l = l0 + lane;
            const uint32_t f = l < nl ? min(st.cursor[l * QP_CURSOR_STRIDE], st.cap) : 0u;
            const uint32_t u = (f + QP_UNIT - 1) / QP_UNIT;
            uint32_t s = u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, s, d);
                if (lane >= (uint32_t)d) s += t;
            }
            if (l < nl) {
                pre[l] = carry + s - u;
                fill[l] = f;
            }
            carry += __shfl_sync(0xffffffffu, s, 31);
        }
        if (lane == 0) pre[nl] = carry;
    }
    __syncthreads();
}
struct QPLoc {
    uint64_t slot;  // first entry of the unit in the stage's arrays (multiple of QP_UNIT)
    uint32_t n;     // entries of the unit that exist
};
// A warp visits its units in increasing order: the list of the next one is found by walking on from the last one's
// (lists hold ~2000 units of a 2^27-key batch, a persistent grid's stride is ~3500: one or two steps).
struct QPWalker {
    uint32_t lo = 0;  // list of the last unit located
    __device__ __forceinline__ QPLoc locate(const QPStage& st, const uint32_t* pre, const uint32_t* fill, uint32_t item) {
        while (pre[lo + 1] <= item) lo++;  // (item < pre[n_lists]: stops at the list that holds it, empty lists are skipped)
        const uint32_t unit = item - pre[lo];
        QPLoc r;
        r.slot = (uint64_t)lo * st.cap + (uint64_t)unit * QP_UNIT;
        r.n = min((uint32_t)QP_UNIT, fill[lo] - unit * QP_UNIT);
        return r;
    }
};

// Positions for the entries of a unit that go on: the lane's entry `it` goes to bucket (cpack >> 4 it) & 15 when bit it
// of `go` is set.  On return word[it] of such an entry = its slot in stage nxt, or QP_NONE32 (and its go bit cleared)
// when the list was full (overflow raised); the other words are left alone.  cnt: the warp's 16 counters in shared
// memory.  (Buckets packed four bits each and one in/out array: the probing kernel has 80 registers to live in.)
__device__ __forceinline__ uint32_t qp_place(uint32_t cpack, uint32_t go, const QPStage& nxt, uint32_t sub,
                                             uint32_t* __restrict__ cnt, uint32_t* __restrict__ overflow,
                                             uint32_t (&word)[QP_KPL]) {
    const uint32_t lane = threadIdx.x & 31;
    if (lane < 16) cnt[lane] = 0;
    __syncwarp();
#pragma unroll
    for (int it = 0; it < QP_KPL; it++)  // position within the unit's run of the bucket
        if ((go >> it) & 1u) word[it] = atomicAdd(cnt + ((cpack >> (4 * it)) & 15u), 1u);
    __syncwarp();
    if (lane < nxt.n_buckets) {
        const uint32_t run = cnt[lane];
        uint32_t start = 0;
        if (run) {
            start = atomicAdd(nxt.cursor + (size_t)(lane * QP_SUBS + sub) * QP_CURSOR_STRIDE, run);
            if ((uint64_t)start + run > nxt.cap) *overflow = 1u;
        }
        cnt[lane] = start;
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < QP_KPL; it++) {
        if ((go >> it) & 1u) {
            const uint32_t c = (cpack >> (4 * it)) & 15u;
            const uint32_t pos = cnt[c] + word[it];
            if (pos < nxt.cap) {
                word[it] = (c * QP_SUBS + sub) * nxt.cap + pos;
            } else {
                word[it] = QP_NONE32;
                go &= ~(1u << it);
            }
        }
    }
    __syncwarp();  // cnt is reused by the warp's next unit
    return go;
}

// The caller's keys (n_first of them) are split by their level nxt.level slot; first.res takes the slot of every key.
// Entry e of a unit sits with lane e % 32 at iteration e / 32 (8-byte loads: the caller's array may be only 8-byte
// aligned).
#ifndef QP_FIRST_CTAS
#define QP_FIRST_CTAS 5  // 48 registers, no spills: 0.58 ms per 2^27 keys against 0.64 at 4 and 0.61 at 6 (spills)
#endif
__global__ void __launch_bounds__(QP_THREADS, QP_FIRST_CTAS)
    qfirst_kernel(QView v, QPStage first, QPStage nxt, uint64_t n_first, uint32_t* __restrict__ overflow) {
    __shared__ QLevelP lv;
    __shared__ uint32_t cnt_all[QP_WARPS][16];
    const uint32_t lane = threadIdx.x & 31;
    uint32_t* __restrict__ cnt = cnt_all[threadIdx.x >> 5];
    if (threadIdx.x == 0) lv = v.lvp[nxt.level];
    __syncthreads();
    const uint32_t total = (uint32_t)((n_first + QP_UNIT - 1) / QP_UNIT);
    const KeyCfg kc = v.kc;
    const uint32_t w_stride = gridDim.x * QP_WARPS;
    for (uint32_t item = blockIdx.x * QP_WARPS + (threadIdx.x >> 5); item < total; item += w_stride) {
        const uint64_t slot0 = (uint64_t)item * QP_UNIT;
        const uint32_t n = (uint32_t)min((uint64_t)QP_UNIT, n_first - slot0);
        uint64_t key[QP_KPL];
        uint32_t slot[QP_KPL];
        uint32_t go = 0;     // bit it: the lane's entry `it` exists
        uint32_t cpack = 0;  // four bits per entry: its bucket
#pragma unroll
        for (int it = 0; it < QP_KPL; it++) {
            const uint32_t e = it * 32 + lane;
            key[it] = e < n ? __ldcs(reinterpret_cast<const unsigned long long*>(first.keys) + slot0 + e) : 0ULL;
            go |= (e < n ? 1u : 0u) << it;
            slot[it] = QP_NONE32;
        }
#pragma unroll
        for (int it = 0; it < QP_KPL; it++) cpack |= (hashmod_small(kc, lv.sp0, key[it], lv.bits) >> nxt.shift) << (4 * it);
        const uint32_t went = qp_place(cpack, go, nxt, item & (QP_SUBS - 1), cnt, overflow, slot);
#pragma unroll
        for (int it = 0; it < QP_KPL; it++) {
            if ((went >> it) & 1u) nxt.keys[slot[it]] = key[it];
            const uint32_t e = it * 32 + lane;
            if (e < n) __stcs(first.res + slot0 + e, slot[it]);
        }
    }
}

// The entries of stage `cur` are probed against level cur.level; hits keep their rank, the misses go on to stage `nxt`.
// A warp's units arrive through the copy engine (TMA bulk copies into a two-deep ring of 2 KB per warp, one unit
// ahead) and stay in shared memory: the probe, the next level's hash and the store each read the key there, so that
// neither the DRAM latency of the key stream nor 16 registers of keys sit in the warp's dependency chain.
// Entry e of a unit sits with lane (e % 64) / 2 at iteration 2 * (e / 64) + e % 2 (16-byte shared loads, 8-byte stores).
__global__ void __launch_bounds__(QP_THREADS, QP_SPLIT_CTAS)
    qprobe_kernel(QView v, QPStage cur, QPStage nxt, uint32_t* __restrict__ overflow) {
#if QP_RING
    __shared__ __align__(128) uint64_t ring_all[QP_WARPS][2][QP_UNIT];
    __shared__ __align__(8) uint64_t mbar_all[QP_WARPS][2];
#endif
    __shared__ QLevelP lv[2];
    __shared__ uint32_t pre[QP_MAX_LISTS + 1], fill[QP_MAX_LISTS];
    __shared__ uint32_t cnt_all[QP_WARPS][16];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* __restrict__ cnt = cnt_all[warp];
    if (threadIdx.x == 0) {
        lv[0] = v.lvp[cur.level];
        lv[1] = v.lvp[nxt.level];
    }
#if QP_RING
    uint64_t* __restrict__ mbar = mbar_all[warp];
    if (lane == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
        fence_proxy_async();
    }
#endif
    qp_prefix(cur, pre, fill);
    const uint32_t total = pre[cur.n_buckets * QP_SUBS];
    const KeyCfg kc = v.kc;
    const uint4* __restrict__ q = v.packed;
    const bool one_bucket = nxt.n_buckets == 1;
    const uint32_t w_stride = gridDim.x * QP_WARPS;
    const uint32_t item0 = blockIdx.x * QP_WARPS + warp;
    const uint32_t n_mine = item0 < total ? (total - item0 + w_stride - 1) / w_stride : 0u;
    QPWalker ahead, here;
#if QP_RING
    auto issue = [&](uint32_t j) {  // unit j of this warp -> ring stage j & 1 (whole units: the arrays are padded)
        if (j < n_mine) {
            const QPLoc u = ahead.locate(cur, pre, fill, item0 + j * w_stride);
            if (lane == 0) {
                mbar_expect_tx(&mbar[j & 1], QP_UNIT * 8);
                bulk_load(ring_all[warp][j & 1], cur.keys + u.slot, QP_UNIT * 8, &mbar[j & 1]);
            }
        }
    };
    issue(0);
    issue(1);
    uint32_t phase = 0;  // bit b: parity of mbar[b]'s next completion
#else
    // no ring: the unit is read from global memory at each use (2 KB per warp: the later reads hit the L1) and the
    // next unit is prefetched into the L1 one unit ahead -- the shared memory a ring takes (32 KB per CTA) is worth more
    // as L1, where the random granule loads wait
    auto issue = [&](uint32_t j) {
        if (j < n_mine) {
            const QPLoc u = ahead.locate(cur, pre, fill, item0 + j * w_stride);
            if (lane < QP_UNIT * 8 / 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(cur.keys + u.slot + lane * 16));
        }
    };
    issue(0);
#endif
    for (uint32_t j = 0; j < n_mine; j++) {
        const uint32_t item = item0 + j * w_stride;
        const QPLoc u = here.locate(cur, pre, fill, item);
#if QP_RING
        const uint32_t b = j & 1;
        mbar_wait(&mbar[b], (phase >> b) & 1u);
        phase ^= 1u << b;
        const ulonglong2* __restrict__ kq = reinterpret_cast<const ulonglong2*>(ring_all[warp][b]);
#else
        issue(j + 1);
        const ulonglong2* __restrict__ kq = reinterpret_cast<const ulonglong2*>(cur.keys + u.slot);
#endif
        uint32_t word[QP_KPL];  // rank of a hit
        uint32_t go = 0;        // bit it: the lane's entry `it` goes on to the next stage
        // probe level cur.level, QP_PROBE_BATCH granules of the lane in flight at a time.  While they are in flight the
        // same entries are hashed for the next level -- for every entry, needed or not: some lane of the warp misses in
        // every iteration anyway, a branch per entry costs more than the hash of the entries that were found, and the
        // arithmetic fills the wait for the granules
        uint32_t cpack = 0;  // four bits per entry: its bucket in the next stage
#pragma unroll
        for (int h = 0; h < QP_KPL; h += QP_PROBE_BATCH) {
            PackedProbe pp[QP_PROBE_BATCH];
            uint4 g[QP_PROBE_BATCH];
            ulonglong2 k2[QP_PROBE_BATCH / 2];
#pragma unroll
            for (int jj = 0; jj < QP_PROBE_BATCH; jj += 2) {
                k2[jj >> 1] = QP_KEY2(kq + ((h + jj) >> 1) * 32 + lane);
                pp[jj] = packed_slot(kc, lv[0], k2[jj >> 1].x);
                g[jj] = __ldg(q + lv[0].gran_off + pp[jj].g);
                pp[jj + 1] = packed_slot(kc, lv[0], k2[jj >> 1].y);
                g[jj + 1] = __ldg(q + lv[0].gran_off + pp[jj + 1].g);
            }
            if (!one_bucket) {
#pragma unroll
                for (int jj = 0; jj < QP_PROBE_BATCH; jj += 2) {
                    cpack |= (hashmod_small(kc, lv[1].sp0, k2[jj >> 1].x, lv[1].bits) >> nxt.shift) << (4 * (h + jj));
                    cpack |= (hashmod_small(kc, lv[1].sp0, k2[jj >> 1].y, lv[1].bits) >> nxt.shift) << (4 * (h + jj + 1));
                }
            }
#pragma unroll
            for (int jj = 0; jj < QP_PROBE_BATCH; jj++) {
                const int it = h + jj;
                const uint32_t e = (it >> 1) * 64 + lane * 2 + (it & 1);
                uint64_t rk;
                const bool hit = packed_hit(lv[0], g[jj], pp[jj].bit, &rk);
                word[it] = (uint32_t)rk;
                go |= ((!hit && e < u.n) ? 1u : 0u) << it;
            }
        }
        const uint32_t went = qp_place(cpack, go, nxt, item & (QP_SUBS - 1), cnt, overflow, word);
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const ulonglong2 k2 = QP_KEY2(kq + jj * 32 + lane);
            if ((went >> (2 * jj)) & 1u) nxt.keys[word[2 * jj]] = k2.x;
            if ((went >> (2 * jj + 1)) & 1u) nxt.keys[word[2 * jj + 1]] = k2.y;
        }
#if QP_RING
        __syncwarp();  // every lane is done with the ring stage
        if (lane == 0) fence_proxy_async();
        issue(j + 2);
#endif
        // (an entry dropped at a full list has QP_NONE32 as its rank: the batch is redone anyway)
        uint2* __restrict__ res = reinterpret_cast<uint2*>(cur.res + u.slot);
        __stcs(cur.went + (u.slot >> 3) + lane, (uint8_t)went);
#pragma unroll
        for (int jj = 0; jj < 4; jj++) __stcs(res + jj * 32 + lane, make_uint2(word[2 * jj], word[2 * jj + 1]));
    }
#if QP_RING
    __syncwarp();
    if (lane == 0) {
        mbar_inval(&mbar[0]);
        mbar_inval(&mbar[1]);
    }
#endif
}

// the units of the walk stage's lists as tiles of query_sync_body
struct QPListTiles {
    const QPStage& st;
    const uint32_t* pre;
    const uint32_t* fill;
    QPWalker w;
    __device__ __forceinline__ uint64_t count() const { return pre[st.n_buckets * QP_SUBS]; }
    __device__ __forceinline__ void locate(uint64_t tile, uint64_t* base, uint32_t* t_n) {
        const QPLoc u = w.locate(st, pre, fill, (uint32_t)tile);
        *base = u.slot;
        *t_n = u.n;
    }
};
// the remaining levels for the keys of the walk stage
__global__ void __launch_bounds__(QS_THREADS, 8) qwalk_kernel(QView v, QPStage st) {
    __shared__ uint32_t pre[QP_SUBS + 1], fill[QP_SUBS];
    qp_prefix(st, pre, fill);
    QPListTiles tiles{st, pre, fill, QPWalker()};
    query_sync_body<true>(v, tiles, st.keys, nullptr, st.res, st.level);
}

// cur.res (FIRST: out, 64-bit, in the caller's order) <- ranks of the entries that went on to stage nxt
template <bool FIRST>
__global__ void __launch_bounds__(QP_THREADS, 8)
    qrestore_kernel(QPStage cur, QPStage nxt, uint64_t n_first, uint64_t* __restrict__ out) {
    __shared__ uint32_t pre[FIRST ? 1 : QP_MAX_LISTS + 1], fill[FIRST ? 1 : QP_MAX_LISTS];
    const uint32_t lane = threadIdx.x & 31;
    uint32_t total;
    if (FIRST) {
        total = (uint32_t)((n_first + QP_UNIT - 1) / QP_UNIT);
    } else {
        qp_prefix(cur, pre, fill);
        total = pre[cur.n_buckets * QP_SUBS];
    }
    const uint32_t* __restrict__ below = nxt.res;
    QPWalker walker;
    const uint32_t w_stride = gridDim.x * QP_WARPS;
    for (uint32_t item = blockIdx.x * QP_WARPS + (threadIdx.x >> 5); item < total; item += w_stride) {
        if (FIRST) {
            const uint64_t slot0 = (uint64_t)item * QP_UNIT;
            const uint32_t n = (uint32_t)min((uint64_t)QP_UNIT, n_first - slot0);
            uint32_t w[QP_KPL];
#pragma unroll
            for (int it = 0; it < QP_KPL; it++) {
                const uint32_t e = it * 32 + lane;
                w[it] = e < n ? __ldcs(cur.res + slot0 + e) : QP_NONE32;
            }
#pragma unroll
            for (int it = 0; it < QP_KPL; it++)
                if (w[it] != QP_NONE32) w[it] = __ldg(below + w[it]);
#pragma unroll
            for (int it = 0; it < QP_KPL; it++) {
                const uint32_t e = it * 32 + lane;
                if (e < n)
                    __stcs(reinterpret_cast<unsigned long long*>(out) + slot0 + e,
                           w[it] == QP_NONE32 ? (unsigned long long)BPHF_NONE : (unsigned long long)w[it]);
            }
        } else {
            const QPLoc u = walker.locate(cur, pre, fill, item);
            uint2* __restrict__ res = reinterpret_cast<uint2*>(cur.res + u.slot);
            const uint32_t went = __ldcs(cur.went + (u.slot >> 3) + lane);
            uint2 w[4];
#pragma unroll
            for (int j = 0; j < 4; j++) w[j] = __ldcs(res + j * 32 + lane);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if ((went >> (2 * j)) & 1u) w[j].x = __ldg(below + w[j].x);
                if ((went >> (2 * j + 1)) & 1u) w[j].y = __ldg(below + w[j].y);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) __stcs(res + j * 32 + lane, w[j]);
        }
    }
}

// after every batch: the plain walk over the batch if a list overflowed, nothing otherwise
__global__ void __launch_bounds__(QS_THREADS, 8)
    qfallback_kernel(QView v, const uint64_t* __restrict__ keys, uint64_t n, uint64_t* __restrict__ out,
                     const uint32_t* __restrict__ overflow, unsigned long long* __restrict__ n_fallbacks) {
    if (*overflow == 0u) return;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_fallbacks) atomicAdd(n_fallbacks, 1ULL);
    query_sync_body<false>(v, QSDenseTiles{n}, keys, out, nullptr, 0);
}

}  // namespace bphf
