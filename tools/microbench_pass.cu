// microbench: the level-1 filter pass (filter(0) + compaction + mark(1)) of a 1e8-key build, standalone, with its
// components switched off one at a time -- which part of the pass keeps it at 0.7 L2 requests per clock and SM when
// plain loads + atomics reach 0.96 (tools/microbench_mix.cu)?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I rust-boomphf_b200/csrc -o tools/microbench_pass tools/microbench_pass.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "hash.cuh"
using namespace bphf;

__device__ __forceinline__ uint64_t splitmix(uint64_t z) {
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t ld_cg(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void red_or(uint32_t* p, uint32_t v) { asm volatile("red.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void gen_keys(uint64_t* k, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) k[i] = splitmix(i);
}
__global__ void mark0(const uint64_t* k, uint64_t n, uint32_t* a, uint32_t* c, uint32_t bits, uint64_t sp0, KeyCfg kc) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t idx = hashmod_small(kc, sp0, k[i], bits);
        const uint32_t m = 1u << (idx & 31);
        if (atomicOr(a + (idx >> 5), m) & m) red_or(c + (idx >> 5), m);
    }
}

constexpr int KPT = 8, TILE = 32 * KPT, RING = 2 * TILE, WARPS = 8;
enum : uint32_t {
    F_KEYS_DRAM = 1,   // keys streamed from memory (else: computed from the index)
    F_PROBE = 2,       // random probe of collide(l-1) (else: survivors decided by a hash bit)
    F_LIST = 4,        // survivors written to the list (reservation + coalesced store)
    F_ATOM = 8,        // fetch_or into a(l)
    F_RED = 16,        // collide(l) update for late arrivals
    F_RESERVE = 32,    // reservation atomicAdd on the append cursor (else: per-warp private region)
};
template <uint32_t F>
__global__ void __launch_bounds__(256, 3) pass(const uint64_t* __restrict__ keys, uint64_t n, const uint32_t* __restrict__ c_pv,
                                               uint32_t p_bits, uint64_t p_sp0, uint32_t* __restrict__ a_lv, uint32_t* __restrict__ c_lv,
                                               uint32_t c_bits, uint64_t c_sp0, uint64_t* __restrict__ dst, unsigned long long* cursor,
                                               KeyCfg kc) {
    __shared__ __align__(16) uint64_t ring_all[WARPS][RING];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint64_t* ring = ring_all[warp];
    uint32_t head = 0, n_st = 0;
    bool resv = false;
    unsigned long long ticket = 0;
    uint32_t aidx[KPT], old[KPT], amask = 0;
#pragma unroll
    for (int j = 0; j < KPT; j++) aidx[j] = old[j] = 0;
    const uint64_t n_gw = (uint64_t)gridDim.x * WARPS, gw = (uint64_t)blockIdx.x * WARPS + warp;
    unsigned long long priv = gw * (n / n_gw + TILE);  // !F_RESERVE: private output region (for measurement only)
    auto settle = [&]() {
        if (F & F_RED) {
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                const uint32_t m = 1u << (aidx[j] & 31);
                if (((amask >> j) & 1u) && (old[j] & m)) red_or(c_lv + (aidx[j] >> 5), m);
            }
        }
        amask = 0;
    };
    auto flush = [&](uint32_t cnt) {
        unsigned long long out;
        if (F & F_RESERVE) out = __shfl_sync(0xffffffffu, ticket, 0);
        else { out = priv; priv += cnt; }
        settle();
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const uint32_t i = lane + 32 * j;
            if (i < cnt) {
                const uint64_t kk = ring[(head + i) & (RING - 1)];
                if (F & F_LIST) dst[out + i] = kk;
                aidx[j] = hashmod_small(kc, c_sp0, kk, c_bits);
                amask |= 1u << j;
            }
        }
        if (F & F_ATOM) {
#pragma unroll
            for (int j = 0; j < KPT; j++)
                if ((amask >> j) & 1u) old[j] = atomicOr(a_lv + (aidx[j] >> 5), 1u << (aidx[j] & 31));
        }
        head = (head + cnt) & (RING - 1);
        n_st -= cnt;
        __syncwarp();
    };
    const uint64_t n_tiles = n / TILE;  // n is a multiple of TILE here
    uint64_t tile = gw;
    uint64_t key[KPT];
    auto load_tile = [&](uint64_t t) {
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const uint64_t i = t * TILE + lane + 32u * j;
            key[j] = (F & F_KEYS_DRAM) ? ld_cg(keys + i) : splitmix(i);
        }
    };
    if (tile < n_tiles) load_tile(tile);
    while (tile < n_tiles) {
        uint32_t cw[KPT];
        uint64_t shs = 0;
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const uint32_t idx = hashmod_small(kc, p_sp0, key[j], p_bits);
            if (F & F_PROBE) cw[j] = __ldcg(c_pv + (idx >> 5));
            else cw[j] = (idx * 0x9E3779B1u) < 1910000000u ? 0xffffffffu : 0u;  // ~44.5 % survive
            shs |= (uint64_t)(idx & 31) << (5 * j);
        }
        if (resv) {
            flush(TILE);
            resv = false;
        }
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const bool redo = (cw[j] >> ((uint32_t)(shs >> (5 * j)) & 31)) & 1u;
            const uint32_t mk = __ballot_sync(0xffffffffu, redo);
            if (redo) ring[(head + n_st + __popc(mk & lt)) & (RING - 1)] = key[j];
            n_st += __popc(mk);
        }
        __syncwarp();
        tile += n_gw;
        if (tile < n_tiles) {
            load_tile(tile);
            if ((F & F_KEYS_DRAM) && lane < 16 && tile + n_gw < n_tiles)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(keys + (tile + n_gw) * TILE + lane * 16));
        }
        if (n_st >= TILE) {
            if ((F & F_RESERVE) && lane == 0) ticket = atomicAdd(cursor, (unsigned long long)TILE);
            resv = true;
        }
    }
    if (resv) flush(TILE);
    if (n_st) {
        if ((F & F_RESERVE) && lane == 0) ticket = atomicAdd(cursor, (unsigned long long)n_st);
        flush(n_st);
    }
    settle();
}

struct Ctx {
    uint64_t *keys, *dst;
    uint32_t *a0, *c0, *a1, *c1;
    unsigned long long* cursor;
    uint64_t n, w0, w1;
    uint32_t bits0, bits1;
    KeyCfg kc;
};
template <uint32_t F>
void run(const Ctx& x, const char* what, int ctas_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9;
    unsigned long long cur = 0;
    for (int r = 0; r < 4; r++) {
        cudaMemset(x.a1, 0, x.w1 * 4);
        cudaMemset(x.c1, 0, x.w1 * 4);
        cudaMemset(x.cursor, 0, 8);
        cudaEventRecord(e0);
        pass<F><<<148 * ctas_per_sm, 256>>>(x.keys, x.n, x.c0, x.bits0, seed_xor_p0(0), x.a1, x.c1, x.bits1, seed_xor_p0(1), x.dst, x.cursor, x.kc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaMemcpy(&cur, x.cursor, 8, cudaMemcpyDeviceToHost);
    printf("  %-58s %d CTAs/SM  %.3f ms   (list %llu keys)  %s\n", what, ctas_per_sm, best, cur, cudaGetErrorString(cudaGetLastError()));
}

// ---- v3: per-warp TMA key ring (2 x 2 KB), 1024-entry survivor ring, list reservation issued a full iteration before use
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_%=;\n}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
template <bool TMA, int DELAY, int CTAS, bool PRIV = false>
__global__ void __launch_bounds__(256, CTAS) pass3(const uint64_t* __restrict__ keys, uint64_t n, const uint32_t* __restrict__ c_pv,
                                                    uint32_t p_bits, uint64_t p_sp0, uint32_t* __restrict__ a_lv, uint32_t* __restrict__ c_lv,
                                                    uint32_t c_bits, uint64_t c_sp0, uint64_t* __restrict__ dst, unsigned long long* cursor,
                                                    KeyCfg kc) {
    constexpr int RING3 = (DELAY == 2 ? 4 : 2) * TILE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* ring_all = reinterpret_cast<uint64_t*>(smem_raw);                       // [WARPS][RING3]
    uint64_t* stage_all = ring_all + WARPS * RING3;                                   // [WARPS][2][TILE]
    uint64_t* mbar_all = stage_all + WARPS * 2 * TILE;                                // [WARPS][2]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint64_t* ring = ring_all + warp * RING3;
    uint64_t* stage = stage_all + warp * 2 * TILE;
    uint64_t* mbar = mbar_all + warp * 2;
    uint32_t head = 0, n_st = 0;
    bool r_old = false, r_new = false;
    unsigned long long t_old = 0, t_new = 0;
    uint32_t aidx[KPT], old[KPT], amask = 0;
#pragma unroll
    for (int j = 0; j < KPT; j++) aidx[j] = old[j] = 0;
    const uint64_t n_gw = (uint64_t)gridDim.x * WARPS, gw = (uint64_t)blockIdx.x * WARPS + warp;
    auto settle = [&]() {
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const uint32_t m = 1u << (aidx[j] & 31);
            if (((amask >> j) & 1u) && (old[j] & m)) red_or(c_lv + (aidx[j] >> 5), m);
        }
        amask = 0;
    };
    unsigned long long priv = gw * (n / n_gw + TILE);  // PRIV: the warp's own list region, no reservation at all
    auto flush = [&](uint32_t cnt, unsigned long long tk) {
        unsigned long long out;
        if (PRIV) { out = priv; priv += cnt; }
        else out = __shfl_sync(0xffffffffu, tk, 0);
        settle();
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const uint32_t i = lane + 32 * j;
            if (i < cnt) {
                const uint64_t kk = ring[(head + i) & (RING3 - 1)];
                dst[out + i] = kk;
                aidx[j] = hashmod_small(kc, c_sp0, kk, c_bits);
                amask |= 1u << j;
            }
        }
#pragma unroll
        for (int j = 0; j < KPT; j++)
            if ((amask >> j) & 1u) old[j] = atomicOr(a_lv + (aidx[j] >> 5), 1u << (aidx[j] & 31));
        head = (head + cnt) & (RING3 - 1);
        n_st -= cnt;
        __syncwarp();
    };
    const uint64_t n_tiles = n / TILE;
    uint64_t tile = gw;
    uint32_t it = 0, phase = 0;
    auto issue = [&](uint64_t t, int b) {
        if (TMA && lane == 0 && t < n_tiles) {
            mbar_expect_tx(&mbar[b], TILE * 8);
            bulk_load(stage + b * TILE, keys + t * TILE, TILE * 8, &mbar[b]);
        }
    };
    if (TMA) {
        if (lane == 0) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncwarp();
        issue(tile, 0);
        issue(tile + n_gw, 1);
    }
    uint64_t key[KPT];
    if (!TMA && tile < n_tiles) {
#pragma unroll
        for (int j = 0; j < KPT; j++) key[j] = ld_cg(keys + tile * TILE + lane + 32u * j);
    }
    while (tile < n_tiles) {
        if (TMA) {
            const int b = it & 1;
            mbar_wait(&mbar[b], (phase >> b) & 1u);
            phase ^= 1u << b;
#pragma unroll
            for (int j = 0; j < KPT; j++) key[j] = stage[b * TILE + lane + 32 * j];
            __syncwarp();
            if (lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(tile + 2 * n_gw, b);
        }
        uint32_t cw[KPT];
        uint64_t shs = 0;
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const uint32_t idx = hashmod_small(kc, p_sp0, key[j], p_bits);
            cw[j] = __ldcg(c_pv + (idx >> 5));
            shs |= (uint64_t)(idx & 31) << (5 * j);
        }
        if (r_old) {
            flush(TILE, t_old);
            r_old = false;
        }
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const bool redo = (cw[j] >> ((uint32_t)(shs >> (5 * j)) & 31)) & 1u;
            const uint32_t mk = __ballot_sync(0xffffffffu, redo);
            if (redo) ring[(head + n_st + __popc(mk & lt)) & (RING3 - 1)] = key[j];
            n_st += __popc(mk);
        }
        __syncwarp();
        tile += n_gw;
        it++;
        if (!TMA && tile < n_tiles) {
#pragma unroll
            for (int j = 0; j < KPT; j++) key[j] = ld_cg(keys + tile * TILE + lane + 32u * j);
            if (lane < 16 && tile + n_gw < n_tiles) asm volatile("prefetch.global.L2 [%0];" ::"l"(keys + (tile + n_gw) * TILE + lane * 16));
        }
        if (DELAY == 2) {
            r_old = r_new;
            t_old = t_new;
            r_new = false;
            if (n_st >= TILE * (r_old ? 2u : 1u)) {
                if (!PRIV && lane == 0) t_new = atomicAdd(cursor, (unsigned long long)TILE);
                r_new = true;
            }
        } else {
            if (n_st >= TILE) {
                if (!PRIV && lane == 0) t_old = atomicAdd(cursor, (unsigned long long)TILE);
                r_old = true;
            }
        }
    }
    if (r_old) flush(TILE, t_old);
    if (DELAY == 2 && r_new) flush(TILE, t_new);
    while (n_st) {
        const uint32_t cnt = n_st < TILE ? n_st : TILE;
        if (!PRIV && lane == 0) t_old = atomicAdd(cursor, (unsigned long long)cnt);
        flush(cnt, t_old);
    }
    settle();
}
template <bool TMA, int DELAY, int CTAS, bool PRIV = false>
void run3(const Ctx& x, const char* what) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const size_t smem = (size_t)WARPS * (DELAY == 2 ? 4 : 2) * TILE * 8 + (TMA ? WARPS * 2 * TILE * 8 : 0) + WARPS * 2 * 8;
    cudaFuncSetAttribute(pass3<TMA, DELAY, CTAS, PRIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float best = 1e9;
    unsigned long long cur = 0;
    for (int r = 0; r < 4; r++) {
        cudaMemset(x.a1, 0, x.w1 * 4);
        cudaMemset(x.c1, 0, x.w1 * 4);
        cudaMemset(x.cursor, 0, 8);
        cudaEventRecord(e0);
        pass3<TMA, DELAY, CTAS, PRIV><<<148 * CTAS, 256, smem>>>(x.keys, x.n, x.c0, x.bits0, seed_xor_p0(0), x.a1, x.c1, x.bits1, seed_xor_p0(1), x.dst,
                                                          x.cursor, x.kc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaMemcpy(&cur, x.cursor, 8, cudaMemcpyDeviceToHost);
    printf("  v3 %-55s %d CTAs/SM  %.3f ms   (list %llu keys)  %s\n", what, CTAS, best, cur, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    Ctx x;
    x.n = 100000000ull / TILE * TILE;
    x.kc = make_key_cfg(BPHF_KIND_U64, 8);
    x.bits0 = (uint32_t)(1.7 * x.n);
    x.bits1 = (uint32_t)(1.7 * 0.4449 * x.n);
    x.w0 = (x.bits0 + 31) / 32 + 16;
    x.w1 = (x.bits1 + 31) / 32 + 16;
    cudaMalloc(&x.keys, x.n * 8);
    cudaMalloc(&x.dst, (x.n + (1 << 22)) * 8);  // room for the private-region variant
    cudaMalloc(&x.a0, x.w0 * 4);
    cudaMalloc(&x.c0, x.w0 * 4);
    cudaMalloc(&x.a1, x.w1 * 4);
    cudaMalloc(&x.c1, x.w1 * 4);
    cudaMalloc(&x.cursor, 8);
    cudaMemset(x.a0, 0, x.w0 * 4);
    cudaMemset(x.c0, 0, x.w0 * 4);
    gen_keys<<<148 * 8, 256>>>(x.keys, x.n);
    mark0<<<148 * 8, 256>>>(x.keys, x.n, x.a0, x.c0, x.bits0, seed_xor_p0(0), x.kc);
    cudaDeviceSynchronize();
    printf("level-1 pass over %llu keys, level 0 = %u bits, level 1 = %u bits: %s\n", (unsigned long long)x.n, x.bits0, x.bits1,
           cudaGetErrorString(cudaGetLastError()));
    constexpr uint32_t ALL = F_KEYS_DRAM | F_PROBE | F_LIST | F_ATOM | F_RED | F_RESERVE;
    run3<false, 1, 3, true>(x, "ld.cg keys, per-warp list regions");
    run3<false, 1, 2, true>(x, "ld.cg keys, per-warp list regions");
    run3<true, 1, 3, true>(x, "TMA keys, per-warp list regions");
    run3<true, 1, 2, true>(x, "TMA keys, per-warp list regions");
    run3<true, 1, 1, true>(x, "TMA keys, per-warp list regions");
    run3<false, 1, 3>(x, "ld.cg keys, reservation 1 iteration ahead");
    run3<false, 2, 3>(x, "ld.cg keys, reservation 2 iterations ahead");
    run3<true, 1, 3>(x, "TMA keys, reservation 1 ahead");
    run3<true, 2, 3>(x, "TMA keys, reservation 2 ahead");
    run3<false, 2, 2>(x, "ld.cg keys, reservation 2 ahead");
    run3<true, 1, 2>(x, "TMA keys, reservation 1 ahead");
    run3<true, 2, 2>(x, "TMA keys, reservation 2 ahead");
    run3<true, 2, 1>(x, "TMA keys, reservation 2 ahead");
    for (int c : {3, 2}) {
        run<ALL>(x, "everything", c);
        run<ALL & ~F_KEYS_DRAM>(x, "keys computed, not streamed", c);
        run<ALL & ~F_PROBE>(x, "no collide probe (survivors by a hash bit)", c);
        run<ALL & ~F_LIST>(x, "no list store (reservation kept)", c);
        run<ALL & ~(F_LIST | F_RESERVE)>(x, "no list store, no reservation", c);
        run<ALL & ~F_RESERVE>(x, "list store into per-warp regions, no reservation", c);
        run<ALL & ~(F_ATOM | F_RED)>(x, "no mark of level 1", c);
        run<ALL & ~F_RED>(x, "no collide update (fetch_or only)", c);
        run<F_KEYS_DRAM | F_PROBE>(x, "keys + probe + compaction only", c);
        run<F_PROBE>(x, "probe + compaction only (keys computed)", c);
        run<F_KEYS_DRAM>(x, "key stream + hashing + compaction only", c);
        run<F_ATOM | F_RED>(x, "mark only (keys computed, survivors by a hash bit)", c);
    }
    return 0;
}
