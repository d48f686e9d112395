// microbench: do random L2 loads and random L2 atomics OVERLAP when different warps of the same SMs issue them, or do
// they queue for one resource?  (VERDICT r1 item 1 assumed two resources -- "SM-side sector limit" for the collide
// probes, "L2-side atomic limit" for the fetch_or's -- and asked for a schedule that loads both at once.)
//
// One launch, 8 warps per CTA: warps with (warp % 8) < LOAD_WARPS issue random 4-byte ld.cg on array B, the others
// random returning atomicOr on array A.  Each class has its own op budget; the kernel ends when both are done.
//   t_load  : loads alone (atomic budget 0)      t_atom : atomics alone      t_mix : both budgets in one launch
//   overlap = (t_load + t_atom - t_mix) / min(t_load, t_atom)     1 = perfectly parallel resources, 0 = one queue
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_mix tools/microbench_mix.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
constexpr int KPT = 8;
template <int LOAD_WARPS, bool RED>
__global__ void __launch_bounds__(256) k(uint32_t* a, const uint32_t* b, uint64_t nwords, uint64_t n_load, uint64_t n_atom,
                                        uint32_t* sink) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool loader = warp < LOAD_WARPS;
    const uint64_t my_n = loader ? n_load : n_atom;
    const uint64_t cls_warps = (uint64_t)gridDim.x * (loader ? LOAD_WARPS : 8 - LOAD_WARPS);
    const uint64_t my_w = (uint64_t)blockIdx.x * (loader ? LOAD_WARPS : 8 - LOAD_WARPS) + (loader ? warp : warp - LOAD_WARPS);
    uint32_t acc = 0;
    for (uint64_t base = my_w * 32 * KPT + lane; base < my_n; base += cls_warps * 32 * KPT) {
        uint64_t w[KPT];
        uint32_t m[KPT], old[KPT];
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const uint64_t h = mix(base + 32ull * j + (loader ? 999 : 12345));
            w[j] = (uint64_t)(((h >> 32) * nwords) >> 32);
            m[j] = 1u << (h & 31);
        }
        if (loader) {
#pragma unroll
            for (int j = 0; j < KPT; j++) old[j] = __ldcg(b + w[j]);
        } else if (RED) {
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                asm volatile("red.global.or.b32 [%0], %1;" ::"l"(a + w[j]), "r"(m[j]) : "memory");
                old[j] = 0;
            }
        } else {
#pragma unroll
            for (int j = 0; j < KPT; j++) old[j] = atomicOr(a + w[j], m[j]);
        }
#pragma unroll
        for (int j = 0; j < KPT; j++) acc += old[j];
    }
    if (acc == 0x12345678) *sink = acc;
}
template <int LW, bool RED>
float run(uint32_t* a, uint32_t* b, uint64_t nwords, uint64_t nl, uint64_t na, uint32_t* sink, int grid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; r++) {
        cudaMemset(a, 0, nwords * 4);
        cudaEventRecord(e0);
        k<LW, RED><<<grid, 256>>>(a, b, nwords, nl, na, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}
template <int LW, bool RED>
void report(const char* what, uint32_t* a, uint32_t* b, uint64_t nwords, uint64_t nl, uint64_t na, uint32_t* sink, int grid) {
    const float tl = run<LW, RED>(a, b, nwords, nl, 0, sink, grid);
    const float ta = run<LW, RED>(a, b, nwords, 0, na, sink, grid);
    const float tm = run<LW, RED>(a, b, nwords, nl, na, sink, grid);
    printf("  %-34s grid %5d  loads %.3f ms (%.0f G/s)  %s %.3f ms (%.0f G/s)  mixed %.3f ms  overlap %.2f\n", what, grid, tl,
           nl / tl / 1e6, RED ? "reds " : "atoms", ta, na / ta / 1e6, tm, (tl + ta - tm) / (tl < ta ? tl : ta));
}
int main() {
    uint32_t *a, *b, *sink;
    const uint64_t maxw = 256ull << 20 >> 2;
    cudaMalloc(&a, maxw * 4);
    cudaMalloc(&b, maxw * 4);
    cudaMalloc(&sink, 4);
    cudaMemset(b, 0, maxw * 4);
    for (double mb : {21.25, 42.5}) {
        const uint64_t nwords = (uint64_t)(mb * 1e6 / 4);
        printf("arrays %.2f MB each (L2 resident):\n", mb);
        for (int grid : {148 * 4, 148 * 8}) {
            // the cascade's level-1 mix: 1e8 probes, 4.45e7 fetch_or
            report<4, false>("4 load warps / 4 atom warps, 1e8 : 4.45e7", a, b, nwords, 100000000ull, 44500000ull, sink, grid);
            report<5, false>("5 / 3, 1e8 : 4.45e7", a, b, nwords, 100000000ull, 44500000ull, sink, grid);
            report<4, false>("4 / 4, 1e8 : 1e8", a, b, nwords, 100000000ull, 100000000ull, sink, grid);
            report<4, true>("4 / 4 (RED), 1e8 : 1e8", a, b, nwords, 100000000ull, 100000000ull, sink, grid);
        }
    }
    return 0;
}
