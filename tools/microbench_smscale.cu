// microbench: do random L2 loads / atomics scale with the number of SMs issuing them (SM-side LSU bound) or not (L2-side bound)?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
template <int MODE>
__global__ void __launch_bounds__(1024) k(uint32_t* arr, uint64_t nwords, uint64_t nops, uint32_t* sink) {
    constexpr int KPT = 8;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * KPT;
    uint32_t acc = 0;
    for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x) * KPT + threadIdx.x; base < nops; base += stride) {
        uint64_t w[KPT]; uint32_t m[KPT], old[KPT];
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            uint64_t h = mix(base + (uint64_t)j * blockDim.x + 12345);
            w[j] = (uint64_t)(((h >> 32) * nwords) >> 32);
            m[j] = 1u << (h & 31);
        }
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < KPT; j++) atomicOr(arr + w[j], m[j]);
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < KPT; j++) old[j] = atomicOr(arr + w[j], m[j]);
#pragma unroll
            for (int j = 0; j < KPT; j++) acc += old[j];
        } else if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < KPT; j++) old[j] = __ldg(arr + w[j]);
#pragma unroll
            for (int j = 0; j < KPT; j++) acc += old[j];
        } else if (MODE == 3) {  // 16-byte random loads
#pragma unroll
            for (int j = 0; j < KPT; j++) { uint4 v = __ldg(reinterpret_cast<const uint4*>(arr) + (w[j] >> 2)); old[j] = v.x ^ v.w; }
#pragma unroll
            for (int j = 0; j < KPT; j++) acc += old[j];
        } else if (MODE == 4) {  // random loads where pairs of lanes share a sector
#pragma unroll
            for (int j = 0; j < KPT; j++) { uint64_t ws = __shfl_sync(0xffffffffu, w[j], threadIdx.x & 30); old[j] = __ldg(arr + (ws & ~7ull) + (threadIdx.x & 1)); }
#pragma unroll
            for (int j = 0; j < KPT; j++) acc += old[j];
        }
    }
    if (acc == 0x12345678) *sink = acc;
}
template <int MODE>
float run(uint32_t* arr, uint64_t nwords, uint64_t nops, uint32_t* sink, int grid) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaMemset(arr, 0, nwords * 4);
        cudaEventRecord(a);
        k<MODE><<<grid, 1024>>>(arr, nwords, nops, sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    const uint64_t nops = 50000000ULL;
    uint32_t *arr, *sink;
    cudaMalloc(&arr, 64 << 20); cudaMalloc(&sink, 4);
    const uint64_t nwords = (uint64_t)(21.25e6 / 4);
    const char* names[] = {"RED.OR", "ATOM.OR(ret)", "LDG 4B", "LDG 16B", "LDG 4B pairs share sector"};
    for (int grid : {18, 37, 74, 148, 296}) {
        float t[5] = {run<0>(arr, nwords, nops, sink, grid), run<1>(arr, nwords, nops, sink, grid), run<2>(arr, nwords, nops, sink, grid),
                      run<3>(arr, nwords, nops, sink, grid), run<4>(arr, nwords, nops, sink, grid)};
        for (int i = 0; i < 5; i++) printf("grid %4d x1024  %-28s %8.3f ms %7.1f Gops/s  %.3f ops/clk/CTA@1.9GHz\n", grid, names[i], t[i], nops / t[i] / 1e6, nops / t[i] / 1e6 / 1.9 / (grid > 148 ? 148 : grid));
    }
    return 0;
}
