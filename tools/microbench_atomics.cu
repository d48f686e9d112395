// microbench: random-address L2/HBM atomic, load and store rates on B200 ("L2-atomic ceiling" of north_star)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mb tools/microbench_atomics.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
template <int MODE, int KPT>
__global__ void __launch_bounds__(256) k(uint32_t* arr, uint32_t* arr2, uint64_t nwords, uint64_t nops, uint32_t* sink) {
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
        if (MODE == 0) {  // RED
#pragma unroll
            for (int j = 0; j < KPT; j++) atomicOr(arr + w[j], m[j]);
        } else if (MODE == 1) {  // ATOM with return
#pragma unroll
            for (int j = 0; j < KPT; j++) old[j] = atomicOr(arr + w[j], m[j]);
#pragma unroll
            for (int j = 0; j < KPT; j++) acc += old[j];
        } else if (MODE == 2) {  // ATOM + conditional RED (the mark step)
#pragma unroll
            for (int j = 0; j < KPT; j++) old[j] = atomicOr(arr + w[j], m[j]);
#pragma unroll
            for (int j = 0; j < KPT; j++) if (old[j] & m[j]) atomicOr(arr2 + w[j], m[j]);
        } else if (MODE == 3) {  // random 4B loads
#pragma unroll
            for (int j = 0; j < KPT; j++) old[j] = __ldg(arr + w[j]);
#pragma unroll
            for (int j = 0; j < KPT; j++) acc += old[j];
        } else if (MODE == 4) {  // random 4B stores
#pragma unroll
            for (int j = 0; j < KPT; j++) arr[w[j]] = m[j];
        } else if (MODE == 5) {  // 64-bit ATOM
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                unsigned long long o = atomicOr((unsigned long long*)arr + (w[j] >> 1), (unsigned long long)m[j] << ((w[j] & 1) * 32));
                acc += (uint32_t)o;
            }
        } else if (MODE == 6) {  // ATOM on interleaved (a,c) pair: a at 2w, c at 2w+1 (same sector)
#pragma unroll
            for (int j = 0; j < KPT; j++) old[j] = atomicOr(arr + 2 * (w[j] >> 1) , m[j]);
#pragma unroll
            for (int j = 0; j < KPT; j++) if (old[j] & m[j]) atomicOr(arr + 2 * (w[j] >> 1) + 1, m[j]);
        } else if (MODE == 7) {  // atomicAdd u32 with return (different ALU op)
#pragma unroll
            for (int j = 0; j < KPT; j++) old[j] = atomicAdd(arr + w[j], 1u);
#pragma unroll
            for (int j = 0; j < KPT; j++) acc += old[j];
        }
    }
    if (acc == 0x12345678) *sink = acc;
}
template <int MODE, int KPT>
float run(uint32_t* arr, uint32_t* arr2, uint64_t nwords, uint64_t nops, uint32_t* sink, int grid) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < 4; r++) {
        cudaMemset(arr, 0, nwords * 4); cudaMemset(arr2, 0, nwords * 4);
        cudaEventRecord(a);
        k<MODE, KPT><<<grid, 256>>>(arr, arr2, nwords, nops, sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    const uint64_t nops = 100000000ULL;
    uint32_t *arr, *arr2, *sink;
    const uint64_t maxw = 512ull << 20 >> 2;
    cudaMalloc(&arr, maxw * 4); cudaMalloc(&arr2, maxw * 4); cudaMalloc(&sink, 4);
    const char* names[] = {"RED.OR", "ATOM.OR(ret)", "ATOM+condRED(mark)", "LDG random 4B", "STG random 4B", "ATOM.OR.64(ret)", "mark interleaved a|c", "ATOM.ADD(ret)"};
    double mbs[] = {2.5, 21.25, 42.5, 94.5, 212.5, 425};
    for (double mb : mbs) {
        uint64_t nwords = (uint64_t)(mb * 1e6 / 4);
        printf("array %.1f MB (x2 for mark):\n", mb);
        for (int grid : {148 * 8, 148 * 16}) {
            float t[8];
            t[0] = run<0, 8>(arr, arr2, nwords, nops, sink, grid);
            t[1] = run<1, 8>(arr, arr2, nwords, nops, sink, grid);
            t[2] = run<2, 8>(arr, arr2, nwords, nops, sink, grid);
            t[3] = run<3, 8>(arr, arr2, nwords, nops, sink, grid);
            t[4] = run<4, 8>(arr, arr2, nwords, nops, sink, grid);
            t[5] = run<5, 8>(arr, arr2, nwords, nops, sink, grid);
            t[6] = run<6, 8>(arr, arr2, nwords, nops, sink, grid);
            t[7] = run<7, 8>(arr, arr2, nwords, nops, sink, grid);
            for (int i = 0; i < 8; i++) printf("  grid %5d KPT 8  %-24s %8.3f ms  %7.1f Gops/s\n", grid, names[i], t[i], nops / t[i] / 1e6);
        }
        float t1 = run<1, 1>(arr, arr2, nwords, nops, sink, 148 * 8);
        float t2 = run<1, 2>(arr, arr2, nwords, nops, sink, 148 * 8);
        float t4 = run<1, 4>(arr, arr2, nwords, nops, sink, 148 * 8);
        float t16 = run<1, 16>(arr, arr2, nwords, nops, sink, 148 * 8);
        printf("  ATOM.OR(ret) KPT 1/2/4/16: %.3f %.3f %.3f %.3f ms\n", t1, t2, t4, t16);
        float r16 = run<0, 16>(arr, arr2, nwords, nops, sink, 148 * 8);
        printf("  RED.OR KPT16: %.3f ms\n", r16);
    }
    return 0;
}
