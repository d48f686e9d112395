// microbench: shared-memory atomicOr throughput per SM (random words in a 160 KB tile), B200
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
template <int MODE, int KPT>
__global__ void __launch_bounds__(1024) k(uint64_t ops_per_cta, uint32_t nwords, uint32_t* sink) {
    extern __shared__ uint32_t sm[];
    uint32_t* a = sm;
    uint32_t* c = sm + nwords;
    for (uint32_t i = threadIdx.x; i < 2 * nwords; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t acc = 0;
    for (uint64_t base = threadIdx.x; base < ops_per_cta; base += (uint64_t)blockDim.x * KPT) {
        uint32_t w[KPT], m[KPT], old[KPT];
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            uint64_t h = mix(base + (uint64_t)j * blockDim.x + blockIdx.x * 0x9E3779B97F4A7C15ULL);
            w[j] = (uint32_t)(((h >> 32) * nwords) >> 32);
            m[j] = 1u << (h & 31);
        }
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < KPT; j++) atomicOr(a + w[j], m[j]);
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < KPT; j++) old[j] = atomicOr(a + w[j], m[j]);
#pragma unroll
            for (int j = 0; j < KPT; j++) acc += old[j];
        } else if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < KPT; j++) old[j] = atomicOr(a + w[j], m[j]);
#pragma unroll
            for (int j = 0; j < KPT; j++) if (old[j] & m[j]) atomicOr(c + w[j], m[j]);
        } else if (MODE == 3) {  // hash only (ALU floor)
#pragma unroll
            for (int j = 0; j < KPT; j++) acc += w[j] ^ m[j];
        } else if (MODE == 4) {  // plain LDS random
#pragma unroll
            for (int j = 0; j < KPT; j++) acc += a[w[j]];
        }
    }
    __syncthreads();
    if (acc == 0x12345678 || sm[threadIdx.x] == 0xdeadbeef) *sink = acc;
}
template <int MODE, int KPT>
float run(uint64_t ops_per_cta, uint32_t nwords, uint32_t* sink, int threads) {
    cudaFuncSetAttribute(k<MODE, KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(a);
        k<MODE, KPT><<<148, threads, 2 * nwords * 4>>>(ops_per_cta, nwords, sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    uint32_t* sink; cudaMalloc(&sink, 4);
    const uint64_t ops = 100000000ULL / 148;
    const char* names[] = {"ATOMS.OR noret", "ATOMS.OR ret", "mark a+c", "hash only", "LDS random"};
    for (uint32_t nwords : {20000u, 10000u, 2048u}) {
        for (int threads : {1024, 512}) {
            printf("smem tile 2 x %u words, %d threads/CTA, 148 CTAs, 1e8 ops total\n", nwords, threads);
            float t0 = run<0, 8>(ops, nwords, sink, threads), t1 = run<1, 8>(ops, nwords, sink, threads), t2 = run<2, 8>(ops, nwords, sink, threads),
                  t3 = run<3, 8>(ops, nwords, sink, threads), t4 = run<4, 8>(ops, nwords, sink, threads);
            float t[] = {t0, t1, t2, t3, t4};
            for (int i = 0; i < 5; i++) printf("   %-16s %7.3f ms  %7.1f Gops/s\n", names[i], t[i], 1e8 / t[i] / 1e6);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
