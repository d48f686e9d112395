// microbench_partial.cu -- do random 8-byte stores merge in the L2 when they are confined to a window?
// n stores into an array of n u64 slots (a permutation-like pattern: every slot about once); store number i goes to
// window i / window_slots, to a pseudo-random slot of it; the whole grid walks the store numbers together, so at any time
// all SMs write into the same window (or two).  Prints ns per store against the window size.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_partial tools/microbench_partial.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__global__ void scatter(uint64_t* out, uint64_t n, uint32_t wbits, int mode) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t p;
        if (mode == 0) p = i;                                                       // coalesced
        else p = ((i >> wbits) << wbits) | (mix(i) & ((1ull << wbits) - 1));        // random inside the window
        if (p < n) out[p] = i;
    }
}
int main() {
    const uint64_t n = 1ull << 29;  // 4 GB
    uint64_t* d;
    cudaMalloc(&d, n * 8);
    cudaMemset(d, 0, n * 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int grid = 148 * 8;
    auto run = [&](uint32_t wbits, int mode, const char* what) {
        scatter<<<grid, 256>>>(d, n, wbits, mode);
        cudaEventRecord(a);
        scatter<<<grid, 256>>>(d, n, wbits, mode);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("%-34s %8.2f ms  %6.1f ps per store  %7.1f GB/s of stored bytes\n", what, ms, ms * 1e9 / n, n * 8 / ms / 1e6);
    };
    run(0, 0, "coalesced");
    char buf[64];
    for (uint32_t w : {10u, 12u, 14u, 16u, 18u, 20u, 22u, 24u, 29u}) {
        snprintf(buf, sizeof buf, "random in windows of 2^%u slots", w);
        run(w, 1, buf);
    }
    return 0;
}
