// microbench: how fast does ONE global word take returning atomicAdds from all SMs (one lane per warp, the pattern of a
// list-append cursor or a tile-claim counter)?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_sameaddr tools/microbench_sameaddr.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
template <bool RET>
__global__ void __launch_bounds__(256) k(unsigned long long* ctr, int n_ctr, int stride_words, int per_warp, unsigned long long* sink) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long acc = 0;
    unsigned long long* c = ctr + (size_t)(gw % n_ctr) * stride_words;
    for (int i = 0; i < per_warp; i++) {
        if (lane == 0) {
            if (RET) acc += atomicAdd(c, 256ull);
            else asm volatile("red.global.add.u64 [%0], %1;" ::"l"(c), "l"(256ull) : "memory");
        }
        acc = __shfl_sync(0xffffffffu, acc, 0);  // the warp waits for the result, as a list reservation does
    }
    if (acc == 0x123456789ull) *sink = acc;
}
int main() {
    unsigned long long *ctr, *sink;
    cudaMalloc(&ctr, 1 << 20);
    cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int grid : {148, 148 * 3, 148 * 4}) {
        for (int n_ctr : {1, 2, 4, 16, 64}) {
            for (int ret = 1; ret >= 0; ret--) {
                const int per_warp = 64;
                float best = 1e9;
                for (int r = 0; r < 3; r++) {
                    cudaMemset(ctr, 0, 1 << 20);
                    cudaEventRecord(e0);
                    if (ret) k<true><<<grid, 256>>>(ctr, n_ctr, 16, per_warp, sink);
                    else k<false><<<grid, 256>>>(ctr, n_ctr, 16, per_warp, sink);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                const double ops = (double)grid * 8 * per_warp;
                printf("grid %4d  counters %3d (128 B apart)  %s  %.0f ops in %.3f ms = %.1f ns per op, %.0f M ops/s\n", grid, n_ctr,
                       ret ? "ATOM(ret)" : "RED      ", ops, best, best * 1e6 / ops, ops / best / 1e3);
            }
        }
    }
    return 0;
}
