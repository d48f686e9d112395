// microbench: tile-local binning scatter (keys -> B bucket regions, key 8 B + loc 4 B), variants of tile shape
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
constexpr uint64_t WY_P0 = 0xa0761d6478bd642fULL, WY_P1 = 0xe7037ed1a0b428dbULL, WY_FIN = 8ULL ^ 0xeb44accab455d165ULL;
__device__ __forceinline__ uint64_t wymum(uint64_t a, uint64_t b) { return __umul64hi(a, b) ^ (a * b); }
__device__ __forceinline__ uint32_t hashmod_small(uint64_t sp0, uint64_t key, uint32_t n) {
    uint64_t h = wymum(wymum(sp0, ((key << 32) | (key >> 32)) ^ WY_P1), WY_FIN);
    return __umulhi((uint32_t)h ^ (uint32_t)(h >> 32), n);
}
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__global__ void keygen(uint64_t* out, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t z = 1 + i * 0x9E3779B97F4A7C15ULL;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; out[i] = z ^ (z >> 31);
    }
}
struct P { uint64_t sp0; uint32_t bits, count, cap, range, mul; uint64_t* keys; uint32_t* loc; uint32_t* cursor; };

// MODE bit0: skip write-out, bit1: skip staging, bit2: skip global reservation, bit3: skip hist atomics
template <int THREADS, int KPT, int MODE>
__global__ void __launch_bounds__(THREADS) scatter(const uint64_t* __restrict__ in, uint64_t n, P d) {
    constexpr int TILE = THREADS * KPT;
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t* stage_key = (uint64_t*)smem;
    uint32_t* stage_loc = (uint32_t*)(smem + (size_t)TILE * 8);
    uint32_t* stage_off = stage_loc + TILE;
    uint32_t* hist = stage_off + TILE;
    uint32_t* sbase = hist + d.count;
    uint32_t* gbase = sbase + d.count;
    uint32_t* warp_tot = gbase + d.count;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t b = tid; b < d.count; b += THREADS) hist[b] = 0;
    __syncthreads();
    const uint64_t n_tiles = n / TILE;
    uint32_t sink = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t key[KPT];
        uint32_t bp[KPT], loc[KPT];
#pragma unroll
        for (int j = 0; j < KPT; j++) key[j] = ld_stream_u64(in + tile * TILE + j * THREADS + tid);
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const uint32_t idx = hashmod_small(d.sp0, key[j], d.bits);
            uint32_t b = __umulhi(idx, d.mul);
            uint32_t l = idx - b * d.range;
            if ((int32_t)l < 0) { b -= 1; l += d.range; }
            loc[j] = l;
            if (MODE & 8) { bp[j] = (b << 16) | (tid & 7); sink += l; }
            else bp[j] = (b << 16) | atomicAdd(&hist[b], 1u);
        }
        __syncthreads();
        const uint32_t per = (d.count + THREADS - 1) / THREADS;
        const uint32_t b0 = min(d.count, (uint32_t)tid * per), b1 = min(d.count, b0 + per);
        uint32_t sum = 0;
        for (uint32_t b = b0; b < b1; b++) sum += hist[b];
        uint32_t incl = sum;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, dd); if (lane >= dd) incl += t; }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; w++) { uint32_t t = warp_tot[w]; if (w < warp) woff += t; total += t; }
        uint32_t run = woff + incl - sum;
        for (uint32_t b = b0; b < b1; b++) {
            const uint32_t h = hist[b];
            sbase[b] = run; run += h;
            if (h) { gbase[b] = (MODE & 4) ? 0 : atomicAdd(d.cursor + (size_t)b * 32, h); hist[b] = 0; }
        }
        __syncthreads();
        if (!(MODE & 2)) {
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                const uint32_t b = bp[j] >> 16, ps = bp[j] & 0xffffu;
                const uint32_t s = sbase[b] + ps, o = gbase[b] + ps;
                stage_key[s] = key[j]; stage_loc[s] = loc[j];
                stage_off[s] = (o < d.cap) ? b * d.cap + o : 0xffffffffu;
            }
            __syncthreads();
            if (!(MODE & 1)) {
                for (uint32_t i = tid; i < total; i += THREADS) {
                    const uint32_t o = stage_off[i];
                    if (o != 0xffffffffu) { d.keys[o] = stage_key[i]; d.loc[o] = stage_loc[i]; }
                }
            }
            __syncthreads();
        } else { sink += bp[0] + total; }
    }
    if (sink == 0x12345) d.cursor[0] = sink;
}
template <int THREADS, int KPT, int MODE>
float run(const uint64_t* in, uint64_t n, P d, int ctas_per_sm) {
    constexpr int TILE = THREADS * KPT;
    size_t smem = (size_t)TILE * 16 + (size_t)d.count * 12 + 256;
    cudaFuncSetAttribute(scatter<THREADS, KPT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaMemset(d.cursor, 0, 4096 * 128);
        cudaEventRecord(a);
        scatter<THREADS, KPT, MODE><<<148 * ctas_per_sm, THREADS, smem>>>(in, n, d);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("ERR %s\n", cudaGetErrorString(e));
    return best;
}
int main() {
    const uint64_t n = 100000000ULL / 32768 * 32768;
    uint64_t* in; cudaMalloc(&in, n * 8);
    keygen<<<148 * 8, 256>>>(in, n);
    for (uint32_t count : {592u, 296u, 148u}) {
        P d; d.sp0 = 1 ^ WY_P0; d.bits = 170000000; d.count = count;
        d.range = ((d.bits + count - 1) / count + 511) & ~511u; d.count = (d.bits + d.range - 1) / d.range;
        d.mul = 0xffffffffu / d.range + 1;
        d.cap = (uint32_t)((double)n * d.range / d.bits * 1.02 + 256) & ~31u;
        cudaMalloc(&d.keys, (size_t)d.count * d.cap * 8); cudaMalloc(&d.loc, (size_t)d.count * d.cap * 4); cudaMalloc(&d.cursor, 4096 * 128);
        printf("buckets %u range %u cap %u\n", d.count, d.range, d.cap);
        printf("  512x8  2/SM full %.3f  nowrite %.3f  nostage %.3f  noreserve+nostage %.3f  hashonly %.3f ms\n",
               run<512, 8, 0>(in, n, d, 2), run<512, 8, 1>(in, n, d, 2), run<512, 8, 2>(in, n, d, 2), run<512, 8, 6>(in, n, d, 2), run<512, 8, 14>(in, n, d, 2));
        printf("  512x8  3/SM full %.3f\n", run<512, 8, 0>(in, n, d, 3));
        printf("  256x16 3/SM full %.3f  4/SM %.3f\n", run<256, 16, 0>(in, n, d, 3), run<256, 16, 0>(in, n, d, 4));
        printf("  256x8  6/SM full %.3f  (tile 2048)\n", run<256, 8, 0>(in, n, d, 6));
        printf("  1024x8 1/SM full %.3f  (tile 8192)\n", run<1024, 8, 0>(in, n, d, 1));
        printf("  1024x4 2/SM full %.3f  (tile 4096)\n", run<1024, 4, 0>(in, n, d, 2));
        printf("  512x16 1/SM full %.3f  (tile 8192)\n", run<512, 16, 0>(in, n, d, 1));
        cudaFree(d.keys); cudaFree(d.loc); cudaFree(d.cursor);
    }
    return 0;
}
