// Does prefetch.global.L1 (CCTL.E.PF1) / .L2 (CCTL.E.PF2) shorten a later dependent load on sm_100a?
// One warp; per round every lane prefetches one random 128-byte line, waits ~3000 cycles, then the load of that line is timed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o prefetch_latency prefetch_latency.cu && ./prefetch_latency
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> // 0: no prefetch, 1: L1, 2: L2
__global__ void k(const int4 *a, size_t lines, long long *cyc, int rounds, int *sink)
{
    unsigned s = 12345u + threadIdx.x * 7919u;
    long long total = 0;
    int acc = 0;
    for (int r = 0; r < rounds; ++r) {
        s = s * 1664525u + 1013904223u;
        const int4 *p = a + (size_t)(s % lines) * 8;
        if (MODE == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
        if (MODE == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        const long long t0 = clock64();
        while (clock64() - t0 < 3000) {}
        __syncwarp();
        const long long t1 = clock64();
        int4 v;
        asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        acc += v.x + v.w;
        asm volatile("" ::"r"(acc));
        const long long t2 = clock64();
        total += t2 - t1;
    }
    if (threadIdx.x == 0) *cyc = total / rounds;
    if (acc == 42) *sink = acc;
}
int main()
{
    const size_t bytes = 1ull << 30, lines = bytes / 128;
    int4 *a; long long *cyc; int *sink;
    cudaMalloc(&a, bytes); cudaMemset(a, 0, bytes); cudaMallocManaged(&cyc, 8); cudaMalloc(&sink, 4);
    for (int pass = 0; pass < 2; ++pass) {
        k<0><<<1, 32>>>(a, lines, cyc, 2000, sink); cudaDeviceSynchronize(); printf("no prefetch      : %lld cycles per dependent load\n", *cyc);
        k<1><<<1, 32>>>(a, lines, cyc, 2000, sink); cudaDeviceSynchronize(); printf("prefetch.global.L1: %lld\n", *cyc);
        k<2><<<1, 32>>>(a, lines, cyc, 2000, sink); cudaDeviceSynchronize(); printf("prefetch.global.L2: %lld\n", *cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
