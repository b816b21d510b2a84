// Micro-benchmark: bandwidth of reading only ONE 32-byte sector of every 64-byte record (the projection would need
// only {L0, L1, vx, vy} if those shared a sector), versus reading whole records.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void read_half(const int4 *__restrict__ in, long long n, int which, double *out)
{
    double acc = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // 2 lanes per record: lane pair reads the two 16-byte halves of the chosen sector
        const long long r = i >> 1;
        const int4 v = __ldcs(in + r * 4 + which * 2 + (i & 1));
        acc += v.x + v.z;
    }
    if (acc == 1.2345) *out = acc;
}
__global__ void read_full(const int4 *__restrict__ in, long long n, double *out)
{
    double acc = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int4 v = __ldcs(in + i);
        acc += v.x + v.z;
    }
    if (acc == 1.2345) *out = acc;
}
int main()
{
    const long long recs = 256ll << 20; // 16 GiB
    int4 *buf; double *out;
    cudaMalloc(&buf, recs * 64); cudaMalloc(&out, 8);
    cudaMemset(buf, 1, recs * 64);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int which = 0; which < 3; ++which) {
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(a);
            if (which < 2) read_half<<<148 * 16, 256>>>(buf, recs * 2, which, out);
            else read_full<<<148 * 16, 256>>>(buf, recs * 4, out);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        const double bytes = which < 2 ? recs * 32.0 : recs * 64.0;
        printf("%s: %.3f ms, %.2f TB/s useful\n", which == 0 ? "sector 0 only" : which == 1 ? "sector 1 only" : "whole records", best, bytes / best / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
