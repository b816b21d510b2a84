// micro-benchmark: cost of the scatter's store pattern.  4 arrays of 16-byte records (or 2 arrays of 32-byte records)
// copied with destination = runs of R records placed at pseudo-random run-aligned / mis-aligned positions.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k_copy16(const int4 *a, const int4 *b, const int4 *c, const int4 *d, int4 *A, int4 *B, int4 *C, int4 *D, const int *dest, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int j = dest[i];
        const int4 x = a[i], y = b[i], z = c[i], w = d[i];
        A[j] = x; B[j] = y; C[j] = z; D[j] = w;
    }
}
struct __align__(32) R32 { int4 lo, hi; };
__global__ void k_copy32(const R32 *a, const R32 *b, R32 *A, R32 *B, const int *dest, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int j = dest[i] % n;
        const R32 x = a[i], y = b[i];
        A[j] = x; B[j] = y;
    }
}
struct __align__(64) R64 { int4 a, b, c, d; };
__global__ void k_copy64(const R64 *a, R64 *A, const int *dest, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int j = dest[i] % n;
        const int4 x = a[i].a, y = a[i].b, z = a[i].c, w = a[i].d;
        A[j].a = x; A[j].b = y; A[j].c = z; A[j].d = w;
    }
}
// 4 lanes per 64-byte record: a warp instruction moves 8 whole records
__global__ void k_copy64c(const int4 *a, int4 *A, const int *dest, int n)
{
    const int lane4 = threadIdx.x & 3;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < 4ll * n; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t >> 2);
        const int j = dest[i] % n;
        A[4ll * j + lane4] = a[4ll * i + lane4];
    }
}
// dest pattern: particle i belongs to run r = i / R; runs are permuted within windows of W runs (local shuffle like the
// real scatter: destinations stay within a window), optional misalignment by `shift` records.
__global__ void k_make_dest(int *dest, int n, int R, int W, int shift)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / R, k = i % R;
        const int win = r / W, rr = r % W;
        const int pr = (int)(((long long)rr * 7919 + 13) % W); // permutation of the window (W prime to 7919)
        long long j = ((long long)win * W + pr) * R + k + shift;
        j %= n;
        dest[i] = (int)j;
    }
}
int main(int argc, char **argv)
{
    const int n = 1 << 27; // 134M records
    int4 *src[4], *dst[4];
    for (int k = 0; k < 4; ++k) { cudaMalloc(&src[k], (size_t)n * 16); cudaMalloc(&dst[k], (size_t)n * 16); cudaMemset(src[k], 1, (size_t)n * 16); }
    int *dest; cudaMalloc(&dest, (size_t)n * 4);
    printf("alloc: %s\n", cudaGetErrorString(cudaGetLastError()));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int Rs[] = {1 << 20, 16, 8, 4, 3, 2, 1};
    for (int R : Rs) for (int shift = 0; shift < 2; ++shift) {
        const int W = 4099; // prime window of runs
        k_make_dest<<<148 * 16, 256>>>(dest, n, R, W, shift);
        float ms16 = 0, ms32 = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k_copy16<<<148 * 16, 256>>>(src[0], src[1], src[2], src[3], dst[0], dst[1], dst[2], dst[3], dest, n);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms16, e0, e1);
        }
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k_copy32<<<148 * 16, 256>>>((R32 *)src[0], (R32 *)src[2], (R32 *)dst[0], (R32 *)dst[2], dest, n / 2);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms32, e0, e1);
        }
        float ms64 = 0, ms64c = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k_copy64<<<148 * 16, 256>>>((R64 *)src[0], (R64 *)dst[0], dest, n / 4);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms64, e0, e1);
        }
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k_copy64c<<<148 * 16, 256>>>(src[0], dst[0], dest, n / 4);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms64c, e0, e1);
        }
        printf("      1x64B: thread/record %.3f ms (%.2f TB/s)   4 lanes/record %.3f ms (%.2f TB/s)\n", ms64, 2.0 * (n / 4) * 64 / ms64 * 1e-9, ms64c,
               2.0 * (n / 4) * 64 / ms64c * 1e-9);
        // k_copy16 moves n*64 B in + out; k_copy32 moves (n/2)*64 B in + out
        { cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; } }
        printf("run %8d shift %d : 4x16B arrays %.3f ms (%.2f TB/s)   2x32B arrays %.3f ms (%.2f TB/s)\n", R, shift, ms16,
               2.0 * n * 64 / ms16 * 1e-9, ms32, 2.0 * (n / 2) * 64 / ms32 * 1e-9);
    }
    return 0;
}
