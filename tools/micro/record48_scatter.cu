// Micro-benchmark: would 48-byte particle records (x, y | vx, vy | cell, id, flags; local coordinates re-derived) pay?
// Copies N records with the destination made of runs of R consecutive records placed pseudo-randomly inside windows
// (like the real re-sort scatter), for 64-byte records (4 lanes per record) and 48-byte records (3 lanes per record,
// 30 of 32 lanes active), and reads the array sequentially at both record sizes.  Reports records/s and TB/s (in + out).
#include <cstdio>
#include <cuda_runtime.h>
template <int LANES> // 16-byte pieces per record: 4 -> 64 B, 3 -> 48 B
__global__ void k_copy(const int4 *a, int4 *A, const int *dest, int n)
{
    const int lane = threadIdx.x & 31;
    constexpr int per_warp = 32 / LANES; // records per warp instruction (8 or 10)
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int q = lane / LANES, f = lane % LANES;
    for (long long base = warp * per_warp; base < n; base += warps * per_warp) {
        const long long i = base + q;
        if (q < per_warp && i < n) {
            const int j = dest[i];
            A[(long long)LANES * j + f] = __ldcs(a + (long long)LANES * i + f);
        }
    }
}
template <int LANES> __global__ void k_read(const int4 *a, long long pieces, int *out)
{
    int acc = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < pieces; t += (long long)gridDim.x * blockDim.x) acc += __ldcs(a + t).x;
    if (acc == 123456789) *out = acc;
}
__global__ void k_make_dest(int *dest, int n, int R, int W, int shift)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / R, k = i % R;
        const int win = r / W, rr = r % W;
        const int pr = (int)(((long long)rr * 7919 + 13) % W);
        long long j = ((long long)win * W + pr) * R + k + shift;
        dest[i] = (int)(j % n);
    }
}
int main()
{
    const int n = 1 << 27; // 134M records
    int4 *src, *dst; int *dest, *out;
    cudaMalloc(&src, (size_t)n * 64); cudaMalloc(&dst, (size_t)n * 64); cudaMalloc(&dest, (size_t)n * 4); cudaMalloc(&out, 4);
    cudaMemset(src, 1, (size_t)n * 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](auto fn) { float best = 1e9, ms; for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); fn(); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; } return best; };
    const int Rs[] = {1 << 20, 8, 4, 3, 2};
    for (int R : Rs) for (int shift = 0; shift < 2; ++shift) {
        k_make_dest<<<148 * 16, 256>>>(dest, n, R, 4099, shift);
        const float t64 = timeit([&] { k_copy<4><<<148 * 16, 256>>>(src, dst, dest, n); });
        const float t48 = timeit([&] { k_copy<3><<<148 * 16, 256>>>(src, dst, dest, n); });
        printf("run %8d shift %d : 64 B %.3f ms (%.2f G rec/s, %.2f TB/s)   48 B %.3f ms (%.2f G rec/s, %.2f TB/s)\n", R, shift, t64, n / t64 * 1e-6,
               2.0 * n * 64 / t64 * 1e-9, t48, n / t48 * 1e-6, 2.0 * n * 48 / t48 * 1e-9);
    }
    const float r64 = timeit([&] { k_read<4><<<148 * 16, 256>>>(src, 4ll * n, out); });
    const float r48 = timeit([&] { k_read<3><<<148 * 16, 256>>>(src, 3ll * n, out); });
    printf("sequential read: 64 B %.3f ms (%.2f G rec/s)   48 B %.3f ms (%.2f G rec/s)\n", r64, n / r64 * 1e-6, r48, n / r48 * 1e-6);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
