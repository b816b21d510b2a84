// Micro-benchmark: can the copy engine do the re-sort's scattered side?  sm_100 has cp.async.bulk.tensor ... tile::gather4 /
// tile::scatter4 (UTMALDG.2D.GATHER4 / UTMASTG.2D.SCATTER4): four arbitrary ROWS of a 2-D tensor per instruction.  With
// 64-byte particle records as rows, a warp could write its 32-record tile to 32 arbitrary destinations with 8 scatter4
// instructions issued by one lane (no LSU stores), or read 32 arbitrary records with 8 gather4 instructions.
// Copies n records per pass, per-warp double-buffered tiles like k_advect_locate_tma:
//   mode 0  tile load (32 rows)  -> tile store (32 rows)            the dense reference
//   mode 1  tile load            -> 8 x scatter4 to dest[]          move pass writing straight to the re-sorted positions
//   mode 2  8 x gather4 by src[] -> tile store                      move pass reading through a permutation (lazy re-sort)
//   mode 3  as mode 2, but lanes 0..7 issue one gather4 each and the 128 bytes of row indices of a tile arrive with ONE coalesced load
//           issued an iteration ahead (the form k_advect_locate_lazy uses since round 1e; mode 2 = one lane, each index load in front of its gather4)
// usage: tma_gather4_scatter4 <mode> <box_rows of the gather/scatter map: 1 or 4> <run length R> <log2 n> [swizzle: 0 none (default), 1 = 64-byte]
// swizzle 1 encodes ALL maps with CU_TENSOR_MAP_SWIZZLE_64B: mode 2 then answers the open question of the lazy re-sort in
// isolation -- do the rows of a gather4 land where the 32-row tile store expects them?  ("wrong pieces 0" = yes)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/_bin/tma_g4s4 tools/micro/tma_gather4_scatter4.cu
// Every configuration runs in its own process (an invalid tensor map / instruction poisons the context).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tile_load(uint32_t dst, const CUtensorMap *m, int row, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(m),
                 "r"(0), "r"(row), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tile_store(const CUtensorMap *m, int row, uint32_t src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(m), "r"(0), "r"(row), "r"(src) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap *m, int4 r, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
                 "l"(m), "r"(0), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void scatter4(const CUtensorMap *m, int4 r, uint32_t src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile::scatter4.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(m), "r"(0), "r"(r.x),
                 "r"(r.y), "r"(r.z), "r"(r.w), "r"(src)
                 : "memory");
}

constexpr int kTile = 32 * 64;
template <int MODE>
__global__ void __launch_bounds__(256, 4)
k_copy(const __grid_constant__ CUtensorMap tile_in, const __grid_constant__ CUtensorMap tile_out, const __grid_constant__ CUtensorMap g_in,
       const __grid_constant__ CUtensorMap s_out, const int4 *__restrict__ perm, int tiles)
{
    extern __shared__ unsigned char raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const uint32_t base = (s32(raw) + 1023u) & ~1023u;
    const uint32_t buf0 = base + warp * 2 * kTile, bar0 = base + wpb * 2 * kTile + warp * 16;
    if (lane == 0) {
        mbar_init(bar0);
        mbar_init(bar0 + 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int w0 = blockIdx.x * wpb + warp, wt = gridDim.x * wpb;
    auto issue_load = [&](int tile, uint32_t b) {
        mbar_expect(bar0 + b * 8, kTile);
        if (MODE == 2) {
            const int4 *p = perm + (size_t)tile * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) gather4(buf0 + b * kTile + k * 256, &g_in, __ldg(p + k), bar0 + b * 8);
        } else {
            tile_load(buf0 + b * kTile, &tile_in, tile << 5, bar0 + b * 8);
        }
    };
    auto load_rows = [&](int tile) {
        int4 r = make_int4(0, 0, 0, 0);
        if (lane < 8 && tile < tiles) r = __ldg(perm + (size_t)tile * 8 + lane);
        return r;
    };
    int4 rows = make_int4(0, 0, 0, 0);
    if (MODE == 3) {
        rows = load_rows(w0);
        if (w0 < tiles) {
            if (lane == 0) mbar_expect(bar0, kTile);
            __syncwarp();
            if (lane < 8) gather4(buf0 + lane * 256, &g_in, rows, bar0);
        }
        rows = load_rows(w0 + wt);
    } else if (lane == 0 && w0 < tiles) issue_load(w0, 0);
    uint32_t b = 0, par = 0;
    for (int tile = w0; tile < tiles; tile += wt) {
        if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (MODE == 3) {
                if (tile + wt < tiles) mbar_expect(bar0 + (b ^ 1) * 8, kTile);
            } else if (tile + wt < tiles) issue_load(tile + wt, b ^ 1);
        }
        if (MODE == 3) {
            __syncwarp();
            if (lane < 8 && tile + wt < tiles) gather4(buf0 + (b ^ 1) * kTile + lane * 256, &g_in, rows, bar0 + (b ^ 1) * 8);
            rows = load_rows(tile + 2 * wt);
        }
        mbar_wait(bar0 + b * 8, par);
        const uint32_t cur = buf0 + b * kTile;
        par ^= b;
        b ^= 1;
        __syncwarp();
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (MODE == 1) {
                const int4 *p = perm + (size_t)tile * 8;
#pragma unroll
                for (int k = 0; k < 8; ++k) scatter4(&s_out, __ldg(p + k), cur + k * 256);
            } else {
                tile_store(&tile_out, tile << 5, cur);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void k_make_perm(int *perm, int n, int R, int W, int shift)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / R, k = i % R;
        const int win = r / W, rr = r % W;
        const int pr = (int)(((long long)rr * 7919 + 13) % W);
        const long long j = ((long long)win * W + pr) * R + k + shift;
        perm[i] = (int)(j % n);
    }
}
__global__ void k_fill(int4 *a, long long pieces)
{
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < pieces; t += (long long)gridDim.x * blockDim.x)
        a[t] = make_int4((int)(t >> 2), (int)(t & 3), 0, 0); // record index in the first word of every 16-byte piece
}
__global__ void k_check(const int4 *src_or_dst, const int *perm, int n, int mode, unsigned long long *bad)
{
    // mode 1: dst[perm[i]] holds record i; modes 2, 3: dst[i] holds record perm[i]; mode 0: dst[i] holds record i
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int row = mode == 1 ? perm[i] : i, want = mode >= 2 ? perm[i] : i;
        for (int f = 0; f < 4; ++f) {
            const int4 v = src_or_dst[4ll * row + f];
            if (v.x != want || v.y != f) atomicAdd(bad, 1ull);
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int mode = argc > 1 ? atoi(argv[1]) : 0, box_rows = argc > 2 ? atoi(argv[2]) : 1, R = argc > 3 ? atoi(argv[3]) : 4;
    const int n = 1 << (argc > 4 ? atoi(argv[4]) : 26);
    const CUtensorMapSwizzle swz = (argc > 5 && atoi(argv[5]) == 1) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
    int4 *src, *dst;
    int *perm;
    unsigned long long *bad;
    CK(cudaMalloc(&src, (size_t)n * 64));
    CK(cudaMalloc(&dst, (size_t)n * 64));
    CK(cudaMalloc(&perm, (size_t)n * 4));
    CK(cudaMalloc(&bad, 8));
    CK(cudaMemset(bad, 0, 8));
    CK(cudaMemset(dst, 0xff, (size_t)n * 64));
    k_fill<<<148 * 8, 256>>>(src, 4ll * n);
    k_make_perm<<<148 * 8, 256>>>(perm, n, R, 4099, R > 1 ? 1 : 0);
    CK(cudaDeviceSynchronize());
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    auto make = [&](CUtensorMap *m, void *base, unsigned rows) {
        const cuuint64_t dims[2] = {16, (cuuint64_t)n};
        const cuuint64_t strides[1] = {64};
        const cuuint32_t box[2] = {16, rows};
        const cuuint32_t es[2] = {1, 1};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUtensorMap t_in, t_out, g_in, s_out;
    CUresult r0 = make(&t_in, src, 32), r1 = make(&t_out, dst, 32), r2 = make(&g_in, src, box_rows), r3 = make(&s_out, dst, box_rows);
    if (r0 || r1 || r2 || r3) { printf("mode %d box_rows %d: tensor map encode failed (%d %d %d %d)\n", mode, box_rows, r0, r1, r2, r3); return 3; }
    const int tiles = n / 32, grid = 148 * 4;
    const size_t smem = 8 * (2 * kTile + 16) + 1024;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k_copy<0><<<grid, 256, smem>>>(t_in, t_out, g_in, s_out, (const int4 *)perm, tiles);
        if (mode == 1) k_copy<1><<<grid, 256, smem>>>(t_in, t_out, g_in, s_out, (const int4 *)perm, tiles);
        if (mode == 2) k_copy<2><<<grid, 256, smem>>>(t_in, t_out, g_in, s_out, (const int4 *)perm, tiles);
        if (mode == 3) k_copy<3><<<grid, 256, smem>>>(t_in, t_out, g_in, s_out, (const int4 *)perm, tiles);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { printf("mode %d box_rows %d R %d: kernel failed: %s\n", mode, box_rows, R, cudaGetErrorString(e)); return 4; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    k_check<<<148 * 8, 256>>>(dst, perm, n, mode, bad);
    unsigned long long hbad = 0;
    CK(cudaMemcpy(&hbad, bad, 8, cudaMemcpyDeviceToHost));
    printf("mode %d box_rows %d run %d n %d swizzle %s : %.3f ms  %.2f G records/s  %.2f TB/s (in + out)  wrong pieces %llu\n", mode, box_rows, R, n, swz == CU_TENSOR_MAP_SWIZZLE_64B ? "64B" : "none", best, n / best * 1e-6,
           2.0 * n * 64 / best * 1e-9, hbad);
    return 0;
}
