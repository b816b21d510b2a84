// pfem2_sort.cuh -- device-wide exclusive scan and stable LSD radix sort of (cell key, index) pairs.
//
// Hand-written for sm_100a (no CUB/Thrust on the product path).  Both primitives read their element
// count from DEVICE memory (`const int *n_ptr`) so the particle step never needs a host round trip:
// grids are sized from the host-known capacity and blocks beyond the live count exit immediately.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace pfem2 {

extern long long g_kernel_launches;
#define PFEM2_LAUNCH(kernel, grid, block, smem, stream, ...)                                                        \
    do {                                                                                                            \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                                 \
        ++::pfem2::g_kernel_launches;                                                                               \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Exclusive scan, three kernels (reduce / spine / downsweep).  T is int or unsigned long long.
// out[i] = sum_{k<i} in[k] for i in [0, n]; out has n + 1 entries (out[n] = total).  in may alias out.
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <class T> __device__ __forceinline__ T warp_inclusive_scan(T v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// block-wide exclusive scan of one value per thread (kScanThreads threads); returns exclusive prefix, total in *total
template <class T> __device__ __forceinline__ T block_exclusive_scan(T v, T *total, T *smem /* 32 entries */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T inc = warp_inclusive_scan(v, lane);
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        T w = lane < nw ? smem[lane] : T(0);
        const T winc = warp_inclusive_scan(w, lane);
        smem[lane] = winc - w; // exclusive warp offsets
        if (lane == nw - 1) smem[32] = winc;
    }
    __syncthreads();
    const T res = smem[warp] + inc - v;
    *total = smem[32];
    __syncthreads();
    return res;
}

template <class T> __global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const T *__restrict__ in, int n, T *__restrict__ block_sums)
{
    __shared__ T sm[33];
    const int base = blockIdx.x * kScanTile;
    T s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int i = base + k * kScanThreads + threadIdx.x;
        if (i < n) s += in[i];
    }
    T total;
    block_exclusive_scan(s, &total, sm);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of the block sums in place; total appended at [nb]
template <class T> __global__ void __launch_bounds__(kScanThreads) k_scan_spine(T *__restrict__ block_sums, int nb)
{
    __shared__ T sm[33];
    T carry = 0;
    for (int base = 0; base < nb; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const T v = i < nb ? block_sums[i] : T(0);
        T total;
        const T ex = block_exclusive_scan(v, &total, sm);
        if (i < nb) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) block_sums[nb] = carry;
}

template <class T>
__global__ void __launch_bounds__(kScanThreads) k_scan_down(const T *in, int n, const T *__restrict__ block_sums, T *out)
{
    __shared__ T sm[33];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems; // blocked arrangement: contiguous items per thread
    T v[kScanItems];
    T s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : T(0);
        s += v[k];
    }
    T total;
    T ex = block_exclusive_scan(s, &total, sm) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = block_sums[gridDim.x]; // total
}

template <class T> inline size_t scan_scratch_elems(int n) { return (size_t)((n + kScanTile - 1) / kScanTile) + 2; }

// in: n entries, out: n + 1 entries, scratch: scan_scratch_elems(n) entries.  n is host-known (cells, histogram bins).
template <class T> inline void exclusive_scan(const T *in, T *out, int n, T *scratch, cudaStream_t st)
{
    if (n <= 0) {
        cudaMemsetAsync(out, 0, sizeof(T), st);
        return;
    }
    const int nb = (n + kScanTile - 1) / kScanTile;
    PFEM2_LAUNCH(k_scan_reduce<T>, nb, kScanThreads, 0, st, in, n, scratch);
    PFEM2_LAUNCH(k_scan_spine<T>, 1, kScanThreads, 0, st, scratch, nb);
    PFEM2_LAUNCH(k_scan_down<T>, nb, kScanThreads, 0, st, in, n, scratch, out);
}

// ------------------------------------------------------------------------------------------------
// Stable LSD radix sort of (key, value) pairs, 8-bit digits.
// Per pass: per-tile digit histogram -> exclusive scan over (digit-major) histograms -> stable scatter.
// ------------------------------------------------------------------------------------------------
constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;                  // keys per thread per tile, processed as kRsItems chunks of 256
constexpr int kRsTile = kRsThreads * kRsItems; // 4096 keys per tile
constexpr int kRsRadix = 256;

inline int rs_num_tiles(int capacity) { return (capacity + kRsTile - 1) / kRsTile; }
// histogram scratch: 256 * tiles + 1 ints, plus scan scratch
inline size_t rs_hist_elems(int capacity) { return (size_t)kRsRadix * rs_num_tiles(capacity) + 1; }

__global__ void __launch_bounds__(kRsThreads)
k_rs_histogram(const unsigned *__restrict__ keys, const int *__restrict__ n_ptr, int shift, int num_tiles, int *__restrict__ hist)
{
    __shared__ int sh[kRsRadix];
    const int n = *n_ptr;
    const int tile = blockIdx.x;
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int base = tile * kRsTile;
    if (base < n) {
#pragma unroll 4
        for (int k = 0; k < kRsItems; ++k) {
            const int i = base + k * kRsThreads + threadIdx.x;
            if (i < n) atomicAdd(&sh[(keys[i] >> shift) & 0xff], 1);
        }
    }
    __syncthreads();
    hist[threadIdx.x * num_tiles + tile] = sh[threadIdx.x]; // digit-major so one scan gives global offsets
}

__global__ void __launch_bounds__(kRsThreads)
k_rs_scatter(const unsigned *__restrict__ keys_in, const unsigned *__restrict__ vals_in, unsigned *__restrict__ keys_out,
             unsigned *__restrict__ vals_out, const int *__restrict__ n_ptr, int shift, int num_tiles,
             const int *__restrict__ hist_scanned)
{
    __shared__ int digit_base[kRsRadix];            // running global offset of each digit for this tile
    __shared__ int warp_count[kRsThreads / 32][kRsRadix];
    const int n = *n_ptr;
    const int tile = blockIdx.x;
    const int base = tile * kRsTile;
    if (base >= n) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    digit_base[threadIdx.x] = hist_scanned[threadIdx.x * num_tiles + tile];
    for (int k = 0; k < kRsItems; ++k) {
        const int chunk = base + k * kRsThreads;
        if (chunk >= n) break;
#pragma unroll
        for (int w = 0; w < kRsThreads / 32; ++w) warp_count[w][threadIdx.x] = 0;
        __syncthreads();
        const int i = chunk + threadIdx.x;
        const bool valid = i < n;
        unsigned key = 0, val = 0;
        unsigned digit = 0xffffffffu; // invalid lanes form their own match group
        if (valid) {
            key = keys_in[i];
            val = vals_in[i];
            digit = (key >> shift) & 0xff;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int rank = __popc(peers & ((1u << lane) - 1));
        if (valid && rank == 0) warp_count[warp][digit] = __popc(peers);
        __syncthreads();
        // thread d: exclusive prefix over warps for digit d, then advance the running base
        {
            int run = digit_base[threadIdx.x];
#pragma unroll
            for (int w = 0; w < kRsThreads / 32; ++w) {
                const int c = warp_count[w][threadIdx.x];
                warp_count[w][threadIdx.x] = run;
                run += c;
            }
            digit_base[threadIdx.x] = run;
        }
        __syncthreads();
        if (valid) {
            const int pos = warp_count[warp][digit] + rank;
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        __syncthreads();
    }
}

// Sorts n (= *n_ptr <= capacity) pairs by the low `key_bits` bits of the key.  Ping-pongs between
// (keys, vals) and (keys_tmp, vals_tmp); returns 1 if the result ended up in the tmp buffers.
// hist: rs_hist_elems(capacity) ints; scan_scratch: scan_scratch_elems<int>(256 * tiles) ints.
inline int radix_sort_pairs(unsigned *keys, unsigned *vals, unsigned *keys_tmp, unsigned *vals_tmp, const int *n_ptr,
                            int capacity, int key_bits, int *hist, int *scan_scratch, cudaStream_t st)
{
    const int tiles = rs_num_tiles(capacity);
    if (tiles == 0) return 0;
    int flip = 0;
    for (int shift = 0; shift < key_bits; shift += 8) {
        unsigned *ki = flip ? keys_tmp : keys, *vi = flip ? vals_tmp : vals;
        unsigned *ko = flip ? keys : keys_tmp, *vo = flip ? vals : vals_tmp;
        PFEM2_LAUNCH(k_rs_histogram, tiles, kRsThreads, 0, st, ki, n_ptr, shift, tiles, hist);
        exclusive_scan<int>(hist, hist, kRsRadix * tiles, scan_scratch, st);
        PFEM2_LAUNCH(k_rs_scatter, tiles, kRsThreads, 0, st, ki, vi, ko, vo, n_ptr, shift, tiles, hist);
        flip ^= 1;
    }
    return flip;
}

} // namespace pfem2
