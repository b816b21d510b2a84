// pfem2_sort.cuh -- device-wide exclusive scan and stable LSD radix sort of (cell key, index) pairs.
//
// Hand-written for sm_100a (no CUB/Thrust on the product path).  Element counts live in DEVICE memory:
// every kernel is launched with a persistent grid sized for the SM count and loops over the tiles the
// live count implies, so the cost follows the actual number of elements (movers per step) and the
// particle step never needs a host round trip.
#pragma once

#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

namespace pfem2 {

extern std::atomic<long long> g_kernel_launches; // (handles may be driven from several host threads)
extern int g_num_sms;
#define PFEM2_LAUNCH(kernel, grid, block, smem, stream, ...)                                                        \
    do {                                                                                                            \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                                 \
        ::pfem2::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);                                                                             \
    } while (0)

inline int persistent_grid(long long max_blocks_needed, int blocks_per_sm)
{
    const long long cap = (long long)g_num_sms * blocks_per_sm;
    return (int)(max_blocks_needed < 1 ? 1 : (max_blocks_needed < cap ? max_blocks_needed : cap));
}

// ------------------------------------------------------------------------------------------------
// Exclusive scan, three kernels (reduce / spine / downsweep).  T is int or unsigned long long.
// out[i] = sum_{k<i} in[k] for i in [0, n]; out has n + 1 entries (out[n] = total).  in may alias out.
// n = *n_ptr * n_mul (device side), bounded by max_n on the host for scratch sizing.
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
template <class T> __device__ __forceinline__ T block_exclusive_scan(T v, T *total, T *smem /* 33 entries */)
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

__device__ __forceinline__ int scan_len(const int *n_ptr, int n_mul, int n_add) { return *n_ptr * n_mul + n_add; }

template <class T>
__global__ void __launch_bounds__(kScanThreads)
k_scan_reduce(const T *__restrict__ in, const int *__restrict__ n_ptr, int n_mul, int n_add, T *__restrict__ block_sums)
{
    __shared__ T sm[33];
    const int n = scan_len(n_ptr, n_mul, n_add);
    const int nb = (n + kScanTile - 1) / kScanTile;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const int base = b * kScanTile;
        T s = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            const int i = base + k * kScanThreads + threadIdx.x;
            if (i < n) s += in[i];
        }
        T total;
        block_exclusive_scan(s, &total, sm);
        if (threadIdx.x == 0) block_sums[b] = total;
    }
}

// single block: exclusive scan of the block sums in place; total appended at [nb]
template <class T>
__global__ void __launch_bounds__(kScanThreads) k_scan_spine(T *__restrict__ block_sums, const int *__restrict__ n_ptr, int n_mul, int n_add)
{
    __shared__ T sm[33];
    const int n = scan_len(n_ptr, n_mul, n_add);
    const int nb = (n + kScanTile - 1) / kScanTile;
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

// Epilogue of the down-sweep: store(i, x) is called with the exclusive prefix x of every element i, finish(total) once with the sum
// of all elements -- lets a caller derive a second array / close its bookkeeping without another pass (and launch) over the result.
struct ScanNoEpilogue {
    template <class T> __device__ __forceinline__ void store(int, T) const {}
    template <class T> __device__ __forceinline__ void finish(T) const {}
};

template <class T, class Epi>
__global__ void __launch_bounds__(kScanThreads)
k_scan_down(const T *in, const int *__restrict__ n_ptr, int n_mul, int n_add, const T *__restrict__ block_sums, T *out, Epi epi)
{
    __shared__ T sm[33];
    const int n = scan_len(n_ptr, n_mul, n_add);
    const int nb = (n + kScanTile - 1) / kScanTile;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const int base = b * kScanTile + threadIdx.x * kScanItems; // blocked arrangement: contiguous items per thread
        T v[kScanItems];
        T s = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            v[k] = (base + k < n) ? in[base + k] : T(0);
            s += v[k];
        }
        T total;
        T ex = block_exclusive_scan(s, &total, sm) + block_sums[b];
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (base + k < n) {
                out[base + k] = ex;
                epi.store(base + k, ex);
            }
            ex += v[k];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out[n] = block_sums[nb]; // total (spine wrote it)
        epi.finish(block_sums[nb]);
    }
}

template <class T> inline size_t scan_scratch_elems(long long max_n) { return (size_t)((max_n + kScanTile - 1) / kScanTile) + 2; }

// n = *n_ptr * n_mul + n_add elements (device side), at most max_n (host side, sizes scratch and grids).
template <class T, class Epi = ScanNoEpilogue>
inline void exclusive_scan_dev(const T *in, T *out, const int *n_ptr, int n_mul, int n_add, long long max_n, T *scratch, cudaStream_t st,
                               Epi epi = Epi())
{
    const long long nb_max = (max_n + kScanTile - 1) / kScanTile;
    const int grid = persistent_grid(nb_max, 8);
    PFEM2_LAUNCH(k_scan_reduce<T>, grid, kScanThreads, 0, st, in, n_ptr, n_mul, n_add, scratch);
    PFEM2_LAUNCH(k_scan_spine<T>, 1, kScanThreads, 0, st, scratch, n_ptr, n_mul, n_add);
    PFEM2_LAUNCH((k_scan_down<T, Epi>), grid, kScanThreads, 0, st, in, n_ptr, n_mul, n_add, scratch, out, epi);
}

// ------------------------------------------------------------------------------------------------
// Stable LSD radix sort of (key, value) pairs, 8-bit digits.
// Per pass: per-tile digit histogram -> exclusive scan over the digit-major histograms -> stable scatter.
// The number of tiles follows the device-side element count.
// ------------------------------------------------------------------------------------------------
constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;                   // keys per thread per tile, processed as kRsItems chunks of 256
constexpr int kRsTile = kRsThreads * kRsItems; // 4096 keys per tile
constexpr int kRsRadix = 256;

inline long long rs_num_tiles(long long capacity) { return (capacity + kRsTile - 1) / kRsTile; }
inline size_t rs_hist_elems(long long capacity) { return (size_t)kRsRadix * rs_num_tiles(capacity) + 1; }
inline size_t rs_scan_scratch_elems(long long capacity) { return scan_scratch_elems<int>(kRsRadix * rs_num_tiles(capacity)); }

// derived device-side lengths of one sort: info[0] = number of tiles
static __global__ void k_rs_prepare(const int *__restrict__ n_ptr, int *__restrict__ info)
{
    info[0] = (*n_ptr + kRsTile - 1) / kRsTile;
}

static __global__ void __launch_bounds__(kRsThreads)
k_rs_histogram(const unsigned *__restrict__ keys, const int *__restrict__ n_ptr, int shift, int *__restrict__ hist)
{
    __shared__ int sh[kRsRadix];
    const int n = *n_ptr;
    const int num_tiles = (n + kRsTile - 1) / kRsTile;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        sh[threadIdx.x] = 0;
        __syncthreads();
        const int base = tile * kRsTile;
#pragma unroll 4
        for (int k = 0; k < kRsItems; ++k) {
            const int i = base + k * kRsThreads + threadIdx.x;
            if (i < n) atomicAdd(&sh[(keys[i] >> shift) & 0xff], 1);
        }
        __syncthreads();
        hist[threadIdx.x * num_tiles + tile] = sh[threadIdx.x]; // digit-major so one scan gives global offsets
        __syncthreads();
    }
}

static __global__ void __launch_bounds__(kRsThreads)
k_rs_scatter(const unsigned *__restrict__ keys_in, const unsigned *__restrict__ vals_in, unsigned *__restrict__ keys_out,
             unsigned *__restrict__ vals_out, const int *__restrict__ n_ptr, int shift, const int *__restrict__ hist_scanned)
{
    __shared__ int digit_base[kRsRadix];            // running global offset of each digit for this tile
    __shared__ int warp_count[kRsThreads / 32][kRsRadix];
    const int n = *n_ptr;
    const int num_tiles = (n + kRsTile - 1) / kRsTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int base = tile * kRsTile;
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
        __syncthreads();
    }
}

// Sorts n (= *n_ptr <= capacity) pairs by the low `key_bits` bits of the key.  Ping-pongs between
// (keys, vals) and (keys_tmp, vals_tmp); returns 1 if the result ended up in the tmp buffers.
// hist: rs_hist_elems(capacity) ints; scan_scratch: rs_scan_scratch_elems(capacity) ints; info: 4 ints.
inline int radix_sort_pairs(unsigned *keys, unsigned *vals, unsigned *keys_tmp, unsigned *vals_tmp, const int *n_ptr,
                            long long capacity, int key_bits, int *hist, int *scan_scratch, int *info, cudaStream_t st)
{
    const long long tiles = rs_num_tiles(capacity);
    if (tiles == 0) return 0;
    const int grid = persistent_grid(tiles, 8);
    PFEM2_LAUNCH(k_rs_prepare, 1, 1, 0, st, n_ptr, info);
    int flip = 0;
    for (int shift = 0; shift < key_bits; shift += 8) {
        unsigned *ki = flip ? keys_tmp : keys, *vi = flip ? vals_tmp : vals;
        unsigned *ko = flip ? keys : keys_tmp, *vo = flip ? vals : vals_tmp;
        PFEM2_LAUNCH(k_rs_histogram, grid, kRsThreads, 0, st, ki, n_ptr, shift, hist);
        exclusive_scan_dev<int>(hist, hist, info, kRsRadix, 0, kRsRadix * tiles, scan_scratch, st);
        PFEM2_LAUNCH(k_rs_scatter, grid, kRsThreads, 0, st, ki, vi, ko, vo, n_ptr, shift, hist);
        flip ^= 1;
    }
    return flip;
}

} // namespace pfem2
