// pfem2_mesh.cu -- mesh preparation helpers of the C ABI (the step right before the path; reference src/mesh_2d.cu) and the
// stand-alone radix sort entry point.
#include "pfem2_handle.cuh"

#include "pfem2_setup.cuh"

#include <algorithm>

using namespace pfem2;
using namespace pfem2::host;

extern "C" {

int pfem2_mesh_inv_jacobi(int n_cells, const double *d_vertices, const unsigned *d_cells, double *d_inv_jacobi, void *stream)
{
    pfem2_handle *h = nullptr;
    if (n_cells <= 0 || !d_vertices || !d_cells || !d_inv_jacobi) return fail(nullptr, PFEM2_EINVAL, "bad argument");
    PFEM2_LAUNCH(k_inv_jacobi, grid_for(n_cells, kThreads, 1 << 30), kThreads, 0, (cudaStream_t)stream, n_cells, (const double2 *)d_vertices,
                 d_cells, d_inv_jacobi);
    CU(cudaGetLastError());
    return PFEM2_OK;
}

int pfem2_mesh_band(int n_cells, const int *d_nbr_offsets, const int *d_nbr_indices, int *band, void *stream)
{
    pfem2_handle *h = nullptr;
    if (n_cells <= 0 || !d_nbr_offsets || !d_nbr_indices || !band) return fail(nullptr, PFEM2_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceTemps tmp;
    int *dev = nullptr;
    CU(tmp.alloc(&dev, 1));
    CU(cudaMemsetAsync(dev, 0, sizeof(int), st));
    PFEM2_LAUNCH(k_band_width, grid_for(n_cells, kThreads, 1 << 30), kThreads, 0, st, n_cells, d_nbr_offsets, d_nbr_indices, dev);
    CU(cudaMemcpyAsync(band, dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return PFEM2_OK;
}

int pfem2_sort_pairs(int n, int key_bits, unsigned *keys, unsigned *vals, unsigned *keys_tmp, unsigned *vals_tmp, int *result_in_tmp,
                     void *stream)
{
    pfem2_handle *h = nullptr;
    if (n < 0 || key_bits < 1 || key_bits > 32 || !keys || !vals || !keys_tmp || !vals_tmp || !result_in_tmp)
        return fail(nullptr, PFEM2_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    DeviceTemps tmp;
    int *n_dev = nullptr, *hist = nullptr, *scratch = nullptr;
    CU(tmp.alloc(&n_dev, 8));
    CU(tmp.alloc(&hist, rs_hist_elems(std::max(n, 1))));
    CU(tmp.alloc(&scratch, rs_scan_scratch_elems(std::max(n, 1))));
    CU(cudaMemcpyAsync(n_dev, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    *result_in_tmp = radix_sort_pairs(keys, vals, keys_tmp, vals_tmp, n_dev, n, key_bits, hist, scratch, n_dev + 4, st);
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return PFEM2_OK;
}

int pfem2_mesh_one_ring(int n_nodes, int n_cells, const unsigned *d_cells, int *d_offsets, int *d_indices, int *nnz, void *stream)
{
    pfem2_handle *h = nullptr;
    if (n_nodes <= 0 || n_cells <= 0 || !d_cells || !d_offsets || !nnz) return fail(nullptr, PFEM2_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int m = 3 * n_cells;
    DeviceTemps tmp; // every temporary is freed on every exit path
    unsigned *k0, *k1, *v0, *v1;
    int *count, *node_off, *n_dev, *hist, *scratch, *scratch2, *err, *info, *len_dev;
    CU(tmp.alloc(&k0, (size_t)m)); CU(tmp.alloc(&k1, (size_t)m));
    CU(tmp.alloc(&v0, (size_t)m)); CU(tmp.alloc(&v1, (size_t)m));
    CU(tmp.alloc(&count, (size_t)n_nodes + 1));
    CU(tmp.alloc(&node_off, (size_t)n_nodes + 1));
    CU(tmp.alloc(&n_dev, 1)); CU(tmp.alloc(&err, 1));
    CU(tmp.alloc(&hist, rs_hist_elems(m)));
    CU(tmp.alloc(&scratch, rs_scan_scratch_elems(m)));
    CU(tmp.alloc(&scratch2, scan_scratch_elems<int>(std::max(n_nodes, n_cells))));
    CU(tmp.alloc(&info, 4));
    CU(tmp.alloc(&len_dev, 2));
    CU(cudaMemsetAsync(count, 0, sizeof(int) * ((size_t)n_nodes + 1), st));
    CU(cudaMemsetAsync(err, 0, sizeof(int), st));
    CU(cudaMemcpyAsync(n_dev, &m, sizeof(int), cudaMemcpyHostToDevice, st));
    PFEM2_LAUNCH(k_incidence_keys, grid_for(m, kThreads, 1 << 30), kThreads, 0, st, n_cells, d_cells, k0, v0, count);
    int bits = 1;
    while ((1ll << bits) < n_nodes) ++bits;
    const int flip = radix_sort_pairs(k0, v0, k1, v1, n_dev, m, bits, hist, scratch, info, st);
    const unsigned *inc = flip ? v1 : v0;
    {
        const int lens[2] = {n_nodes, n_cells};
        CU(cudaMemcpyAsync(len_dev, lens, sizeof lens, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st)); // `lens` is a stack temporary
    }
    exclusive_scan_dev<int>(count, node_off, len_dev, 1, 0, n_nodes, scratch2, st);
    if (!d_indices) {
        int *counts = flip ? (int *)k0 : (int *)k1; // the key buffer the sort did not end in: free, >= n_cells ints
        PFEM2_LAUNCH(k_one_ring, grid_for(n_cells, 128, 1 << 30), 128, 0, st, n_cells, d_cells, node_off, inc, counts, nullptr, nullptr, err);
        exclusive_scan_dev<int>(counts, d_offsets, len_dev + 1, 1, 0, n_cells, scratch2, st);
    } else {
        PFEM2_LAUNCH(k_one_ring, grid_for(n_cells, 128, 1 << 30), 128, 0, st, n_cells, d_cells, node_off, inc, nullptr, d_offsets, d_indices, err);
    }
    int herr = 0;
    CU(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(nnz, d_offsets + n_cells, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    if (herr) return fail(nullptr, PFEM2_EINVAL, "a cell has more than 96 one-ring neighbours");
    return PFEM2_OK;
}

} // extern "C"
