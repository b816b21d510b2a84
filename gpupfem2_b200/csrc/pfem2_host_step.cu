// pfem2_host_step.cu -- pfem2_step_host: one whole particle step with HOST nodal buffers (the end-to-end form of
//   advect(F) ; project(W) ; correct(F, W)
// of cases/Cylinder2D/main.cu:797-808,863-865), pipelined over PCIe: the upload of the nodal field overlaps the move pass and the
// download of the projected field overlaps the projection, chunk by chunk of the cell range.  The dependencies between chunks
// and node slices are derived from the mesh itself (band width of the one-ring lists x substeps), so any numbering is handled:
// a numbering without locality simply yields "wait for the whole upload" and "download at the end".
#include "pfem2_handle.cuh"

#include "pfem2_setup.cuh"

#include <algorithm>

using namespace pfem2;
using namespace pfem2::host;

namespace {

// Plan for K chunks (made once per (K, substeps)): cell chunk bounds, node slices of the upload, and per chunk the upload
// slices it depends on / the node prefix that is final after its projection.
int plan_host_pipe(pfem2_handle *h, int K, int substeps)
{
    pfem2_handle::HostPipe &pp = h->pipe;
    if (pp.K == K && pp.substeps == substeps) return PFEM2_OK;
    const int C = h->mesh.n_cells, N = h->mesh.n_nodes;
    cudaStream_t st = h->stream;
    if (!pp.copy) CU(cudaStreamCreateWithFlags(&pp.copy, cudaStreamNonBlocking));
    while ((int)pp.up_ev.size() < K) {
        cudaEvent_t a = nullptr, b = nullptr;
        CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        pp.up_ev.push_back(a);
        pp.dn_ev.push_back(b);
    }
    int rc;
    if ((rc = mesh_band(h))) return rc;
    pp.cb.resize(K + 1);
    pp.ns.resize(K + 1);
    for (int j = 0; j <= K; ++j) pp.cb[j] = (int)((long long)C * j / K);
    DeviceTemps tmp;
    int *dev = nullptr; // [cb (K+1) | up_need (K) | dn_ready (K)]
    CU(tmp.alloc(&dev, (size_t)(3 * K + 1)));
    std::vector<int> init(3 * K + 1, 0);
    for (int j = 0; j <= K; ++j) init[j] = pp.cb[j];
    for (int j = 0; j < K; ++j) init[1 + 2 * K + j] = N; // dn_ready starts at "everything"
    CU(cudaMemcpyAsync(dev, init.data(), sizeof(int) * init.size(), cudaMemcpyHostToDevice, st));
    const long long ext = std::min<long long>((long long)h->band * substeps, C);
    PFEM2_LAUNCH(k_chunk_node_ranges, grid_for(C, kThreads, 1 << 30), kThreads, 0, st, C, h->geom, K, dev, (int)ext, dev + 1 + K,
                 dev + 1 + 2 * K);
    std::vector<int> out(3 * K + 1);
    CU(cudaMemcpyAsync(out.data(), dev, sizeof(int) * out.size(), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // upload slice j = exactly the node prefix chunk j needs on top of what the chunks before it needed (chunks run in
    // order, so the dependency only grows); download prefix after chunk j likewise
    pp.up_slice.assign(K, 0);
    pp.dn_ready.assign(K, 0);
    int prev_up = 0, prev_dn = 0;
    pp.ns[0] = 0;
    for (int j = 0; j < K; ++j) {
        const int need = (j == K - 1) ? N : std::min(std::max(out[1 + K + j], 1), N); // node prefix [0, need) must have landed
        prev_up = std::max(prev_up, need);
        pp.ns[j + 1] = prev_up;
        pp.up_slice[j] = j;
        const int ready = (j == K - 1) ? N : std::min(out[1 + 2 * K + j], N);
        prev_dn = std::max(prev_dn, ready);
        pp.dn_ready[j] = prev_dn;
    }
    pp.K = K;
    pp.substeps = substeps;
    return PFEM2_OK;
}

int host_pipe_chunks(const pfem2_handle *h)
{
    if (h->opt.host_pipeline == 1 || h->opt.stable_order) return 1;
    if (h->own_lo != 0 || h->own_hi != h->mesh.n_cells) return 1; // multi-GPU strips exchange particles between the phases
    if (h->opt.host_pipeline > 1) return std::min(h->opt.host_pipeline, 16);
    return h->mesh.n_cells < (1 << 18) ? 1 : 8; // small meshes are launch-bound: one chunk (sweep on channel16m: 1 chunk 23.5 ms,
                                                // 2: 20.3, 4: 19.3, 6: 18.9, 8: 18.8, 12: 18.7; device time alone 17.8)
}

} // namespace

extern "C" int pfem2_step_host(pfem2_handle *h, const double *fx, const double *fy, double *wx, double *wy, double dt, int substeps,
                               int *count_out)
{
    if (!h || !fx || !fy || !wx || !wy) return PFEM2_EINVAL;
    CU(cudaSetDevice(h->device));
    const int N = h->mesh.n_nodes;
    const size_t nb = sizeof(double) * (size_t)N;
    for (double *&p : h->nodal)
        if (!p) CU(cudaMalloc((void **)&p, nb));
    cudaStream_t st = h->stream;
    int rc;
    const int K = host_pipe_chunks(h);
    if (K <= 1 || substeps < 1) {
        CU(cudaMemcpyAsync(h->nodal[0], fx, nb, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(h->nodal[1], fy, nb, cudaMemcpyHostToDevice, st));
        if ((rc = pfem2_advect(h, h->nodal[0], h->nodal[1], dt, substeps))) return rc;
        if ((rc = pfem2_project(h, h->nodal[2], h->nodal[3]))) return rc;
        if ((rc = pfem2_correct(h, h->nodal[0], h->nodal[1], h->nodal[2], h->nodal[3]))) return rc;
        CU(cudaMemcpyAsync(wx, h->nodal[2], nb, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(wy, h->nodal[3], nb, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if ((rc = sync_counters(h))) return rc;
        if (count_out) *count_out = h->host_count;
        return PFEM2_OK;
    }
    // Pipelined form: upload slices on the copy stream -> chunked move pass; chunked projection -> download slices on the
    // copy stream.  Same kernels, same arithmetic, same results as the three calls above.
    if ((rc = plan_host_pipe(h, K, substeps))) return rc;
    pfem2_handle::HostPipe &pp = h->pipe;
    {   // the nodal buffers may still be read by work of the caller's stream (previous step): order the uploads behind it
        CU(cudaEventRecord(pp.dn_ev[0], st));
        CU(cudaStreamWaitEvent(pp.copy, pp.dn_ev[0], 0));
    }
    for (int s = 0; s < K; ++s) {
        const size_t o = (size_t)pp.ns[s], len = (size_t)(pp.ns[s + 1] - pp.ns[s]) * sizeof(double);
        if (len) {
            CU(cudaMemcpyAsync(h->nodal[0] + o, fx + o, len, cudaMemcpyHostToDevice, pp.copy));
            CU(cudaMemcpyAsync(h->nodal[1] + o, fy + o, len, cudaMemcpyHostToDevice, pp.copy));
        }
        CU(cudaEventRecord(pp.up_ev[s], pp.copy));
    }
    pp.packed_slices = 0;
    pp.active = true;
    rc = pfem2_advect(h, h->nodal[0], h->nodal[1], dt, substeps); // the move pass runs chunk by chunk (advect_move)
    pp.active = false;
    if (rc) return rc;
    for (; pp.packed_slices < K; ++pp.packed_slices) CU(cudaStreamWaitEvent(st, pp.up_ev[pp.packed_slices], 0));
    {
        int done = 0; // nodes [0, done) are final and on their way to the host
        for (int j = 0; j < K; ++j) {
            {
                PhaseScope ps(h, PFEM2_PHASE_PROJECT_CELLS);
                launch_project_cells(h, pp.cb[j], pp.cb[j + 1]);
            }
            const int ready = pp.dn_ready[j];
            if (ready > done) {
                {
                    PhaseScope ps(h, PFEM2_PHASE_PROJECT_NODES);
                    launch_project_nodes(h, done, ready, h->nodal[2], h->nodal[3], nullptr, nullptr, nullptr, nullptr);
                }
                CU(cudaEventRecord(pp.dn_ev[j], st));
                CU(cudaStreamWaitEvent(pp.copy, pp.dn_ev[j], 0));
                const size_t len = (size_t)(ready - done) * sizeof(double);
                CU(cudaMemcpyAsync(wx + done, h->nodal[2] + done, len, cudaMemcpyDeviceToHost, pp.copy));
                CU(cudaMemcpyAsync(wy + done, h->nodal[3] + done, len, cudaMemcpyDeviceToHost, pp.copy));
                done = ready;
            }
        }
        h->partials_valid = true;
    }
    CU(cudaGetLastError());
    if ((rc = pfem2_correct(h, h->nodal[0], h->nodal[1], h->nodal[2], h->nodal[3]))) return rc;
    CU(cudaStreamSynchronize(pp.copy));
    CU(cudaStreamSynchronize(st));
    if ((rc = sync_counters(h))) return rc;
    if (count_out) *count_out = h->host_count;
    return PFEM2_OK;
}
