// pfem2_host_step.cu -- the particle step with HOST nodal buffers (the end-to-end form of
//   advect(F) ; project(W) ; correct(F, W)
// of cases/Cylinder2D/main.cu:797-808,863-865), pipelined over PCIe: the upload of the nodal field overlaps the move pass and the
// download of the projected field overlaps the projection, chunk by chunk of the cell range.  The dependencies between chunks
// and node slices are derived from the mesh itself (band width of the one-ring lists x substeps), so any numbering is handled:
// a numbering without locality simply yields "wait for the whole upload" and "download at the end".
//   pfem2_step_host       one GPU (the whole mesh)
//   pfem2_step_host_p2p   one strip of a partitioned multi-GPU run: the same pipeline over the strip's own cells, with the P2P
//                         migration between the move pass and the rank pass and the P2P halo sum in front of the interface nodes'
//                         division -- the C-side driver of the multi-GPU step (no Python in the loop)
#include "pfem2_handle.cuh"

#include "pfem2_setup.cuh"

#include <algorithm>
#include <cstdlib>

using namespace pfem2;
using namespace pfem2::host;

namespace {

// Plan for K chunks of the OWNED cell range (made once per (K, substeps)): cell chunk bounds, node slices of the upload, and per
// chunk the upload slices it depends on / the node prefix that is final after its projection.  Nodes: the advect reads
// [in_lo, in_hi) (pfem2_node_ranges), the projection completes [own_node_lo, own_node_hi).
int plan_host_pipe(pfem2_handle *h, int K, int substeps)
{
    pfem2_handle::HostPipe &pp = h->pipe;
    if (pp.K == K && pp.substeps == substeps) return PFEM2_OK;
    const int C = h->mesh.n_cells;
    cudaStream_t st = h->stream;
    if (!pp.copy) CU(cudaStreamCreateWithFlags(&pp.copy, cudaStreamNonBlocking));
    while ((int)pp.up_ev.size() < K + 2) {
        cudaEvent_t a = nullptr, b = nullptr;
        CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        pp.up_ev.push_back(a);
        pp.dn_ev.push_back(b);
    }
    int rc;
    if ((rc = mesh_band(h))) return rc;
    if ((rc = ensure_v2_node_range(h, substeps))) return rc;
    const int in_lo = h->v2_node_lo, in_hi = h->v2_node_hi, out_lo = h->own_node_lo, out_hi = h->own_node_hi;
    pp.cb.resize(K + 1);
    pp.ns.resize(K + 1);
    // tapered chunks (weights 1, 2, 3, .. up to the middle and down again): the first upload slice and the last download slice are
    // the copies nothing can hide, so the chunks at both ends are the small ones
    const long long own_n = h->own_hi - h->own_lo;
    {
        std::vector<long long> w(K);
        long long total = 0;
        for (int j = 0; j < K; ++j) total += (w[j] = std::min(j, K - 1 - j) + 1);
        long long acc = 0;
        pp.cb[0] = h->own_lo;
        for (int j = 0; j < K; ++j) {
            acc += w[j];
            pp.cb[j + 1] = h->own_lo + (int)(own_n * acc / total);
        }
    }
    DeviceTemps tmp;
    int *dev = nullptr; // [cb (K+1) | up_need (K) | dn_ready (K)]
    CU(tmp.alloc(&dev, (size_t)(3 * K + 1)));
    std::vector<int> init(3 * K + 1, 0);
    for (int j = 0; j <= K; ++j) init[j] = pp.cb[j];
    for (int j = 0; j < K; ++j) init[1 + 2 * K + j] = out_hi; // dn_ready starts at "everything"
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
    int prev_up = in_lo, prev_dn = out_lo;
    pp.ns[0] = in_lo;
    for (int j = 0; j < K; ++j) {
        const int need = (j == K - 1) ? in_hi : std::min(std::max(out[1 + K + j], in_lo), in_hi); // nodes [in_lo, need) must have landed
        prev_up = std::max(prev_up, need);
        pp.ns[j + 1] = prev_up;
        pp.up_slice[j] = j;
        const int ready = (j == K - 1) ? out_hi : std::min(std::max(out[1 + 2 * K + j], out_lo), out_hi);
        prev_dn = std::max(prev_dn, ready);
        pp.dn_ready[j] = prev_dn;
    }
    pp.K = K;
    pp.substeps = substeps;
    return PFEM2_OK;
}

int host_pipe_chunks(const pfem2_handle *h, bool strip)
{
    if (h->opt.host_pipeline == 1 || h->opt.stable_order) return 1;
    if (!strip && (h->own_lo != 0 || h->own_hi != h->mesh.n_cells)) return 1; // strips go through pfem2_step_host_p2p
    if (h->opt.host_pipeline > 1) return std::min(h->opt.host_pipeline, 16);
    const long long own = h->own_hi - h->own_lo;
    if (own < (1 << 18)) return 1; // small meshes are launch-bound: one chunk
    // about 256K cells per chunk, 2..8 chunks.  Sweeps: channel16m on one GPU, uniform chunks: 1 chunk 23.5 ms, 2: 20.3, 4: 19.3, 6: 18.9,
    // 8: 18.8, 12: 18.7 against 17.8 of device time (tapered since: 8 chunks 14.78 against 14.47).  A strip of an 8-GPU run (2M cells,
    // 1.9 ms of device time): 2 chunks 3.12 ms, 4: 2.98, 8: 2.88 (the move chunks overlap on two streams, so more chunks cost little
    // and shorten the first upload / last download slice); PFEM2_PIPE_CELLS_LOG2 overrides the chunk size for A/B runs
    static const int shift = getenv("PFEM2_PIPE_CELLS_LOG2") ? atoi(getenv("PFEM2_PIPE_CELLS_LOG2")) : 18;
    return (int)std::min<long long>(8, std::max<long long>(2, own >> shift));
}

// strip == true: this handle is one strip of a partitioned run with connected P2P inboxes (rank = its index)
int step_host(pfem2_handle *h, bool strip, int rank, const double *fx, const double *fy, double *wx, double *wy, double dt, int substeps,
              int *count_out)
{
    if (!h || !fx || !fy || !wx || !wy) return PFEM2_EINVAL;
    if (substeps < 1) return fail(h, PFEM2_EINVAL, "particleSubsteps must be >= 1");
    CU(cudaSetDevice(h->device));
    const int N = h->mesh.n_nodes;
    const size_t nb = sizeof(double) * (size_t)N;
    for (double *&p : h->nodal)
        if (!p) CU(cudaMalloc((void **)&p, nb));
    if (strip && !h->acc3) {
        CU(cudaMalloc((void **)&h->acc3, 3 * nb));
        CU(cudaMemsetAsync(h->acc3, 0, 3 * nb, h->stream)); // nodes no owned cell touches are never written
    }
    cudaStream_t st = h->stream;
    int rc;
    if ((rc = ensure_v2_node_range(h, substeps))) return rc;
    const int in_lo = h->v2_node_lo, in_hi = h->v2_node_hi, out_lo = h->own_node_lo, out_hi = h->own_node_hi;
    auto advect = [&]() -> int { // advectParticles; strips: with the migration between the move pass and the rank pass
        if (!strip) return pfem2_advect(h, h->nodal[0], h->nodal[1], dt, substeps);
        int r;
        if ((r = pfem2_advect_move(h, h->nodal[0], h->nodal[1], dt, substeps))) return r;
        if ((r = pfem2_emigrants_send_p2p(h, rank))) return r;
        if ((r = pfem2_immigrants_recv_p2p(h))) return r;
        return pfem2_advect_finish(h, h->nodal[0], h->nodal[1]);
    };
    // strips: the interface nodes' sums get the neighbour's share before the division; their slices are finalised and downloaded last
    auto finish_interfaces = [&](cudaStream_t copy) -> int {
        int r;
        if ((r = pfem2_project_halo_p2p(h, h->acc3))) return r;
        for (int side = 0; side < 2; ++side) {
            const int lo = std::max(h->p2p.idx_lo[side], out_lo), hi = std::min(h->p2p.idx_hi[side], out_hi);
            if (!h->p2p.peer[side] || hi <= lo) continue;
            launch_project_finalize_range(h, lo, hi, h->acc3, h->nodal[2], h->nodal[3]);
            if (copy != st) {
                CU(cudaEventRecord(h->pipe.dn_ev[h->pipe.K + side], st));
                CU(cudaStreamWaitEvent(copy, h->pipe.dn_ev[h->pipe.K + side], 0));
            }
            const size_t len = (size_t)(hi - lo) * sizeof(double);
            CU(cudaMemcpyAsync(wx + lo, h->nodal[2] + lo, len, cudaMemcpyDeviceToHost, copy));
            CU(cudaMemcpyAsync(wy + lo, h->nodal[3] + lo, len, cudaMemcpyDeviceToHost, copy));
        }
        return PFEM2_OK;
    };
    const int K = host_pipe_chunks(h, strip);
    if (K <= 1) {
        const size_t ilen = sizeof(double) * (size_t)(in_hi - in_lo), olen = sizeof(double) * (size_t)(out_hi - out_lo);
        if (ilen) {
            CU(cudaMemcpyAsync(h->nodal[0] + in_lo, fx + in_lo, ilen, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(h->nodal[1] + in_lo, fy + in_lo, ilen, cudaMemcpyHostToDevice, st));
        }
        if ((rc = advect())) return rc;
        if (!strip) {
            if ((rc = pfem2_project(h, h->nodal[2], h->nodal[3]))) return rc;
        } else {
            if ((rc = flush_correct(h))) return rc;
            {
                PhaseScope ps(h, PFEM2_PHASE_PROJECT_CELLS);
                launch_project_cells(h);
            }
            PhaseScope ps(h, PFEM2_PHASE_PROJECT_NODES);
            launch_project_nodes_acc_range(h, out_lo, out_hi, h->acc3);
            if ((rc = pfem2_project_halo_p2p(h, h->acc3))) return rc;
            launch_project_finalize_range(h, out_lo, out_hi, h->acc3, h->nodal[2], h->nodal[3]);
        }
        if ((rc = pfem2_correct(h, h->nodal[0], h->nodal[1], h->nodal[2], h->nodal[3]))) return rc;
        if (olen) {
            CU(cudaMemcpyAsync(wx + out_lo, h->nodal[2] + out_lo, olen, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(wy + out_lo, h->nodal[3] + out_lo, olen, cudaMemcpyDeviceToHost, st));
        }
        CU(cudaStreamSynchronize(st));
        if ((rc = sync_counters(h))) return rc;
        if (count_out) *count_out = h->host_count;
        return PFEM2_OK;
    }
    // Pipelined form: upload slices on the copy stream -> chunked move pass; chunked projection -> download slices on the
    // copy stream.  Same kernels, same arithmetic, same results as the plain calls.
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
    rc = advect(); // the move pass runs chunk by chunk (advect_move)
    pp.active = false;
    if (rc) return rc;
    for (; pp.packed_slices < K; ++pp.packed_slices) CU(cudaStreamWaitEvent(st, pp.up_ev[pp.packed_slices], 0));
    {
        int done = out_lo; // nodes [out_lo, done) are final (strips: but for the interface nodes) and on their way to the host
        for (int j = 0; j < K; ++j) {
            {
                PhaseScope ps(h, PFEM2_PHASE_PROJECT_CELLS);
                launch_project_cells(h, pp.cb[j], pp.cb[j + 1]);
            }
            const int ready = pp.dn_ready[j];
            if (ready > done) {
                {
                    PhaseScope ps(h, PFEM2_PHASE_PROJECT_NODES);
                    if (!strip) {
                        launch_project_nodes(h, done, ready, h->nodal[2], h->nodal[3], nullptr, nullptr, nullptr, nullptr);
                    } else { // (interface nodes get a provisional value here; finish_interfaces overwrites it on both sides of the bus)
                        launch_project_nodes_acc_range(h, done, ready, h->acc3);
                        launch_project_finalize_range(h, done, ready, h->acc3, h->nodal[2], h->nodal[3]);
                    }
                }
                CU(cudaEventRecord(pp.dn_ev[j], st));
                CU(cudaStreamWaitEvent(pp.copy, pp.dn_ev[j], 0));
                const size_t len = (size_t)(ready - done) * sizeof(double);
                CU(cudaMemcpyAsync(wx + done, h->nodal[2] + done, len, cudaMemcpyDeviceToHost, pp.copy));
                CU(cudaMemcpyAsync(wy + done, h->nodal[3] + done, len, cudaMemcpyDeviceToHost, pp.copy));
                done = ready;
            }
        }
        if (strip) {
            PhaseScope ps(h, PFEM2_PHASE_PROJECT_NODES);
            if ((rc = finish_interfaces(pp.copy))) return rc;
        }
        h->partials_valid = !strip;
    }
    CU(cudaGetLastError());
    if ((rc = pfem2_correct(h, h->nodal[0], h->nodal[1], h->nodal[2], h->nodal[3]))) return rc;
    CU(cudaStreamSynchronize(pp.copy));
    CU(cudaStreamSynchronize(st));
    if ((rc = sync_counters(h))) return rc;
    if (count_out) *count_out = h->host_count;
    return PFEM2_OK;
}

} // namespace

extern "C" {

int pfem2_step_host(pfem2_handle *h, const double *fx, const double *fy, double *wx, double *wy, double dt, int substeps, int *count_out)
{
    return step_host(h, false, 0, fx, fy, wx, wy, dt, substeps, count_out);
}

int pfem2_step_host_p2p(pfem2_handle *h, int rank, const double *fx, const double *fy, double *wx, double *wy, double dt, int substeps,
                        int *count_out)
{
    if (!h) return PFEM2_EINVAL;
    if (rank < 0 || rank >= h->mg_ranks) return fail(h, PFEM2_EINVAL, "rank outside the rank bounds (pfem2_set_rank_bounds first)");
    if ((rank > 0 && !h->p2p.peer[0]) || (rank + 1 < h->mg_ranks && !h->p2p.peer[1]))
        return fail(h, PFEM2_ESTATE, "a neighbour strip exists but is not connected (pfem2_p2p_connect)");
    return step_host(h, true, rank, fx, fy, wx, wy, dt, substeps, count_out);
}

} // extern "C"
