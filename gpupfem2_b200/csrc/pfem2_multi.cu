// pfem2_multi.cu -- multi-GPU building blocks of the C ABI (strip partition of the cell index range; the reference is
// single-GPU, so nothing here has a reference counterpart).  One handle per GPU; each handle owns the cells [cell_lo, cell_hi)
// and the particles inside them.  Two real exchanges per step: migration of the particles that left the strip (between the move
// pass and the re-sort) and the projection halo (interface-node accumulators).  Three transports: P2P (NVLink peer memory through
// CUDA IPC, default), neighbour (fixed-size buffers over ncclSend / ncclRecv, driven by the Python layer) and exact (host counts).
#include "pfem2_handle.cuh"

#include "pfem2_multi.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

using namespace pfem2;
using namespace pfem2::host;

namespace {

// device + host copy of the strip bounds (cells [bounds[r], bounds[r + 1]) belong to rank r)
int store_rank_bounds(pfem2_handle *h, const int *h_bounds, int n_ranks)
{
    if (h->mg_ranks != n_ranks) {
        cudaFree(h->mg_bounds); cudaFree(h->mg_rank_count);
        h->mg_bounds = h->mg_rank_count = nullptr;
        CU(cudaMalloc((void **)&h->mg_bounds, sizeof(int) * (n_ranks + 1)));
        CU(cudaMalloc((void **)&h->mg_rank_count, sizeof(int) * (n_ranks + 1)));
        h->mg_ranks = n_ranks;
    }
    h->mg_host_bounds.assign(h_bounds, h_bounds + n_ranks + 1);
    // the device copy is in THIS strip's cell numbering (global - cell_base; may be negative / beyond n_cells for far strips)
    std::vector<int> local(h_bounds, h_bounds + n_ranks + 1);
    for (int &b : local) b -= h->cell_base;
    CU(cudaMemcpyAsync(h->mg_bounds, local.data(), sizeof(int) * (n_ranks + 1), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream)); // `local` is a host temporary
    return PFEM2_OK;
}

// shared tail of immigrants_append_device / immigrants_recv_p2p: append one [header | records] block, count it, grow the array
int append_migration_block(pfem2_handle *h, const int4 *buf, int capacity_records, int from_left)
{
    cudaStream_t st = h->stream;
    const int grid = grid_for(capacity_records, kThreads, g_num_sms * 2);
    // lazy re-sort: the rows go behind the dense output of the move pass; they also get their entry in the dense key array of the
    // rank pass, and everybody is summed into stay[] (the fast order keeps no separate arrival counts)
    unsigned *keys = h->lazy_move ? h->keys[1] : nullptr;
    int *arrive = h->lazy_move ? nullptr : h->arrive;
    PFEM2_LAUNCH(k_immigrant_append_dev, grid, kThreads, 0, st, h->soa[h->cur], h->ctr, buf, capacity_records, keys, h->cell_base);
    PFEM2_LAUNCH(k_count_appended_dev, grid, kThreads, 0, st, h->soa[h->cur], h->ctr, buf, capacity_records, h->opt.subcell_mode ? 1 : 0,
                 h->mesh.n_cells, h->ppc, h->level, h->sub_step, h->stay, arrive, h->cell_mask);
    PFEM2_LAUNCH(k_add_count_dev, 1, 1, 0, st, h->ctr, buf, capacity_records, h->cell_mask, h->own_lo, h->own_hi, from_left ? 1 : 0);
    CU(cudaGetLastError());
    return PFEM2_OK;
}

} // namespace

extern "C" {

int pfem2_advect_move(pfem2_handle *h, const double *vx, const double *vy, double dt, int substeps)
{
    return advect_move(h, nodal(vx, vy, nullptr), dt, substeps, 0, true, false);
}

int pfem2_advect_finish(pfem2_handle *h, const double *vx, const double *vy) { return advect_finish(h, nodal(vx, vy, nullptr), 1); }

int pfem2_emigrants_count(pfem2_handle *h, const int *h_bounds, int n_ranks, int *h_counts)
{
    if (!h || !h_bounds || !h_counts || n_ranks < 1 || n_ranks > 64) return PFEM2_EINVAL;
    if (!h->move_pending) return fail(h, PFEM2_ESTATE, "emigrants_count outside advect_move / advect_finish");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    if (h->mg_fused && h->mg_ranks == n_ranks && std::equal(h_bounds, h_bounds + n_ranks + 1, h->mg_host_bounds.begin())) {
        // the move pass counted them (k_advect_locate_tma): rank_count[0..n_ranks) per destination, [n_ranks] = total
        h->mg_host_counts.assign(n_ranks + 1, 0);
        CU(cudaMemcpyAsync(h->mg_host_counts.data(), h->mg_rank_count, sizeof(int) * (n_ranks + 1), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        h->mg_fused_total = h->mg_host_counts[n_ranks];
        for (int r = 0; r < n_ranks; ++r) h_counts[r] = h->mg_host_counts[r];
        return PFEM2_OK;
    }
    if (h->mg_fused) { // different bounds than the move pass used: the statistics stand, the emigrants are searched the old way
        h->mg_fused_total = -1;
    }
    {
        const int rcb = store_rank_bounds(h, h_bounds, n_ranks);
        if (rcb) return rcb;
    }
    CU(cudaMemsetAsync(h->mg_rank_count, 0, sizeof(int) * (n_ranks + 1), st));
    PFEM2_LAUNCH(k_emigrant_count, grid_for(h->capacity), kThreads, 0, st, h->soa[h->cur], h->ctr, h->own_lo, h->own_hi, h->mg_bounds,
                 n_ranks, h->mg_rank_count);
    h->mg_host_counts.assign(n_ranks, 0);
    CU(cudaMemcpyAsync(h->mg_host_counts.data(), h->mg_rank_count, sizeof(int) * n_ranks, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int r = 0; r < n_ranks; ++r) h_counts[r] = h->mg_host_counts[r];
    return PFEM2_OK;
}

int pfem2_emigrants_pack(pfem2_handle *h, void *d_records, long long capacity_records)
{
    if (h) h->partials_valid = false;
    if (!h || !d_records) return PFEM2_EINVAL;
    if (!h->move_pending || h->mg_ranks == 0) return fail(h, PFEM2_ESTATE, "emigrants_pack before emigrants_count");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    std::vector<int> off(h->mg_ranks + 1, 0);
    for (int r = 0; r < h->mg_ranks; ++r) off[r + 1] = off[r] + h->mg_host_counts[r];
    if (off[h->mg_ranks] > capacity_records) return fail(h, PFEM2_ECAPACITY, "emigrant buffer too small");
    CU(cudaMemcpyAsync(h->mg_rank_count, off.data(), sizeof(int) * (h->mg_ranks + 1), cudaMemcpyHostToDevice, st)); // cursors
    if (h->mg_fused && h->mg_fused_total >= 0) {
        if (h->mg_fused_total > 0)
            PFEM2_LAUNCH(k_emigrant_pack_list, grid_for(h->mg_fused_total), kThreads, 0, st, h->soa[h->cur], h->keys[0], h->mg_fused_total,
                         h->mg_bounds, h->mg_ranks, h->mg_rank_count, (int4 *)d_records, h->cell_base);
    } else
    PFEM2_LAUNCH(k_emigrant_pack, grid_for(h->capacity), kThreads, 0, st, h->soa[h->cur], h->ctr, h->own_lo, h->own_hi, h->mg_bounds,
                 h->mg_ranks, h->mg_rank_count, (int4 *)d_records, h->cell_base);
    CU(cudaStreamSynchronize(st)); // `off` is a host temporary
    return PFEM2_OK;
}

int pfem2_immigrants_append(pfem2_handle *h, const void *d_records, int n)
{
    if (h) h->partials_valid = false;
    if (!h || n < 0 || (n > 0 && !d_records)) return PFEM2_EINVAL;
    if (!h->move_pending) return fail(h, PFEM2_ESTATE, "immigrants_append outside advect_move / advect_finish");
    if (n == 0) return PFEM2_OK;
    CU(cudaSetDevice(h->device));
    if ((long long)h->host_count + n > h->capacity) return fail(h, PFEM2_ECAPACITY, "no room for the immigrants");
    PFEM2_LAUNCH(k_immigrant_append, grid_for(n), kThreads, 0, h->stream, h->soa[h->cur], h->ctr, (const int4 *)d_records, n,
                 h->lazy_move ? h->keys[1] : (unsigned *)nullptr, h->cell_base);
    if (h->mg_fused) { // the move pass counted the residents; the immigrants are counted here (no pass over everybody later)
        PFEM2_LAUNCH(k_count_appended, grid_for(n), kThreads, 0, h->stream, h->soa[h->cur], h->ctr, n, h->opt.subcell_mode ? 1 : 0,
                     h->mesh.n_cells, h->ppc, h->level, h->sub_step, h->stay, h->lazy_move ? (int *)nullptr : h->arrive, h->cell_mask);
    }
    PFEM2_LAUNCH(k_add_count, 1, 1, 0, h->stream, h->ctr, n);
    h->host_count += n;
    CU(cudaGetLastError());
    return PFEM2_OK;
}

int pfem2_set_global_cell_offset(pfem2_handle *h, int cell_offset)
{
    if (!h || cell_offset < 0) return PFEM2_EINVAL;
    if (h->seeded || h->mg_ranks) return fail(h, PFEM2_ESTATE, "set_global_cell_offset after seed / set_rank_bounds");
    h->cell_base = cell_offset;
    return PFEM2_OK;
}

int pfem2_set_rank_bounds(pfem2_handle *h, const int *h_bounds, int n_ranks)
{
    if (!h || !h_bounds || n_ranks < 1 || n_ranks > 64) return PFEM2_EINVAL;
    if (h->move_pending) return fail(h, PFEM2_ESTATE, "set_rank_bounds between advect_move and advect_finish");
    for (int r = 0; r < n_ranks; ++r)
        if (h_bounds[r] > h_bounds[r + 1]) return fail(h, PFEM2_EINVAL, "rank bounds must be ascending");
    CU(cudaSetDevice(h->device));
    return store_rank_bounds(h, h_bounds, n_ranks);
}

int pfem2_emigrants_pack_neighbours(pfem2_handle *h, int rank, void *d_left, void *d_right, int capacity_records)
{
    if (h) h->partials_valid = false;
    if (!h || capacity_records < 1 || rank < 0) return PFEM2_EINVAL;
    if (!h->move_pending) return fail(h, PFEM2_ESTATE, "emigrants_pack_neighbours outside advect_move / advect_finish");
    if (!h->mg_fused) // stable order, one-lane-per-record kernels or no rank bounds yet: use emigrants_count / emigrants_pack
        return fail(h, PFEM2_ESTATE, "the move pass did not list its emigrants (call pfem2_set_rank_bounds before pfem2_advect_move; "
                                     "fast order and TMA-tiled kernels only)");
    if (rank >= h->mg_ranks) return fail(h, PFEM2_EINVAL, "rank outside the rank bounds");
    if ((rank > 0 && !d_left) || (rank + 1 < h->mg_ranks && !d_right))
        return fail(h, PFEM2_EINVAL, "a neighbour strip exists but its migration buffer is NULL");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    if (d_left) CU(cudaMemsetAsync(d_left, 0, sizeof(MigrationHeader), st));
    if (d_right) CU(cudaMemsetAsync(d_right, 0, sizeof(MigrationHeader), st));
    // the number of emigrants lives on the device (rank_count[n_ranks]): a fixed grid strides over the list
    PFEM2_LAUNCH(k_emigrant_pack_nbr, grid_for(capacity_records, kThreads, g_num_sms * 2), kThreads, 0, st, h->soa[h->cur], h->keys[0],
                 h->mg_rank_count, h->mg_ranks, h->mg_bounds, rank, (int4 *)d_left, (int4 *)d_right, capacity_records, h->ctr,
                 h->cell_mask, h->own_hi, h->mesh.n_cells, h->cell_base);
    h->mg_fused_total = -1; // consumed
    CU(cudaGetLastError());
    return PFEM2_OK;
}

int pfem2_immigrants_append_device(pfem2_handle *h, const void *d_buffer, int capacity_records, int from_left)
{
    if (h) h->partials_valid = false;
    if (!h || !d_buffer || capacity_records < 1) return PFEM2_EINVAL;
    if (!h->move_pending) return fail(h, PFEM2_ESTATE, "immigrants_append_device outside advect_move / advect_finish");
    if (!h->mg_fused) return fail(h, PFEM2_ESTATE, "immigrants_append_device needs the fused move pass (see pfem2_emigrants_pack_neighbours)");
    CU(cudaSetDevice(h->device));
    return append_migration_block(h, (const int4 *)d_buffer, capacity_records, from_left);
}

// ---- P2P transport (NVLink peer memory through CUDA IPC) ----
// watchdog of the device-side waits: 20 s, PFEM2_P2P_TIMEOUT_S overrides (ranks that reach a step far apart in time)
static unsigned long long p2p_timeout_ns()
{
    static const unsigned long long ns = [] {
        const char *e = getenv("PFEM2_P2P_TIMEOUT_S");
        const double s = e ? atof(e) : 20.0;
        return (unsigned long long)((s > 0.0 ? s : 20.0) * 1e9);
    }();
    return ns;
}
int pfem2_p2p_inbox_create(pfem2_handle *h, int side, int capacity_records, int n_interface_nodes, const int *h_interface_nodes,
                           void *ipc_handle_out)
{
    if (!h || side < 0 || side > 1 || capacity_records < 1 || n_interface_nodes < 0 || (n_interface_nodes && !h_interface_nodes) ||
        !ipc_handle_out)
        return PFEM2_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI passes IPC handles as 64 opaque bytes");
    if (h->p2p.inbox[side]) return fail(h, PFEM2_ESTATE, "inbox already created for this side");
    if (h->p2p.cap && h->p2p.cap != capacity_records) return fail(h, PFEM2_EINVAL, "both inboxes must have the same capacity");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    // a multiple of 2 MiB so that the block is an allocation of its own (an IPC handle names a whole allocation)
    const size_t bytes = (p2p_inbox_bytes(capacity_records, n_interface_nodes) + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    CU(cudaMalloc(&h->p2p.inbox[side], bytes));
    CU(cudaMemsetAsync(h->p2p.inbox[side], 0, bytes, st));
    PFEM2_LAUNCH(k_p2p_init_head, 1, 1, 0, st, (P2PInboxHead *)h->p2p.inbox[side], capacity_records, n_interface_nodes);
    CU(cudaMalloc((void **)&h->p2p.idx[side], sizeof(int) * (size_t)std::max(n_interface_nodes, 1)));
    if (n_interface_nodes)
        CU(cudaMemcpyAsync(h->p2p.idx[side], h_interface_nodes, sizeof(int) * (size_t)n_interface_nodes, cudaMemcpyHostToDevice, st));
    if (!h->p2p.cursors) {
        CU(cudaMalloc((void **)&h->p2p.cursors, 4 * sizeof(int)));
        CU(cudaMemsetAsync(h->p2p.cursors, 0, 4 * sizeof(int), st));
    }
    CU(cudaStreamSynchronize(st)); // the head is initialised before anybody can map the inbox; the host index list may go away
    h->p2p.cap = capacity_records;
    h->p2p.n_idx[side] = n_interface_nodes;
    h->p2p.idx_lo[side] = h->p2p.idx_hi[side] = 0;
    if (n_interface_nodes) {
        h->p2p.idx_lo[side] = *std::min_element(h_interface_nodes, h_interface_nodes + n_interface_nodes);
        h->p2p.idx_hi[side] = *std::max_element(h_interface_nodes, h_interface_nodes + n_interface_nodes) + 1;
    }
    cudaIpcMemHandle_t hd;
    CU(cudaIpcGetMemHandle(&hd, h->p2p.inbox[side]));
    memcpy(ipc_handle_out, &hd, sizeof hd);
    return PFEM2_OK;
}

int pfem2_p2p_connect(pfem2_handle *h, int side, const void *ipc_handle)
{
    if (!h || side < 0 || side > 1 || !ipc_handle) return PFEM2_EINVAL;
    if (!h->p2p.inbox[side]) return fail(h, PFEM2_ESTATE, "create this side's inbox before connecting to the neighbour's");
    if (h->p2p.peer[side]) return fail(h, PFEM2_ESTATE, "already connected on this side");
    CU(cudaSetDevice(h->device));
    cudaIpcMemHandle_t hd;
    memcpy(&hd, ipc_handle, sizeof hd);
    void *peer = nullptr;
    CU(cudaIpcOpenMemHandle(&peer, hd, cudaIpcMemLazyEnablePeerAccess));
    P2PInboxHead head;
    cudaError_t e = cudaMemcpy(&head, peer, sizeof head, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess || head.magic != kP2PMagic || head.capacity_records != h->p2p.cap || head.n_halo_nodes != h->p2p.n_idx[side]) {
        cudaIpcCloseMemHandle(peer);
        cudaGetLastError();
        return fail(h, PFEM2_EINVAL, "the neighbour's inbox does not match (magic / capacity / interface size): cannot use the P2P transport");
    }
    h->p2p.peer[side] = peer;
    return PFEM2_OK;
}

int pfem2_emigrants_send_p2p(pfem2_handle *h, int rank)
{
    if (h) h->partials_valid = false;
    if (!h || rank < 0) return PFEM2_EINVAL;
    if (!h->move_pending) return fail(h, PFEM2_ESTATE, "emigrants_send_p2p outside advect_move / advect_finish");
    if (!h->mg_fused)
        return fail(h, PFEM2_ESTATE, "the move pass did not list its emigrants (call pfem2_set_rank_bounds before pfem2_advect_move; "
                                     "fast order and TMA-tiled kernels only)");
    if (rank >= h->mg_ranks) return fail(h, PFEM2_EINVAL, "rank outside the rank bounds");
    if ((rank > 0 && !h->p2p.peer[0]) || (rank + 1 < h->mg_ranks && !h->p2p.peer[1]))
        return fail(h, PFEM2_ESTATE, "a neighbour strip exists but is not connected (pfem2_p2p_connect)");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int cap = h->p2p.cap;
    const unsigned seq = ++h->p2p.mig_seq;
    const int parity = (int)(seq & 1u);
    unsigned char *pl = (unsigned char *)h->p2p.peer[0], *pr = (unsigned char *)h->p2p.peer[1];
    MigrationHeader *hl = pl ? (MigrationHeader *)(pl + p2p_block_offset(cap, parity)) : nullptr;
    MigrationHeader *hr = pr ? (MigrationHeader *)(pr + p2p_block_offset(cap, parity)) : nullptr;
    // pack + publish in one launch: the block that finishes last writes the headers, fences (system scope) and releases the flags
    PFEM2_LAUNCH(k_p2p_send, grid_for(cap, kThreads, g_num_sms * 2), kThreads, 0, st, h->soa[h->cur], h->keys[0], h->mg_rank_count, h->mg_ranks,
                 h->mg_bounds, rank, hl, pl ? &((P2PInboxHead *)pl)->flag_mig : nullptr, hr, pr ? &((P2PInboxHead *)pr)->flag_mig : nullptr, cap,
                 h->ctr, h->p2p.cursors, h->cell_base, h->cell_mask, h->own_hi, h->mesh.n_cells, seq, (unsigned *)(h->p2p.cursors + 3));
    h->mg_fused_total = -1; // consumed
    CU(cudaGetLastError());
    return PFEM2_OK;
}

int pfem2_immigrants_recv_p2p(pfem2_handle *h)
{
    if (h) h->partials_valid = false;
    if (!h) return PFEM2_EINVAL;
    if (!h->move_pending || !h->mg_fused) return fail(h, PFEM2_ESTATE, "immigrants_recv_p2p outside advect_move / advect_finish");
    if (!h->p2p.mig_seq) return fail(h, PFEM2_ESTATE, "immigrants_recv_p2p before emigrants_send_p2p");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int cap = h->p2p.cap;
    const unsigned seq = h->p2p.mig_seq;
    const int parity = (int)(seq & 1u);
    unsigned char *il = h->p2p.peer[0] ? (unsigned char *)h->p2p.inbox[0] : nullptr; // a neighbour delivers only if it is connected
    unsigned char *ir = h->p2p.peer[1] ? (unsigned char *)h->p2p.inbox[1] : nullptr;
    if (!il && !ir) return PFEM2_OK;
    // wait + append both blocks + their per-cell statistics + the new count in one launch.  Lazy re-sort: the rows go behind the
    // dense output of the move pass, they also get their entry in the dense key array of the rank pass, and everybody is summed
    // into stay[] (the fast order keeps no separate arrival counts)
    PFEM2_LAUNCH(k_p2p_receive, grid_for(cap, kThreads, g_num_sms * 2), kThreads, 0, st, h->soa[h->cur], h->ctr,
                 il ? &((const P2PInboxHead *)il)->flag_mig : nullptr, ir ? &((const P2PInboxHead *)ir)->flag_mig : nullptr, seq, p2p_timeout_ns(),
                 il ? (const int4 *)(il + p2p_block_offset(cap, parity)) : nullptr, ir ? (const int4 *)(ir + p2p_block_offset(cap, parity)) : nullptr,
                 cap, h->lazy_move ? h->keys[1] : (unsigned *)nullptr, h->cell_base, h->opt.subcell_mode ? 1 : 0, h->mesh.n_cells, h->ppc, h->level,
                 h->sub_step, h->stay, h->lazy_move ? (int *)nullptr : h->arrive, h->cell_mask, h->own_lo, h->own_hi,
                 (unsigned *)(h->p2p.cursors + 3));
    CU(cudaGetLastError());
    return PFEM2_OK;
}

// interface sums -> the neighbours' halo blocks + release (one launch) / wait for theirs + add (one launch)
static int halo_send(pfem2_handle *h, double *d_acc3)
{
    cudaStream_t st = h->stream;
    const int cap = h->p2p.cap;
    const unsigned seq = ++h->p2p.halo_seq;
    const int parity = (int)(seq & 1u);
    unsigned char *peer[2] = {(unsigned char *)h->p2p.peer[0], (unsigned char *)h->p2p.peer[1]};
    const int n[2] = {peer[0] ? h->p2p.n_idx[0] : 0, peer[1] ? h->p2p.n_idx[1] : 0};
    const int total = std::max(n[0] + n[1], 1);
    PFEM2_LAUNCH(k_halo_send2, grid_for(total, kThreads, 1 << 30), kThreads, 0, st, d_acc3, h->p2p.idx[0], n[0],
                 peer[0] ? (double *)(peer[0] + p2p_halo_offset(cap, h->p2p.n_idx[0], parity)) : nullptr,
                 peer[0] ? &((P2PInboxHead *)peer[0])->flag_halo : nullptr, h->p2p.idx[1], n[1],
                 peer[1] ? (double *)(peer[1] + p2p_halo_offset(cap, h->p2p.n_idx[1], parity)) : nullptr,
                 peer[1] ? &((P2PInboxHead *)peer[1])->flag_halo : nullptr, seq, (unsigned *)(h->p2p.cursors + 3));
    CU(cudaGetLastError());
    return PFEM2_OK;
}

static int halo_recv(pfem2_handle *h, double *d_acc3)
{
    cudaStream_t st = h->stream;
    const int cap = h->p2p.cap;
    const unsigned seq = h->p2p.halo_seq;
    const int parity = (int)(seq & 1u);
    unsigned char *peer[2] = {(unsigned char *)h->p2p.peer[0], (unsigned char *)h->p2p.peer[1]};
    unsigned char *inbox[2] = {(unsigned char *)h->p2p.inbox[0], (unsigned char *)h->p2p.inbox[1]};
    const int n[2] = {peer[0] ? h->p2p.n_idx[0] : 0, peer[1] ? h->p2p.n_idx[1] : 0};
    const int total = std::max(n[0] + n[1], 1);
    // (two contributions per shared node: a + b == b + a bit for bit)
    PFEM2_LAUNCH(k_halo_recv2, grid_for(total, kThreads, 1 << 30), kThreads, 0, st, d_acc3,
                 peer[0] ? &((const P2PInboxHead *)inbox[0])->flag_halo : nullptr, peer[1] ? &((const P2PInboxHead *)inbox[1])->flag_halo : nullptr,
                 seq, h->ctr, p2p_timeout_ns(), h->p2p.idx[0], n[0],
                 peer[0] ? (const double *)(inbox[0] + p2p_halo_offset(cap, h->p2p.n_idx[0], parity)) : nullptr, h->p2p.idx[1], n[1],
                 peer[1] ? (const double *)(inbox[1] + p2p_halo_offset(cap, h->p2p.n_idx[1], parity)) : nullptr);
    CU(cudaGetLastError());
    return PFEM2_OK;
}

int pfem2_project_halo_p2p(pfem2_handle *h, double *d_acc3)
{
    if (!h || !d_acc3) return PFEM2_EINVAL;
    CU(cudaSetDevice(h->device));
    if (!h->p2p.peer[0] && !h->p2p.peer[1]) return PFEM2_OK;
    const int rc = halo_send(h, d_acc3);
    return rc ? rc : halo_recv(h, d_acc3);
}

// PFEM2_P2P_SPLIT=1 (A/B; default off): hide the exchange behind interior work -- the cells within reach of the strip boundaries
// are moved / reduced first, their emigrants / interface sums are sent, the interior follows while the delivery travels.  Measured
// on channel16m (profiles/r02_summary.md §7): 2.007 against 1.937 ms per step at 8 GPUs, 3.76 against 3.72 at four -- the five extra
// launches and the tails of the small boundary kernels cost more than the exchange latency they hide (the strips run in lockstep,
// so a delivery is rarely late), hence off.
static bool p2p_split()
{
    static const bool on = [] {
        const char *e = getenv("PFEM2_P2P_SPLIT");
        return e && atoi(e) != 0;
    }();
    return on;
}

// advectParticles of one strip, P2P transport, in one call: move pass (lists its emigrants), send, append the immigrants, rank pass
int pfem2_advect_p2p(pfem2_handle *h, int rank, const double *vx, const double *vy, double dt, int substeps)
{
    if (!h) return PFEM2_EINVAL;
    int rc;
    CU(cudaSetDevice(h->device));
    if (p2p_split() && (rc = mesh_band(h))) return rc;
    if ((rc = advect_move(h, nodal(vx, vy, nullptr), dt, substeps, 0, true, p2p_split()))) return rc;
    if ((rc = pfem2_emigrants_send_p2p(h, rank))) return rc;
    if ((rc = advect_move_interior(h))) return rc;
    if ((rc = pfem2_immigrants_recv_p2p(h))) return rc;
    return advect_finish(h, nodal(vx, vy, nullptr), 1);
}

// projectVelocityOntoGrid of one strip, P2P transport, in one call: cell pass, node sums, halo exchange, division
int pfem2_project_p2p(pfem2_handle *h, double *d_acc3, double *d_vx, double *d_vy)
{
    if (!h || !d_acc3 || !d_vx || !d_vy) return PFEM2_EINVAL;
    if (!h->seeded) return fail(h, PFEM2_ESTATE, "project before seed");
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = flush_correct(h))) return rc;
    const bool connected = h->p2p.peer[0] || h->p2p.peer[1];
    if (!p2p_split()) {
        {
            PhaseScope ps(h, PFEM2_PHASE_PROJECT_CELLS);
            launch_project_cells(h);
        }
        PhaseScope ps(h, PFEM2_PHASE_PROJECT_NODES);
        launch_project_nodes_acc_range(h, h->own_node_lo, h->own_node_hi, d_acc3);
        if (connected && (rc = halo_send(h, d_acc3))) return rc;
        if (connected && (rc = halo_recv(h, d_acc3))) return rc;
        launch_project_finalize_range(h, h->own_node_lo, h->own_node_hi, d_acc3, d_vx, d_vy);
        CU(cudaGetLastError());
        return PFEM2_OK;
    }
    // split form: the cells that touch an interface node first, their sums sent, the interior while they travel
    if ((rc = mesh_band(h))) return rc;
    const int bl = std::min(h->own_lo + h->band + 1, h->own_hi), br = std::max(h->own_hi - h->band - 1, bl);
    {
        PhaseScope ps(h, PFEM2_PHASE_PROJECT_CELLS);
        launch_project_cells(h, h->own_lo, bl);
        if (br < h->own_hi) launch_project_cells(h, br, h->own_hi);
    }
    if (connected) {
        PhaseScope ps(h, PFEM2_PHASE_PROJECT_NODES);
        for (int side = 0; side < 2; ++side)
            if (h->p2p.peer[side]) launch_project_nodes_acc_range(h, h->p2p.idx_lo[side], h->p2p.idx_hi[side], d_acc3);
        if ((rc = halo_send(h, d_acc3))) return rc;
    }
    if (br > bl) {
        PhaseScope ps(h, PFEM2_PHASE_PROJECT_CELLS);
        launch_project_cells(h, bl, br);
    }
    PhaseScope ps(h, PFEM2_PHASE_PROJECT_NODES);
    launch_project_nodes_acc_range(h, h->own_node_lo, h->own_node_hi, d_acc3); // (interface nodes again: the same sums)
    if (connected && (rc = halo_recv(h, d_acc3))) return rc;
    launch_project_finalize_range(h, h->own_node_lo, h->own_node_hi, d_acc3, d_vx, d_vy);
    CU(cudaGetLastError());
    return PFEM2_OK;
}

int pfem2_p2p_last_sent(pfem2_handle *h, int *out)
{
    if (!h || !out) return PFEM2_EINVAL;
    *out = 0;
    if (!h->p2p.cursors) return PFEM2_OK;
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(out, h->p2p.cursors + 2, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return PFEM2_OK;
}

int pfem2_project_accumulate(pfem2_handle *h, double *d_acc3)
{
    if (!h || !d_acc3) return PFEM2_EINVAL;
    if (!h->seeded) return fail(h, PFEM2_ESTATE, "project before seed");
    CU(cudaSetDevice(h->device));
    {
        const int rcf = flush_correct(h);
        if (rcf) return rcf;
    }
    cudaStream_t st = h->stream;
    const int N = h->mesh.n_nodes;
    {
        PhaseScope ps(h, PFEM2_PHASE_PROJECT_CELLS);
        launch_project_cells(h);
    }
    PhaseScope ps(h, PFEM2_PHASE_PROJECT_NODES);
    const int nl = h->node_list ? h->n_node_list : N;
    PFEM2_LAUNCH(k_project_nodes_acc, grid_for(nl, kThreads, 1 << 30), kThreads, 0, st, nl, h->node_list, h->node_off,
                 (const int *)h->node_inc, h->partial, d_acc3);
    CU(cudaGetLastError());
    return PFEM2_OK;
}

int pfem2_project_finalize(pfem2_handle *h, const double *d_acc3, double *d_vx, double *d_vy)
{
    if (!h || !d_acc3 || !d_vx || !d_vy) return PFEM2_EINVAL;
    CU(cudaSetDevice(h->device));
    const int N = h->mesh.n_nodes;
    PhaseScope ps(h, PFEM2_PHASE_PROJECT_NODES);
    const int nl = h->node_list ? h->n_node_list : N;
    PFEM2_LAUNCH(k_project_finalize, grid_for(nl, kThreads, 1 << 30), kThreads, 0, h->stream, nl, h->node_list, d_acc3, d_vx, d_vy);
    CU(cudaGetLastError());
    return PFEM2_OK;
}

} // extern "C"
