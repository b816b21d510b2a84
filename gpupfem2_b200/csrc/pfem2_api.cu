// pfem2_api.cu -- host side of libpfem2_b200.so: the handle, memory management and the C ABI
// declared in include/pfem2_b200.h.  Replaces the host methods of the reference's ParticleHandler2D
// (src/particles/particle_handler_2d.cu:238-423); see DESIGN.md for the pipeline.
#include "../../include/pfem2_b200.h"

#include "pfem2_kernels.cuh"
#include "pfem2_lazy.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace pfem2 {
long long g_kernel_launches = 0;
int g_num_sms = 148;
}

using namespace pfem2;

namespace {

thread_local std::string g_create_error;

struct DeviceBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

} // namespace

struct pfem2_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string error;
    pfem2_options opt{};

    // mesh (borrowed) + private repack
    pfem2_mesh_view mesh{};
    CellGeom *geom = nullptr;
    int *node_off = nullptr;      // n_nodes + 1
    unsigned *node_inc = nullptr; // 3 * n_cells, (3c + i) ascending per node
    int level = 1, ppc = 1;
    double sub_step = 1.0;
    int own_lo = 0, own_hi = 0;              // owned cell range [own_lo, own_hi): seeding / re-seeding / emigration (multi-GPU)
    int *own_len_dev = nullptr;              // device int: own_hi - own_lo (scan length)
    int *node_list = nullptr;                // nodes of the owned cells (nullptr = all nodes), multi-GPU
    int n_node_list = 0;
    int own_node_lo = 0, own_node_hi = 0;    // node id range of the owned cells: what a deferred correction can touch
    int v2_node_lo = 0, v2_node_hi = 0;      // node id range of the cells a particle of the owned range can reach in one advect call
    int v2_range_substeps = -1;              // ... computed for this many substeps (-1: not yet)
    int *mg_bounds = nullptr;                // device copy of the rank cell bounds (n_ranks + 1)
    int *mg_rank_count = nullptr;            // device, per destination rank
    int mg_ranks = 0;
    std::vector<int> mg_host_counts;
    std::vector<int> mg_host_bounds;         // host copy of mg_bounds (what a fused move pass used)
    bool mg_fused = false;                   // the move pass in flight listed its emigrants and counted the per-cell statistics
    int mg_fused_total = 0;                  // emigrants listed by that pass
    bool move_pending = false;               // advect_move done, advect_finish outstanding
    double *dv[2] = {nullptr, nullptr};      // deferred velocity correction: nodal increment snapshot (n_nodes each)
    double2 *dv2 = nullptr;                  // the same increment interleaved (x, y) per node, for the TMA-tiled advect pass
    double2 *v2 = nullptr;                   // nodal velocity of the advect in flight, interleaved (packed per call)
    bool dv_pending = false;
    CUtensorMap tmap[2];                     // [rows x 64 B] view of the two record buffers (32-row boxes, 64-byte swizzle)
    void *tmap_base[2] = {nullptr, nullptr}; // what the maps were encoded for
    int tmap_rows[2] = {0, 0};
    double *centers = nullptr; // 3 * ppc
    int key_bits = 1;

    // particles: double-buffered SoA
    int capacity = 0;
    ParticleSoA soa[2]{};
    int cur = 0;
    bool seeded = false;

    // per-step scratch
    Counters *ctr = nullptr;
    Counters *host_ctr = nullptr; // pinned mirror
    cudaEvent_t readback = nullptr;
    bool readback_pending = false;
    int host_count = 0; // last count known on the host
    int host_added = 0;
    bool scatter_attr_set = false;           // scatter_tma: the kernel's dynamic shared-memory limit has been raised on this device
    unsigned *keys[2]{}, *vals[2]{};         // (new cell, array index) of the movers, ping-pong for the radix sort
    unsigned *stay_bits = nullptr;           // capacity / 32 + 2: ballot of particles that stayed in their cell
    int *warp_movers = nullptr;              // capacity / 32 + 2: movers per warp, scanned in place
    int *warp_scan_scratch = nullptr;
    int *stay = nullptr, *arrive = nullptr, *cursor = nullptr; // n_cells + 1 each (one allocation, zeroed together)
    unsigned long long *cell_mask = nullptr; // n_cells + 1
    unsigned long long *packed = nullptr;    // n_cells + 2 (scan in place)
    unsigned long long *scan_scratch64 = nullptr;
    int *cell_start[2] = {nullptr, nullptr}; // n_cells + 1, ping-pong (old / new segment table)
    int cs = 0;
    int *n_cells_dev = nullptr;              // device copy of n_cells (scan length)
    int *rs_hist = nullptr;
    int *rs_scan_scratch = nullptr;
    int *rs_info = nullptr;
    double *partial = nullptr; // 9 * n_cells
    int4 *edge_nbr = nullptr;  // n_cells: cells across the three edges (-1 = boundary)

    // lazily allocated
    void *aos = nullptr;
    size_t aos_bytes = 0;
    double *nodal[4] = {nullptr, nullptr, nullptr, nullptr}; // F.x F.y W.x W.y for pfem2_step_host

    // trailing projection (pfem2_options.fuse_project): the projection's cell pass runs concurrently with the re-sort
    // scatter inside advectParticles and leaves the per-cell sums in `partial`
    int last_substeps = 1;             // substeps of the move pass in flight (reach of a particle = band x substeps)
    int band = -1;                     // max |neighbour - cell| over the one-ring lists (-1: not computed yet)
    bool partials_valid = false;       // `partial` holds the nine sums of the current particle state
    cudaStream_t trail_stream = nullptr;
    cudaEvent_t trail_ev[2] = {nullptr, nullptr};
    int *trail_prog = nullptr;         // per producer block: slabs done; [n] = the consumer's chunk counter
    int trail_prog_n = 0;

    // pfem2_step_host pipeline: the step runs in K chunks of the cell range so that the host <-> device copies of the nodal
    // fields overlap the move pass (upload) and the projection (download)
    struct HostPipe {
        int K = 0, substeps = 0;          // what the plan was made for (0 = none yet)
        std::vector<int> cb, ns;          // cell chunk bounds (K + 1), node slice bounds of the upload (K + 1)
        std::vector<int> up_slice;        // chunk j may start once upload slices 0..up_slice[j] have landed
        std::vector<int> dn_ready;        // after projecting chunk j the nodes [0, dn_ready[j]) are final
        cudaStream_t copy = nullptr;      // non-blocking copy stream
        std::vector<cudaEvent_t> up_ev, dn_ev;
        bool active = false;              // a pipelined step is being issued
        int packed_slices = 0;            // upload slices already interleaved into v2
    } pipe;

    // lazy re-sort (pfem2_options.lazy_sort): the current buffer is dense but in the order of the PREVIOUS step's cells; vals[perm_buf]
    // maps sorted position -> record index (padded to a multiple of 32 with a valid row), keys[1] holds the new cells of the last move pass
    bool permuted = false;
    int perm_buf = 0;
    int *tail_cursor = nullptr;              // device int: re-seeded records appended behind the dense array
    bool lazy_swizzle = true;                // 64-byte swizzle of the lazy move pass's tiles (PFEM2_LAZY_SWIZZLE=0: linear tiles, the fallback)
    bool lazy_nsub3 = true;                  // PFEM2_LAZY_NSUB3=0: runtime-S form of the lazy move pass also for S = 3 (A/B; ptxas allocates the
                                             // S = 3 specialisation without spills, the runtime-S form with 4 / 8 bytes)
    CUtensorMap gmap[2], omap[2];            // lazy move pass: gather maps (box {16, 1}) and tile-store maps (box {16, 32}) of the two buffers
    void *lzmap_base[2] = {nullptr, nullptr};
    int lzmap_rows[2] = {0, 0};

    // P2P transport of the neighbour protocol (multi-GPU): inboxes in this GPU's memory the neighbours store into, and the
    // neighbours' inboxes mapped through CUDA IPC.  side 0 = left neighbour (rank - 1), side 1 = right neighbour (rank + 1)
    struct P2P {
        int cap = 0;                            // records per migration block (the same on every strip)
        void *inbox[2] = {nullptr, nullptr};    // mine: written by neighbour `side`
        void *peer[2] = {nullptr, nullptr};     // theirs: the inbox neighbour `side` keeps for me (IPC mapping)
        int *idx[2] = {nullptr, nullptr};       // interface node ids shared with neighbour `side` (ascending), device
        int n_idx[2] = {0, 0};
        int *cursors = nullptr;                 // device: [0], [1] pack cursors per side, [2] records handed over by the last send
        unsigned mig_seq = 0, halo_seq = 0;     // deliveries made so far (block parity = seq & 1)
    } p2p;

    // optional per-phase CUDA-event timing (pfem2_set_profiling)
    bool profiling = false;
    struct PhaseRec { int phase; cudaEvent_t a, b; };
    std::vector<PhaseRec> phase_recs;
    std::vector<cudaEvent_t> event_pool;
    double phase_ms[PFEM2_NUM_PHASES] = {0};
    long long phase_calls[PFEM2_NUM_PHASES] = {0};
};

namespace {

#define CU(call)                                                                                                    \
    do {                                                                                                            \
        cudaError_t e_ = (call);                                                                                    \
        if (e_ != cudaSuccess) {                                                                                    \
            char buf_[512];                                                                                         \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            if (h) h->error = buf_; else g_create_error = buf_;                                                    \
            return PFEM2_ECUDA;                                                                                     \
        }                                                                                                           \
    } while (0)

int fail(pfem2_handle *h, int code, const char *msg)
{
    if (h) h->error = msg; else g_create_error = msg;
    return code;
}

cudaEvent_t take_event(pfem2_handle *h)
{
    if (!h->event_pool.empty()) {
        cudaEvent_t e = h->event_pool.back();
        h->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// RAII: brackets the launches of one pipeline phase with CUDA events on the handle's stream
struct PhaseScope {
    pfem2_handle *h;
    pfem2_handle::PhaseRec rec;
    PhaseScope(pfem2_handle *h_, int phase) : h(h_)
    {
        if (!h->profiling) return;
        rec.phase = phase;
        rec.a = take_event(h);
        rec.b = take_event(h);
        cudaEventRecord(rec.a, h->stream);
    }
    ~PhaseScope()
    {
        if (!h->profiling) return;
        cudaEventRecord(rec.b, h->stream);
        h->phase_recs.push_back(rec);
    }
};

template <class T> int dev_alloc(pfem2_handle *h, T **p, size_t n)
{
    *p = nullptr;
    CU(cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)));
    return PFEM2_OK;
}

int grid_for(long long n, int threads = kThreads, int max_blocks = 0)
{
    if (max_blocks <= 0) max_blocks = g_num_sms * 16; // persistent grid-stride kernels: a multiple of the SM count
    long long b = (n + threads - 1) / threads;
    return (int)std::max<long long>(1, std::min<long long>(b, max_blocks));
}

int alloc_soa(pfem2_handle *h, ParticleSoA &s, int cap)
{
    ParticleRec *r = nullptr;
    const int rc = dev_alloc(h, &r, cap);
    if (rc) return rc;
    s.bind(r);
    return PFEM2_OK;
}

void free_soa(ParticleSoA &s)
{
    cudaFree(s.records());
    s = ParticleSoA{};
}

void free_particle_scratch(pfem2_handle *h)
{
    for (int k = 0; k < 2; ++k) {
        cudaFree(h->keys[k]); cudaFree(h->vals[k]);
        h->keys[k] = h->vals[k] = nullptr;
    }
    cudaFree(h->rs_hist); cudaFree(h->rs_scan_scratch);
    cudaFree(h->stay_bits); cudaFree(h->warp_movers); cudaFree(h->warp_scan_scratch);
    h->rs_hist = h->rs_scan_scratch = h->warp_movers = h->warp_scan_scratch = nullptr;
    h->stay_bits = nullptr;
}

int alloc_particle_scratch(pfem2_handle *h, int cap)
{
    int rc;
    for (int k = 0; k < 2; ++k) {
        if ((rc = dev_alloc(h, &h->keys[k], (size_t)cap + 32))) return rc; // + one tile: the lazy re-sort pads its index arrays to 32
        if ((rc = dev_alloc(h, &h->vals[k], (size_t)cap + 32))) return rc;
    }
    if ((rc = dev_alloc(h, &h->rs_hist, rs_hist_elems(cap)))) return rc;
    if ((rc = dev_alloc(h, &h->rs_scan_scratch, rs_scan_scratch_elems(cap)))) return rc;
    if ((rc = dev_alloc(h, &h->stay_bits, (size_t)cap / 32 + 2))) return rc;
    if ((rc = dev_alloc(h, &h->warp_movers, (size_t)cap / 32 + 2))) return rc;
    if ((rc = dev_alloc(h, &h->warp_scan_scratch, scan_scratch_elems<int>((long long)cap / 32 + 2)))) return rc;
    return PFEM2_OK;
}

int alloc_particle_storage(pfem2_handle *h, int cap)
{
    int rc;
    for (int k = 0; k < 2; ++k)
        if ((rc = alloc_soa(h, h->soa[k], cap))) return rc;
    // the incidence sort at create() reuses keys/vals, so they hold at least 3 * n_cells entries (ensured by caller)
    if ((rc = alloc_particle_scratch(h, cap))) return rc;
    h->capacity = cap;
    return PFEM2_OK;
}

// grow particle storage to new_cap, preserving the current buffer's live prefix
int grow(pfem2_handle *h, int new_cap)
{
    ParticleSoA old = h->soa[h->cur];
    ParticleSoA other = h->soa[h->cur ^ 1];
    const int n = h->host_count;
    free_soa(other);
    free_particle_scratch(h);
    h->soa[h->cur ^ 1] = ParticleSoA{};
    ParticleSoA fresh{};
    int rc;
    if ((rc = alloc_soa(h, fresh, new_cap))) return rc;
    if (n) CU(cudaMemcpyAsync(fresh.records(), old.records(), (size_t)n * sizeof(ParticleRec), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    free_soa(old);
    h->soa[h->cur] = fresh;
    if ((rc = alloc_soa(h, h->soa[h->cur ^ 1], new_cap))) return rc;
    if ((rc = alloc_particle_scratch(h, new_cap))) return rc;
    h->capacity = new_cap;
    if (h->aos) { cudaFree(h->aos); h->aos = nullptr; h->aos_bytes = 0; }
    return PFEM2_OK;
}

// wait for the counter read-back of the last advect (if any) and refresh the host-side view
int sync_counters(pfem2_handle *h)
{
    if (h->readback_pending) {
        CU(cudaEventSynchronize(h->readback));
        h->readback_pending = false;
        h->host_count = h->host_ctr->count;
        h->host_added = h->host_ctr->added;
        if (h->host_ctr->overflow & kOverflowP2PTimeout)
            return fail(h, PFEM2_ECUDA, "a neighbour strip's P2P delivery did not arrive within the watchdog time; state is invalid");
        if (h->host_ctr->overflow & kOverflowMigration)
            return fail(h, PFEM2_ECAPACITY, "a migration buffer overflowed or a particle left for a non-adjacent strip (raise the migration "
                                            "capacity / use wider strips); state is invalid");
        if (h->host_ctr->overflow) return fail(h, PFEM2_ECAPACITY, "particle capacity exceeded during advect; state is invalid");
    }
    return PFEM2_OK;
}

int queue_readback(pfem2_handle *h)
{
    CU(cudaMemcpyAsync(h->host_ctr, h->ctr, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->readback, h->stream));
    h->readback_pending = true;
    return PFEM2_OK;
}

NodalVel nodal(const double *x, const double *y, double *const *table)
{
    NodalVel v;
    v.x = x;
    v.y = y;
    v.table = table;
    return v;
}

// Re-establish the cell-sorted order in the other buffer: stayers keep their relative order, the movers listed in
// keys[0]/vals[0] (n = ctr->n_movers, array order) are radix-sorted by new cell and appended behind the stayers of
// their cell, lost particles are dropped and (optionally) every empty sub-cell is re-seeded.
// band width of the cell numbering (one-time): a particle's cell index changes by at most this much per substep
int mesh_band(pfem2_handle *h)
{
    if (h->band >= 0) return PFEM2_OK;
    int *dev = nullptr;
    CU(cudaMalloc((void **)&dev, sizeof(int)));
    CU(cudaMemsetAsync(dev, 0, sizeof(int), h->stream));
    const int C = h->mesh.n_cells;
    PFEM2_LAUNCH(k_band_width, grid_for(C, kThreads, 1 << 30), kThreads, 0, h->stream, C, h->mesh.d_nbr_offsets, h->mesh.d_nbr_indices, dev);
    int band = 0;
    CU(cudaMemcpyAsync(&band, dev, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(dev);
    h->band = band;
    return PFEM2_OK;
}

// node id range [lo, hi) of the cells [cell_lo, cell_hi) (one small kernel + an 8-byte read-back; multi-GPU set-up only)
int node_range_of_cells(pfem2_handle *h, int cell_lo, int cell_hi, int &lo, int &hi)
{
    lo = hi = 0;
    if (cell_hi <= cell_lo) return PFEM2_OK;
    int *dev = nullptr;
    const int init[2] = {0x7fffffff, -1};
    int out[2] = {0, 0};
    CU(cudaMalloc((void **)&dev, 2 * sizeof(int)));
    CU(cudaMemcpyAsync(dev, init, sizeof init, cudaMemcpyHostToDevice, h->stream));
    PFEM2_LAUNCH(k_node_minmax, grid_for(cell_hi - cell_lo, kThreads, 1 << 30), kThreads, 0, h->stream, cell_lo, cell_hi, h->geom, dev);
    CU(cudaMemcpyAsync(out, dev, sizeof out, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(dev);
    if (out[1] >= out[0]) {
        lo = out[0];
        hi = out[1] + 1;
    }
    return PFEM2_OK;
}

// Multi-GPU: the nodal arrays the move pass gathers from (interleaved velocity v2) only need the nodes of the cells a particle
// of the owned range can reach in one call: its cell index changes by at most the band width of the one-ring lists per substep.
int ensure_v2_node_range(pfem2_handle *h, int substeps)
{
    const int C = h->mesh.n_cells;
    if (h->own_lo == 0 && h->own_hi == C) {
        h->v2_node_lo = 0;
        h->v2_node_hi = h->mesh.n_nodes;
        return PFEM2_OK;
    }
    if (h->v2_range_substeps == substeps) return PFEM2_OK;
    int rc;
    if ((rc = mesh_band(h))) return rc;
    const long long reach = (long long)h->band * substeps;
    const int c0 = (int)std::max<long long>(0, h->own_lo - reach), c1 = (int)std::min<long long>(C, h->own_hi + reach);
    if ((rc = node_range_of_cells(h, c0, c1, h->v2_node_lo, h->v2_node_hi))) return rc;
    h->v2_range_substeps = substeps;
    return PFEM2_OK;
}

bool trailing_projection_enabled(const pfem2_handle *h, bool reseed, bool stable)
{
    return h->opt.fuse_project == 1 && reseed && !stable && h->opt.lane_per_record == 0 && !h->opt.scatter_tma && h->own_lo == 0 &&
           h->own_hi == h->mesh.n_cells && h->band >= 0;
}

int reorder(pfem2_handle *h, bool reseed, bool have_stayers, bool stable, NodalVel vel)
{
    cudaStream_t st = h->stream;
    const int C = h->mesh.n_cells;
    int flip = 0;
    if (stable) {
        PhaseScope ps(h, PFEM2_PHASE_SORT);
        flip = radix_sort_pairs(h->keys[0], h->vals[0], h->keys[1], h->vals[1], &h->ctr->n_movers, h->capacity, h->key_bits,
                                h->rs_hist, h->rs_scan_scratch, h->rs_info, st);
    }
    PhaseScope ps(h, PFEM2_PHASE_REORDER);
    const int lo = h->own_lo, hi = h->own_hi, own_n = hi - lo; // cell-wise work only over the owned range
    PFEM2_LAUNCH(k_plan_cells, grid_for(own_n, kThreads, 1 << 30), kThreads, 0, st, C, lo, hi, h->ppc, reseed ? 1 : 0, h->stay, h->arrive,
                 h->cell_mask, h->packed, h->ctr);
    exclusive_scan_dev<unsigned long long>(h->packed + lo, h->packed + lo, h->own_len_dev, 1, 0, own_n, h->scan_scratch64, st);
    PFEM2_LAUNCH(k_plan_finish, 1, 1, 0, st, hi, h->packed, h->ctr);
    ParticleSoA src = h->soa[h->cur], dst = h->soa[h->cur ^ 1];
    if (stable) {
        if (have_stayers)
            PFEM2_LAUNCH(k_scatter_stayers, grid_for(h->capacity), kThreads, 0, st, src, dst, &h->ctr->n_old, h->stay_bits,
                         h->cell_start[h->cs], h->packed, h->ctr);
        PFEM2_LAUNCH(k_scatter_movers, grid_for(h->capacity), kThreads, 0, st, src, dst, C, &h->ctr->n_movers, h->keys[flip],
                     h->vals[flip], h->stay, h->packed, h->ctr);
    } else {
        PFEM2_LAUNCH(k_init_cursor, grid_for(own_n, kThreads, 1 << 30), kThreads, 0, st, lo, hi, h->packed, h->cursor);
        if (h->opt.scatter_tma) {
            constexpr int kStages = 3;
            const size_t smem = scatter_smem_bytes<kStages>(kThreads);
            if (!h->scatter_attr_set) { // a function attribute is per device: remembered per handle, not per process
                CU(cudaFuncSetAttribute(k_scatter_all_tma<kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                h->scatter_attr_set = true;
            }
            PFEM2_LAUNCH(k_scatter_all_tma<kStages>, grid_for(h->capacity, kThreads, g_num_sms * 4), kThreads, smem, st, src, dst,
                         &h->ctr->n_old, h->cursor, h->ctr);
        } else {
            if (trailing_projection_enabled(h, reseed, stable)) {
                // scatter (this stream) and re-seed + projection cell pass (side stream) run concurrently; the consumer trails the
                // producer by the reach of a particle (band width x substeps) and so reads the new records from L2
                constexpr int kU = 12, kSlab = (kThreads / 32) * 8 * kU;
                const int producers = std::max(1, std::min(g_num_sms * 2, (int)(((long long)h->capacity + kSlab - 1) / kSlab)));
                if (!h->trail_stream) {
                    CU(cudaStreamCreateWithFlags(&h->trail_stream, cudaStreamNonBlocking));
                    CU(cudaEventCreateWithFlags(&h->trail_ev[0], cudaEventDisableTiming));
                    CU(cudaEventCreateWithFlags(&h->trail_ev[1], cudaEventDisableTiming));
                }
                // [0] completed producer iterations, [1 .. iters] per-iteration block counters, [last] the consumer's chunk counter
                const int iters_max = (int)(((long long)h->capacity + kSlab - 1) / kSlab / producers) + 2;
                if (h->trail_prog_n < iters_max + 2) {
                    cudaFree(h->trail_prog);
                    h->trail_prog = nullptr;
                    CU(cudaMalloc((void **)&h->trail_prog, sizeof(int) * (size_t)(iters_max + 2)));
                    h->trail_prog_n = iters_max + 2;
                }
                CU(cudaMemsetAsync(h->trail_prog, 0, sizeof(int) * (size_t)h->trail_prog_n, st));
                CU(cudaEventRecord(h->trail_ev[0], st));
                CU(cudaStreamWaitEvent(h->trail_stream, h->trail_ev[0], 0));
                PFEM2_LAUNCH((k_scatter_quads_ordered<kU, 2>), producers, kThreads, 0, st, src, dst, &h->ctr->n_old, h->cursor, h->ctr,
                             h->trail_prog);
                const int consumers = g_num_sms; // one 256-thread block per SM next to the two producer blocks
                const int reach = (int)std::min<long long>((long long)h->band * std::max(h->last_substeps, 1), C);
#define PFEM2_TRAIL(G)                                                                                                               \
    PFEM2_LAUNCH((k_reseed_project_trailing<G>), consumers, kThreads, 0, h->trail_stream, C, h->ppc, reach, kSlab, producers,          \
                 (const int *)h->trail_prog, h->trail_prog + h->trail_prog_n - 1, (const int *)h->cell_start[h->cs],                   \
                 (const double2 *)h->mesh.d_vertices, h->geom, h->centers, vel, h->cell_mask, h->stay, h->arrive, h->packed, dst,      \
                 h->cell_start[h->cs ^ 1], h->partial, h->ctr)
                if (h->ppc <= 4) PFEM2_TRAIL(2);
                else if (h->ppc <= 16) PFEM2_TRAIL(4);
                else if (h->ppc <= 36) PFEM2_TRAIL(8);
                else PFEM2_TRAIL(16);
#undef PFEM2_TRAIL
                CU(cudaEventRecord(h->trail_ev[1], h->trail_stream));
                CU(cudaStreamWaitEvent(st, h->trail_ev[1], 0));
                h->cur ^= 1;
                h->cs ^= 1;
                h->partials_valid = true;
                CU(cudaGetLastError());
                return PFEM2_OK;
            }
            if (h->opt.lane_per_record == 0)
                // 12 eight-record groups in flight per warp, 2 blocks per SM (measured on channel16m: U x blocks = 8x3 6.40 ms,
                // 12x2 6.16, 16x2 6.40, 20x2 7.4, 24x1 7.3, 8x4 6.7, 4x6 6.7 for the whole reorder phase)
                PFEM2_LAUNCH((k_scatter_all_quads<12, 2>), grid_for(h->capacity), kThreads, 0, st, src, dst, &h->ctr->n_old, h->cursor, h->ctr);
            else
                PFEM2_LAUNCH(k_scatter_all_regs, grid_for(h->capacity), kThreads, 0, st, src, dst, &h->ctr->n_old, h->cursor, h->ctr);
        }
    }
    PFEM2_LAUNCH(k_reseed, grid_for(own_n + 1, kThreads, 1 << 30), kThreads, 0, st, lo, hi, h->ppc, (const double2 *)h->mesh.d_vertices,
                 h->geom, h->centers, vel, h->cell_mask, h->stay, h->arrive, h->packed, dst, h->cell_start[h->cs ^ 1], h->ctr);
    h->cur ^= 1;
    h->cs ^= 1;
    CU(cudaGetLastError());
    return PFEM2_OK;
}

// Tensor map of a record buffer for the TMA-tiled advect pass.  cuTensorMapEncodeTiled is a driver entry point; it is
// resolved through the runtime so that the library carries no link-time dependency on libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int record_tensor_map(pfem2_handle *h, int k)
{
    void *base = h->soa[k].records();
    if (h->tmap_base[k] == base && h->tmap_rows[k] == h->capacity) return PFEM2_OK;
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(h, PFEM2_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = (EncodeTiledFn)fn;
    }
    const cuuint64_t dims[2] = {16, (cuuint64_t)h->capacity}; // int32 elements per record, records
    const cuuint64_t strides[1] = {sizeof(ParticleRec)};     // bytes between records
    const cuuint32_t box[2] = {16, 32};                      // one warp tile: 32 whole records
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(&h->tmap[k], CU_TENSOR_MAP_DATA_TYPE_INT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return fail(h, PFEM2_ECUDA, buf);
    }
    h->tmap_base[k] = base;
    h->tmap_rows[k] = h->capacity;
    return PFEM2_OK;
}

bool advect_tma_enabled(const pfem2_handle *h) { return h->opt.lane_per_record == 0; }

// ---- lazy re-sort (pfem2_options.lazy_sort, EXPERIMENTAL; kernels in pfem2_lazy.cuh) ----
bool lazy_enabled(const pfem2_handle *h)
{
    return h->opt.lazy_sort != 0 && h->own_lo == 0 && h->own_hi == h->mesh.n_cells; // (incompatible options are refused at create)
}

// maps of a record buffer for the lazy move pass: the same [rows x 64 B] tensor as record_tensor_map, once with box {16, 1} (tile::gather4
// takes four row indices) and once with box {16, 32} (the dense tile store), both with the handle's swizzle mode
int lazy_record_maps(pfem2_handle *h, int k)
{
    void *base = h->soa[k].records();
    if (h->lzmap_base[k] == base && h->lzmap_rows[k] == h->capacity) return PFEM2_OK;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(h, PFEM2_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {16, (cuuint64_t)h->capacity};
    const cuuint64_t strides[1] = {sizeof(ParticleRec)};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = h->lazy_swizzle ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
    for (int which = 0; which < 2; ++which) {
        const cuuint32_t box[2] = {16, which ? 32u : 1u};
        const CUresult r = ((EncodeTiledFn)fn)(which ? &h->omap[k] : &h->gmap[k], CU_TENSOR_MAP_DATA_TYPE_INT32, 2, base, dims, strides, box, estr,
                                                CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(h, PFEM2_ECUDA, which ? "cuTensorMapEncodeTiled (lazy store map) failed" : "cuTensorMapEncodeTiled (gather map) failed");
    }
    h->lzmap_base[k] = base;
    h->lzmap_rows[k] = h->capacity;
    return PFEM2_OK;
}

// back to the ordinary state: the sorted order is made physical in the other buffer.  Every reader of the physical order calls this.
int materialize(pfem2_handle *h)
{
    if (!h->permuted) return PFEM2_OK;
    PFEM2_LAUNCH(k_materialize, grid_for(h->capacity), kThreads, 0, h->stream, h->soa[h->cur], h->soa[h->cur ^ 1],
                 (const unsigned *)h->vals[h->perm_buf], h->ctr);
    h->cur ^= 1;
    h->permuted = false;
    CU(cudaGetLastError());
    return PFEM2_OK;
}

#ifdef PFEM2_NO_FAST // A/B build (make variants): the move pass of rounds 1a-1d, start cell carried through the substep loop
constexpr bool kNoFastForm = true;
#else
constexpr bool kNoFastForm = false;
#endif

template <int MODE, bool WALK, bool MASK64>
void launch_advect(pfem2_handle *h, NodalVel vel, double hsub, int substeps, int do_count)
{
    const int C = h->mesh.n_cells;
    if (advect_tma_enabled(h)) { // default: particle tiles moved by the copy engine, nodal velocity interleaved
        unsigned *sb = h->opt.stable_order ? h->stay_bits : nullptr;
        const int N = h->mesh.n_nodes;
        const size_t smem = advect_tma_smem_bytes(kAdvThreads);
        const int *cstart = nullptr;
        int c_lo = 0, c_hi = 0;
#define PFEM2_ADV_TMA(NSUB)                                                                                                          \
    do {                                                                                                                             \
        if (sb || !WALK || kNoFastForm) PFEM2_ADV_TMA_(NSUB, false); /* FAST is a register-allocation matter: taken where ptxas spills less */ \
        else PFEM2_ADV_TMA_(NSUB, true);                                                                                             \
    } while (0)
#define PFEM2_ADV_TMA_(NSUB, FAST)                                                                                                   \
    PFEM2_LAUNCH((k_advect_locate_tma<MODE, WALK, MASK64, NSUB, FAST>), grid, kAdvThreads, smem, h->stream, h->tmap[h->cur], h->geom, h->edge_nbr,   \
                 h->mesh.d_nbr_offsets, h->mesh.d_nbr_indices, h->v2, hsub, substeps, C, h->ppc, h->level, h->sub_step, h->ctr, sb,         \
                 h->warp_movers, h->stay, h->opt.stable_order ? h->arrive : (int *)nullptr, h->cell_mask, do_count, h->dv_pending ? h->dv2 : nullptr, h->own_lo, h->own_hi,          \
                 h->mg_bounds, h->mg_ranks, h->mg_rank_count, h->mg_fused ? h->keys[0] : (unsigned *)nullptr, cstart, c_lo, c_hi)
        if (!h->pipe.active) {
            (void)N;
            const int n0 = h->v2_node_lo, n1 = h->v2_node_hi; // all nodes on a single GPU; a strip's reach otherwise
            if (n1 > n0) PFEM2_LAUNCH(k_pack_nodal, grid_for(n1 - n0, kThreads, 1 << 30), kThreads, 0, h->stream, n0, n1, vel, h->v2);
            const int grid = grid_for(h->capacity, kAdvThreads, g_num_sms * kAdvBlocksPerSM); // persistent: all resident blocks
            if (substeps == 3)
                PFEM2_ADV_TMA(3);
            else
                PFEM2_ADV_TMA(0);
        } else {
            // pfem2_step_host: chunk j of the cell range starts as soon as the slices of the nodal field it can touch have
            // landed (events recorded on the copy stream) and have been interleaved into v2
            pfem2_handle::HostPipe &pp = h->pipe;
            cstart = h->cell_start[h->cs];
            const int grid = grid_for((long long)h->capacity / pp.K + 1, kAdvThreads, g_num_sms * kAdvBlocksPerSM);
            for (int j = 0; j < pp.K; ++j) {
                for (; pp.packed_slices <= pp.up_slice[j]; ++pp.packed_slices) {
                    const int s0 = pp.ns[pp.packed_slices], s1 = pp.ns[pp.packed_slices + 1];
                    cudaStreamWaitEvent(h->stream, pp.up_ev[pp.packed_slices], 0);
                    if (s1 > s0) PFEM2_LAUNCH(k_pack_nodal, grid_for(s1 - s0, kThreads, 1 << 30), kThreads, 0, h->stream, s0, s1, vel, h->v2);
                }
                c_lo = pp.cb[j];
                c_hi = pp.cb[j + 1];
                if (substeps == 3)
                    PFEM2_ADV_TMA(3);
                else
                    PFEM2_ADV_TMA(0);
            }
            for (; pp.packed_slices < pp.K; ++pp.packed_slices) // (not reached: the last chunk needs every slice)
                cudaStreamWaitEvent(h->stream, pp.up_ev[pp.packed_slices], 0);
        }
#undef PFEM2_ADV_TMA
#undef PFEM2_ADV_TMA_
        return;
    }
    unsigned *sbits = h->opt.stable_order ? h->stay_bits : nullptr; // only the stable-order path consumes the ballots
#define PFEM2_ADV_LAUNCH(NSUB)                                                                                                     \
    PFEM2_LAUNCH((k_advect_locate<MODE, WALK, MASK64, NSUB>), grid_for(h->capacity), kThreads, 0, h->stream, h->soa[h->cur], h->geom,   \
                 h->edge_nbr, h->mesh.d_nbr_offsets, h->mesh.d_nbr_indices, vel, hsub, substeps, C, h->ppc, h->level, h->sub_step,   \
                 h->ctr, sbits, h->warp_movers, h->stay, h->opt.stable_order ? h->arrive : (int *)nullptr, h->cell_mask, do_count, h->dv_pending ? h->dv[0] : nullptr,            \
                 h->dv_pending ? h->dv[1] : nullptr)
    if (substeps == 3)
        PFEM2_ADV_LAUNCH(3);
    else
        PFEM2_ADV_LAUNCH(0);
#undef PFEM2_ADV_LAUNCH
}

// first half of advectParticles: S x (advect + locate); with do_count the per-cell statistics are fused in
int advect_move(pfem2_handle *h, NodalVel vel, double dt, int substeps, int do_count, bool mg_move = false)
{
    if (!h) return PFEM2_EINVAL;
    if (!h->seeded) return fail(h, PFEM2_ESTATE, "advect before seed");
    if (h->move_pending) return fail(h, PFEM2_ESTATE, "advect_move called twice without advect_finish");
    if (substeps < 1) return fail(h, PFEM2_EINVAL, "particleSubsteps must be >= 1");
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = sync_counters(h))) return rc;
    if ((rc = materialize(h))) return rc;
    // capacity policy: keep room for the growth seen so far (re-seeding only ever adds)
    {
        const long long margin = std::max<long long>({(long long)h->host_count / 16, 4ll * h->host_added, 4096ll});
        if ((long long)h->host_count + margin > h->capacity) {
            const long long want = std::max<long long>((long long)(1.25 * h->host_count), (long long)h->host_count + 2 * margin);
            if (want > 2147483647ll - 1024) return fail(h, PFEM2_ECAPACITY, "particle count exceeds 32-bit indexing");
            if ((rc = grow(h, (int)want))) return rc;
        }
    }
    cudaStream_t st = h->stream;
    const int C = h->mesh.n_cells;
    h->partials_valid = false;
    h->last_substeps = substeps;
    if (h->opt.fuse_project == 1 && h->band < 0 && !mg_move && (rc = mesh_band(h))) return rc;
    const double hsub = dt / substeps; // particle_handler_2d.cu:330, host double
    {   // per-cell scratch of the owned range (+ a few cells for the tolerance-band spill of the occupancy bits)
        const size_t lo = (size_t)h->own_lo, len = (size_t)std::min(C, h->own_hi + 4) - lo + 1;
        CU(cudaMemsetAsync(h->stay + lo, 0, sizeof(int) * len, st));
        CU(cudaMemsetAsync(h->arrive + lo, 0, sizeof(int) * len, st));
        CU(cudaMemsetAsync(h->cursor + lo, 0, sizeof(int) * len, st));
        CU(cudaMemsetAsync(h->cell_mask + lo, 0, sizeof(unsigned long long) * len, st));
    }
    PFEM2_LAUNCH(k_begin_advect, 1, 1, 0, st, h->ctr, h->capacity);
    // multi-GPU, from the second step on (the rank bounds arrive with the first emigrants_count call): the move pass lists the
    // emigrants and counts the per-cell statistics of everybody else itself
    h->mg_fused = mg_move && advect_tma_enabled(h) && !h->opt.stable_order && h->mg_ranks > 0 &&
                  !(getenv("PFEM2_MG_FUSED") && atoi(getenv("PFEM2_MG_FUSED")) == 0); // env: A/B measurements only
    if (h->mg_fused) {
        do_count = 1;
        CU(cudaMemsetAsync(h->mg_rank_count, 0, sizeof(int) * (h->mg_ranks + 1), st));
    }
    if (advect_tma_enabled(h)) {
        if ((rc = record_tensor_map(h, h->cur))) return rc;
        if (!h->v2) CU(cudaMalloc((void **)&h->v2, sizeof(double2) * (size_t)h->mesh.n_nodes));
        if ((rc = ensure_v2_node_range(h, substeps))) return rc;
    }
    {
        PhaseScope ps(h, PFEM2_PHASE_ADVECT);
        const bool m64 = h->ppc > 32, walk = h->opt.exact_search == 0;
        const int mode = h->opt.subcell_mode ? 1 : 0;
#define PFEM2_ADV(M, W, B) launch_advect<M, W, B>(h, vel, hsub, substeps, do_count)
        if (mode == 0) {
            if (walk) { if (m64) PFEM2_ADV(0, true, true); else PFEM2_ADV(0, true, false); }
            else      { if (m64) PFEM2_ADV(0, false, true); else PFEM2_ADV(0, false, false); }
        } else {
            if (walk) { if (m64) PFEM2_ADV(1, true, true); else PFEM2_ADV(1, true, false); }
            else      { if (m64) PFEM2_ADV(1, false, true); else PFEM2_ADV(1, false, false); }
        }
#undef PFEM2_ADV
    }
    CU(cudaGetLastError());
    h->dv_pending = false; // the move pass applied the deferred correction
    h->move_pending = true;
    return PFEM2_OK;
}

// second half: (statistics, if not fused) + re-sort by cell + distribution check / re-seed
int advect_finish(pfem2_handle *h, NodalVel vel, int need_count)
{
    if (!h) return PFEM2_EINVAL;
    if (!h->move_pending) return fail(h, PFEM2_ESTATE, "advect_finish without advect_move");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    int rc;
    bool stable = h->opt.stable_order != 0;
    if (need_count && h->mg_fused) {
        need_count = 0; // the move pass and immigrants_append already counted everybody who is still here
    }
    h->mg_fused = false;
    if (need_count) {
        // multi-GPU: particles came and went since the move pass; count everybody (all "arrived": no stayer shortcut)
        PhaseScope ps(h, PFEM2_PHASE_REORDER);
        const int C = h->mesh.n_cells;
        const bool m64 = h->ppc > 32;
        const int mode = h->opt.subcell_mode ? 1 : 0;
#define PFEM2_CNT(M, B)                                                                                                              \
    PFEM2_LAUNCH((k_count_all<M, B>), grid_for(h->capacity), kThreads, 0, st, h->soa[h->cur], h->ctr, C, h->ppc, h->level, h->sub_step, \
                 h->stay, h->arrive, h->cell_mask)
        if (mode == 0) { if (m64) PFEM2_CNT(0, true); else PFEM2_CNT(0, false); }
        else           { if (m64) PFEM2_CNT(1, true); else PFEM2_CNT(1, false); }
#undef PFEM2_CNT
        if (stable) // everybody is a mover: (cell, index) pairs of the whole array, then the stable radix sort
            PFEM2_LAUNCH(k_all_movers, grid_for(h->capacity), kThreads, 0, st, h->soa[h->cur], C, h->ctr, h->keys[0], h->vals[0],
                         (int *)nullptr /* already counted by k_count_all */, &h->ctr->n_movers);
    } else if (stable) {
        PhaseScope ps(h, PFEM2_PHASE_SORT);
        exclusive_scan_dev<int>(h->warp_movers, h->warp_movers, &h->ctr->n_warps, 1, 0, (long long)h->capacity / 32 + 1,
                                h->warp_scan_scratch, st);
        PFEM2_LAUNCH(k_emit_movers, grid_for(h->capacity), kThreads, 0, st, h->soa[h->cur], h->ctr, h->stay_bits, h->warp_movers,
                     h->keys[0], h->vals[0]);
    }
    h->move_pending = false;
    if ((rc = reorder(h, true, !need_count, stable, vel))) return rc;
    if ((rc = queue_readback(h))) return rc;
    if (h->opt.verbose) {
        if ((rc = sync_counters(h))) return rc;
        printf("Particle handler contains %d particles\n", h->host_count); // particle_handler_2d.cu:341
    }
    return PFEM2_OK;
}

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
    // pageable host source: the copy is staged before the call returns, so the vector may change afterwards
    CU(cudaMemcpyAsync(h->mg_bounds, h->mg_host_bounds.data(), sizeof(int) * (n_ranks + 1), cudaMemcpyHostToDevice, h->stream));
    return PFEM2_OK;
}

// shared tail of immigrants_append_device / immigrants_recv_p2p: append one [header | records] block, count it, grow the array
int append_migration_block(pfem2_handle *h, const int4 *buf, int capacity_records, int from_left)
{
    cudaStream_t st = h->stream;
    const int grid = grid_for(capacity_records, kThreads, g_num_sms * 2);
    PFEM2_LAUNCH(k_immigrant_append_dev, grid, kThreads, 0, st, h->soa[h->cur], h->ctr, buf, capacity_records);
    const int C = h->mesh.n_cells;
    const bool m64 = h->ppc > 32;
#define PFEM2_CNTD(M, B)                                                                                                             \
    PFEM2_LAUNCH((k_count_appended_dev<M, B>), grid, kThreads, 0, st, h->soa[h->cur], h->ctr, buf, capacity_records, C, h->ppc, h->level, \
                 h->sub_step, h->stay, h->arrive, h->cell_mask)
    if (h->opt.subcell_mode == 0) { if (m64) PFEM2_CNTD(0, true); else PFEM2_CNTD(0, false); }
    else                          { if (m64) PFEM2_CNTD(1, true); else PFEM2_CNTD(1, false); }
#undef PFEM2_CNTD
    PFEM2_LAUNCH(k_add_count_dev, 1, 1, 0, st, h->ctr, buf, capacity_records, h->cell_mask, h->own_lo, h->own_hi, from_left ? 1 : 0);
    CU(cudaGetLastError());
    return PFEM2_OK;
}

// advectParticles in the lazy re-sort: gathered move pass -> plan -> rank pass -> appended re-seeds (pfem2_lazy.cuh)
int advect_lazy(pfem2_handle *h, NodalVel vel, double dt, int substeps)
{
    if (!h->seeded) return fail(h, PFEM2_ESTATE, "advect before seed");
    if (h->move_pending) return fail(h, PFEM2_ESTATE, "advect while a multi-GPU move is pending");
    if (substeps < 1) return fail(h, PFEM2_EINVAL, "particleSubsteps must be >= 1");
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = sync_counters(h))) return rc;
    {   // the dense array holds the lost particles of the pass as well and the re-seeds are appended behind it: keep twice the margin
        const long long margin = 2 * std::max<long long>({(long long)h->host_count / 16, 4ll * h->host_added, 4096ll});
        if ((long long)h->host_count + margin > h->capacity) {
            const long long want = std::max<long long>((long long)(1.25 * h->host_count), (long long)h->host_count + 2 * margin);
            if (want > 2147483647ll - 1024) return fail(h, PFEM2_ECAPACITY, "particle count exceeds 32-bit indexing");
            if ((rc = materialize(h))) return rc;
            if ((rc = grow(h, (int)want))) return rc;
        }
    }
    cudaStream_t st = h->stream;
    const int C = h->mesh.n_cells, N = h->mesh.n_nodes;
    h->partials_valid = false;
    h->last_substeps = substeps;
    const double hsub = dt / substeps;
    {
        const size_t len = (size_t)C + 1;
        CU(cudaMemsetAsync(h->stay, 0, sizeof(int) * len, st));
        CU(cudaMemsetAsync(h->arrive, 0, sizeof(int) * len, st));
        CU(cudaMemsetAsync(h->cursor, 0, sizeof(int) * len, st));
        CU(cudaMemsetAsync(h->cell_mask, 0, sizeof(unsigned long long) * len, st));
    }
    PFEM2_LAUNCH(k_begin_advect, 1, 1, 0, st, h->ctr, h->capacity);
    if (!h->tail_cursor) CU(cudaMalloc((void **)&h->tail_cursor, sizeof(int)));
    CU(cudaMemsetAsync(h->tail_cursor, 0, sizeof(int), st));
    if (!h->permuted) { // physically sorted (seed, upload, materialize): the identity permutation
        const int padded = (h->host_count + 31) & ~31;
        PFEM2_LAUNCH(k_iota, grid_for(padded), kThreads, 0, st, h->vals[h->perm_buf], h->ctr, padded);
    }
    if (!h->v2) CU(cudaMalloc((void **)&h->v2, sizeof(double2) * (size_t)N));
    if ((rc = lazy_record_maps(h, h->cur))) return rc;
    if ((rc = lazy_record_maps(h, h->cur ^ 1))) return rc;
    {
        PhaseScope ps(h, PFEM2_PHASE_ADVECT);
        const size_t smem = advect_tma_smem_bytes(kAdvThreads);
        const bool m64 = h->ppc > 32, walk = h->opt.exact_search == 0;
        const int mode = h->opt.subcell_mode ? 1 : 0;
        const int *cstart = nullptr;
        int c_lo = 0, c_hi = C, grid = 1;
#define PFEM2_LAZY_ADV(M, W, B, NSUB, SWZ)                                                                                                    \
    PFEM2_LAUNCH((k_advect_locate_lazy<M, W, B, NSUB, SWZ>), grid, kAdvThreads, smem, st, h->gmap[h->cur], h->omap[h->cur ^ 1],                 \
                 (const int4 *)h->vals[h->perm_buf], h->keys[1], h->geom, h->edge_nbr, h->mesh.d_nbr_offsets, h->mesh.d_nbr_indices, h->v2, hsub, \
                 substeps, C, h->ppc, h->level, h->sub_step, h->ctr, h->stay, h->cell_mask, h->dv_pending ? h->dv2 : (const double2 *)nullptr,  \
                 cstart, c_lo, c_hi)
#define PFEM2_LAZY_ADV_N(M, W, B)                                                                                                             \
    do {                                                                                                                                      \
        if (!h->lazy_swizzle) PFEM2_LAZY_ADV(M, W, B, 0, false);                                                                              \
        else if (substeps == 3 && h->lazy_nsub3) PFEM2_LAZY_ADV(M, W, B, 3, true);                                                            \
        else PFEM2_LAZY_ADV(M, W, B, 0, true);                                                                                                \
    } while (0)
        auto launch = [&]() {
            if (mode == 0) {
                if (walk) { if (m64) PFEM2_LAZY_ADV_N(0, true, true); else PFEM2_LAZY_ADV_N(0, true, false); }
                else      { if (m64) PFEM2_LAZY_ADV_N(0, false, true); else PFEM2_LAZY_ADV_N(0, false, false); }
            } else {
                if (walk) { if (m64) PFEM2_LAZY_ADV_N(1, true, true); else PFEM2_LAZY_ADV_N(1, true, false); }
                else      { if (m64) PFEM2_LAZY_ADV_N(1, false, true); else PFEM2_LAZY_ADV_N(1, false, false); }
            }
        };
        if (!h->pipe.active) {
            PFEM2_LAUNCH(k_pack_nodal, grid_for(N, kThreads, 1 << 30), kThreads, 0, st, 0, N, vel, h->v2);
            grid = grid_for(h->capacity, kAdvThreads, g_num_sms * kAdvBlocksPerSM);
            launch();
        } else {
            // pfem2_step_host: chunk j of the cell range starts as soon as the slices of the nodal field it can touch have landed (events
            // recorded on the copy stream) and have been interleaved into v2 -- the schedule of launch_advect, over whole tiles
            pfem2_handle::HostPipe &pp = h->pipe;
            cstart = h->cell_start[h->cs];
            grid = grid_for((long long)h->capacity / pp.K + 1, kAdvThreads, g_num_sms * kAdvBlocksPerSM);
            for (int j = 0; j < pp.K; ++j) {
                for (; pp.packed_slices <= pp.up_slice[j]; ++pp.packed_slices) {
                    const int s0 = pp.ns[pp.packed_slices], s1 = pp.ns[pp.packed_slices + 1];
                    cudaStreamWaitEvent(st, pp.up_ev[pp.packed_slices], 0);
                    if (s1 > s0) PFEM2_LAUNCH(k_pack_nodal, grid_for(s1 - s0, kThreads, 1 << 30), kThreads, 0, st, s0, s1, vel, h->v2);
                }
                c_lo = pp.cb[j];
                c_hi = pp.cb[j + 1];
                launch();
            }
        }
#undef PFEM2_LAZY_ADV_N
#undef PFEM2_LAZY_ADV
    }
    CU(cudaGetLastError());
    h->dv_pending = false; // the move pass applied the deferred correction
    h->cur ^= 1;           // the dense output is the current buffer now (order of the previous step's cells, lost particles included)
    {
        PhaseScope ps(h, PFEM2_PHASE_REORDER);
        PFEM2_LAUNCH(k_plan_cells, grid_for(C, kThreads, 1 << 30), kThreads, 0, st, C, 0, C, h->ppc, 1, h->stay, h->arrive, h->cell_mask, h->packed,
                     h->ctr);
        exclusive_scan_dev<unsigned long long>(h->packed, h->packed, h->own_len_dev, 1, 0, C, h->scan_scratch64, st);
        PFEM2_LAUNCH(k_plan_finish, 1, 1, 0, st, C, h->packed, h->ctr);
        PFEM2_LAUNCH(k_init_cursor, grid_for(C, kThreads, 1 << 30), kThreads, 0, st, 0, C, h->packed, h->cursor);
        unsigned *src_new = h->vals[h->perm_buf ^ 1];
        PFEM2_LAUNCH(k_rank, grid_for(h->capacity), kThreads, 0, st, (const unsigned *)h->keys[1], (const int *)&h->ctr->n_old, h->cursor, src_new,
                     h->ctr);
        PFEM2_LAUNCH(k_reseed_lazy, grid_for(C + 1, kThreads, 1 << 30), kThreads, 0, st, 0, C, h->ppc, (const double2 *)h->mesh.d_vertices, h->geom,
                     h->centers, vel, h->cell_mask, h->stay, h->packed, h->soa[h->cur], (const int *)&h->ctr->n_old, h->tail_cursor, src_new,
                     h->cell_start[h->cs ^ 1], h->ctr);
        h->cs ^= 1;
        h->perm_buf ^= 1;
        h->permuted = true;
    }
    CU(cudaGetLastError());
    if ((rc = queue_readback(h))) return rc;
    if (h->opt.verbose) {
        if ((rc = sync_counters(h))) return rc;
        printf("Particle handler contains %d particles\n", h->host_count); // particle_handler_2d.cu:341
    }
    return PFEM2_OK;
}

int do_advect(pfem2_handle *h, NodalVel vel, double dt, int substeps)
{
    int rc;
    if (h && lazy_enabled(h)) return advect_lazy(h, vel, dt, substeps);
    if ((rc = advect_move(h, vel, dt, substeps, 1))) return rc;
    return advect_finish(h, vel, 0);
}

int flush_correct(pfem2_handle *h);

void launch_project_cells(pfem2_handle *h, const ParticleSoA &p, int c_lo = -1, int c_hi = -1)
{
    cudaStream_t st = h->stream;
    const int ppc = h->ppc;
    if (c_lo < 0) { // the owned range
        c_lo = h->own_lo;
        c_hi = h->own_hi;
    }
    const long long nc = c_hi - c_lo;
    if (h->permuted) { // lazy re-sort: the segment [cell_start[c], cell_start[c + 1]) names its records through the permutation
        const unsigned *src = h->vals[h->perm_buf];
        if (ppc <= 4)
            PFEM2_LAUNCH(k_project_cells_lazy<2>, grid_for(nc * 2), kThreads, 0, st, c_lo, c_hi, p, src, h->cell_start[h->cs], h->partial);
        else if (ppc <= 16)
            PFEM2_LAUNCH(k_project_cells_lazy<4>, grid_for(nc * 4), kThreads, 0, st, c_lo, c_hi, p, src, h->cell_start[h->cs], h->partial);
        else if (ppc <= 36)
            PFEM2_LAUNCH(k_project_cells_lazy<8>, grid_for(nc * 8), kThreads, 0, st, c_lo, c_hi, p, src, h->cell_start[h->cs], h->partial);
        else
            PFEM2_LAUNCH(k_project_cells_lazy<16>, grid_for(nc * 16), kThreads, 0, st, c_lo, c_hi, p, src, h->cell_start[h->cs], h->partial);
        return;
    }
    // lanes per cell: about a quarter of the nominal segment length, so each lane keeps several loads in flight
    if (ppc <= 4)
        PFEM2_LAUNCH(k_project_cells<2>, grid_for(nc * 2), kThreads, 0, st, c_lo, c_hi, p, h->cell_start[h->cs], h->partial);
    else if (ppc <= 16)
        PFEM2_LAUNCH(k_project_cells<4>, grid_for(nc * 4), kThreads, 0, st, c_lo, c_hi, p, h->cell_start[h->cs], h->partial);
    else if (ppc <= 36)
        PFEM2_LAUNCH(k_project_cells<8>, grid_for(nc * 8), kThreads, 0, st, c_lo, c_hi, p, h->cell_start[h->cs], h->partial);
    else
        PFEM2_LAUNCH(k_project_cells<16>, grid_for(nc * 16), kThreads, 0, st, c_lo, c_hi, p, h->cell_start[h->cs], h->partial);
}

int do_project(pfem2_handle *h, double *vx, double *vy, double *const *table, double *cx = nullptr, double *cy = nullptr,
               double *const *table_copy = nullptr)
{
    if (!h) return PFEM2_EINVAL;
    if (!h->seeded) return fail(h, PFEM2_ESTATE, "project before seed");
    CU(cudaSetDevice(h->device));
    {
        const int rcf = flush_correct(h);
        if (rcf) return rcf;
    }
    cudaStream_t st = h->stream;
    const int C = h->mesh.n_cells, N = h->mesh.n_nodes;
    ParticleSoA p = h->soa[h->cur];
    // lanes per cell: enough to cover the typical segment in one or two strides
    const int ppc = h->ppc;
    if (!h->partials_valid) { // (the trailing projection of the last advect already left the per-cell sums in `partial`)
        PhaseScope ps(h, PFEM2_PHASE_PROJECT_CELLS);
        launch_project_cells(h, p);
        h->partials_valid = true;
    }
    PhaseScope ps(h, PFEM2_PHASE_PROJECT_NODES);
    PFEM2_LAUNCH(k_project_nodes, grid_for(N, kThreads, 1 << 30), kThreads, 0, st, 0, N, h->node_off, (const int *)h->node_inc, h->partial,
                 vx, vy, table, cx, cy, table_copy);
    CU(cudaGetLastError());
    return PFEM2_OK;
}

int apply_correct_now(pfem2_handle *h, NodalVel v, NodalVel vold, bool has_old)
{
    if (h) h->partials_valid = false; // particle velocities change
    {
        const int rcm = materialize(h); // the eager kernel walks the physical order
        if (rcm) return rcm;
    }
    ParticleSoA p = h->soa[h->cur];
    const int grid = grid_for(h->capacity);
    PhaseScope ps(h, PFEM2_PHASE_CORRECT);
    if (has_old)
        PFEM2_LAUNCH(k_correct<true>, grid, kThreads, 0, h->stream, p, h->geom, v, vold, h->ctr);
    else
        PFEM2_LAUNCH(k_correct<false>, grid, kThreads, 0, h->stream, p, h->geom, v, vold, h->ctr);
    CU(cudaGetLastError());
    return PFEM2_OK;
}

// apply a deferred correction now (before anything reads particle velocities other than the next advect)
int flush_correct(pfem2_handle *h)
{
    if (!h->dv_pending) return PFEM2_OK;
    h->dv_pending = false;
    return apply_correct_now(h, nodal(h->dv[0], h->dv[1], nullptr), nodal(nullptr, nullptr, nullptr), false);
}

int do_correct(pfem2_handle *h, NodalVel v, NodalVel vold, bool has_old)
{
    if (!h) return PFEM2_EINVAL;
    if (!h->seeded) return fail(h, PFEM2_ESTATE, "correct before seed");
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = flush_correct(h))) return rc; // two corrections in a row: the first one is applied eagerly
    if (!h->opt.defer_correct || h->move_pending) return apply_correct_now(h, v, vold, has_old);
    const int N = h->mesh.n_nodes;
    for (double *&d : h->dv)
        if (!d) CU(cudaMalloc((void **)&d, sizeof(double) * (size_t)N));
    if (!h->dv2) CU(cudaMalloc((void **)&h->dv2, sizeof(double2) * (size_t)N));
    PhaseScope ps(h, PFEM2_PHASE_CORRECT);
    // the increment is only ever read at the nodes of owned cells (the particles this handle holds at the next move pass)
    const int n0 = h->own_node_lo, n1 = h->own_node_hi;
    if (n1 > n0)
        PFEM2_LAUNCH(k_snapshot_dv, grid_for(n1 - n0, kThreads, 1 << 30), kThreads, 0, h->stream, n0, n1, v, vold, has_old ? 1 : 0, h->dv[0],
                     h->dv[1], h->dv2);
    CU(cudaGetLastError());
    h->dv_pending = true;
    return PFEM2_OK;
}

// node -> incidence CSR on the device with the library's own radix sort; uses keys/vals as scratch
int build_node_incidence(pfem2_handle *h)
{
    cudaStream_t st = h->stream;
    const int C = h->mesh.n_cells, N = h->mesh.n_nodes;
    const int m = 3 * C;
    int *count = nullptr, *n_dev = nullptr, *scratch = nullptr;
    int rc;
    if ((rc = dev_alloc(h, &count, (size_t)N + 1))) return rc;
    if ((rc = dev_alloc(h, &n_dev, 1))) return rc;
    if ((rc = dev_alloc(h, &scratch, scan_scratch_elems<int>(N)))) return rc;
    CU(cudaMemsetAsync(count, 0, sizeof(int) * ((size_t)N + 1), st));
    CU(cudaMemcpyAsync(n_dev, &m, sizeof(int), cudaMemcpyHostToDevice, st));
    PFEM2_LAUNCH(k_incidence_keys, grid_for(m, kThreads, 1 << 30), kThreads, 0, st, C, h->mesh.d_cells, h->keys[0], h->vals[0], count);
    int bits = 1;
    while ((1ll << bits) < N) ++bits;
    const int flip = radix_sort_pairs(h->keys[0], h->vals[0], h->keys[1], h->vals[1], n_dev, h->capacity, bits, h->rs_hist,
                                      h->rs_scan_scratch, h->rs_info, st);
    CU(cudaMemcpyAsync(h->node_inc, h->vals[flip], sizeof(unsigned) * (size_t)m, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(n_dev, &N, sizeof(int), cudaMemcpyHostToDevice, st));
    exclusive_scan_dev<int>(count, h->node_off, n_dev, 1, 0, N, scratch, st);
    CU(cudaStreamSynchronize(st));
    cudaFree(count); cudaFree(n_dev); cudaFree(scratch);
    return PFEM2_OK;
}

} // namespace

extern "C" {

void pfem2_default_options(pfem2_options *o)
{
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->struct_size = (int)sizeof *o;
    o->subcell_mode = 0;
    o->max_division_level = 4;
    o->capacity_factor = 1.5;
    o->stream = nullptr;
    o->device = -1;
    o->verbose = 0;
    o->defer_correct = 1;
}

const char *pfem2_last_error(const pfem2_handle *h) { return h ? h->error.c_str() : g_create_error.c_str(); }

const char *pfem2_version(void) { return "pfem2_b200 0.1 (sm_100a)"; }

long long pfem2_kernel_launches(void) { return g_kernel_launches; }

int pfem2_create(pfem2_handle **out, const pfem2_mesh_view *mesh, int cell_division_level, const pfem2_options *opt_in)
{
    pfem2_handle *h = nullptr;
    if (!out || !mesh) return fail(nullptr, PFEM2_EINVAL, "null argument");
    *out = nullptr;
    if (mesh->n_cells <= 0 || mesh->n_nodes <= 0 || !mesh->d_vertices || !mesh->d_cells || !mesh->d_inv_jacobi ||
        !mesh->d_nbr_offsets || !mesh->d_nbr_indices)
        return fail(nullptr, PFEM2_EINVAL, "incomplete mesh view");
    pfem2_options opt;
    pfem2_default_options(&opt);
    if (opt_in) memcpy(&opt, opt_in, std::min<size_t>(sizeof opt, (size_t)std::max(opt_in->struct_size, 0)));
    if (opt.max_division_level <= 0) opt.max_division_level = 4;
    if (opt.max_division_level > kMaxLevel) return fail(nullptr, PFEM2_EINVAL, "max_division_level > 8");
    if (opt.capacity_factor < 1.05) opt.capacity_factor = 1.5;
    if (opt.lazy_sort && (opt.stable_order || opt.lane_per_record || opt.fuse_project || opt.scatter_tma))
        return fail(nullptr, PFEM2_EINVAL, "lazy_sort works with the default kernels only (no stable_order / lane_per_record / fuse_project / scatter_tma)");

    int dev = opt.device;
    if (dev < 0) CU(cudaGetDevice(&dev));
    CU(cudaSetDevice(dev));
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) g_num_sms = sms;
    }
    h = new pfem2_handle;
    h->device = dev;
    h->opt = opt;
    h->stream = (cudaStream_t)opt.stream;
    h->mesh = *mesh;
    {
        const char *e = getenv("PFEM2_LAZY_SWIZZLE"); // env: hardware bring-up of the lazy move pass only
        h->lazy_swizzle = !(e && atoi(e) == 0);
        e = getenv("PFEM2_LAZY_NSUB3");
        h->lazy_nsub3 = !(e && atoi(e) == 0);
    }
    // :241-243
    const int n = std::max(std::min(cell_division_level, opt.max_division_level), 1);
    h->level = n;
    h->ppc = n * n;
    h->sub_step = 1.0 / n;
    const int C = mesh->n_cells, N = mesh->n_nodes;
    if ((long long)C * h->ppc * opt.capacity_factor > 2147483000.0) {
        delete h;
        return fail(nullptr, PFEM2_EINVAL, "cells * particlesPerCell * capacity_factor exceeds 32-bit indexing");
    }
    h->key_bits = 1;
    while ((1ll << h->key_bits) <= C) ++h->key_bits; // keys 0..C (C = lost)
    h->own_lo = 0;
    h->own_hi = C;
    h->own_node_lo = h->v2_node_lo = 0;
    h->own_node_hi = h->v2_node_hi = mesh->n_nodes;

    // sub-cell centres (:248-274), host arithmetic without contraction
    std::vector<double> cen(3 * (size_t)h->ppc);
    {
        int num = -1;
        const volatile double dx = 1.0 / n;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < 2 * i + 1; ++j) {
                const volatile double xmin = (j / 2) * dx;
                const volatile double xmax = xmin + dx;
                const volatile double ymin = (n - 1 - i) * dx;
                const volatile double ymax = ymin + dx;
                volatile double v0[3], v1[3], v2[3];
                v0[0] = xmin; v0[1] = ymax; { volatile double t = 1.0 - xmin; v0[2] = t - ymax; }
                v1[0] = (j % 2 == 0) ? xmin : xmax;
                v1[1] = (j % 2 == 0) ? ymin : ymax;
                { volatile double t = 1.0 - v1[0]; v1[2] = t - v1[1]; }
                v2[0] = xmax; v2[1] = ymin; { volatile double t = 1.0 - xmax; v2[2] = t - ymin; }
                ++num;
                for (int k = 0; k < 3; ++k) {
                    volatile double s = v0[k] + v1[k];
                    s = s + v2[k];
                    cen[3 * (size_t)num + k] = s * 0.3333333333333333; // CONSTANTS::ONE_THIRD
                }
            }
    }

    int rc;
#define TRY(x) do { if ((rc = (x))) { std::string e = h->error; pfem2_destroy(h); g_create_error = e; return rc; } } while (0)
    TRY(dev_alloc(h, &h->geom, (size_t)C));
    TRY(dev_alloc(h, &h->node_off, (size_t)N + 1));
    TRY(dev_alloc(h, &h->node_inc, 3 * (size_t)C));
    TRY(dev_alloc(h, &h->centers, cen.size()));
    TRY(dev_alloc(h, &h->ctr, 1));
    TRY(dev_alloc(h, &h->stay, 3 * ((size_t)C + 1)));
    h->arrive = h->stay + ((size_t)C + 1);
    h->cursor = h->arrive + ((size_t)C + 1);
    TRY(dev_alloc(h, &h->n_cells_dev, 1));
    TRY(dev_alloc(h, &h->own_len_dev, 1));
    TRY(dev_alloc(h, &h->rs_info, 4));
    TRY(dev_alloc(h, &h->edge_nbr, (size_t)C));
    TRY(dev_alloc(h, &h->cell_mask, (size_t)C + 1));
    TRY(dev_alloc(h, &h->packed, (size_t)C + 2));
    TRY(dev_alloc(h, &h->scan_scratch64, scan_scratch_elems<unsigned long long>(C)));
    TRY(dev_alloc(h, &h->cell_start[0], (size_t)C + 1));
    TRY(dev_alloc(h, &h->cell_start[1], (size_t)C + 1));
    TRY(dev_alloc(h, &h->partial, 9 * (size_t)C));
    {
        cudaError_t e = cudaMallocHost((void **)&h->host_ctr, sizeof(Counters));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->readback, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            h->error = std::string("pinned/event allocation failed: ") + cudaGetErrorString(e);
            TRY(PFEM2_ECUDA);
        }
        memset(h->host_ctr, 0, sizeof(Counters));
    }
    const long long want = std::max<long long>((long long)std::ceil((double)C * h->ppc * opt.capacity_factor), 3ll * C);
    TRY(alloc_particle_storage(h, (int)want));
    {
        cudaError_t e = cudaMemcpyAsync(h->centers, cen.data(), cen.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { h->error = cudaGetErrorString(e); TRY(PFEM2_ECUDA); }
    }
    PFEM2_LAUNCH(k_build_geom, grid_for(C, kThreads, 1 << 30), kThreads, 0, h->stream, C, (const double2 *)mesh->d_vertices, mesh->d_cells,
                 mesh->d_inv_jacobi, h->geom);
    PFEM2_LAUNCH(k_set_counters, 1, 1, 0, h->stream, h->ctr, 0, h->capacity);
    {
        // locate acceleration data (edge neighbours, strict-interior margins); partial[] doubles as scratch
        double *hmin = h->partial;
        unsigned long long *dmax = h->packed;
        cudaError_t e = cudaMemsetAsync(dmax, 0, sizeof(unsigned long long), h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h->n_cells_dev, &C, sizeof(int), cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) { h->error = cudaGetErrorString(e); TRY(PFEM2_ECUDA); }
        PFEM2_LAUNCH(k_cell_metrics, grid_for(C, kThreads, 1 << 30), kThreads, 0, h->stream, C, (const double2 *)mesh->d_vertices, h->geom,
                     hmin, dmax);
        PFEM2_LAUNCH(k_build_locate_data, grid_for(C, kThreads, 1 << 30), kThreads, 0, h->stream, C, h->geom, mesh->d_nbr_offsets,
                     mesh->d_nbr_indices, hmin, dmax, h->edge_nbr);
    }
    TRY(build_node_incidence(h));
    {
        cudaError_t e = cudaMemsetAsync(h->partial, 0, sizeof(double) * 9 * (size_t)C, h->stream); // was scratch above
        if (e == cudaSuccess) e = cudaMemcpyAsync(h->own_len_dev, &C, sizeof(int), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { h->error = cudaGetErrorString(e); TRY(PFEM2_ECUDA); }
    }
#undef TRY
    *out = h;
    return PFEM2_OK;
}

int pfem2_destroy(pfem2_handle *h)
{
    if (!h) return PFEM2_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_soa(h->soa[0]);
    free_soa(h->soa[1]);
    free_particle_scratch(h);
    cudaFree(h->geom); cudaFree(h->node_off); cudaFree(h->node_inc); cudaFree(h->centers); cudaFree(h->ctr);
    cudaFree(h->stay); cudaFree(h->cell_mask); cudaFree(h->packed); cudaFree(h->scan_scratch64);
    cudaFree(h->cell_start[0]); cudaFree(h->cell_start[1]); cudaFree(h->partial); cudaFree(h->aos);
    cudaFree(h->n_cells_dev); cudaFree(h->rs_info); cudaFree(h->edge_nbr);
    cudaFree(h->mg_bounds); cudaFree(h->mg_rank_count); cudaFree(h->own_len_dev); cudaFree(h->node_list);
    cudaFree(h->dv[0]); cudaFree(h->dv[1]);
    cudaFree(h->dv2); cudaFree(h->v2);
    for (int k = 0; k < 2; ++k) {
        if (h->p2p.peer[k]) cudaIpcCloseMemHandle(h->p2p.peer[k]);
        cudaFree(h->p2p.inbox[k]);
        cudaFree(h->p2p.idx[k]);
    }
    cudaFree(h->p2p.cursors);
    cudaFree(h->tail_cursor);
    if (h->trail_stream) cudaStreamDestroy(h->trail_stream);
    for (cudaEvent_t e : h->trail_ev) if (e) cudaEventDestroy(e);
    cudaFree(h->trail_prog);
    if (h->pipe.copy) cudaStreamDestroy(h->pipe.copy);
    for (cudaEvent_t e : h->pipe.up_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : h->pipe.dn_ev) cudaEventDestroy(e);
    for (double *p : h->nodal) cudaFree(p);
    for (auto &r : h->phase_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
    if (h->host_ctr) cudaFreeHost(h->host_ctr);
    if (h->readback) cudaEventDestroy(h->readback);
    delete h;
    return PFEM2_OK;
}

int pfem2_seed(pfem2_handle *h)
{
    if (h) h->partials_valid = false;
    if (!h) return PFEM2_EINVAL;
    CU(cudaSetDevice(h->device));
    const int C = h->mesh.n_cells;
    h->cur = 0;
    h->cs = 0;
    h->permuted = false;
    PFEM2_LAUNCH(k_set_counters, 1, 1, 0, h->stream, h->ctr, 0, h->capacity);
    PFEM2_LAUNCH(k_seed, grid_for(std::max<long long>((long long)(h->own_hi - h->own_lo) * h->ppc, C + 1)), kThreads, 0, h->stream, C,
                 h->own_lo, h->own_hi, h->ppc, (const double2 *)h->mesh.d_vertices, h->geom, h->centers, h->soa[0], h->cell_start[0],
                 h->ctr);
    CU(cudaGetLastError());
    h->seeded = true;
    h->dv_pending = false;
    h->host_count = (h->own_hi - h->own_lo) * h->ppc;
    h->host_added = 0;
    h->readback_pending = false;
    if (h->opt.verbose) {
        CU(cudaStreamSynchronize(h->stream)); // the reference synchronises here too (:315)
        printf("Created %d particles\n", h->host_count); // :319
    }
    return PFEM2_OK;
}

int pfem2_init_velocity(pfem2_handle *h, const double *vx, const double *vy)
{
    return do_correct(h, nodal(vx, vy, nullptr), nodal(nullptr, nullptr, nullptr), false);
}
int pfem2_init_velocity_ptrs(pfem2_handle *h, double *const *t)
{
    return do_correct(h, nodal(nullptr, nullptr, t), nodal(nullptr, nullptr, nullptr), false);
}
int pfem2_advect(pfem2_handle *h, const double *vx, const double *vy, double dt, int substeps)
{
    return do_advect(h, nodal(vx, vy, nullptr), dt, substeps);
}
int pfem2_advect_ptrs(pfem2_handle *h, double *const *t, double dt, int substeps)
{
    return do_advect(h, nodal(nullptr, nullptr, t), dt, substeps);
}
int pfem2_project(pfem2_handle *h, double *vx, double *vy) { return do_project(h, vx, vy, nullptr); }
int pfem2_project_ptrs(pfem2_handle *h, double *const *t) { return do_project(h, nullptr, nullptr, t); }
int pfem2_project_dual(pfem2_handle *h, double *vx, double *vy, double *cx, double *cy)
{
    if (!vx || !vy || !cx || !cy) return h ? fail(h, PFEM2_EINVAL, "project_dual: null nodal array") : PFEM2_EINVAL;
    return do_project(h, vx, vy, nullptr, cx, cy, nullptr);
}
int pfem2_project_dual_ptrs(pfem2_handle *h, double *const *t, double *const *t_copy)
{
    if (!t || !t_copy) return h ? fail(h, PFEM2_EINVAL, "project_dual: null pointer table") : PFEM2_EINVAL;
    return do_project(h, nullptr, nullptr, t, nullptr, nullptr, t_copy);
}
int pfem2_correct(pfem2_handle *h, const double *vx, const double *vy, const double *ox, const double *oy)
{
    return do_correct(h, nodal(vx, vy, nullptr), nodal(ox, oy, nullptr), true);
}
int pfem2_correct_ptrs(pfem2_handle *h, double *const *t, double *const *told)
{
    return do_correct(h, nodal(nullptr, nullptr, t), nodal(nullptr, nullptr, told), true);
}

int pfem2_particle_count(pfem2_handle *h, int *out)
{
    if (!h || !out) return PFEM2_EINVAL;
    int rc;
    if ((rc = sync_counters(h))) return rc;
    *out = h->host_count;
    return PFEM2_OK;
}

int pfem2_get_stats(pfem2_handle *h, pfem2_stats *out)
{
    if (!h || !out) return PFEM2_EINVAL;
    CU(cudaSetDevice(h->device));
    const int rc = sync_counters(h);
    Counters c;
    CU(cudaMemcpyAsync(&c, h->ctr, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    out->count = c.count; out->lost = c.lost; out->added = c.added; out->movers = c.movers;
    out->capacity = h->capacity; out->overflow = c.overflow;
    return rc;
}

int pfem2_export_aos(pfem2_handle *h, const void **d_particles96, int *count)
{
    if (!h || !d_particles96) return PFEM2_EINVAL;
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = sync_counters(h))) return rc;
    if ((rc = materialize(h))) return rc;
    if ((rc = flush_correct(h))) return rc;
    const size_t need = (size_t)std::max(h->capacity, 1) * 96;
    if (h->aos_bytes < need) {
        if (h->aos) cudaFree(h->aos);
        h->aos = nullptr;
        CU(cudaMalloc(&h->aos, need));
        h->aos_bytes = need;
    }
    PFEM2_LAUNCH(k_export_aos, grid_for(h->capacity), kThreads, 0, h->stream, h->soa[h->cur], h->ctr, (uint4 *)h->aos);
    CU(cudaGetLastError());
    *d_particles96 = h->aos;
    if (count) *count = h->host_count;
    return PFEM2_OK;
}

// Plan of the pipelined pfem2_step_host for K chunks (made once per (K, substeps)): cell chunk bounds, node slices of the
// upload, and per chunk the upload slices it depends on / the node prefix that is final after its projection.  The
// dependencies are derived from the mesh itself (band width of the one-ring lists x substeps), so any numbering is handled:
// a numbering without locality simply yields "wait for the whole upload" and "download at the end".
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
    pp.cb.resize(K + 1);
    pp.ns.resize(K + 1);
    for (int j = 0; j <= K; ++j) pp.cb[j] = (int)((long long)C * j / K);
    int *dev = nullptr; // [band | cb (K+1) | up_need (K) | dn_ready (K)]
    CU(cudaMalloc((void **)&dev, sizeof(int) * (size_t)(3 * K + 2)));
    std::vector<int> init(3 * K + 2, 0);
    for (int j = 0; j <= K; ++j) init[1 + j] = pp.cb[j];
    for (int j = 0; j < K; ++j) init[2 + 2 * K + j] = N; // dn_ready starts at "everything"
    CU(cudaMemcpyAsync(dev, init.data(), sizeof(int) * init.size(), cudaMemcpyHostToDevice, st));
    PFEM2_LAUNCH(k_band_width, grid_for(C, kThreads, 1 << 30), kThreads, 0, st, C, h->mesh.d_nbr_offsets, h->mesh.d_nbr_indices, dev);
    int band = 0;
    CU(cudaMemcpyAsync(&band, dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const long long ext = std::min<long long>((long long)band * substeps, C);
    PFEM2_LAUNCH(k_chunk_node_ranges, grid_for(C, kThreads, 1 << 30), kThreads, 0, st, C, h->geom, K, dev + 1, (int)ext, dev + 2 + K,
                 dev + 2 + 2 * K);
    std::vector<int> out(3 * K + 2);
    CU(cudaMemcpyAsync(out.data(), dev, sizeof(int) * out.size(), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(dev);
    // upload slice j = exactly the node prefix chunk j needs on top of what the chunks before it needed (chunks run in
    // order, so the dependency only grows); download prefix after chunk j likewise
    pp.up_slice.assign(K, 0);
    pp.dn_ready.assign(K, 0);
    int prev_up = 0, prev_dn = 0;
    pp.ns[0] = 0;
    for (int j = 0; j < K; ++j) {
        const int need = (j == K - 1) ? N : std::min(std::max(out[2 + K + j], 1), N); // node prefix [0, need) must have landed
        prev_up = std::max(prev_up, need);
        pp.ns[j + 1] = prev_up;
        pp.up_slice[j] = j;
        const int ready = (j == K - 1) ? N : std::min(out[2 + 2 * K + j], N);
        prev_dn = std::max(prev_dn, ready);
        pp.dn_ready[j] = prev_dn;
    }
    pp.K = K;
    pp.substeps = substeps;
    return PFEM2_OK;
}

int host_pipe_chunks(const pfem2_handle *h)
{
    if (h->opt.host_pipeline == 1 || !advect_tma_enabled(h) || h->opt.stable_order) return 1;
    if (h->own_lo != 0 || h->own_hi != h->mesh.n_cells) return 1; // multi-GPU strips exchange particles between the phases
    if (h->opt.host_pipeline > 1) return std::min(h->opt.host_pipeline, 16);
    return h->mesh.n_cells < (1 << 18) ? 1 : 8; // small meshes are launch-bound: one chunk (sweep on channel16m: 1 chunk 23.5 ms,
                                                // 2: 20.3, 4: 19.3, 6: 18.9, 8: 18.8, 12: 18.7; device time alone 17.8)
}

int pfem2_step_host(pfem2_handle *h, const double *fx, const double *fy, double *wx, double *wy, double dt, int substeps,
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
    rc = pfem2_advect(h, h->nodal[0], h->nodal[1], dt, substeps); // the move pass runs chunk by chunk (launch_advect)
    pp.active = false;
    if (rc) return rc;
    for (; pp.packed_slices < K; ++pp.packed_slices) CU(cudaStreamWaitEvent(st, pp.up_ev[pp.packed_slices], 0));
    if ((rc = flush_correct(h))) return rc; // nothing pending after an advect; kept for symmetry with do_project
    {
        ParticleSoA p = h->soa[h->cur];
        int done = 0; // nodes [0, done) are final and on their way to the host
        for (int j = 0; j < K; ++j) {
            if (!h->partials_valid) {
                PhaseScope ps(h, PFEM2_PHASE_PROJECT_CELLS);
                launch_project_cells(h, p, pp.cb[j], pp.cb[j + 1]);
            }
            const int ready = pp.dn_ready[j];
            if (ready > done) {
                {
                    PhaseScope ps(h, PFEM2_PHASE_PROJECT_NODES);
                    PFEM2_LAUNCH(k_project_nodes, grid_for(ready - done, kThreads, 1 << 30), kThreads, 0, st, done, ready, h->node_off,
                                 (const int *)h->node_inc, h->partial, h->nodal[2], h->nodal[3], (double *const *)nullptr);
                }
                CU(cudaEventRecord(pp.dn_ev[j], st));
                CU(cudaStreamWaitEvent(pp.copy, pp.dn_ev[j], 0));
                const size_t len = (size_t)(ready - done) * sizeof(double);
                CU(cudaMemcpyAsync(wx + done, h->nodal[2] + done, len, cudaMemcpyDeviceToHost, pp.copy));
                CU(cudaMemcpyAsync(wy + done, h->nodal[3] + done, len, cudaMemcpyDeviceToHost, pp.copy));
                done = ready;
            }
        }
    }
    CU(cudaGetLastError());
    if ((rc = pfem2_correct(h, h->nodal[0], h->nodal[1], h->nodal[2], h->nodal[3]))) return rc;
    CU(cudaStreamSynchronize(pp.copy));
    CU(cudaStreamSynchronize(st));
    if ((rc = sync_counters(h))) return rc;
    if (count_out) *count_out = h->host_count;
    return PFEM2_OK;
}

int pfem2_download(pfem2_handle *h, double *x, double *y, double *l0, double *l1, double *l2, double *vx, double *vy,
                   unsigned *cell, unsigned *id)
{
    if (!h) return PFEM2_EINVAL;
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = sync_counters(h))) return rc;
    if ((rc = materialize(h))) return rc;
    if ((rc = flush_correct(h))) return rc;
    const size_t n = (size_t)h->host_count;
    const ParticleSoA &p = h->soa[h->cur];
    std::vector<ParticleRec> hr(n);
    if (n) CU(cudaMemcpyAsync(hr.data(), p.records(), n * sizeof(ParticleRec), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < n; ++i) {
        if (x) x[i] = hr[i].pos.x;
        if (y) y[i] = hr[i].pos.y;
        if (l0) l0[i] = hr[i].lab.x;
        if (l1) l1[i] = hr[i].lab.y;
        if (l2) l2[i] = hr[i].tail.l2;
        if (vx) vx[i] = hr[i].vel.x;
        if (vy) vy[i] = hr[i].vel.y;
        if (cell) cell[i] = hr[i].tail.cell;
        if (id) id[i] = hr[i].tail.id;
    }
    return PFEM2_OK;
}

int pfem2_upload(pfem2_handle *h, int n, const double *x, const double *y, const double *l0, const double *l1, const double *l2,
                 const double *vx, const double *vy, const unsigned *cell, const unsigned *id)
{
    if (h) h->partials_valid = false;
    if (!h || n < 0 || !x || !y || !l0 || !l1 || !l2 || !vx || !vy || !cell) return PFEM2_EINVAL;
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = sync_counters(h))) return rc;
    h->dv_pending = false; // the uploaded state replaces everything, including a correction not yet applied
    h->permuted = false;   // ... and a permutation of the replaced state
    if (n > h->capacity) {
        h->host_count = 0;
        if ((rc = grow(h, (int)std::min<long long>(2147483000ll, (long long)(1.25 * n) + 4096)))) return rc;
    }
    cudaStream_t st = h->stream;
    const int C = h->mesh.n_cells;
    ParticleSoA &p = h->soa[h->cur];
    {
        std::vector<ParticleRec> hr(n);
        for (int i = 0; i < n; ++i) {
            if (cell[i] >= (unsigned)C) return fail(h, PFEM2_EINVAL, "upload: cell index out of range");
            hr[i].pos = make_double2(x[i], y[i]);
            hr[i].lab = make_double2(l0[i], l1[i]);
            hr[i].vel = make_double2(vx[i], vy[i]);
            hr[i].tail.l2 = l2[i];
            hr[i].tail.cell = cell[i];
            hr[i].tail.id = id ? id[i] : 0u;
        }
        if (n) CU(cudaMemcpyAsync(p.records(), hr.data(), (size_t)n * sizeof(ParticleRec), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
    }
    PFEM2_LAUNCH(k_set_counters, 1, 1, 0, st, h->ctr, n, h->capacity);
    {   // per-cell scratch of the owned range (+ a few cells for the tolerance-band spill of the occupancy bits)
        const size_t lo = (size_t)h->own_lo, len = (size_t)std::min(C, h->own_hi + 4) - lo + 1;
        CU(cudaMemsetAsync(h->stay + lo, 0, sizeof(int) * len, st));
        CU(cudaMemsetAsync(h->arrive + lo, 0, sizeof(int) * len, st));
        CU(cudaMemsetAsync(h->cursor + lo, 0, sizeof(int) * len, st));
        CU(cudaMemsetAsync(h->cell_mask + lo, 0, sizeof(unsigned long long) * len, st));
    }
    PFEM2_LAUNCH(k_all_movers, grid_for(h->capacity), kThreads, 0, st, p, C, h->ctr, h->keys[0], h->vals[0], h->arrive, &h->ctr->n_movers);
    if ((rc = reorder(h, false, false, true, nodal(nullptr, nullptr, nullptr)))) return rc;
    h->seeded = true;
    if ((rc = queue_readback(h))) return rc;
    return sync_counters(h);
}

int pfem2_device_records(pfem2_handle *h, const void **d_records)
{
    if (!h || !d_records) return PFEM2_EINVAL;
    CU(cudaSetDevice(h->device));
    {   // the records expose the physical order and the particle velocities: a pending permutation (lazy re-sort) is made
        // physical and a deferred correction is applied first
        int rc = materialize(h);
        if (!rc) rc = flush_correct(h);
        if (rc) return rc;
    }
    *d_records = h->soa[h->cur].records();
    return PFEM2_OK;
}

int pfem2_cell_starts(pfem2_handle *h, const int **d_cell_start)
{
    if (!h || !d_cell_start) return PFEM2_EINVAL;
    *d_cell_start = h->cell_start[h->cs];
    return PFEM2_OK;
}

int pfem2_set_profiling(pfem2_handle *h, int enabled)
{
    if (!h) return PFEM2_EINVAL;
    h->profiling = enabled != 0;
    return PFEM2_OK;
}

int pfem2_get_phase_times(pfem2_handle *h, double *ms, long long *calls, int reset)
{
    if (!h) return PFEM2_EINVAL;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    for (auto &r : h->phase_recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
            h->phase_ms[r.phase] += t;
            h->phase_calls[r.phase] += 1;
        }
        h->event_pool.push_back(r.a);
        h->event_pool.push_back(r.b);
    }
    h->phase_recs.clear();
    for (int k = 0; k < PFEM2_NUM_PHASES; ++k) {
        if (ms) ms[k] = h->phase_ms[k];
        if (calls) calls[k] = h->phase_calls[k];
        if (reset) { h->phase_ms[k] = 0; h->phase_calls[k] = 0; }
    }
    return PFEM2_OK;
}

int pfem2_set_owned_cells(pfem2_handle *h, int cell_lo, int cell_hi)
{
    if (!h) return PFEM2_EINVAL;
    if (cell_lo < 0 || cell_hi > h->mesh.n_cells || cell_lo > cell_hi) return fail(h, PFEM2_EINVAL, "bad owned cell range");
    if (h->seeded) return fail(h, PFEM2_ESTATE, "set_owned_cells after seed");
    if (h->opt.lazy_sort && (cell_lo != 0 || cell_hi != h->mesh.n_cells))
        return fail(h, PFEM2_EINVAL, "lazy_sort is single-GPU only (the strip-partitioned calls work on the physical order)");
    h->own_lo = cell_lo;
    h->own_hi = cell_hi;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int own_n = cell_hi - cell_lo, N = h->mesh.n_nodes;
    CU(cudaMemcpyAsync(h->own_len_dev, &own_n, sizeof(int), cudaMemcpyHostToDevice, st));
    // compact list of the nodes the owned cells touch
    int *flag = nullptr, *pos = nullptr, *scratch = nullptr, *n_dev = nullptr;
    CU(cudaMalloc((void **)&flag, sizeof(int) * ((size_t)N + 1)));
    CU(cudaMalloc((void **)&pos, sizeof(int) * ((size_t)N + 1)));
    CU(cudaMalloc((void **)&scratch, sizeof(int) * scan_scratch_elems<int>(N)));
    CU(cudaMalloc((void **)&n_dev, sizeof(int)));
    CU(cudaMemsetAsync(flag, 0, sizeof(int) * ((size_t)N + 1), st));
    CU(cudaMemcpyAsync(n_dev, &N, sizeof(int), cudaMemcpyHostToDevice, st));
    if (own_n > 0) PFEM2_LAUNCH(k_mark_nodes, grid_for(own_n, kThreads, 1 << 30), kThreads, 0, st, cell_lo, cell_hi, h->mesh.d_cells, flag);
    exclusive_scan_dev<int>(flag, pos, n_dev, 1, 0, N, scratch, st);
    int total = 0;
    CU(cudaMemcpyAsync(&total, pos + N, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(h->node_list);
    h->node_list = nullptr;
    h->n_node_list = total;
    if (own_n < h->mesh.n_cells) {
        CU(cudaMalloc((void **)&h->node_list, sizeof(int) * (size_t)std::max(total, 1)));
        PFEM2_LAUNCH(k_compact_nodes, grid_for(N, kThreads, 1 << 30), kThreads, 0, st, N, flag, pos, h->node_list);
        CU(cudaStreamSynchronize(st));
    }
    cudaFree(flag); cudaFree(pos); cudaFree(scratch); cudaFree(n_dev);
    CU(cudaGetLastError());
    h->v2_range_substeps = -1;
    if (own_n < h->mesh.n_cells) {
        const int rcn = node_range_of_cells(h, cell_lo, cell_hi, h->own_node_lo, h->own_node_hi);
        if (rcn) return rcn;
    } else {
        h->own_node_lo = 0;
        h->own_node_hi = N;
    }
    return PFEM2_OK;
}

int pfem2_advect_move(pfem2_handle *h, const double *vx, const double *vy, double dt, int substeps)
{
    return advect_move(h, nodal(vx, vy, nullptr), dt, substeps, 0, true);
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
                         h->mg_bounds, h->mg_ranks, h->mg_rank_count, (int4 *)d_records);
    } else
    PFEM2_LAUNCH(k_emigrant_pack, grid_for(h->capacity), kThreads, 0, st, h->soa[h->cur], h->ctr, h->own_lo, h->own_hi, h->mg_bounds,
                 h->mg_ranks, h->mg_rank_count, (int4 *)d_records);
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
    PFEM2_LAUNCH(k_immigrant_append, grid_for(n), kThreads, 0, h->stream, h->soa[h->cur], h->ctr, (const int4 *)d_records, n);
    if (h->mg_fused) { // the move pass counted the residents; the immigrants are counted here (no pass over everybody later)
        const int C = h->mesh.n_cells;
        const bool m64 = h->ppc > 32;
#define PFEM2_CNTA(M, B)                                                                                                             \
    PFEM2_LAUNCH((k_count_appended<M, B>), grid_for(n), kThreads, 0, h->stream, h->soa[h->cur], h->ctr, n, C, h->ppc, h->level,        \
                 h->sub_step, h->stay, h->arrive, h->cell_mask)
        if (h->opt.subcell_mode == 0) { if (m64) PFEM2_CNTA(0, true); else PFEM2_CNTA(0, false); }
        else                          { if (m64) PFEM2_CNTA(1, true); else PFEM2_CNTA(1, false); }
#undef PFEM2_CNTA
    }
    PFEM2_LAUNCH(k_add_count, 1, 1, 0, h->stream, h->ctr, n);
    h->host_count += n;
    CU(cudaGetLastError());
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
                 h->cell_mask, h->own_hi, h->mesh.n_cells);
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
    PFEM2_LAUNCH(k_emigrant_pack_p2p, grid_for(cap, kThreads, g_num_sms * 2), kThreads, 0, st, h->soa[h->cur], h->keys[0], h->mg_rank_count,
                 h->mg_ranks, h->mg_bounds, rank, hl ? (int4 *)(hl + 1) : nullptr, hr ? (int4 *)(hr + 1) : nullptr, cap, h->ctr,
                 h->p2p.cursors);
    PFEM2_LAUNCH(k_p2p_publish_migration, 1, 1, 0, st, hl, pl ? &((P2PInboxHead *)pl)->flag_mig : nullptr, hr,
                 pr ? &((P2PInboxHead *)pr)->flag_mig : nullptr, h->p2p.cursors, cap, h->cell_mask, h->own_hi, h->mesh.n_cells, seq);
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
    PFEM2_LAUNCH(k_p2p_wait, 1, 1, 0, st, il ? &((const P2PInboxHead *)il)->flag_mig : nullptr,
                 ir ? &((const P2PInboxHead *)ir)->flag_mig : nullptr, seq, h->ctr, p2p_timeout_ns());
    int rc;
    if (il && (rc = append_migration_block(h, (const int4 *)(il + p2p_block_offset(cap, parity)), cap, 1))) return rc;
    if (ir && (rc = append_migration_block(h, (const int4 *)(ir + p2p_block_offset(cap, parity)), cap, 0))) return rc;
    return PFEM2_OK;
}

int pfem2_project_halo_p2p(pfem2_handle *h, double *d_acc3)
{
    if (!h || !d_acc3) return PFEM2_EINVAL;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int cap = h->p2p.cap;
    if (!h->p2p.peer[0] && !h->p2p.peer[1]) return PFEM2_OK;
    const unsigned seq = ++h->p2p.halo_seq;
    const int parity = (int)(seq & 1u);
    unsigned *flags[2] = {nullptr, nullptr};
    for (int k = 0; k < 2; ++k) {
        if (!h->p2p.peer[k]) continue;
        unsigned char *peer = (unsigned char *)h->p2p.peer[k];
        const int n = h->p2p.n_idx[k];
        if (n)
            PFEM2_LAUNCH(k_halo_send, grid_for(n, kThreads, 1 << 30), kThreads, 0, st, d_acc3, h->p2p.idx[k], n,
                         (double *)(peer + p2p_halo_offset(cap, n, parity)));
        flags[k] = &((P2PInboxHead *)peer)->flag_halo;
    }
    PFEM2_LAUNCH(k_p2p_publish_flag, 1, 1, 0, st, flags[0], flags[1], seq);
    PFEM2_LAUNCH(k_p2p_wait, 1, 1, 0, st, h->p2p.peer[0] ? &((const P2PInboxHead *)h->p2p.inbox[0])->flag_halo : nullptr,
                 h->p2p.peer[1] ? &((const P2PInboxHead *)h->p2p.inbox[1])->flag_halo : nullptr, seq, h->ctr, p2p_timeout_ns());
    for (int k = 0; k < 2; ++k) {
        if (!h->p2p.peer[k]) continue;
        const int n = h->p2p.n_idx[k];
        if (n)
            PFEM2_LAUNCH(k_halo_add, grid_for(n, kThreads, 1 << 30), kThreads, 0, st, d_acc3, h->p2p.idx[k], n,
                         (const double *)((unsigned char *)h->p2p.inbox[k] + p2p_halo_offset(cap, n, parity)));
    }
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
    const int C = h->mesh.n_cells, N = h->mesh.n_nodes;
    ParticleSoA p = h->soa[h->cur];
    {
        PhaseScope ps(h, PFEM2_PHASE_PROJECT_CELLS);
        launch_project_cells(h, p);
    }
    PhaseScope ps(h, PFEM2_PHASE_PROJECT_NODES);
    const int nl = h->node_list ? h->n_node_list : N;
    PFEM2_LAUNCH(k_project_nodes_acc, grid_for(nl, kThreads, 1 << 30), kThreads, 0, st, nl, h->node_list, h->node_off,
                 (const int *)h->node_inc, h->partial, d_acc3);
    (void)C;
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

int pfem2_mesh_inv_jacobi(int n_cells, const double *d_vertices, const unsigned *d_cells, double *d_inv_jacobi, void *stream)
{
    pfem2_handle *h = nullptr;
    if (n_cells <= 0 || !d_vertices || !d_cells || !d_inv_jacobi) return fail(nullptr, PFEM2_EINVAL, "bad argument");
    PFEM2_LAUNCH(k_inv_jacobi, grid_for(n_cells, kThreads, 1 << 30), kThreads, 0, (cudaStream_t)stream, n_cells, (const double2 *)d_vertices,
                 d_cells, d_inv_jacobi);
    CU(cudaGetLastError());
    return PFEM2_OK;
}

int pfem2_sort_pairs(int n, int key_bits, unsigned *keys, unsigned *vals, unsigned *keys_tmp, unsigned *vals_tmp, int *result_in_tmp,
                     void *stream)
{
    pfem2_handle *h = nullptr;
    if (n < 0 || key_bits < 1 || key_bits > 32 || !keys || !vals || !keys_tmp || !vals_tmp || !result_in_tmp)
        return fail(nullptr, PFEM2_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    int *n_dev = nullptr, *hist = nullptr, *scratch = nullptr;
    CU(cudaMalloc((void **)&n_dev, 8 * sizeof(int)));
    CU(cudaMalloc((void **)&hist, sizeof(int) * rs_hist_elems(std::max(n, 1))));
    CU(cudaMalloc((void **)&scratch, sizeof(int) * rs_scan_scratch_elems(std::max(n, 1))));
    CU(cudaMemcpyAsync(n_dev, &n, sizeof(int), cudaMemcpyHostToDevice, st));
    *result_in_tmp = radix_sort_pairs(keys, vals, keys_tmp, vals_tmp, n_dev, n, key_bits, hist, scratch, n_dev + 4, st);
    CU(cudaStreamSynchronize(st));
    cudaFree(n_dev); cudaFree(hist); cudaFree(scratch);
    CU(cudaGetLastError());
    return PFEM2_OK;
}

int pfem2_mesh_one_ring(int n_nodes, int n_cells, const unsigned *d_cells, int *d_offsets, int *d_indices, int *nnz, void *stream)
{
    pfem2_handle *h = nullptr;
    if (n_nodes <= 0 || n_cells <= 0 || !d_cells || !d_offsets || !nnz) return fail(nullptr, PFEM2_EINVAL, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int m = 3 * n_cells;
    unsigned *k0, *k1, *v0, *v1;
    int *count, *node_off, *n_dev, *hist, *scratch, *scratch2, *err;
    CU(cudaMalloc((void **)&k0, sizeof(unsigned) * (size_t)m)); CU(cudaMalloc((void **)&k1, sizeof(unsigned) * (size_t)m));
    CU(cudaMalloc((void **)&v0, sizeof(unsigned) * (size_t)m)); CU(cudaMalloc((void **)&v1, sizeof(unsigned) * (size_t)m));
    CU(cudaMalloc((void **)&count, sizeof(int) * ((size_t)n_nodes + 1)));
    CU(cudaMalloc((void **)&node_off, sizeof(int) * ((size_t)n_nodes + 1)));
    CU(cudaMalloc((void **)&n_dev, sizeof(int))); CU(cudaMalloc((void **)&err, sizeof(int)));
    CU(cudaMalloc((void **)&hist, sizeof(int) * rs_hist_elems(m)));
    CU(cudaMalloc((void **)&scratch, sizeof(int) * rs_scan_scratch_elems(m)));
    CU(cudaMalloc((void **)&scratch2, sizeof(int) * scan_scratch_elems<int>(std::max(n_nodes, n_cells))));
    CU(cudaMemsetAsync(count, 0, sizeof(int) * ((size_t)n_nodes + 1), st));
    CU(cudaMemsetAsync(err, 0, sizeof(int), st));
    CU(cudaMemcpyAsync(n_dev, &m, sizeof(int), cudaMemcpyHostToDevice, st));
    PFEM2_LAUNCH(k_incidence_keys, grid_for(m, kThreads, 1 << 30), kThreads, 0, st, n_cells, d_cells, k0, v0, count);
    int bits = 1;
    while ((1ll << bits) < n_nodes) ++bits;
    int *info;
    CU(cudaMalloc((void **)&info, 4 * sizeof(int)));
    const int flip = radix_sort_pairs(k0, v0, k1, v1, n_dev, m, bits, hist, scratch, info, st);
    const unsigned *inc = flip ? v1 : v0;
    int *len_dev;
    CU(cudaMalloc((void **)&len_dev, 2 * sizeof(int)));
    {
        const int lens[2] = {n_nodes, n_cells};
        CU(cudaMemcpyAsync(len_dev, lens, sizeof lens, cudaMemcpyHostToDevice, st));
    }
    exclusive_scan_dev<int>(count, node_off, len_dev, 1, 0, n_nodes, scratch2, st);
    if (!d_indices) {
        int *counts = (int *)k0 == (int *)inc ? (int *)k1 : (int *)k0; // any free buffer of >= n_cells ints
        counts = flip ? (int *)k0 : (int *)k1;
        PFEM2_LAUNCH(k_one_ring, grid_for(n_cells, 128, 1 << 30), 128, 0, st, n_cells, d_cells, node_off, inc, counts, nullptr, nullptr, err);
        exclusive_scan_dev<int>(counts, d_offsets, len_dev + 1, 1, 0, n_cells, scratch2, st);
    } else {
        PFEM2_LAUNCH(k_one_ring, grid_for(n_cells, 128, 1 << 30), 128, 0, st, n_cells, d_cells, node_off, inc, nullptr, d_offsets, d_indices, err);
    }
    int herr = 0;
    CU(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(nnz, d_offsets + n_cells, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(count); cudaFree(node_off); cudaFree(n_dev);
    cudaFree(err); cudaFree(hist); cudaFree(scratch); cudaFree(scratch2); cudaFree(len_dev); cudaFree(info);
    CU(cudaGetLastError());
    if (herr) return fail(nullptr, PFEM2_EINVAL, "a cell has more than 96 one-ring neighbours");
    return PFEM2_OK;
}

} // extern "C"
