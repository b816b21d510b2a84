// pfem2_handle.cuh -- the handle behind the C ABI (include/pfem2_b200.h) and the host helpers the translation units of
// libpfem2_b200.so share:
//   pfem2_api.cu        life cycle, memory, the three public calls (advect / project / correct), state exchange with the host
//   pfem2_host_step.cu  pfem2_step_host: the step with HOST nodal buffers, pipelined over PCIe
//   pfem2_multi.cu      strip-partitioned multi-GPU building blocks (migration, halo, P2P transport)
//   pfem2_mesh.cu       mesh preparation helpers and the stand-alone radix sort entry point
#pragma once

#include "../../include/pfem2_b200.h"

#include "pfem2_common.cuh"

#include <string>
#include <vector>

struct pfem2_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string error;
    pfem2_options opt{};

    // mesh (borrowed) + private repack
    pfem2_mesh_view mesh{};
    pfem2::CellGeom *geom = nullptr;
    int *node_off = nullptr;      // n_nodes + 1
    unsigned *node_inc = nullptr; // 3 * n_cells, (3c + i) ascending per node
    int4 *edge_nbr = nullptr;     // n_cells: cells across the three edges (-1 = boundary)
    int level = 1, ppc = 1;
    double sub_step = 1.0;
    double *centers = nullptr; // 3 * ppc
    int key_bits = 1;
    int band = -1;             // max |neighbour - cell| over the one-ring lists (-1: not computed yet)

    // owned cell range [own_lo, own_hi) (the whole mesh on a single GPU): seeding / re-seeding / emigration
    int own_lo = 0, own_hi = 0;
    int *own_len_dev = nullptr;           // device int: own_hi - own_lo (scan length)
    int *node_list = nullptr;             // nodes of the owned cells (nullptr = all nodes), multi-GPU
    int n_node_list = 0;
    int own_node_lo = 0, own_node_hi = 0; // node id range of the owned cells: what a deferred correction can touch
    int v2_node_lo = 0, v2_node_hi = 0;   // node id range of the cells a particle of the owned range can reach in one advect call
    int v2_range_substeps = -1;           // ... computed for this many substeps (-1: not yet)

    // particles: two record buffers.  `cur` is the current one; `permuted` says how to read it (see below)
    int capacity = 0;
    pfem2::ParticleSoA soa[2]{};
    int cur = 0;
    bool seeded = false;
    CUtensorMap tmap[2];                     // [rows x 64 B] view of the two record buffers (32-row boxes, 64-byte swizzle)
    void *tmap_base[2] = {nullptr, nullptr}; // what the maps were encoded for
    int tmap_rows[2] = {0, 0};

    // Lazy re-sort (pfem2_options.lazy_sort, default): after an advect the current buffer is DENSE but in the order of the previous
    // step's cells; vals[perm_buf] maps sorted position -> row (padded to a multiple of 32 with a valid row), keys[1] holds the new
    // cells of the last move pass.  materialize() makes the sorted order physical again for every reader that wants it.
    bool permuted = false;
    int perm_buf = 0;
    bool lazy_move = false;                  // the move pass in flight was the gathered one (advect_finish ranks instead of scattering)
    int *tail_cursor = nullptr;              // device ints: [0] re-seeded records appended behind the dense array, then the tile cursors of the move pass
    int mv_launches = 0;                     // launches of the gathered move pass inside the advect in flight (each has its own tile cursor)
    bool lazy_swizzle = true;                // PFEM2_LAZY_SWIZZLE=0: linear tiles (layout cross-check of the tests)
    bool lazy_nsub3 = true;                  // PFEM2_LAZY_NSUB3=0: runtime-S form of the move pass also for S = 3 (A/B)
    CUtensorMap gmap[2], omap[2];            // gather maps (box {16, 1}) and tile-store maps (box {16, 32}) of the two buffers
    void *lzmap_base[2] = {nullptr, nullptr};
    int lzmap_rows[2] = {0, 0};

    // deferred velocity correction: nodal increment snapshot, folded into the next move pass
    double *dv[2] = {nullptr, nullptr}; // n_nodes each
    double2 *dv2 = nullptr;             // the same increment interleaved (x, y) per node
    double2 *v2 = nullptr;              // nodal velocity of the advect in flight, interleaved (packed per call)
    bool dv_pending = false;

    // per-step scratch
    pfem2::Counters *ctr = nullptr;
    pfem2::Counters *host_ctr = nullptr; // pinned mirror
    cudaEvent_t readback = nullptr;
    bool readback_pending = false;
    int host_count = 0; // last count known on the host
    int host_added = 0;
    unsigned *keys[2]{}, *vals[2]{};         // (new cell, array index) pairs: movers of the stable order / permutation of the lazy re-sort
    unsigned *stay_bits = nullptr;           // capacity / 32 + 2: ballot of particles that stayed in their cell (stable order)
    int *warp_movers = nullptr;              // capacity / 32 + 2: movers per warp, scanned in place (stable order)
    int *warp_scan_scratch = nullptr;
    int *stay = nullptr, *arrive = nullptr, *cursor = nullptr; // n_cells + 1 each (one allocation)
    bool arrive_dirty = false;               // arrive[] may hold counts of a physical re-sort (the lazy path needs it zero and never writes it)
    unsigned long long *cell_mask = nullptr; // n_cells + 1
    unsigned long long *packed = nullptr;    // n_cells + 2 (scan in place)
    unsigned long long *scan_scratch64 = nullptr;
    int *cell_start[2] = {nullptr, nullptr}; // n_cells + 1, ping-pong (old / new segment table)
    int cs = 0;
    int *n_cells_dev = nullptr;              // device copy of n_cells (scan length)
    int *rs_hist = nullptr;
    int *rs_scan_scratch = nullptr;
    int *rs_info = nullptr;
    double *partial = nullptr; // 9 * n_cells: the nine projection sums per cell
    bool partials_valid = false;
    int last_substeps = 1;

    // lazily allocated
    void *aos = nullptr;
    size_t aos_bytes = 0;
    double *nodal[4] = {nullptr, nullptr, nullptr, nullptr}; // F.x F.y W.x W.y for pfem2_step_host
    double *acc3 = nullptr;                                  // 3 * n_nodes node accumulators of pfem2_step_host_p2p

    // multi-GPU (strip partition)
    int cell_base = 0;            // the mesh view is the slice [cell_base, cell_base + n_cells) of the global cell numbering (partitioned mesh)
    int *mg_bounds = nullptr;     // device copy of the rank cell bounds (n_ranks + 1), in this strip's numbering
    int *mg_rank_count = nullptr; // device, per destination rank (+ total)
    int mg_ranks = 0;
    std::vector<int> mg_host_counts;
    std::vector<int> mg_host_bounds;
    bool mg_fused = false;    // the move pass in flight listed its emigrants and counted the per-cell statistics
    int mg_fused_total = 0;   // emigrants listed by that pass
    bool move_pending = false; // advect_move done, advect_finish outstanding
    bool mv_interior_pending = false; // split move pass: the boundary layers are moved, the interior is outstanding (its parameters:)
    double mv_hsub = 0.0;
    int mv_substeps = 0, mv_do_count = 0, mv_bl = 0, mv_br = 0;
    bool mv_dv_pending = false;

    // pfem2_step_host pipeline: the step runs in K chunks of the cell range so that the host <-> device copies of the nodal
    // fields overlap the move pass (upload) and the projection (download)
    struct HostPipe {
        int K = 0, substeps = 0;          // what the plan was made for (0 = none yet)
        std::vector<int> cb, ns;          // cell chunk bounds (K + 1), node slice bounds of the upload (K + 1)
        std::vector<int> up_slice;        // chunk j may start once upload slices 0..up_slice[j] have landed
        std::vector<int> dn_ready;        // after projecting chunk j the nodes [0, dn_ready[j]) are final
        cudaStream_t copy = nullptr;      // non-blocking copy stream
        cudaStream_t mv[2] = {nullptr, nullptr}; // the chunks of the move pass alternate between two streams: the head of chunk j + 1 fills the SMs the tail of chunk j leaves idle
        cudaEvent_t mv_ev[3] = {nullptr, nullptr, nullptr}; // "packs so far done" (main stream), "chunks done" (the two move streams)
        std::vector<cudaEvent_t> up_ev, dn_ev;
        bool active = false;              // a pipelined step is being issued
        int packed_slices = 0;            // upload slices already interleaved into v2
    } pipe;

    // P2P transport of the neighbour protocol (multi-GPU): inboxes in this GPU's memory the neighbours store into, and the
    // neighbours' inboxes mapped through CUDA IPC.  side 0 = left neighbour (rank - 1), side 1 = right neighbour (rank + 1)
    struct P2P {
        int cap = 0;                            // records per migration block (the same on every strip)
        void *inbox[2] = {nullptr, nullptr};    // mine: written by neighbour `side`
        void *peer[2] = {nullptr, nullptr};     // theirs: the inbox neighbour `side` keeps for me (IPC mapping)
        int *idx[2] = {nullptr, nullptr};       // interface node ids shared with neighbour `side` (ascending), device
        int n_idx[2] = {0, 0};
        int idx_lo[2] = {0, 0}, idx_hi[2] = {0, 0}; // node id range [lo, hi) of the interface with neighbour `side`
        int *cursors = nullptr;                 // device: [0], [1] pack cursors per side, [2] records handed over by the last send,
                                                // [3] block counter of the fused kernels ("last block done")
        unsigned mig_seq = 0, halo_seq = 0;     // deliveries made so far (block parity = seq & 1)
    } p2p;

    // CUDA graphs of advectParticles for launch-bound (small) cases: pfem2_api.cu, advect_graphed.  One graph per parity of the buffer
    // ping-pong (cur, cs, perm_buf), replayed while the arguments of the call stay what they were captured for
    struct AdvectGraph {
        cudaGraphExec_t exec = nullptr;
        const void *vx = nullptr, *vy = nullptr, *table = nullptr, *buf0 = nullptr, *buf1 = nullptr;
        double dt = 0.0;
        int substeps = 0, capacity = 0;
        bool dv = false;
        long long kernels = 0; // kernel nodes of the graph (pfem2_kernel_launches counts a replay as these)
    } graphs[8];
    cudaStream_t graph_stream = nullptr; // private capture stream (the legacy default stream cannot be captured)
    bool capturing = false;
    bool graphs_broken = false;          // capture failed once in this process: do not try again
    long long graph_replays = 0;

    // optional per-phase CUDA-event timing (pfem2_set_profiling)
    bool profiling = false;
    struct PhaseRec { int phase; cudaEvent_t a, b; };
    std::vector<PhaseRec> phase_recs;
    std::vector<cudaEvent_t> event_pool;
    double phase_ms[PFEM2_NUM_PHASES] = {0};
    long long phase_calls[PFEM2_NUM_PHASES] = {0};
};

namespace pfem2 {
namespace host {

extern thread_local std::string g_create_error;

#define CU(call)                                                                                                    \
    do {                                                                                                            \
        cudaError_t e_ = (call);                                                                                    \
        if (e_ != cudaSuccess) {                                                                                    \
            char buf_[512];                                                                                         \
            snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            if (h) h->error = buf_; else ::pfem2::host::g_create_error = buf_;                                      \
            return PFEM2_ECUDA;                                                                                     \
        }                                                                                                           \
    } while (0)

int fail(pfem2_handle *h, int code, const char *msg);

// RAII: brackets the launches of one pipeline phase with CUDA events on the handle's stream
struct PhaseScope {
    pfem2_handle *h;
    pfem2_handle::PhaseRec rec;
    PhaseScope(pfem2_handle *h_, int phase);
    ~PhaseScope();
};

// RAII: device temporaries of a set-up function, freed on every exit path
struct DeviceTemps {
    std::vector<void *> ptrs;
    ~DeviceTemps()
    {
        for (void *p : ptrs) cudaFree(p);
    }
    template <class T> cudaError_t alloc(T **p, size_t n)
    {
        *p = nullptr;
        const cudaError_t e = cudaMalloc((void **)p, (n ? n : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

template <class T> int dev_alloc(pfem2_handle *h, T **p, size_t n)
{
    *p = nullptr;
    CU(cudaMalloc((void **)p, (n ? n : 1) * sizeof(T)));
    return PFEM2_OK;
}

// persistent grid-stride kernels: a multiple of the SM count
inline int grid_for(long long n, int threads = kThreads, int max_blocks = 0)
{
    if (max_blocks <= 0) max_blocks = g_num_sms * 16;
    const long long b = (n + threads - 1) / threads;
    return (int)(b < 1 ? 1 : (b < max_blocks ? b : max_blocks));
}

inline NodalVel nodal(const double *x, const double *y, double *const *table)
{
    NodalVel v;
    v.x = x;
    v.y = y;
    v.table = table;
    return v;
}

// ---- shared between the translation units (defined in pfem2_api.cu) ----
int sync_counters(pfem2_handle *h);  // wait for the counter read-back of the last advect (if any), refresh the host-side view
int queue_readback(pfem2_handle *h);
int materialize(pfem2_handle *h);    // lazy re-sort: make the sorted order physical (no-op in the physical state)
int flush_correct(pfem2_handle *h);  // apply a deferred velocity correction now
int mesh_band(pfem2_handle *h);      // band width of the cell numbering (one-time)
bool lazy_enabled(const pfem2_handle *h);
int advect_move(pfem2_handle *h, NodalVel vel, double dt, int substeps, int do_count, bool mg_move, bool split);
int advect_move_interior(pfem2_handle *h);
int advect_finish(pfem2_handle *h, NodalVel vel, int need_count);
void launch_project_cells(pfem2_handle *h, int c_lo = -1, int c_hi = -1);
void launch_project_nodes(pfem2_handle *h, int node_lo, int node_hi, double *vx, double *vy, double *const *table, double *cx, double *cy,
                          double *const *table_copy);
void launch_pack_nodal(pfem2_handle *h, int node_lo, int node_hi, NodalVel vel, bool begin);
void launch_project_nodes_acc_range(pfem2_handle *h, int node_lo, int node_hi, double *acc3);
void launch_project_finalize_range(pfem2_handle *h, int node_lo, int node_hi, const double *acc3, double *vx, double *vy);
int ensure_v2_node_range(pfem2_handle *h, int substeps);
int node_range_of_cells(pfem2_handle *h, int cell_lo, int cell_hi, int &lo, int &hi);

} // namespace host
} // namespace pfem2
