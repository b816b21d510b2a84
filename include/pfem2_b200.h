/*
 * pfem2_b200.h -- C ABI of the B200-native PFEM-2 particle step (libpfem2_b200.so).
 *
 * This is the drop-in boundary for gpuPfem2's `ParticleHandler2D`
 * (reference src/particles/particle_handler_2d.cuh:9-54).  The reference has no FFI: its "operator
 * API" is that C++ class, called by cases/Cylinder2D/main.cu:652-653,766,798,802,864 and
 * cases/PoiseuilleFlow2D/main.cu:513-514,622,657,661,723, and read by DataExport
 * (src/data_export.cu:97-104).  Each entry point below names the member it replaces; the header-
 * compatible C++ shim that forwards to them is gpupfem2_b200/shim/pfem2_particle_handler_2d.cuh
 * (see INTEGRATION.md).
 *
 * Conventions
 *  - plain C: pointers and sizes only, no C++/torch types, no exceptions, never exit().
 *  - every function returns 0 on success or a PFEM2_E* code; pfem2_last_error() gives the text.
 *  - all `d_` pointers are DEVICE pointers on the handle's device; `h_` pointers are host pointers.
 *  - mesh arrays are borrowed for the handle's lifetime and never written.
 *  - all work is issued on the handle's stream (default: the legacy null stream, like the
 *    reference, so ordering with the caller's FEM kernels is preserved).
 *  - nodal velocity arguments come in two flavours: `*_ptrs` takes the reference's
 *    `deviceVector<double*>::data` (a DEVICE array of two device pointers, x then y component,
 *    cases/Cylinder2D/main.cu:597-604); the plain form takes the two component pointers directly.
 */
#ifndef PFEM2_B200_H
#define PFEM2_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFEM2_OK 0
#define PFEM2_EINVAL 1    /* bad argument */
#define PFEM2_ECUDA 2     /* CUDA runtime error (text in pfem2_last_error) */
#define PFEM2_ECAPACITY 3 /* particle storage overflowed its capacity */
#define PFEM2_ESTATE 4    /* call order violated (e.g. advect before seed) */

typedef struct pfem2_handle pfem2_handle;

/* Read-only view of the reference's Mesh2D device arrays (src/mesh_2d.cuh:18-44). */
typedef struct pfem2_mesh_view {
    int n_nodes;               /* getVertices().size */
    int n_cells;               /* getCells().size */
    const double *d_vertices;  /* Point2 = double2 per node      (getVertices().data) */
    const unsigned *d_cells;   /* uint3 per cell, file order     (getCells().data) */
    const double *d_inv_jacobi;/* Matrix2x2 = 4 doubles per cell (getInvJacobi().data) */
    const int *d_nbr_offsets;  /* n_cells + 1                    (getCellNeighborsOffsets().data) */
    const int *d_nbr_indices;  /* ascending one-ring per cell    (getCellNeighborIndices().data) */
} pfem2_mesh_view;

typedef struct pfem2_options {
    int struct_size;        /* = sizeof(pfem2_options), for forward compatibility */
    int subcell_mode;       /* 0: reference-exact flat sub-cell index with spill (default); 1: clamped */
    int max_division_level; /* 0 -> 4 = CONSTANTS::MAX_CELL_DIVISION_LEVEL (constants.h:15); <= 8 */
    double capacity_factor; /* 0 -> 1.5: particle capacity = factor * cells * ppc (reference: 1.1 + realloc) */
    void *stream;           /* cudaStream_t; NULL = legacy default stream */
    int device;             /* CUDA device ordinal; -1 = current device */
    int verbose;            /* 1: print the reference's two stdout lines from the library itself */
    int exact_search;       /* 1: always use the reference's ordered one-ring scan; 0 (default): edge-walk fast path
                               that returns the same cell (falls back to the ordered scan in the tolerance band) */
    int stable_order;       /* 0 (default): fast order, the slot order inside a cell depends on atomic retirement order;
                               1: deterministic order (physical re-sort: stayers keep their order, movers radix-sorted by cell) */
    int reserved_scatter_tma; /* (round-1 A/B kernel variant, removed: must be 0) */
    int defer_correct;      /* 1 (default): correctParticleVelocity snapshots the nodal increment and the particle update is
                               folded into the next advect pass (bit-identical; applied eagerly before any other reader);
                               0: eager kernel */
    int reserved_lane_per_record; /* (round-1 A/B kernel variant, removed: must be 0) */
    int host_pipeline;      /* pfem2_step_host: 0 (default) = 8 chunks on meshes of >= 262144 cells: the upload of the nodal field
                               overlaps the move pass and the download of the projected field overlaps the projection, chunk by
                               chunk of the cell range (dependencies derived from the mesh numbering); 1 = no pipelining; n > 1 =
                               n chunks */
    int reserved_fuse_project; /* (round-1 experiment, measured break-even and removed: must be 0) */
    int lazy_sort;          /* 1 (default): lazy re-sort -- the records move once per step.  advectParticles gathers its tiles through the
                               permutation left by the previous step's counting sort (cp.async.bulk.tensor tile::gather4), writes them
                               dense and ends with a 4-byte rank pass instead of a 128-byte-per-particle scatter; the projection reads
                               through the permutation; every other reader of the physical order (getParticles, download, device
                               records, the eager correction) materialises the sorted array first.  Same results; 14.3 instead of
                               17.7 ms per step on the 16M-triangle channel.  0: physical re-sort in every advect.  Ignored (off)
                               with stable_order */
    int graph_advect;       /* CUDA graph of advectParticles: 0 (default) = on meshes below 262144 cells (launch-bound: the shipped cases
                               run ~10 kernels of a few microseconds per call), 1 = always, -1 = never.  The first call of a buffer
                               parity captures the enqueue sequence, later calls with the same arguments replay it as ONE launch */
} pfem2_options;

/* counters of the last pfem2_advect call (device-resident, read back on demand) */
typedef struct pfem2_stats {
    int count;         /* live particles */
    int lost;          /* particles deleted by the last advect (all substeps) */
    int added;         /* particles re-seeded by the last advect */
    int movers;        /* particle-substeps that left their cell in the last advect */
    int capacity;      /* particle slots allocated */
    int overflow;      /* non-zero if a step needed more slots than capacity (state is then invalid) */
} pfem2_stats;

void pfem2_default_options(pfem2_options *opt);
const char *pfem2_last_error(const pfem2_handle *h); /* h may be NULL: last create() error */
const char *pfem2_version(void);

/* ParticleHandler2D::ParticleHandler2D(const Mesh2D*, int)   particle_handler_2d.cu:238-294 */
int pfem2_create(pfem2_handle **out, const pfem2_mesh_view *mesh, int cell_division_level, const pfem2_options *opt);
/* ParticleHandler2D::~ParticleHandler2D()                     particle_handler_2d.cu:296-302 */
int pfem2_destroy(pfem2_handle *h);

/* seedParticles()                                             particle_handler_2d.cu:304-320 */
int pfem2_seed(pfem2_handle *h);
/* initParticleVelocity(velocitySolution)                      particle_handler_2d.cu:322-326 */
int pfem2_init_velocity(pfem2_handle *h, const double *d_vx, const double *d_vy);
int pfem2_init_velocity_ptrs(pfem2_handle *h, double *const *d_vel2);
/* advectParticles(velocitySolution, timeStep, particleSubsteps)   particle_handler_2d.cu:328-342
 * (S x [advect + own-cell test + one-ring search + delete], then distribution check + re-seed;
 *  afterwards the particle arrays are physically sorted by owning cell) */
int pfem2_advect(pfem2_handle *h, const double *d_vx, const double *d_vy, double dt, int substeps);
int pfem2_advect_ptrs(pfem2_handle *h, double *const *d_vel2, double dt, int substeps);
/* projectVelocityOntoGrid(velocity): WRITES the two nodal arrays   particle_handler_2d.cu:350-361 */
int pfem2_project(pfem2_handle *h, double *d_vx, double *d_vy);
int pfem2_project_ptrs(pfem2_handle *h, double *const *d_vel2);
/* Extension: the same projection, additionally written into a second pair of nodal arrays.  The cases copy the
 * projected field into their "old" solution right after the call (copy_d2d, cases/Cylinder2D/main.cu:804-805,
 * cases/PoiseuilleFlow2D/main.cu:664-665); here the node pass of the projection writes both, so the two copies go away. */
int pfem2_project_dual(pfem2_handle *h, double *d_vx, double *d_vy, double *d_vx_copy, double *d_vy_copy);
int pfem2_project_dual_ptrs(pfem2_handle *h, double *const *d_vel2, double *const *d_vel_copy2);
/* correctParticleVelocity(velocitySolution, velocitySolutionOld)   particle_handler_2d.cu:344-348 */
int pfem2_correct(pfem2_handle *h, const double *d_vx, const double *d_vy, const double *d_vx_old, const double *d_vy_old);
int pfem2_correct_ptrs(pfem2_handle *h, double *const *d_vel2, double *const *d_vel_old2);

/* getParticleCount()  particle_handler_2d.cuh:28-30  (synchronises the handle's stream) */
int pfem2_particle_count(pfem2_handle *h, int *out);
int pfem2_get_stats(pfem2_handle *h, pfem2_stats *out);
/* getParticles()      particle_handler_2d.cuh:24-26: device pointer to the reference's 96-byte AoS
 * Particle2D records (particle_2d.cuh:51-57: ID@0 position@16 localPosition@32 velocity@64 cellID@80),
 * materialised lazily from the SoA storage; valid until the next mutating call. */
int pfem2_export_aos(pfem2_handle *h, const void **d_particles96, int *count);

/* One whole particle step with HOST nodal buffers (pinned or pageable), the end-to-end form of
 *   advect(F) ; project(W) ; correct(F, W)
 * h_f* : frozen/solved nodal velocity in (n_nodes each), h_w* : projected nodal velocity out.
 * Copies F host->device, runs the three calls, copies W device->host and returns the live count. */
int pfem2_step_host(pfem2_handle *h, const double *h_fx, const double *h_fy, double *h_wx, double *h_wy,
                    double dt, int substeps, int *count_out);

/* particle state download / upload (parity tests, checkpoint/restart).  Host SoA arrays of
 * pfem2_particle_count() elements; any pointer may be NULL.  Upload re-sorts by cell. */
int pfem2_download(pfem2_handle *h, double *h_x, double *h_y, double *h_l0, double *h_l1, double *h_l2,
                   double *h_vx, double *h_vy, unsigned *h_cell, unsigned *h_id);
int pfem2_upload(pfem2_handle *h, int n, const double *h_x, const double *h_y, const double *h_l0, const double *h_l1,
                 const double *h_l2, const double *h_vx, const double *h_vy, const unsigned *h_cell, const unsigned *h_id);
/* device pointer to the live particle array: 64-byte records
 *   { double x, y; double L0, L1; double L2; unsigned cell; unsigned id; double vx, vy; }
 * sorted by cell after every advect (valid until the next mutating call) */
int pfem2_device_records(pfem2_handle *h, const void **d_records);
/* per-cell segment table of the sorted storage: particles of cell c are [start[c], start[c+1]) */
int pfem2_cell_starts(pfem2_handle *h, const int **d_cell_start);

/* ---- multi-GPU building blocks (strip partition of the cell index range; no reference counterpart:
 * the reference is single-GPU).  One handle per GPU over the SAME global mesh; each handle owns the cells
 * [cell_lo, cell_hi) and the particles inside them.  A step is
 *     advect_move ; emigrants_count ; emigrants_pack ; <exchange records over NCCL> ; immigrants_append ; advect_finish ;
 *     project_accumulate ; <sum interface-node accumulators over NCCL> ; project_finalize ; correct
 * Records are the library's 64-byte particle records: {x, y | L0, L1 | L2, cell, id | vx, vy}. ---- */
int pfem2_set_owned_cells(pfem2_handle *h, int cell_lo, int cell_hi); /* before pfem2_seed; seeds / re-seeds only these cells */
/* node id ranges of a strip (host-side staging of nodal fields per rank): out4 = { in_lo, in_hi, own_lo, own_hi }.  An advect with
 * `substeps` substeps reads the nodal velocity only at the nodes [in_lo, in_hi) (the cells a particle of the owned range can reach:
 * band width of the one-ring lists x substeps); the projection writes and the correction reads the nodes [own_lo, own_hi) of the
 * owned cells.  The whole node range on a single GPU. */
int pfem2_node_ranges(pfem2_handle *h, int substeps, int *out4);
int pfem2_advect_move(pfem2_handle *h, const double *d_vx, const double *d_vy, double dt, int substeps);
/* per destination rank (cells [h_bounds[r], h_bounds[r+1])) the number of live particles that left the owned range */
int pfem2_emigrants_count(pfem2_handle *h, const int *h_bounds, int n_ranks, int *h_counts);
/* packs them grouped by destination rank in rank order and removes them locally */
int pfem2_emigrants_pack(pfem2_handle *h, void *d_records, long long capacity_records);
int pfem2_immigrants_append(pfem2_handle *h, const void *d_records, int n);
int pfem2_advect_finish(pfem2_handle *h, const double *d_vx, const double *d_vy);
/* Neighbour protocol: the same exchange without any host round trip (strips only exchange with adjacent strips).
 * A migration buffer is  [64-byte header | capacity_records 64-byte records]  in device memory, of a capacity both sides
 * agree on, so that the transfer (ncclSend / ncclRecv of the whole buffer) needs no size negotiation; the number of valid
 * records travels in the header and is only read on the device.  Header: { int count; int flags; int pad[2];
 * unsigned long long spill[4]; int pad[4]; } -- `spill` carries the sub-cell occupancy bits that particles in the tolerance band
 * of this strip's last cells put into the next strip's first cells, so re-seeding decisions match a single GPU.
 *   set_rank_bounds (once, before the first advect_move) ; advect_move ; emigrants_pack_neighbours ; <send / recv the buffers> ;
 *   immigrants_append_device (per received buffer) ; advect_finish
 * Requires the default kernels (fast order, TMA-tiled move pass): otherwise emigrants_pack_neighbours returns PFEM2_ESTATE and
 * the caller uses emigrants_count / emigrants_pack.  An overflowing buffer or an emigrant bound for a non-adjacent strip sets
 * the sticky overflow flag: the next call that synchronises the counters fails with PFEM2_ECAPACITY. */
int pfem2_set_rank_bounds(pfem2_handle *h, const int *h_bounds, int n_ranks);
/* Partitioned mesh: the handle's mesh view holds only the slice [cell_offset, cell_offset + n_cells) of a GLOBAL cell numbering (the
 * strip's own cells plus the halo cells a particle can reach in one advect: pfem2_mesh_band x substeps on either side), with node
 * ids relative to the slice's first node.  Everything inside the handle works in the slice's numbering; the rank bounds stay
 * global and particle records that cross strips carry global cell ids (translated by the pack / append kernels).  Call before
 * pfem2_set_rank_bounds and pfem2_seed.  0 (default): the view is the global mesh. */
int pfem2_set_global_cell_offset(pfem2_handle *h, int cell_offset);
/* d_left / d_right: migration buffers for rank - 1 / rank + 1 (NULL where there is no such strip) */
int pfem2_emigrants_pack_neighbours(pfem2_handle *h, int rank, void *d_left, void *d_right, int capacity_records);
/* from_left != 0: the buffer came from rank - 1 (its spill words are OR-ed into this strip's first cells) */
int pfem2_immigrants_append_device(pfem2_handle *h, const void *d_buffer, int capacity_records, int from_left);
/* P2P transport of the neighbour protocol: NVLink peer memory instead of ncclSend / ncclRecv, no host in the loop.
 * Every strip creates one inbox per neighbour in its own HBM (side 0: data from rank - 1, side 1: data from rank + 1) and
 * hands the 64-byte CUDA IPC handle to that neighbour (any out-of-band channel: the Python layer uses all_gather_object);
 * the neighbour maps it with pfem2_p2p_connect.  From then on
 *   emigrants_send_p2p      the pack kernel stores the emigrants' records straight into the neighbours' inboxes, a publish
 *                           kernel writes the header, fences (system scope) and releases a sequence number;
 *   immigrants_recv_p2p     this strip's stream waits for both neighbours' sequence numbers (device-side acquire spin with a
 *                           20 s watchdog (PFEM2_P2P_TIMEOUT_S overrides): a dead peer raises an error instead of hanging the GPU) and appends from its own memory;
 *   project_halo_p2p        between project_accumulate and project_finalize: interface-node accumulators are stored into the
 *                           neighbours' inboxes, awaited and added (a + b == b + a: both strips get the same bits).
 * h_interface_nodes: ascending ids of the nodes shared with that neighbour (both strips must pass the same list).  All
 * strips use the same capacity_records.  Requirements as for emigrants_pack_neighbours. */
int pfem2_p2p_inbox_create(pfem2_handle *h, int side, int capacity_records, int n_interface_nodes, const int *h_interface_nodes,
                           void *ipc_handle_out /* 64 bytes */);
/* side 0: ipc_handle = the inbox rank - 1 created with side 1;  side 1: the inbox rank + 1 created with side 0 */
int pfem2_p2p_connect(pfem2_handle *h, int side, const void *ipc_handle);
int pfem2_emigrants_send_p2p(pfem2_handle *h, int rank);
int pfem2_immigrants_recv_p2p(pfem2_handle *h);
int pfem2_project_halo_p2p(pfem2_handle *h, double *d_acc3);
int pfem2_p2p_last_sent(pfem2_handle *h, int *out); /* records handed over by the last emigrants_send_p2p (synchronises) */
/* The two halves of a strip's step as single calls (what DistributedParticleHandler2D uses with the P2P transport; results
 * identical to the piecewise sequence above):
 *   pfem2_advect_p2p   advect_move ; emigrants_send_p2p ; immigrants_recv_p2p ; advect_finish
 *   pfem2_project_p2p  cell pass ; node sums ; project_halo_p2p ; division.  d_acc3: 3 x n_nodes doubles of scratch, zero-initialised
 *                      once by the caller (nodes no owned cell touches are never written).
 * PFEM2_P2P_SPLIT=1 selects a form that moves / reduces the cells next to the strip boundaries first and hides the exchange behind
 * the interior (measured slower on B200: more, smaller launches; kept for A/B). */
int pfem2_advect_p2p(pfem2_handle *h, int rank, const double *d_vx, const double *d_vy, double dt, int substeps);
int pfem2_project_p2p(pfem2_handle *h, double *d_acc3, double *d_vx, double *d_vy);
/* The whole step of ONE STRIP with host nodal buffers, driven from C (the multi-GPU form of pfem2_step_host; every rank calls it in
 * the same step):  upload of the node slice the strip's advect reads (in chunks under the move pass) ; advect_move ;
 * emigrants_send_p2p ; immigrants_recv_p2p ; advect_finish ; projection in chunks of the own cell range (interior nodes are divided
 * and downloaded under the cell pass) ; project_halo_p2p ; interface nodes divided and downloaded ; correct.
 * h_f* / h_w*: host arrays over the handle's node numbering (n_nodes of its mesh view); read / written only inside the ranges
 * pfem2_node_ranges reports.  Needs connected P2P inboxes (pfem2_p2p_connect) and pfem2_set_rank_bounds. */
int pfem2_step_host_p2p(pfem2_handle *h, int rank, const double *h_fx, const double *h_fy, double *h_wx, double *h_wy, double dt,
                        int substeps, int *count_out);
/* d_acc3: n_nodes x {sum L v_x, sum L v_y, sum L} of the particles this handle holds (no division) */
int pfem2_project_accumulate(pfem2_handle *h, double *d_acc3);
int pfem2_project_finalize(pfem2_handle *h, const double *d_acc3, double *d_vx, double *d_vy);

/* ---- mesh preparation helpers (the step right before the path; reference src/mesh_2d.cu) ---- */
/* kCalculateInvJacobi, mesh_2d.cu:21-34, same operation order -> same bits */
int pfem2_mesh_inv_jacobi(int n_cells, const double *d_vertices, const unsigned *d_cells, double *d_inv_jacobi, void *stream);
/* vertex-sharing one-ring CSR, ascending (Mesh2D::fillCellNeighborIndices, mesh_2d.cu:107-139), built on
 * the device in O(C).  Call once with d_indices == NULL to get offsets (and *nnz), then again with
 * d_indices of *nnz ints. */
int pfem2_mesh_one_ring(int n_nodes, int n_cells, const unsigned *d_cells, int *d_offsets, int *d_indices, int *nnz,
                        void *stream);

/* band width of a cell numbering: max |neighbour - cell| over the one-ring lists.  A particle's cell index changes by at most this
 * much per substep, which bounds the halo a strip needs (band x substeps cells on either side of its range). */
int pfem2_mesh_band(int n_cells, const int *d_nbr_offsets, const int *d_nbr_indices, int *band, void *stream);

/* ---- stand-alone device radix sort of (cell key, particle index) pairs (exposed for tests) ---- */
int pfem2_sort_pairs(int n, int key_bits, unsigned *d_keys, unsigned *d_vals, unsigned *d_keys_tmp, unsigned *d_vals_tmp,
                     int *result_in_tmp, void *stream);

/* ---- per-phase device timing (bench.py roofline): CUDA events on the handle's stream around each phase ---- */
#define PFEM2_PHASE_ADVECT 0        /* k_advect_locate: S substeps of advect + own-cell test + one-ring search */
#define PFEM2_PHASE_SORT 1          /* radix sort of (cell key, index) pairs */
#define PFEM2_PHASE_REORDER 2       /* distribution-check plan + scan + gather into cell order + re-seed */
#define PFEM2_PHASE_PROJECT_CELLS 3 /* per-cell segmented reduction */
#define PFEM2_PHASE_PROJECT_NODES 4 /* per-node gather + division */
#define PFEM2_PHASE_CORRECT 5       /* velocity correction */
#define PFEM2_NUM_PHASES 6
int pfem2_set_profiling(pfem2_handle *h, int enabled);
/* accumulated milliseconds and launch-group counts per phase since the last reset (synchronises the stream) */
int pfem2_get_phase_times(pfem2_handle *h, double *ms, long long *calls, int reset);

/* launch accounting: kernels launched by this library since process start (bench.py gpu_launches) */
long long pfem2_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* PFEM2_B200_H */
