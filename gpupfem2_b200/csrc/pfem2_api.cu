// pfem2_api.cu -- host side of libpfem2_b200.so: the handle's life cycle, memory management and the three public calls
// of the C ABI declared in include/pfem2_b200.h.  Replaces the host methods of the reference's ParticleHandler2D
// (src/particles/particle_handler_2d.cu:238-423); see DESIGN.md for the pipeline.
#include "pfem2_handle.cuh"

#include "pfem2_move.cuh"
#include "pfem2_project.cuh"
#include "pfem2_resort.cuh"
#include "pfem2_setup.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace pfem2 {
std::atomic<long long> g_kernel_launches{0};
int g_num_sms = 148;
namespace host {
thread_local std::string g_create_error;

int fail(pfem2_handle *h, int code, const char *msg)
{
    if (h) h->error = msg; else g_create_error = msg;
    return code;
}

static cudaEvent_t take_event(pfem2_handle *h)
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

PhaseScope::PhaseScope(pfem2_handle *h_, int phase) : h(h_)
{
    if (!h->profiling) return;
    rec.phase = phase;
    rec.a = take_event(h);
    rec.b = take_event(h);
    cudaEventRecord(rec.a, h->stream);
}
PhaseScope::~PhaseScope()
{
    if (!h->profiling) return;
    cudaEventRecord(rec.b, h->stream);
    h->phase_recs.push_back(rec);
}

// ------------------------------------------------------------------------------------------------
// particle storage
// ------------------------------------------------------------------------------------------------
static int alloc_soa(pfem2_handle *h, ParticleSoA &s, int cap)
{
    ParticleRec *r = nullptr;
    const int rc = dev_alloc(h, &r, cap);
    if (rc) return rc;
    s.bind(r);
    return PFEM2_OK;
}

static void free_soa(ParticleSoA &s)
{
    cudaFree(s.records());
    s = ParticleSoA{};
}

static void free_particle_scratch(pfem2_handle *h)
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

static int alloc_particle_scratch(pfem2_handle *h, int cap)
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

static int alloc_particle_storage(pfem2_handle *h, int cap)
{
    int rc;
    for (int k = 0; k < 2; ++k)
        if ((rc = alloc_soa(h, h->soa[k], cap))) return rc;
    // the incidence sort at create() reuses keys/vals, so they hold at least 3 * n_cells entries (ensured by caller)
    if ((rc = alloc_particle_scratch(h, cap))) return rc;
    h->capacity = cap;
    return PFEM2_OK;
}

// grow particle storage to new_cap, preserving the current buffer's live prefix (physical state only: callers materialize first)
static int grow(pfem2_handle *h, int new_cap)
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
    if (h->capturing) // (an event-record NODE: the host may wait on the event after every replay of the graph)
        CU(cudaEventRecordWithFlags(h->readback, h->stream, cudaEventRecordExternal));
    else
        CU(cudaEventRecord(h->readback, h->stream));
    h->readback_pending = true;
    return PFEM2_OK;
}

// ------------------------------------------------------------------------------------------------
// mesh-derived ranges
// ------------------------------------------------------------------------------------------------
// band width of the cell numbering (one-time): a particle's cell index changes by at most this much per substep
int mesh_band(pfem2_handle *h)
{
    if (h->band >= 0) return PFEM2_OK;
    DeviceTemps tmp;
    int *dev = nullptr;
    CU(tmp.alloc(&dev, 1));
    CU(cudaMemsetAsync(dev, 0, sizeof(int), h->stream));
    const int C = h->mesh.n_cells;
    PFEM2_LAUNCH(k_band_width, grid_for(C, kThreads, 1 << 30), kThreads, 0, h->stream, C, h->mesh.d_nbr_offsets, h->mesh.d_nbr_indices, dev);
    int band = 0;
    CU(cudaMemcpyAsync(&band, dev, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->band = band;
    return PFEM2_OK;
}

// node id range [lo, hi) of the cells [cell_lo, cell_hi) (one small kernel + an 8-byte read-back; multi-GPU set-up only)
int node_range_of_cells(pfem2_handle *h, int cell_lo, int cell_hi, int &lo, int &hi)
{
    lo = hi = 0;
    if (cell_hi <= cell_lo) return PFEM2_OK;
    DeviceTemps tmp;
    int *dev = nullptr;
    const int init[2] = {0x7fffffff, -1};
    int out[2] = {0, 0};
    CU(tmp.alloc(&dev, 2));
    CU(cudaMemcpyAsync(dev, init, sizeof init, cudaMemcpyHostToDevice, h->stream));
    PFEM2_LAUNCH(k_node_minmax, grid_for(cell_hi - cell_lo, kThreads, 1 << 30), kThreads, 0, h->stream, cell_lo, cell_hi, h->geom, dev);
    CU(cudaMemcpyAsync(out, dev, sizeof out, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
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

// ------------------------------------------------------------------------------------------------
// tensor maps of a record buffer.  cuTensorMapEncodeTiled is a driver entry point; it is resolved through the runtime so
// that the library carries no link-time dependency on libcuda.
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int tensor_map_encoder(pfem2_handle *h, EncodeTiledFn &encode)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(h, PFEM2_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    encode = (EncodeTiledFn)fn;
    return PFEM2_OK;
}

// [rows x 64 B] tensor over buffer k: `box_rows` rows per box (32 = one warp tile, 1 = the box of tile::gather4)
static int encode_record_map(pfem2_handle *h, CUtensorMap *out, int k, unsigned box_rows, bool swizzle)
{
    EncodeTiledFn encode = nullptr;
    const int rc = tensor_map_encoder(h, encode);
    if (rc) return rc;
    const cuuint64_t dims[2] = {16, (cuuint64_t)h->capacity}; // int32 elements per record, records
    const cuuint64_t strides[1] = {sizeof(ParticleRec)};     // bytes between records
    const cuuint32_t box[2] = {16, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, h->soa[k].records(), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return fail(h, PFEM2_ECUDA, buf);
    }
    return PFEM2_OK;
}

// in-place move pass: 32-row boxes, 64-byte swizzle
static int record_tensor_map(pfem2_handle *h, int k)
{
    void *base = h->soa[k].records();
    if (h->tmap_base[k] == base && h->tmap_rows[k] == h->capacity) return PFEM2_OK;
    const int rc = encode_record_map(h, &h->tmap[k], k, 32, true);
    if (rc) return rc;
    h->tmap_base[k] = base;
    h->tmap_rows[k] = h->capacity;
    return PFEM2_OK;
}

// gathered move pass: the same tensor once with box {16, 1} (tile::gather4 takes four row indices) and once with box {16, 32}
// (the dense tile store), both with the handle's swizzle mode
static int lazy_record_maps(pfem2_handle *h, int k)
{
    void *base = h->soa[k].records();
    if (h->lzmap_base[k] == base && h->lzmap_rows[k] == h->capacity) return PFEM2_OK;
    int rc;
    if ((rc = encode_record_map(h, &h->gmap[k], k, 1, h->lazy_swizzle))) return rc;
    if ((rc = encode_record_map(h, &h->omap[k], k, 32, h->lazy_swizzle))) return rc;
    h->lzmap_base[k] = base;
    h->lzmap_rows[k] = h->capacity;
    return PFEM2_OK;
}

// ------------------------------------------------------------------------------------------------
// lazy re-sort state
// ------------------------------------------------------------------------------------------------
bool lazy_enabled(const pfem2_handle *h) { return h->opt.lazy_sort != 0 && h->opt.stable_order == 0; }

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

// ------------------------------------------------------------------------------------------------
// advectParticles, first half: S x (advect + locate) fused into one pass
// ------------------------------------------------------------------------------------------------
// interleave the nodal velocity; begin = true: the launch also opens the advect (counters, emigrant counters, re-seed cursor)
void launch_pack_nodal(pfem2_handle *h, int node_lo, int node_hi, NodalVel vel, bool begin)
{
    if (node_hi <= node_lo && !begin) return;
    PFEM2_LAUNCH(k_pack_nodal, grid_for(std::max(node_hi - node_lo, 1), kThreads, 1 << 30), kThreads, 0, h->stream, node_lo, node_hi, vel, h->v2,
                 begin ? h->ctr : (Counters *)nullptr, h->capacity, h->mg_rank_count, h->mg_fused ? h->mg_ranks + 1 : 0, h->tail_cursor);
}

// one launch of the move pass over the cells [c_lo, c_hi) of the segment table `cstart` (nullptr: everybody)
// src = the buffer the pass reads (h->cur, or h->cur ^ 1 for the interior part of a split pass after the flip)
static void launch_move(pfem2_handle *h, bool lazy, double hsub, int substeps, int do_count, int grid, const int *cstart, int c_lo, int c_hi,
                        int src = -1, int part = 0)
{
    if (src < 0) src = h->cur;
    const int C = h->mesh.n_cells;
    const size_t smem = advect_tma_smem_bytes(kAdvThreads);
    const bool walk = h->opt.exact_search == 0;
    const int mode = h->opt.subcell_mode ? 1 : 0;
    const double2 *dv2 = h->dv_pending ? h->dv2 : nullptr;
    unsigned *emig = h->mg_fused ? h->keys[0] : nullptr;
    cudaStream_t st = h->stream;
    if (lazy) {
        // tiles per claim of the global tile cursor: groups when every warp of the grid gets several of them, single tiles otherwise (a
        // small pass -- the shipped meshes, one chunk of a small strip -- would leave most warps without a group)
        const long long tiles = ((long long)h->host_count + 31) / 32 / std::max(1, h->pipe.active ? h->pipe.K : 1);
        constexpr int kGroup = PFEM2_MOVE_GDYN > 0 ? PFEM2_MOVE_GDYN : 1;
        const bool groups = part != 1 && part != 3 && tiles > (long long)grid * (kAdvThreads / 32) * kGroup;
#define PFEM2_MOVE_GATHER(W, NSUB, SWZ, CLAIM)                                                                                            \
    PFEM2_LAUNCH((k_move_gather<W, NSUB, SWZ, CLAIM>), grid, kAdvThreads, smem, st, h->gmap[src], h->omap[src ^ 1],                         \
                 (const int4 *)h->vals[h->perm_buf], h->keys[1], h->geom, h->edge_nbr, h->mesh.d_nbr_offsets, h->mesh.d_nbr_indices, h->v2, \
                 hsub, substeps, mode, C, h->ppc, h->level, h->sub_step, h->ctr, h->stay, h->cell_mask, dv2, h->own_lo, h->own_hi,         \
                 h->mg_bounds, h->mg_ranks, h->mg_rank_count, emig, cstart, c_lo, c_hi, part,                                              \
                 h->tail_cursor + kTileCursor0 + 32 * (h->mv_launches++ & 31))
#define PFEM2_MOVE_GATHER_C(W, CLAIM)                                                                                                     \
    do {                                                                                                                                  \
        if (!h->lazy_swizzle) PFEM2_MOVE_GATHER(W, 0, false, CLAIM);                                                                      \
        else if (substeps == 3 && h->lazy_nsub3) PFEM2_MOVE_GATHER(W, 3, true, CLAIM);                                                    \
        else PFEM2_MOVE_GATHER(W, 0, true, CLAIM);                                                                                        \
    } while (0)
#define PFEM2_MOVE_GATHER_W(W)                                                                                                            \
    do {                                                                                                                                  \
        if (groups) PFEM2_MOVE_GATHER_C(W, kGroup);                                                                                       \
        else PFEM2_MOVE_GATHER_C(W, 1);                                                                                                   \
    } while (0)
        if (walk) PFEM2_MOVE_GATHER_W(true);
        else PFEM2_MOVE_GATHER_W(false);
#undef PFEM2_MOVE_GATHER_C
#undef PFEM2_MOVE_GATHER_W
#undef PFEM2_MOVE_GATHER
        return;
    }
    const bool stable = h->opt.stable_order != 0;
    unsigned *sb = stable ? h->stay_bits : nullptr;
#define PFEM2_MOVE_TILES(W, NSUB, FAST)                                                                                                   \
    PFEM2_LAUNCH((k_move_tiles<W, NSUB, FAST>), grid, kAdvThreads, smem, st, h->tmap[src], h->geom, h->edge_nbr, h->mesh.d_nbr_offsets,  \
                 h->mesh.d_nbr_indices, h->v2, hsub, substeps, mode, C, h->ppc, h->level, h->sub_step, h->ctr, sb, h->warp_movers, h->stay, \
                 stable ? h->arrive : (int *)nullptr, h->cell_mask, do_count, dv2, h->own_lo, h->own_hi, h->mg_bounds, h->mg_ranks,        \
                 h->mg_rank_count, emig, cstart, c_lo, c_hi)
    // FAST is a register-allocation matter (the start cell is not carried through the substep loop): taken where ptxas spills less
    if (walk && !stable) {
        if (substeps == 3) PFEM2_MOVE_TILES(true, 3, true);
        else PFEM2_MOVE_TILES(true, 0, true);
    } else if (walk) {
        if (substeps == 3) PFEM2_MOVE_TILES(true, 3, false);
        else PFEM2_MOVE_TILES(true, 0, false);
    } else {
        if (substeps == 3) PFEM2_MOVE_TILES(false, 3, false);
        else PFEM2_MOVE_TILES(false, 0, false);
    }
#undef PFEM2_MOVE_TILES
}

// split (strips, lazy + fused only): this call moves only the cells within reach of the strip boundaries -- every emigrant comes from
// there --, so that the caller can send them and move the interior (advect_move_interior) while the delivery travels
// Rows an advect may need beyond the live count: the re-seeds of the step (extrapolated from the last one), in a strip the
// immigrants of both inboxes, and 3 % for good measure.  (The dense array of the lazy re-sort needs exactly count + immigrants +
// re-seeds rows: the particles lost in the pass sit inside the first `count` rows.)
static long long capacity_margin(const pfem2_handle *h)
{
    return std::max<long long>({(long long)h->host_count / 32, 4ll * h->host_added, 4096ll}) + 2ll * h->p2p.cap;
}

int advect_move(pfem2_handle *h, NodalVel vel, double dt, int substeps, int do_count, bool mg_move, bool split)
{
    if (!h) return PFEM2_EINVAL;
    if (!h->seeded) return fail(h, PFEM2_ESTATE, "advect before seed");
    if (h->move_pending) return fail(h, PFEM2_ESTATE, "advect_move called twice without advect_finish");
    if (substeps < 1) return fail(h, PFEM2_EINVAL, "particleSubsteps must be >= 1");
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = sync_counters(h))) return rc;
    const bool stable = h->opt.stable_order != 0;
    // multi-GPU, once the rank bounds are known (pfem2_set_rank_bounds / the first emigrants_count call): the move pass lists the
    // emigrants and counts the per-cell statistics of everybody else itself
    const bool fused = mg_move && !stable && h->mg_ranks > 0 && !(getenv("PFEM2_MG_FUSED") && atoi(getenv("PFEM2_MG_FUSED")) == 0);
    // the gathered (lazy) pass needs the emigrant list in the strip-partitioned step: the unfused protocol searches the physical array
    const bool lazy = lazy_enabled(h) && (!mg_move || fused);
    if (!lazy && (rc = materialize(h))) return rc;
    {   // capacity policy: keep room for the growth seen so far (re-seeding only ever adds)
        const long long margin = capacity_margin(h);
        if ((long long)h->host_count + margin > h->capacity) {
            const long long want = std::max<long long>((long long)(1.25 * h->host_count), (long long)h->host_count + 2 * margin);
            // (sorted positions are ints; the move pass looks up to a few tile claims of every warp beyond the count: 2^24 of headroom)
            if (want > 2147483647ll - (1ll << 24)) return fail(h, PFEM2_ECAPACITY, "particle count exceeds 32-bit indexing");
            if ((rc = materialize(h))) return rc;
            if ((rc = grow(h, (int)want))) return rc;
        }
    }
    cudaStream_t st = h->stream;
    const int C = h->mesh.n_cells;
    h->partials_valid = false;
    h->last_substeps = substeps;
    const double hsub = dt / substeps; // particle_handler_2d.cu:330, host double
    {   // per-cell scratch of the owned range (+ a few cells for the tolerance-band spill of the occupancy bits).  arrive[] is only
        // written by the stable order (it stays zero otherwise) and cursor[] is initialised by the plan's scan (PlanEpilogue)
        const size_t lo = (size_t)h->own_lo, len = (size_t)std::min(C, h->own_hi + 4) - lo + 1;
        CU(cudaMemsetAsync(h->stay + lo, 0, sizeof(int) * len, st));
        if (!lazy || h->arrive_dirty) CU(cudaMemsetAsync(h->arrive + lo, 0, sizeof(int) * len, st));
        h->arrive_dirty = !lazy; // (the physical paths count arrivals separately)
        CU(cudaMemsetAsync(h->cell_mask + lo, 0, sizeof(unsigned long long) * len, st));
    }
    h->mg_fused = fused;
    h->mv_launches = 0; // (tile cursors of the gathered pass: zeroed by the launch that opens the advect, one per move launch)
    if (fused) do_count = 1;
    if (!h->v2) CU(cudaMalloc((void **)&h->v2, sizeof(double2) * (size_t)h->mesh.n_nodes));
    if ((rc = ensure_v2_node_range(h, substeps))) return rc;
    if (lazy) {
        if (!h->tail_cursor) CU(cudaMalloc((void **)&h->tail_cursor, sizeof(int) * (kTileCursor0 + kTileCursors)));
        if (!h->permuted) { // physically sorted (seed, upload, materialize): the identity permutation
            const int padded = (h->host_count + 31) & ~31;
            PFEM2_LAUNCH(k_iota, grid_for(padded), kThreads, 0, st, h->vals[h->perm_buf], h->ctr, padded);
        }
        if ((rc = lazy_record_maps(h, h->cur))) return rc;
        if ((rc = lazy_record_maps(h, h->cur ^ 1))) return rc;
    } else {
        if ((rc = record_tensor_map(h, h->cur))) return rc;
    }
    {
        PhaseScope ps(h, PFEM2_PHASE_ADVECT);
        h->mv_interior_pending = false;
        if (!h->pipe.active && split && lazy && fused && h->band >= 0) {
            launch_pack_nodal(h, h->v2_node_lo, h->v2_node_hi, vel, true);
            const long long reach = (long long)h->band * substeps;
            const int bl = (int)std::min<long long>(h->own_lo + reach, h->own_hi), br = (int)std::max<long long>(h->own_hi - reach, bl);
            const int *cstart = h->cell_start[h->cs];
            const int gb = grid_for((long long)(reach + 1) * h->ppc * 2, kAdvThreads, g_num_sms * kAdvBlocksPerSM);
            launch_move(h, lazy, hsub, substeps, do_count, gb, cstart, bl, br, -1, 1); // the cells next to the left strip boundary ...
            launch_move(h, lazy, hsub, substeps, do_count, gb, cstart, bl, br, -1, 3); // ... and to the right one
            h->mv_interior_pending = true;
            h->mv_hsub = hsub; h->mv_substeps = substeps; h->mv_do_count = do_count; h->mv_bl = bl; h->mv_br = br;
            h->mv_dv_pending = h->dv_pending; // the interior particles get the same deferred correction
        } else if (!h->pipe.active) {
            launch_pack_nodal(h, h->v2_node_lo, h->v2_node_hi, vel, true); // all nodes on a single GPU; a strip's reach otherwise
            const int grid = grid_for(h->capacity, kAdvThreads, g_num_sms * kAdvBlocksPerSM); // persistent: all resident blocks
            launch_move(h, lazy, hsub, substeps, do_count, grid, nullptr, 0, C);
        } else {
            // pfem2_step_host: chunk j of the cell range starts as soon as the slices of the nodal field it can touch have
            // landed (events recorded on the copy stream) and have been interleaved into v2.  The gathered pass moves whole tiles
            // of the sorted order per chunk (k_move_gather)
            // The nodal slices are interleaved on the main stream as they land; the move chunks alternate between two helper streams
            // (each behind the packs it needs), so that the first blocks of chunk j + 1 run on the SMs the last blocks of chunk j
            // have left -- on one stream every chunk boundary cost the tail of one launch plus the head of the next (~30 us, 7
            // boundaries per step).  The chunks touch disjoint tiles, their counters are atomics, each has its own tile cursor.
            pfem2_handle::HostPipe &pp = h->pipe;
            const int grid = grid_for((long long)h->capacity / pp.K + 1, kAdvThreads, g_num_sms * kAdvBlocksPerSM);
            static const bool two_streams = !(getenv("PFEM2_PIPE_STREAMS") && atoi(getenv("PFEM2_PIPE_STREAMS")) == 1); // A/B switch
            if (two_streams && !pp.mv[0]) {
                for (cudaStream_t &m : pp.mv) CU(cudaStreamCreateWithFlags(&m, cudaStreamNonBlocking));
                for (cudaEvent_t &e : pp.mv_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            }
            for (int j = 0; j < pp.K; ++j) {
                for (; pp.packed_slices <= pp.up_slice[j]; ++pp.packed_slices) {
                    cudaStreamWaitEvent(st, pp.up_ev[pp.packed_slices], 0);
                    launch_pack_nodal(h, pp.ns[pp.packed_slices], pp.ns[pp.packed_slices + 1], vel, pp.packed_slices == 0);
                }
                if (two_streams) {
                    cudaStream_t m = pp.mv[j & 1];
                    CU(cudaEventRecord(pp.mv_ev[0], st)); // everything issued on the main stream so far, the packs of this chunk included
                    CU(cudaStreamWaitEvent(m, pp.mv_ev[0], 0));
                    h->stream = m;
                    launch_move(h, lazy, hsub, substeps, do_count, grid, h->cell_start[h->cs], pp.cb[j], pp.cb[j + 1]);
                    h->stream = st;
                } else {
                    launch_move(h, lazy, hsub, substeps, do_count, grid, h->cell_start[h->cs], pp.cb[j], pp.cb[j + 1]);
                }
            }
            if (two_streams)
                for (int k = 0; k < 2; ++k) { // join: the rest of the step follows on the main stream
                    CU(cudaEventRecord(pp.mv_ev[1 + k], pp.mv[k]));
                    CU(cudaStreamWaitEvent(st, pp.mv_ev[1 + k], 0));
                }
            for (; pp.packed_slices < pp.K; ++pp.packed_slices) // (not reached: the last chunk needs every slice)
                cudaStreamWaitEvent(st, pp.up_ev[pp.packed_slices], 0);
        }
    }
    CU(cudaGetLastError());
    if (lazy) h->cur ^= 1; // the dense output is the current buffer now (order of the previous step's cells, lost particles included)
    h->lazy_move = lazy;
    h->dv_pending = false; // the move pass applied the deferred correction
    h->move_pending = true;
    return PFEM2_OK;
}

// second part of a split move pass: the cells between the two boundary layers (reads the buffer the first part read: cur ^ 1 now)
int advect_move_interior(pfem2_handle *h)
{
    if (!h->mv_interior_pending) return PFEM2_OK;
    h->mv_interior_pending = false;
    CU(cudaSetDevice(h->device));
    PhaseScope ps(h, PFEM2_PHASE_ADVECT);
    const int grid = grid_for(h->capacity, kAdvThreads, g_num_sms * kAdvBlocksPerSM);
    const bool keep = h->dv_pending;
    h->dv_pending = h->mv_dv_pending; // (launch_move hands dv2 to the kernel iff a correction is pending)
    launch_move(h, true, h->mv_hsub, h->mv_substeps, h->mv_do_count, grid, h->cell_start[h->cs], h->mv_bl, h->mv_br, h->cur ^ 1, 2);
    h->dv_pending = keep;
    CU(cudaGetLastError());
    return PFEM2_OK;
}

// ------------------------------------------------------------------------------------------------
// advectParticles, second half: distribution check (plan) + re-sort by owning cell + re-seed
// ------------------------------------------------------------------------------------------------
// plan: packed[c] = survivors + missing of cell c; scan -> segment starts; count / overflow
// with_cursor: the cursors of the counting sort / rank pass are initialised too (by the scan's epilogue, which also closes the plan)
static void launch_plan(pfem2_handle *h, bool reseed, bool with_cursor)
{
    cudaStream_t st = h->stream;
    const int C = h->mesh.n_cells, lo = h->own_lo, hi = h->own_hi, own_n = hi - lo; // cell-wise work only over the owned range
    PFEM2_LAUNCH(k_plan_cells, grid_for(own_n, kThreads, 1 << 30), kThreads, 0, st, C, lo, hi, h->ppc, reseed ? 1 : 0, h->stay, h->arrive,
                 h->cell_mask, h->packed, h->ctr);
    // (the scan's down-sweep also writes the cursors of the counting sort / rank pass and closes the plan: PlanEpilogue)
    exclusive_scan_dev<unsigned long long, PlanEpilogue>(h->packed + lo, h->packed + lo, h->own_len_dev, 1, 0, own_n, h->scan_scratch64, st,
                                                         PlanEpilogue{with_cursor ? h->cursor + lo : (int *)nullptr, h->ctr});
}

// Physical re-sort into the other buffer.  stable: stayers keep their relative order, the movers listed in keys[0]/vals[0]
// (n = ctr->n_movers, array order) are radix-sorted by new cell and appended behind the stayers of their cell; otherwise one
// counting-sort scatter of everybody.  Lost particles are dropped and (optionally) every empty sub-cell is re-seeded.
static int reorder(pfem2_handle *h, bool reseed, bool have_stayers, bool stable, NodalVel vel)
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
    const int lo = h->own_lo, hi = h->own_hi, own_n = hi - lo;
    launch_plan(h, reseed, !stable);
    ParticleSoA src = h->soa[h->cur], dst = h->soa[h->cur ^ 1];
    if (stable) {
        if (have_stayers)
            PFEM2_LAUNCH(k_scatter_stayers, grid_for(h->capacity), kThreads, 0, st, src, dst, &h->ctr->n_old, h->stay_bits,
                         h->cell_start[h->cs], h->packed, h->ctr);
        PFEM2_LAUNCH(k_scatter_movers, grid_for(h->capacity), kThreads, 0, st, src, dst, C, &h->ctr->n_movers, h->keys[flip],
                     h->vals[flip], h->stay, h->packed, h->ctr);
    } else {
        // 12 eight-record groups in flight per warp, 2 blocks per SM (measured on channel16m: U x blocks = 8x3 6.40 ms,
        // 12x2 6.16, 16x2 6.40, 20x2 7.4, 24x1 7.3, 8x4 6.7, 4x6 6.7 for the whole reorder phase)
        PFEM2_LAUNCH((k_scatter_all_quads<12, 2>), grid_for(h->capacity), kThreads, 0, st, src, dst, &h->ctr->n_old, h->cursor, h->ctr);
    }
    PFEM2_LAUNCH(k_reseed, grid_for(own_n + 1, kThreads, 1 << 30), kThreads, 0, st, lo, hi, h->ppc, (const double2 *)h->mesh.d_vertices,
                 h->geom, h->centers, vel, h->cell_mask, h->stay, h->arrive, h->packed, dst, h->cell_start[h->cs ^ 1], h->ctr);
    h->cur ^= 1;
    h->cs ^= 1;
    CU(cudaGetLastError());
    return PFEM2_OK;
}

// Lazy re-sort: the records stay where the move pass wrote them; a rank pass over the dense key array builds the permutation
// sorted position -> row, re-seeds are appended behind the array
static int rank_and_reseed(pfem2_handle *h, NodalVel vel)
{
    cudaStream_t st = h->stream;
    PhaseScope ps(h, PFEM2_PHASE_REORDER);
    const int lo = h->own_lo, hi = h->own_hi, own_n = hi - lo;
    launch_plan(h, true, true);
    unsigned *src_new = h->vals[h->perm_buf ^ 1];
    // one pass per warp (1024 keys per block) in the hardware's block order, like the projection: 0.1 ms faster than 16 blocks per SM
    // striding over the array (the blocks in flight write one window of the permutation)
    PFEM2_LAUNCH(k_rank, grid_for(((long long)h->host_count + h->host_added + 3) / 4 + kThreads, kThreads, 1 << 30), kThreads, 0, st, (const unsigned *)h->keys[1], (const int *)&h->ctr->n_old, h->cursor, src_new,
                 h->ctr);
    PFEM2_LAUNCH(k_reseed_lazy, grid_for(own_n + 1, kThreads, 1 << 30), kThreads, 0, st, lo, hi, h->ppc, (const double2 *)h->mesh.d_vertices,
                 h->geom, h->centers, vel, h->cell_mask, h->stay, h->packed, h->soa[h->cur], (const int *)&h->ctr->n_old, h->tail_cursor,
                 src_new, h->cell_start[h->cs ^ 1], h->ctr);
    h->cs ^= 1;
    h->perm_buf ^= 1;
    h->permuted = true;
    CU(cudaGetLastError());
    return PFEM2_OK;
}

int advect_finish(pfem2_handle *h, NodalVel vel, int need_count)
{
    if (!h) return PFEM2_EINVAL;
    if (!h->move_pending) return fail(h, PFEM2_ESTATE, "advect_finish without advect_move");
    if (h->mv_interior_pending) return fail(h, PFEM2_ESTATE, "advect_finish before the interior part of a split move pass");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    int rc;
    const bool stable = h->opt.stable_order != 0;
    if (need_count && h->mg_fused) need_count = 0; // the move pass and immigrants_append already counted everybody who is still here
    h->mg_fused = false;
    h->move_pending = false;
    if (h->lazy_move) {
        h->lazy_move = false;
        if ((rc = rank_and_reseed(h, vel))) return rc;
    } else {
        if (need_count) {
            // multi-GPU, unfused: particles came and went since the move pass; count everybody (all "arrived": no stayer shortcut)
            PhaseScope ps(h, PFEM2_PHASE_REORDER);
            const int C = h->mesh.n_cells;
            PFEM2_LAUNCH(k_count_all, grid_for(h->capacity), kThreads, 0, st, h->soa[h->cur], h->ctr, h->opt.subcell_mode ? 1 : 0, C, h->ppc,
                         h->level, h->sub_step, h->stay, h->arrive, h->cell_mask);
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
        if ((rc = reorder(h, true, !need_count, stable, vel))) return rc;
    }
    if ((rc = queue_readback(h))) return rc;
    if (h->opt.verbose) {
        if ((rc = sync_counters(h))) return rc;
        printf("Particle handler contains %d particles\n", h->host_count); // particle_handler_2d.cu:341
    }
    return PFEM2_OK;
}

// ---- advectParticles as ONE graph launch (small, launch-bound cases: the shipped meshes run ~10 kernels of a few microseconds) ----
// Eligible: single GPU, lazy re-sort in its steady (permuted) state, no profiling / verbose output, buffers allocated, no growth due.
// The first eligible call of a parity captures the plain enqueue sequence on a private stream (the host-side state transitions happen
// as usual) and launches the graph on the handle's stream; later calls with the same arguments replay it and apply the same
// transitions by hand.  pfem2_options.graph_advect: 0 = auto (meshes below 2^18 cells), 1 = always, -1 = never.
static bool graph_wanted(const pfem2_handle *h)
{
    if (h->graphs_broken || h->opt.graph_advect < 0) return false;
    return h->opt.graph_advect > 0 || h->mesh.n_cells < (1 << 18);
}

static void drop_graphs(pfem2_handle *h)
{
    for (auto &g : h->graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g = pfem2_handle::AdvectGraph{};
    }
}

static int advect_graphed(pfem2_handle *h, NodalVel vel, double dt, int substeps, bool &done)
{
    done = false;
    if (!graph_wanted(h) || !lazy_enabled(h) || !h->permuted || h->profiling || h->opt.verbose || h->pipe.active || h->move_pending ||
        h->own_lo != 0 || h->own_hi != h->mesh.n_cells || !h->v2 || !h->tail_cursor || !h->seeded || substeps < 1)
        return PFEM2_OK;
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = sync_counters(h))) return rc;
    {   // the capacity policy of advect_move: a call that has to grow runs the plain way
        if ((long long)h->host_count + capacity_margin(h) > h->capacity) return PFEM2_OK;
    }
    pfem2_handle::AdvectGraph &g = h->graphs[h->cur | (h->cs << 1) | (h->perm_buf << 2)];
    const bool match = g.exec && g.vx == vel.x && g.vy == vel.y && g.table == (const void *)vel.table && g.dt == dt && g.substeps == substeps &&
                       g.capacity == h->capacity && g.dv == h->dv_pending && g.buf0 == h->soa[0].records() && g.buf1 == h->soa[1].records();
    if (match) {
        // the host-side transitions of advect_move + advect_finish (lazy, single GPU)
        h->partials_valid = false;
        h->last_substeps = substeps;
        h->arrive_dirty = false;
        h->mg_fused = false;
        h->dv_pending = false;
        h->cur ^= 1;
        h->cs ^= 1;
        h->perm_buf ^= 1;
        h->permuted = true;
        CU(cudaGraphLaunch(g.exec, h->stream));
        g_kernel_launches.fetch_add(g.kernels, std::memory_order_relaxed);
        h->readback_pending = true;
        ++h->graph_replays;
        done = true;
        return PFEM2_OK;
    }
    // capture
    if (!h->graph_stream && cudaStreamCreateWithFlags(&h->graph_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        h->graphs_broken = true;
        return PFEM2_OK;
    }
    if (g.exec) {
        cudaGraphExecDestroy(g.exec);
        g = pfem2_handle::AdvectGraph{};
    }
    pfem2_handle::AdvectGraph key;
    key.vx = vel.x; key.vy = vel.y; key.table = (const void *)vel.table; key.dt = dt; key.substeps = substeps; key.capacity = h->capacity;
    key.dv = h->dv_pending; key.buf0 = h->soa[0].records(); key.buf1 = h->soa[1].records();
    if ((rc = lazy_record_maps(h, 0)) || (rc = lazy_record_maps(h, 1))) return rc; // (host-side encodes: before the capture starts)
    cudaStream_t user = h->stream;
    if (cudaStreamBeginCapture(h->graph_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        h->graphs_broken = true;
        return PFEM2_OK;
    }
    h->stream = h->graph_stream;
    h->capturing = true;
    const long long launches0 = g_kernel_launches.load(std::memory_order_relaxed);
    rc = advect_move(h, vel, dt, substeps, 1, false, false);
    if (!rc) rc = advect_finish(h, vel, 0);
    h->capturing = false;
    h->stream = user;
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(h->graph_stream, &graph);
    if (rc || e != cudaSuccess || !graph) {
        // the enqueue sequence could not be captured here: the state transitions have happened but no work was enqueued -> the state
        // is unusable; report it (this does not happen on the supported stack; graphs can be switched off with graph_advect = -1)
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        h->graphs_broken = true;
        return rc ? rc : fail(h, PFEM2_ECUDA, "CUDA graph capture of advectParticles failed; set pfem2_options.graph_advect = -1");
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess || !exec) {
        cudaGetLastError();
        h->graphs_broken = true;
        return fail(h, PFEM2_ECUDA, "CUDA graph instantiation of advectParticles failed; set pfem2_options.graph_advect = -1");
    }
    key.exec = exec;
    key.kernels = g_kernel_launches.load(std::memory_order_relaxed) - launches0;
    g = key;
    CU(cudaGraphLaunch(g.exec, h->stream));
    done = true;
    return PFEM2_OK;
}

static int do_advect(pfem2_handle *h, NodalVel vel, double dt, int substeps)
{
    int rc;
    if (h) {
        bool done = false;
        if ((rc = advect_graphed(h, vel, dt, substeps, done))) return rc;
        if (done) return PFEM2_OK;
    }
    if ((rc = advect_move(h, vel, dt, substeps, 1, false, false))) return rc;
    return advect_finish(h, vel, 0);
}

// ------------------------------------------------------------------------------------------------
// projectVelocityOntoGrid / correctParticleVelocity
// ------------------------------------------------------------------------------------------------
void launch_project_cells(pfem2_handle *h, int c_lo, int c_hi)
{
    cudaStream_t st = h->stream;
    const int ppc = h->ppc;
    const ParticleSoA p = h->soa[h->cur];
    if (c_lo < 0) { // the owned range
        c_lo = h->own_lo;
        c_hi = h->own_hi;
    }
    const long long nc = c_hi - c_lo;
    const int *cs = h->cell_start[h->cs];
    // lanes per cell: about a quarter of the nominal segment length, so each lane keeps several loads in flight
    // One block per 256 / G cells, in the hardware's block order: the blocks in flight work on one contiguous window of the sorted
    // order (3.0 ms on channel16m; as 16 resident-size blocks per SM striding over the cells 3.5 ms -- 3.2 waves with a 20 % tail -- and
    // 3.3 ms with exactly 5 or 10 per SM: profiles/r03_summary.md §4)
    const int max_blocks = 1 << 30;
#define PFEM2_PROJECT(G)                                                                                                                   \
    do {                                                                                                                                   \
        if (h->permuted) /* lazy re-sort: the segment [cell_start[c], cell_start[c + 1]) names its records through the permutation */      \
            PFEM2_LAUNCH(k_project_cells_lazy<G>, grid_for(nc * G, kThreads, max_blocks), kThreads, 0, st, c_lo, c_hi, p,                  \
                         (const unsigned *)h->vals[h->perm_buf], cs, h->partial);                                                          \
        else                                                                                                                               \
            PFEM2_LAUNCH(k_project_cells<G>, grid_for(nc * G, kThreads, max_blocks), kThreads, 0, st, c_lo, c_hi, p, cs, h->partial);      \
    } while (0)
    if (ppc <= 4) PFEM2_PROJECT(2);
    else if (ppc <= 16) PFEM2_PROJECT(4);
    else if (ppc <= 36) PFEM2_PROJECT(8);
    else PFEM2_PROJECT(16);
#undef PFEM2_PROJECT
}

void launch_project_nodes(pfem2_handle *h, int node_lo, int node_hi, double *vx, double *vy, double *const *table, double *cx, double *cy,
                          double *const *table_copy)
{
    if (node_hi > node_lo)
        PFEM2_LAUNCH(k_project_nodes, grid_for(node_hi - node_lo, kThreads, 1 << 30), kThreads, 0, h->stream, node_lo, node_hi, h->node_off,
                     (const int *)h->node_inc, h->partial, vx, vy, table, cx, cy, table_copy);
}

void launch_project_nodes_acc_range(pfem2_handle *h, int node_lo, int node_hi, double *acc3)
{
    if (node_hi > node_lo)
        PFEM2_LAUNCH(k_project_nodes_acc_range, grid_for(node_hi - node_lo, kThreads, 1 << 30), kThreads, 0, h->stream, node_lo, node_hi, h->own_lo,
                     h->own_hi, h->node_off, (const int *)h->node_inc, h->partial, acc3);
}

void launch_project_finalize_range(pfem2_handle *h, int node_lo, int node_hi, const double *acc3, double *vx, double *vy)
{
    if (node_hi > node_lo)
        PFEM2_LAUNCH(k_project_finalize_range, grid_for(node_hi - node_lo, kThreads, 1 << 30), kThreads, 0, h->stream, node_lo, node_hi, h->own_lo,
                     h->own_hi, h->node_off, (const int *)h->node_inc, acc3, vx, vy);
}

static int do_project(pfem2_handle *h, double *vx, double *vy, double *const *table, double *cx = nullptr, double *cy = nullptr,
                      double *const *table_copy = nullptr)
{
    if (!h) return PFEM2_EINVAL;
    if (!h->seeded) return fail(h, PFEM2_ESTATE, "project before seed");
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = flush_correct(h))) return rc;
    if (!h->partials_valid) {
        PhaseScope ps(h, PFEM2_PHASE_PROJECT_CELLS);
        launch_project_cells(h);
        h->partials_valid = true;
    }
    PhaseScope ps(h, PFEM2_PHASE_PROJECT_NODES);
    launch_project_nodes(h, 0, h->mesh.n_nodes, vx, vy, table, cx, cy, table_copy);
    CU(cudaGetLastError());
    return PFEM2_OK;
}

static int apply_correct_now(pfem2_handle *h, NodalVel v, NodalVel vold, bool has_old)
{
    h->partials_valid = false; // particle velocities change
    const int rcm = materialize(h); // the eager kernel walks the physical order
    if (rcm) return rcm;
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

static int do_correct(pfem2_handle *h, NodalVel v, NodalVel vold, bool has_old)
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
    h->partials_valid = false;
    return PFEM2_OK;
}

// node -> incidence CSR on the device with the library's own radix sort; uses keys/vals as scratch
static int build_node_incidence(pfem2_handle *h)
{
    cudaStream_t st = h->stream;
    const int C = h->mesh.n_cells, N = h->mesh.n_nodes;
    const int m = 3 * C;
    DeviceTemps tmp;
    int *count = nullptr, *n_dev = nullptr, *scratch = nullptr;
    CU(tmp.alloc(&count, (size_t)N + 1));
    CU(tmp.alloc(&n_dev, 1));
    CU(tmp.alloc(&scratch, scan_scratch_elems<int>(N)));
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
    return PFEM2_OK;
}

} // namespace host
} // namespace pfem2

using namespace pfem2;
using namespace pfem2::host;

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
    o->lazy_sort = 1;
}

const char *pfem2_last_error(const pfem2_handle *h) { return h ? h->error.c_str() : g_create_error.c_str(); }

const char *pfem2_version(void) { return "pfem2_b200 0.2 (sm_100a)"; }

long long pfem2_kernel_launches(void) { return g_kernel_launches.load(std::memory_order_relaxed); }

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
    if (opt.reserved_scatter_tma || opt.reserved_lane_per_record || opt.reserved_fuse_project)
        return fail(nullptr, PFEM2_EINVAL, "scatter_tma / lane_per_record / fuse_project were A/B kernel variants of round 1 and have been removed "
                                           "(the fields are reserved and must be 0)");

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
        const char *e = getenv("PFEM2_LAZY_SWIZZLE"); // env: tile-layout cross-check of the gathered move pass (tests)
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
#define BAIL(code) do { std::string e = h->error; pfem2_destroy(h); g_create_error = e; return (code); } while (0)
#define TRY(x) do { if ((rc = (x))) BAIL(rc); } while (0)
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
            BAIL(PFEM2_ECUDA);
        }
        memset(h->host_ctr, 0, sizeof(Counters));
    }
    const long long want = std::max<long long>((long long)std::ceil((double)C * h->ppc * opt.capacity_factor), 3ll * C);
    TRY(alloc_particle_storage(h, (int)want));
    {
        cudaError_t e = cudaMemcpyAsync(h->centers, cen.data(), cen.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { h->error = cudaGetErrorString(e); BAIL(PFEM2_ECUDA); }
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
        if (e != cudaSuccess) { h->error = cudaGetErrorString(e); BAIL(PFEM2_ECUDA); }
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
        if (e != cudaSuccess) { h->error = cudaGetErrorString(e); BAIL(PFEM2_ECUDA); }
    }
#undef TRY
#undef BAIL
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
    drop_graphs(h);
    if (h->graph_stream) cudaStreamDestroy(h->graph_stream);
    if (h->pipe.copy) cudaStreamDestroy(h->pipe.copy);
    for (cudaStream_t m : h->pipe.mv)
        if (m) cudaStreamDestroy(m);
    for (cudaEvent_t e : h->pipe.mv_ev)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->pipe.up_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : h->pipe.dn_ev) cudaEventDestroy(e);
    for (double *p : h->nodal) cudaFree(p);
    cudaFree(h->acc3);
    for (auto &r : h->phase_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
    if (h->host_ctr) cudaFreeHost(h->host_ctr);
    if (h->readback) cudaEventDestroy(h->readback);
    delete h;
    return PFEM2_OK;
}

int pfem2_seed(pfem2_handle *h)
{
    if (!h) return PFEM2_EINVAL;
    if (h->move_pending) return fail(h, PFEM2_ESTATE, "seed between advect_move and advect_finish");
    CU(cudaSetDevice(h->device));
    const int C = h->mesh.n_cells;
    h->partials_valid = false;
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
        h->aos_bytes = 0;
        CU(cudaMalloc(&h->aos, need));
        h->aos_bytes = need;
    }
    PFEM2_LAUNCH(k_export_aos, grid_for(h->capacity), kThreads, 0, h->stream, h->soa[h->cur], h->ctr, (uint4 *)h->aos);
    CU(cudaGetLastError());
    *d_particles96 = h->aos;
    if (count) *count = h->host_count;
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
    if (!h || n < 0 || !x || !y || !l0 || !l1 || !l2 || !vx || !vy || !cell) return PFEM2_EINVAL;
    if (h->move_pending) return fail(h, PFEM2_ESTATE, "upload between advect_move and advect_finish");
    CU(cudaSetDevice(h->device));
    int rc;
    if ((rc = sync_counters(h))) return rc;
    // validate and stage first: a rejected upload leaves the handle's state (pending correction, permutation) untouched
    const int C = h->mesh.n_cells;
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
    if (n > h->capacity) {
        if ((rc = materialize(h))) return rc; // (grow keeps the physical prefix; it is overwritten right below)
        const int keep = h->host_count;
        h->host_count = 0;
        rc = grow(h, (int)std::min<long long>(2147483000ll, (long long)(1.25 * n) + 4096));
        if (rc) {
            h->host_count = keep;
            return rc;
        }
    }
    // commit: the uploaded state replaces everything, including a correction not yet applied and a permutation of the replaced state
    h->partials_valid = false;
    h->arrive_dirty = true;
    h->dv_pending = false;
    h->permuted = false;
    cudaStream_t st = h->stream;
    ParticleSoA &p = h->soa[h->cur];
    if (n) CU(cudaMemcpyAsync(p.records(), hr.data(), (size_t)n * sizeof(ParticleRec), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    PFEM2_LAUNCH(k_set_counters, 1, 1, 0, st, h->ctr, n, h->capacity);
    {   // per-cell scratch of the owned range (+ a few cells for the tolerance-band spill of the occupancy bits)
        const size_t lo = (size_t)h->own_lo, len = (size_t)std::min(C, h->own_hi + 4) - lo + 1;
        CU(cudaMemsetAsync(h->stay + lo, 0, sizeof(int) * len, st));
        CU(cudaMemsetAsync(h->arrive + lo, 0, sizeof(int) * len, st));
        CU(cudaMemsetAsync(h->cursor + lo, 0, sizeof(int) * len, st));
        CU(cudaMemsetAsync(h->cell_mask + lo, 0, sizeof(unsigned long long) * len, st));
    }
    // an arbitrary (unsorted) array is handled as "everybody is a mover" of the stable path: radix sort by cell
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

int pfem2_node_ranges(pfem2_handle *h, int substeps, int *out4)
{
    if (!h || !out4 || substeps < 1) return PFEM2_EINVAL;
    CU(cudaSetDevice(h->device));
    const int rc = ensure_v2_node_range(h, substeps);
    if (rc) return rc;
    out4[0] = h->v2_node_lo;
    out4[1] = h->v2_node_hi;
    out4[2] = h->own_node_lo;
    out4[3] = h->own_node_hi;
    return PFEM2_OK;
}

int pfem2_set_owned_cells(pfem2_handle *h, int cell_lo, int cell_hi)
{
    if (!h) return PFEM2_EINVAL;
    if (cell_lo < 0 || cell_hi > h->mesh.n_cells || cell_lo > cell_hi) return fail(h, PFEM2_EINVAL, "bad owned cell range");
    if (h->seeded) return fail(h, PFEM2_ESTATE, "set_owned_cells after seed");
    h->own_lo = cell_lo;
    h->own_hi = cell_hi;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int own_n = cell_hi - cell_lo, N = h->mesh.n_nodes;
    CU(cudaMemcpyAsync(h->own_len_dev, &own_n, sizeof(int), cudaMemcpyHostToDevice, st));
    // compact list of the nodes the owned cells touch
    DeviceTemps tmp;
    int *flag = nullptr, *pos = nullptr, *scratch = nullptr, *n_dev = nullptr;
    CU(tmp.alloc(&flag, (size_t)N + 1));
    CU(tmp.alloc(&pos, (size_t)N + 1));
    CU(tmp.alloc(&scratch, scan_scratch_elems<int>(N)));
    CU(tmp.alloc(&n_dev, 1));
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

#ifdef PFEM2_MOVE_TRACE
// diagnosis build only (tools/trace_move.py): buffers for the per-warp / per-tile timing of the gathered move pass
int pfem2_debug_move_trace(unsigned long long *warp_buf, uint4 *tile_buf)
{
    pfem2_handle *h = nullptr;
    CU(cudaMemcpyToSymbol(g_trace_warp, &warp_buf, sizeof(warp_buf)));
    CU(cudaMemcpyToSymbol(g_trace_tile, &tile_buf, sizeof(tile_buf)));
    return PFEM2_OK;
}
#endif

} // extern "C"
