// pfem2_move.cuh -- the move pass: S substeps of advect + locate fused, per-cell statistics (sm_100a, TMA tiles).
#pragma once

#include "pfem2_common.cuh"

namespace pfem2 {

// ---------------------------------------------------------------------------------------------
// locate helpers
// ---------------------------------------------------------------------------------------------
// Exact rule of kCheckParticleInNeighbors (:133-162): ascending one-ring of cell c, first accepting cell wins.
__device__ __forceinline__ bool ring_scan(const CellGeom *__restrict__ geom, const int *__restrict__ nbr_off,
                                          const int *__restrict__ nbr_idx, unsigned &c, double x, double y, double &L0,
                                          double &L1, double &L2)
{
    const int k1 = __ldg(nbr_off + c + 1);
    for (int k = __ldg(nbr_off + c); k < k1; ++k) {
        const unsigned nb = (unsigned)__ldg(nbr_idx + k);
        const CellGeom gn = load_geom(geom, nb);
        double a0, a1, a2;
        to_local(gn, x, y, a0, a1, a2);
        if (inside_unit(a0, a1, a2)) {
            c = nb;
            L0 = a0; L1 = a1; L2 = a2;
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ bool shares_vertex(const CellGeom &a, const CellGeom &b)
{
    return a.n0 == b.n0 || a.n0 == b.n1 || a.n0 == b.n2 || a.n1 == b.n0 || a.n1 == b.n1 || a.n1 == b.n2 || a.n2 == b.n0 ||
           a.n2 == b.n1 || a.n2 == b.n2;
}

// Locate a particle that left cell c (own-cell test failed with local coordinates a0..a2).
// Fast path: walk across the edge with the most negative coordinate, up to kWalkHops cells.  A cell T in which
// the point is STRICTLY interior by the per-cell margin (CellGeom.pad, set at create so that no other cell of a
// non-overlapping triangulation can accept the point even with the +-2e-6 tolerance: DESIGN.md "locate") is the
// unique acceptor, so the reference's ordered scan would return exactly T if T is in the one-ring of c, and would
// delete the particle otherwise.  Everything else (tolerance band, domain boundary, hop limit) falls back to the
// ordered scan itself, so the result is always the reference's rule.
constexpr int kWalkHops = 3;
template <bool WALK>
__device__ __forceinline__ bool locate_mover(const CellGeom *__restrict__ geom, const int4 *__restrict__ edge_nbr,
                                             const int *__restrict__ nbr_off, const int *__restrict__ nbr_idx,
                                             CellGeom &g, int4 &e, unsigned &c, double x, double y, double &L0, double &L1,
                                             double &L2)
{
    // on entry L0..L2 are the (rejected) local coordinates in cell c; on success they are those in the new cell
    if (WALK) {
        const unsigned o0 = g.n0, o1 = g.n1, o2 = g.n2; // nodes of the cell the particle started the substep in
#pragma unroll 1
        for (int hop = 0; hop < kWalkHops; ++hop) {
            const int nxt = (L0 <= L1 && L0 <= L2) ? e.x : ((L1 <= L2) ? e.y : e.z);
            if (nxt < 0) break; // domain boundary: let the ordered scan decide
            // the candidate's record replaces g / e in place (every exit below either keeps it or reloads)
            g = load_geom(geom, (unsigned)nxt);
            e = __ldg(edge_nbr + nxt);
            to_local(g, x, y, L0, L1, L2);
            const double m = (double)__uint_as_float(g.pad);
            if (L0 > m && L1 > m && L2 > m) { // strictly interior: unique acceptor
                // the first hop crosses an edge of the start cell (two shared vertices): always inside the one-ring
                if (hop > 0) {
                    const bool in_ring = g.n0 == o0 || g.n0 == o1 || g.n0 == o2 || g.n1 == o0 || g.n1 == o1 || g.n1 == o2 ||
                                         g.n2 == o0 || g.n2 == o1 || g.n2 == o2;
                    if (!in_ring) return false; // acceptor outside the one-ring -> deleted
                }
                c = (unsigned)nxt;
                return true;
            }
            if (inside_unit(L0, L1, L2)) break; // accepted inside the tolerance band: ties possible -> ordered scan
        }
    }
    if (!ring_scan(geom, nbr_off, nbr_idx, c, x, y, L0, L1, L2)) return false;
    g = load_geom(geom, c);
    e = __ldg(edge_nbr + c);
    return true;
}

// Per-cell survivor counts and sub-cell occupancy (kCountParticlesInSubcells :173-181) for one warp of particles:
// lanes that end in the same cell form one group (match_any); its leader issues one atomic per counter for the
// whole group.  sb / mb are the warp ballots of "stayed in its cell" / "changed cell".
// subcell_mode (pfem2_options) and "ppc > 32" are warp-uniform run-time values here: as template parameters they multiplied the
// instantiations of the move pass by four for two instructions' worth of difference.
__device__ __forceinline__ void accumulate_cell_stats(int subcell_mode, bool live, unsigned c, double L0, double L1, double L2, unsigned sb,
                                                      unsigned mb, int lane, int n_cells, int ppc, int level, double sub_step,
                                                      int *__restrict__ stay, int *__restrict__ arrive,
                                                      unsigned long long *__restrict__ cell_mask)
{
    unsigned fc = 0xffffffffu;
    unsigned long long bit = 0;
    if (live) {
        const int sub = subcell_mode == 0 ? subcell_index(L0, L1, L2, level, sub_step) : subcell_index_clamped(L0, L1, L2, level, sub_step);
        // reference: ++hist[cellID * ppc + sub] in unsigned arithmetic, unchecked
        if ((unsigned)sub < (unsigned)ppc) { // the regular case: the particle's own word (no division)
            fc = c;
            bit = 1ull << sub;
        } else { // tolerance band: the index spills into a neighbouring cell's word, or out of the histogram
            const unsigned long long flat = (unsigned long long)(unsigned)(c * (unsigned)ppc + (unsigned)sub);
            if (flat < (unsigned long long)n_cells * ppc) {
                fc = (unsigned)(flat / (unsigned)ppc);
                bit = 1ull << (unsigned)(flat - (unsigned long long)fc * ppc);
            }
        }
    }
    const bool own_word = live && fc == c; // false only for the tolerance-band spill into another cell's word
    const unsigned peers = __match_any_sync(0xffffffffu, live ? c : 0xffffffffu);
    const unsigned long long gbit = own_word ? bit : 0ull;
    unsigned lo = __reduce_or_sync(peers, (unsigned)gbit);
    unsigned hi = 0;
    if (ppc > 32) hi = __reduce_or_sync(peers, (unsigned)(gbit >> 32));
    if (live && (peers & ((1u << lane) - 1)) == 0) {
        const int ns = __popc(peers & sb), na = __popc(peers & mb);
        if (arrive) { // the stable-order path places stayers and movers separately
            if (ns) atomicAdd(stay + c, ns);
            if (na) atomicAdd(arrive + c, na);
        } else { // fast order: only the sum is needed (arrive[] stays zero)
            atomicAdd(stay + c, ns + na);
        }
        const unsigned long long word = (unsigned long long)lo | ((unsigned long long)hi << 32);
        if (word) atomicOr(cell_mask + c, word);
    }
    if (live && !own_word && fc != 0xffffffffu) atomicOr(cell_mask + fc, bit);
}

// destination rank of a cell: bounds[r] <= cell < bounds[r + 1]   (n_ranks <= 64)
__device__ __forceinline__ int rank_of_cell(unsigned c, const int *__restrict__ bounds, int n_ranks)
{
    int r = 0;
    while (r + 1 < n_ranks && (int)c >= bounds[r + 1]) ++r;
    return r;
}

// ---------------------------------------------------------------------------------------------
// The move pass: advect + locate, all S substeps fused in one pass over the particles
//   kAdvectParticles :54-70, kCheckParticleInCell :117-131, kCheckParticleInNeighbors :133-162.
// The nodal field is frozen inside advectParticles and particles do not interact, so the S substeps need no global
// barrier between them.  A particle with no accepting cell in own ∪ one-ring is marked lost.
//
// Data movement: every warp owns two 2 KB tiles of shared memory and lets the copy engine move whole 32-record tiles
// global <-> shared (cp.async.bulk.tensor, 64-byte swizzle so that "lane r reads record r" is bank-conflict free): the
// LSU sees four shared loads and four shared stores per lane and iteration and nothing else for the particle state;
// the next tile lands while the current one is computed (ping-pong, one mbarrier per tile).  With one lane per 64-byte
// record and 128-bit global accesses a warp instruction touches 16 cache lines, and the L1 data pipe -- not HBM, not
// instruction issue -- bounded that form (76 %, profiles/r01b_summary.md).  Nodal velocities come interleaved
// (double2 per node: one 128-bit gather per node instead of two 64-bit ones).
//
// Two kernels share the per-particle body (move_record) and the statistics tail (finish_tile):
//   k_move_gather   (default, lazy re-sort)  tile j of the SORTED order is gathered through the permutation src[] left
//                   by the previous step's rank pass (8 x tile::gather4 per 32 records) and stored DENSE into the other
//                   buffer at rows 32 j ..; the new cells also go to a dense key array for the rank pass;
//   k_move_tiles    (physical order: stable_order, lazy_sort = 0) tile j of the array, in place.
// Per warp-tile both accumulate, with one atomic per distinct cell (match_any aggregation):
//   stay[c] (+ arrive[c])  survivors that end in cell c,
//   cell_mask[c]           sub-cell occupancy bits, flat unclamped index like kCountParticlesInSubcells (:173-181).
// Multi-GPU (emig_idx != nullptr): a particle whose new cell lies outside the owned range [own_lo, own_hi) is an
// emigrant.  It is written back like everybody else (the pack kernel reads it there), but it is kept out of the per-cell
// statistics, its row is appended to emig_idx and rank_count[destination] / rank_count[n_ranks] (total) are bumped, so
// that neither a counting nor a search pass over the whole array is needed afterwards.
// ---------------------------------------------------------------------------------------------
constexpr int kAdvTileBytes = 32 * (int)sizeof(ParticleRec); // one warp tile
// device ints behind pfem2_handle::tail_cursor: [0] the cursor of the appended re-seeds, [kTileCursor0 + i] the tile cursor of the
// i-th launch of the gathered move pass inside one advect (chunks of pfem2_step_host, parts of a split strip pass), 128 bytes apart
constexpr int kTileCursor0 = 32, kTileCursors = 32 * 32;

#ifdef PFEM2_MOVE_TRACE
// diagnosis build (make variant NAME=trace DEFS=-DPFEM2_MOVE_TRACE, tools/trace_move.py): per-warp start / end time + SM of the gathered move
// pass and the duration of every tile iteration, written to buffers handed in through pfem2_debug_move_trace
static __device__ unsigned long long *g_trace_warp = nullptr; // [warps][4]: t_start, t_end (globaltimer ns), smid, tiles
static __device__ uint4 *g_trace_tile = nullptr;              // [tiles]: ns in the four parts of the tile's iteration (top + issue, wait for the tile, move, store + statistics)
__device__ __forceinline__ unsigned long long trace_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned trace_smid()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}
#endif
constexpr size_t advect_tma_smem_bytes(int threads) { return (size_t)(threads / 32) * (2 * kAdvTileBytes + 2 * sizeof(uint64_t)) + 1024; }

// Block size / resident blocks per SM of the move pass, compile-time.  Swept on B200 (channel16m, ms per pass):
// 256x4 8.42, 128x8 8.38, 64x16 8.45 (all 64 registers, 32 warps per SM); 192x5 8.67 (30 warps); 128x7 8.97 (72 registers, 28 warps);
// 128x6 9.14 and 256x3 9.24 (80 registers, no spills, 24 warps): the pass is latency-bound, resident warps beat registers.
// How the warps of the gathered move pass get their tiles: PFEM2_MOVE_GDYN = G > 0 (default 4): groups of G consecutive tiles through one
// global cursor, every warp works at the front of the pass; 0: the fixed warp-strided share of round 2 (A/B: profiles/r03_summary.md).
#ifndef PFEM2_MOVE_GDYN
#define PFEM2_MOVE_GDYN 4
#endif
static_assert(PFEM2_MOVE_GDYN >= 0 && (PFEM2_MOVE_GDYN & (PFEM2_MOVE_GDYN - 1)) == 0, "PFEM2_MOVE_GDYN: 0 or a power of two");
#ifndef PFEM2_ADV_THREADS
#define PFEM2_ADV_THREADS 256
#define PFEM2_ADV_MINB 4
#endif
constexpr int kAdvThreads = PFEM2_ADV_THREADS, kAdvBlocksPerSM = PFEM2_ADV_MINB;

// One particle through the S substeps.  Its record sits in the warp's shared-memory tile: the 16-byte field f at sa0 ^ (f << 4).
// On return position / local coordinates / cell have been written back into the tile (a lost particle only gets its position and
// cell = kLostCell: it stores no local position, like the reference) and c, L0..L2 hold the final cell and local position.
template <bool WALK, int NSUB>
__device__ __forceinline__ void move_record(uint32_t sa0, const CellGeom *__restrict__ geom, const int4 *__restrict__ edge_nbr,
                                            const int *__restrict__ nbr_off, const int *__restrict__ nbr_idx,
                                            const double2 *__restrict__ V2, const double2 *__restrict__ dV2, double h, int substeps,
                                            unsigned &c, double &L0, double &L1, double &L2, int &moved, bool &lost)
{
    const uint32_t sa1 = sa0 ^ 16u, sa2 = sa0 ^ 32u, sa3 = sa0 ^ 48u;
    const int4 r0 = lds128(sa0), r1 = lds128(sa1), r2 = lds128(sa2);
    c = (unsigned)r2.z;
    double x = __hiloint2double(r0.y, r0.x);
    double y = __hiloint2double(r0.w, r0.z);
    L0 = __hiloint2double(r1.y, r1.x);
    L1 = __hiloint2double(r1.w, r1.z);
    L2 = __hiloint2double(r2.y, r2.x);
    // cell record and its three nodal velocities stay in registers while the particle stays in the cell
    CellGeom g = load_geom(geom, c);
    int4 e = __ldg(edge_nbr + c); // prefetched with the cell record so the first walk hop has no extra dependent load
    double2 a0 = __ldg(V2 + g.n0), a1 = __ldg(V2 + g.n1), a2 = __ldg(V2 + g.n2);
    if (dV2) { // pending velocity correction, with the cell / local position the particle had at the correct call
        const int4 r3 = lds128(sa3);
        const double2 d0 = __ldg(dV2 + g.n0), d1 = __ldg(dV2 + g.n1), d2 = __ldg(dV2 + g.n2);
        const double vx = __dadd_rn(__hiloint2double(r3.y, r3.x), interp3(L0, L1, L2, d0.x, d1.x, d2.x));
        const double vy = __dadd_rn(__hiloint2double(r3.w, r3.z), interp3(L0, L1, L2, d0.y, d1.y, d2.y));
        sts128(sa3, make_int4(__double2loint(vx), __double2hiint(vx), __double2loint(vy), __double2hiint(vy)));
    }
    const int nsub = NSUB > 0 ? NSUB : substeps;
    // kept rolled on purpose: unrolling lets stayers run ahead into the next substep's code while the movers of the
    // warp are still in the walk, which costs more issue slots than it saves
#pragma unroll 1
    for (int s = 0; s < nsub; ++s) {
        // kAdvectParticles: velocity from the STORED local position and cell
        const double ux = interp3(L0, L1, L2, a0.x, a1.x, a2.x);
        const double uy = interp3(L0, L1, L2, a0.y, a1.y, a2.y);
        x = __fma_rn(ux, h, x);
        y = __fma_rn(uy, h, y);
        // own cell first (wins even if a neighbour would also accept)
        to_local(g, x, y, L0, L1, L2);
        if (inside_unit(L0, L1, L2)) continue;
        ++moved;
        if (!locate_mover<WALK>(geom, edge_nbr, nbr_off, nbr_idx, g, e, c, x, y, L0, L1, L2)) {
            lost = true;
            break;
        }
        if (s + 1 < nsub) {
            a0 = __ldg(V2 + g.n0);
            a1 = __ldg(V2 + g.n1);
            a2 = __ldg(V2 + g.n2);
        }
    }
    sts128(sa0, make_int4(__double2loint(x), __double2hiint(x), __double2loint(y), __double2hiint(y)));
    if (lost) {
        sts32(sa2 + 8, kLostCell);
    } else {
        sts128(sa1, make_int4(__double2loint(L0), __double2hiint(L0), __double2loint(L1), __double2hiint(L1)));
        sts128(sa2, make_int4(__double2loint(L2), __double2hiint(L2), (int)c, r2.w));
    }
}

// One thread's atomic add whose result is NOT needed right away.  atomicAdd() inside `if (lane == 0)` is compiled as a warp-aggregated
// atomic -- vote, leader's ATOMG, SHFL of the result -- and the shuffle waits for the round trip on the spot; the plain instruction
// leaves the result in flight until its first use (measured neutral on channel16m: the wait overlapped the statistics of the tile).
__device__ __forceinline__ int atom_add_in_flight(int *p, int v)
{
    int r;
    asm volatile("atom.relaxed.gpu.global.add.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
    return r;
}

// multi-GPU: list an emigrant (rare: a handful per tile row at a strip interface); returns true if the particle left the owned range
__device__ __forceinline__ bool list_emigrant(bool live, unsigned c, int row, int own_lo, int own_hi, const int *__restrict__ rank_bounds,
                                              int n_ranks, int *__restrict__ rank_count, unsigned *__restrict__ emig_idx)
{
    if (!(emig_idx && live && ((int)c < own_lo || (int)c >= own_hi))) return false;
    const int slot = atomicAdd(rank_count + n_ranks, 1);
    emig_idx[slot] = (unsigned)row;
    atomicAdd(rank_count + rank_of_cell(c, rank_bounds, n_ranks), 1);
    return true;
}

// ---- physical order, in place ----
// FAST (the fast order, i.e. not pfem2_options.stable_order): only the number of survivors per cell is needed (stayers + arrivals are
// summed, accumulate_cell_stats with arrive == nullptr, and no stay bits are written), so the cell a particle started in is not carried
// through the substep loop -- one register less at the 64-register cap: ptxas then allocates the S = 3 form without a single spill.
// The stable order additionally gets, per warp of 32 consecutive particles,
//   stay_bits[w]    ballot of particles that end in the cell they started in (they keep their array order),
//   warp_movers[w]  number of particles that changed cell (sorted separately by new cell).
template <bool WALK, int NSUB, bool FAST>
__global__ void __launch_bounds__(kAdvThreads, kAdvBlocksPerSM)
k_move_tiles(const __grid_constant__ CUtensorMap tmap, const CellGeom *__restrict__ geom, const int4 *__restrict__ edge_nbr,
             const int *__restrict__ nbr_off, const int *__restrict__ nbr_idx, const double2 *__restrict__ V2, double h,
             int substeps, int subcell_mode, int n_cells, int ppc, int level, double sub_step, Counters *ctr, unsigned *__restrict__ stay_bits,
             int *__restrict__ warp_movers, int *__restrict__ stay, int *__restrict__ arrive,
             unsigned long long *__restrict__ cell_mask, int do_count, const double2 *__restrict__ dV2, int own_lo, int own_hi,
             const int *__restrict__ rank_bounds, int n_ranks, int *__restrict__ rank_count, unsigned *__restrict__ emig_idx,
             const int *__restrict__ chunk_start, int chunk_lo, int chunk_hi)
{
    // chunk_start != nullptr: only the particles of the cells [chunk_lo, chunk_hi) are moved (chunk_start = the segment table
    // of the sorted array): pfem2_step_host launches the pass chunk by chunk while the nodal field is still arriving over PCIe.
    extern __shared__ unsigned char adv_smem_raw[];
    __shared__ int s_mov, s_lost;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    // 32-bit shared-window addresses throughout (ld.shared / st.shared, no generic pointers).  Tiles are 1 KB aligned: the
    // swizzle pattern is a function of the shared-memory address bits 4..8.
    const uint32_t smem = (smem_u32(adv_smem_raw) + 1023u) & ~1023u;
    const uint32_t tile0 = smem + (uint32_t)warp * (2 * kAdvTileBytes);
    const uint32_t bar0 = smem + (uint32_t)warps_per_block * (2 * kAdvTileBytes) + (uint32_t)warp * 16;
    if (threadIdx.x == 0) s_mov = s_lost = 0;
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        fence_barrier_init();
    }
    __syncthreads();
    int p_lo = 0, n = ctr->count; // particle range [p_lo, n) of this launch
    if (chunk_start) {
        p_lo = min(__ldg(chunk_start + chunk_lo), n);
        n = min(__ldg(chunk_start + chunk_hi), n);
    }
    const int tiles = (n + 31) >> 5;
    const int warp_global = (p_lo >> 5) + blockIdx.x * warps_per_block + warp;
    const int warps_total = gridDim.x * warps_per_block;
    // my record inside a tile: the 16-byte field f of row `lane` sits at  lane * 64 + ((f ^ sw) << 4)  =  my0 ^ (f << 4)
    const uint32_t my0 = (uint32_t)lane * 64 + ((((uint32_t)lane >> 1) & 3) << 4);
    if (lane == 0 && warp_global < tiles) {
        mbar_arrive_expect_tx(bar0, kAdvTileBytes);
        tma_load_tile_2d(tile0, &tmap, 0, warp_global << 5, bar0);
    }
    uint32_t b = 0, par = 0; // buffer of this iteration (0 / 1) and the phase parity of its mbarrier
    for (int tile = warp_global; tile < tiles; tile += warps_total) {
        const uint32_t buf = tile0 + b * kAdvTileBytes;
        // the other buffer: its store (previous iteration) must have finished reading shared memory, then the next tile
        // is fetched into it
        if (lane == 0) {
            bulk_wait_group_read<0>();
            const int nxt = tile + warps_total;
            if (nxt < tiles) {
                mbar_arrive_expect_tx(bar0 + (b ^ 1) * 8, kAdvTileBytes);
                tma_load_tile_2d(tile0 + (b ^ 1) * kAdvTileBytes, &tmap, 0, nxt << 5, bar0 + (b ^ 1) * 8);
            }
        }
        mbar_wait(bar0 + b * 8, par);
        par ^= b; // each barrier is used every other iteration: its parity flips after the odd buffer's turn
        b ^= 1;
        const int base = tile << 5;
        const int i = base + lane;
        const bool valid = i >= p_lo && i < n;
        unsigned c0 = 0, c = 0;
        double L0 = 0, L1 = 0, L2 = 0;
        int moved = 0; // substeps in which the particle left its cell
        bool lost = false;
        if (valid) {
            if (!FAST) c0 = (unsigned)lds128((buf + my0) ^ 32u).z;
            move_record<WALK, NSUB>(buf + my0, geom, edge_nbr, nbr_off, nbr_idx, V2, dV2, h, substeps, c, L0, L1, L2, moved, lost);
        }
        // hand the tile back: generic-proxy writes -> visible to the async proxy -> one lane issues the store
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_tile_2d(&tmap, 0, base, buf);
            bulk_commit_group();
        }
        bool live = valid && !lost;
        if (list_emigrant(live, c, i, own_lo, own_hi, rank_bounds, n_ranks, rank_count, emig_idx)) live = false;
        const bool stays = live && (FAST || c == c0);
        const unsigned sb = __ballot_sync(0xffffffffu, stays);
        const unsigned mb = __ballot_sync(0xffffffffu, live && !stays);
        const unsigned lb = __ballot_sync(0xffffffffu, lost);
        const int wm = __reduce_add_sync(0xffffffffu, moved);
        if (lane == 0) {
            if (wm) atomicAdd(&s_mov, wm);
            if (lb) atomicAdd(&s_lost, __popc(lb));
            if (stay_bits) {
                stay_bits[base >> 5] = sb;
                warp_movers[base >> 5] = __popc(mb);
            }
        }
        if (do_count)
            accumulate_cell_stats(subcell_mode, live, c, L0, L1, L2, sb, mb, lane, n_cells, ppc, level, sub_step, stay, arrive,
                                                        cell_mask);
    }
    if (lane == 0) bulk_wait_group<0>(); // the last stores still read this block's shared memory
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_mov) atomicAdd(&ctr->movers, s_mov);
        if (s_lost) atomicAdd(&ctr->lost, s_lost);
    }
}

// ---- sorted order through the permutation (lazy re-sort, the default) ----
// The records move ONCE per step.  A step keeps a permutation src[] (sorted position j -> row of the dense array) instead of
// physically re-sorting the array:
//   advect   k_move_gather          tile j of the SORTED order is gathered through src[] from buffer A (8 x gather4 per 32 records),
//                                   moved, and stored DENSE to buffer B at rows 32 j .. 32 j + 31 (B = "sorted by the cell of the
//                                   previous step"); the new cell also goes to the dense key array.
//            k_plan_cells, scan (+ PlanEpilogue: cursors, counts)  new segment table from the counts (pfem2_resort.cuh)
//            k_rank                 slot = cursor[cell]++ (one atomic per (warp, cell) group) ; src_new[slot] = i    (4 + 4 bytes per
//                                   particle instead of the 128-byte-per-particle scatter)
//            k_reseed_lazy          new particles are appended behind the array; their rows fill the tail of their cell's src_new range
//   project  k_project_cells_lazy   the nine sums of a cell from records[src_new[j]], j in the cell's segment
//   anything else that wants the sorted array (getParticles, download, eager correct, grow)
//            k_materialize          out[j] = in[src[j]] (four lanes per record), then the handle is in the physical state again
// Lost particles are never referenced by src_new, so the next move pass compacts them away for free.  Measured on channel16m
// (profiles/r02a_*): 14.3 ms per step against 17.7 ms with the physical re-sort.
//
// Differences to k_move_tiles:
//   * input tile = rows src[32 t .. 32 t + 31] of `gmap` (the current buffer), fetched with 8 gather4 operations (lanes 0..7, one each);
//   * output tile = rows 32 t .. of `tmap_out` (the OTHER buffer): the pass is not in place;
//   * keys[32 t + lane] = new cell (kLostCell for lost, emigrated and padding lanes): what k_rank consumes;
//   * src[] is padded with a valid row index up to a multiple of 32.
// The swizzle of `gmap` is the one of `tmap_out` (64-byte): TMA swizzling is a function of the shared-memory address, so the four
// rows a gather4 drops at tile + 256 k land exactly where a 32-row tile load would have put rows 4 k .. 4 k + 3 (verified on B200:
// profiles/r02a_gather4_swizzle.log and the parity tests in both layouts).
// SWZ = false (PFEM2_LAZY_SWIZZLE=0): both maps are encoded without swizzle and lane r reads its record at r * 64 (4-way bank
// conflicts on the shared-memory side; 15.4 instead of 14.3 ms per step): kept as the layout-independent cross-check of the tests.
template <bool WALK, int NSUB, bool SWZ, int CLAIM>
__global__ void __launch_bounds__(kAdvThreads, kAdvBlocksPerSM)
k_move_gather(const __grid_constant__ CUtensorMap gmap, const __grid_constant__ CUtensorMap tmap_out, const int4 *__restrict__ src,
              unsigned *__restrict__ keys, const CellGeom *__restrict__ geom, const int4 *__restrict__ edge_nbr,
              const int *__restrict__ nbr_off, const int *__restrict__ nbr_idx, const double2 *__restrict__ V2, double h, int substeps,
              int subcell_mode, int n_cells, int ppc, int level, double sub_step, Counters *ctr, int *__restrict__ stay,
              unsigned long long *__restrict__ cell_mask, const double2 *__restrict__ dV2, int own_lo, int own_hi,
              const int *__restrict__ rank_bounds, int n_ranks, int *__restrict__ rank_count, unsigned *__restrict__ emig_idx,
              const int *__restrict__ chunk_start, int chunk_lo, int chunk_hi, int part, int *__restrict__ tile_cursor)
{
    // part = 1, 2, 3: split pass of a strip (multi-GPU).  With t1 = start of cell chunk_lo rounded UP to a tile and t2 = start of cell
    // chunk_hi rounded down (>= t1), part 1 moves the positions [0, t1) (every particle of the cells next to the left strip boundary),
    // part 3 the positions [t2, count) (the cells next to the right boundary) and part 2 the interior [t1, t2): the host launches
    // 1 and 3, sends the emigrants -- they all come from there -- and launches 2 while the delivery travels.
    // chunk_start != nullptr (pfem2_step_host, chunked): this launch moves the tiles whose FIRST sorted position lies in the segment range
    // of the cells [chunk_lo, chunk_hi) -- whole tiles, because the pass is not in place (a tile shared by two launches would be written
    // twice from the unmoved source).  Rounding the chunk's particle range down to tiles only moves work to an EARLIER chunk's neighbour on
    // the low side and to the LATER chunk on the high side; the upload dependencies are node prefixes that grow with the chunk index, so
    // they still hold.  The launches of one step partition the tiles exactly (the last chunk ends at the last, possibly partial, tile).
    extern __shared__ unsigned char adv_smem_raw[];
    __shared__ int s_mov, s_lost;
    constexpr int warps_per_block = kAdvThreads / 32; // (the launch uses kAdvThreads: compile-time, so the strides below fold)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t smem = (smem_u32(adv_smem_raw) + 1023u) & ~1023u;
    const uint32_t tile0 = smem + (uint32_t)warp * (2 * kAdvTileBytes);
    const uint32_t bar0 = smem + (uint32_t)warps_per_block * (2 * kAdvTileBytes) + (uint32_t)warp * 16;
#if PFEM2_MOVE_GDYN
    // Tiles are handed out through ONE cursor in global memory (zeroed by the launch that opens the advect): every warp works at the
    // front of the pass, where the cell records and nodal values it gathers were fetched moments ago by the warps on the neighbouring
    // tiles.  With a fixed share per warp a warp that falls a few rounds behind that front loses those hits, gets slower (every level
    // of its dependent gathers becomes a DRAM access: +2.2 us per tile measured), falls further behind and ends the pass up to 2.5 ms
    // after everybody else (profiles/r03_summary.md).  Here a slow warp simply takes fewer tiles.  A claim is a group of G consecutive
    // tiles (single tiles: 21.7 instead of 9.7 ms, the one address takes ~0.4 atomics per ns); the first group of a warp is a fixed one.
    // A claim is made at least two iterations before its gather is issued, so the atomic's round trip is off the critical path: the
    // raw claim waits in a register of lane 0, the two pending positions of the warp in shared memory (all lanes read them with one
    // broadcast load each).
    __shared__ int s_next[warps_per_block][2];
#endif
    if (threadIdx.x == 0) s_mov = s_lost = 0;
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        fence_barrier_init();
    }
    __syncthreads();
    // sorted positions [p_lo, n) of this launch, p_lo a multiple of 32; n = the live count, or the (tile-aligned) end of the chunk: one
    // bound serves the loop and the validity of a lane
    int n = ctr->count, p_lo = 0;
    if (chunk_start) {
        if (part) {
            const int t1 = min((__ldg(chunk_start + chunk_lo) + 31) & ~31, n);
            const int t2 = max(min(__ldg(chunk_start + chunk_hi), n) & ~31, t1);
            p_lo = part == 1 ? 0 : (part == 2 ? t1 : t2);
            n = part == 1 ? t1 : (part == 2 ? t2 : n);
        } else {
            p_lo = min(__ldg(chunk_start + chunk_lo), n) & ~31;
            if (chunk_hi < own_hi) n = min(__ldg(chunk_start + chunk_hi), n) & ~31; // (the last chunk of the owned range ends at the count)
        }
    }
    const int stride = (int)gridDim.x * (warps_per_block * 32);
#if PFEM2_MOVE_GDYN
    // CLAIM consecutive tiles per claim: PFEM2_MOVE_GDYN for passes of many tiles per warp, single tiles (host's choice) when the pass
    // has so few tiles that groups would leave warps without work
    constexpr int kClaim = CLAIM, kClaimLog2 = CLAIM == 1 ? 0 : (CLAIM == 2 ? 1 : (CLAIM == 4 ? 2 : 3));
    static_assert((1 << kClaimLog2) == kClaim, "CLAIM: 1, 2, 4 or 8 tiles");
    constexpr int kClaimLast = (kClaim - 1) << 5;            // (position - p_lo) & kClaimLast == kClaimLast: last tile of its group
    const int first = p_lo + (((blockIdx.x * warps_per_block + warp) * kClaim) << 5); // group number `global warp` is the fixed one
    const int dyn0 = p_lo + stride * kClaim;                 // sorted position of claim 0
    auto follow = [&](int pos, int claim) { return ((pos - p_lo) & kClaimLast) != kClaimLast ? pos + 32 : dyn0 + claim * (kClaim * 32); };
#else
    const int first = p_lo + ((blockIdx.x * warps_per_block + warp) << 5);
#endif
    const uint32_t my0 = (uint32_t)lane * 64 + (SWZ ? ((((uint32_t)lane >> 1) & 3) << 4) : 0u);
    // Fetch of a tile = 8 gather4 operations.  The eight lanes 0..7 each hold the four row indices of "their" quarter-kilobyte and issue
    // one gather4 (the hardware instruction is warp-uniform: ptxas wraps a divergent issue in an elect / R2UR loop over the active lanes,
    // so this costs the same 8 UTMALDG as one lane issuing all eight), but the indices arrive with ONE coalesced 128-byte load per tile
    // that is issued an iteration ahead -- in the one-lane form each 16-byte index load sat right in front of the gather4 that consumed it
    // (the asm statements are memory barriers to the compiler): eight DRAM latencies in a row on the warp's critical path.
#if PFEM2_MOVE_GDYN
    const uint32_t slot = smem_u32(&s_next[warp][0]);
    int pending = 0; // lane 0: the claimed group the warp turns to when it has taken the last tile of the one it is in
    if (lane == 0) {
        // the claims the first two pending positions use up (groups of 1 / 2 / >= 4 tiles: 2 / 1 / 0) and the one that waits, in ONE atomic
        pending = atom_add_in_flight(tile_cursor, kClaim == 1 ? 3 : (kClaim == 2 ? 2 : 1));
        const int p1 = kClaim > 1 ? first + 32 : dyn0 + pending * 32;
        const int p2 = kClaim > 2 ? first + 64 : (kClaim == 2 ? dyn0 + pending * 64 : dyn0 + (pending + 1) * 32);
        if (kClaim <= 2) pending += kClaim == 1 ? 2 : 1;
        sts32(slot, (unsigned)p1);
        sts32(slot + 4u, (unsigned)p2);
    }
    __syncwarp();
#endif
    auto load_rows = [&](int base) { // indices of the tile at sorted position `base` (lanes 0..7), zeros otherwise
        int4 r = make_int4(0, 0, 0, 0);
        if (lane < 8 && base < n) r = __ldg(src + ((size_t)base >> 2) + lane);
        return r;
    };
    auto issue = [&](int4 r, uint32_t buf, uint32_t bar) { // all lanes call; expect_tx was armed by lane 0 before the __syncwarp
        if (lane < 8) tma_gather4_rows(buf + (uint32_t)lane * 256u, &gmap, r, bar);
    };
#ifdef PFEM2_MOVE_TRACE
    const int trace_w = blockIdx.x * warps_per_block + warp;
    if (lane == 0 && g_trace_warp) {
        g_trace_warp[4 * trace_w + 0] = trace_now();
        g_trace_warp[4 * trace_w + 2] = trace_smid();
        g_trace_warp[4 * trace_w + 3] = 0;
    }
#endif
    int4 rows = load_rows(first);
    if (first < n) {
        if (lane == 0) mbar_arrive_expect_tx(bar0, kAdvTileBytes);
        __syncwarp();
        issue(rows, tile0, bar0);
    }
#if PFEM2_MOVE_GDYN
    rows = load_rows((int)lds32(slot));
#else
    rows = load_rows(first + stride); // for the fetch at the top of the first iteration
#endif
    uint32_t b = 0, par = 0;
#if PFEM2_MOVE_GDYN
    for (int base = first, nxt_base; base < n; base = nxt_base) {
#else
    for (int base = first; base < n; base += stride) {
#endif
#ifdef PFEM2_MOVE_TRACE
        const unsigned trace_t0 = (unsigned)trace_now();
#endif
        const uint32_t buf = tile0 + b * kAdvTileBytes;
#if PFEM2_MOVE_GDYN
        const int nxt = (int)lds32(slot);
#else
        const int nxt = base + stride;
#endif
        if (lane == 0) {
            bulk_wait_group_read<0>(); // the other buffer's store (previous iteration) has finished reading shared memory
            if (nxt < n) mbar_arrive_expect_tx(bar0 + (b ^ 1) * 8, kAdvTileBytes);
        }
        __syncwarp(); // lanes 1..7 write into that buffer too: behind lane 0's wait
        if (nxt < n) issue(rows, tile0 + (b ^ 1) * kAdvTileBytes, bar0 + (b ^ 1) * 8);
#ifdef PFEM2_MOVE_TRACE
        const unsigned trace_t1 = (unsigned)trace_now();
#endif
        mbar_wait(bar0 + b * 8, par);
#ifdef PFEM2_MOVE_TRACE
        const unsigned trace_t2 = (unsigned)trace_now();
#endif
        par ^= b;
        b ^= 1;
        const int i = base + lane;
        const bool valid = i < n;
        unsigned c = 0;
        double L0 = 0, L1 = 0, L2 = 0;
        int moved = 0;
        bool lost = false;
        if (valid) move_record<WALK, NSUB>(buf + my0, geom, edge_nbr, nbr_off, nbr_idx, V2, dV2, h, substeps, c, L0, L1, L2, moved, lost);
#ifdef PFEM2_MOVE_TRACE
        __syncwarp();
        const unsigned trace_t3 = (unsigned)trace_now();
#endif
        bool live = valid && !lost;
        if (list_emigrant(live, c, i, own_lo, own_hi, rank_bounds, n_ranks, rank_count, emig_idx)) live = false; // row i of the dense output
        // the dense key array of the rank pass: one coalesced 128-byte store per tile (padding lanes, lost particles and emigrants: kLostCell)
        keys[i] = live ? c : kLostCell;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { // dense store into the OTHER buffer at the sorted position of the tile
            tma_store_tile_2d(&tmap_out, 0, base, buf);
            bulk_commit_group();
        }
#if PFEM2_MOVE_GDYN
        {   // indices of the tile after the next; the pending positions move up, the claim of the previous iteration becomes a position
            // and one more tile is claimed
            nxt_base = (int)lds32(slot);
            const int nn = (int)lds32(slot + 4u);
            rows = load_rows(nn);
            __syncwarp(); // (every lane has read the slots)
            if (lane == 0) {
                sts32(slot, (unsigned)nn);
                sts32(slot + 4u, (unsigned)follow(nn, pending));
                if (((nn - p_lo) & kClaimLast) == kClaimLast) pending = atom_add_in_flight(tile_cursor, 1); // (used one group of iterations from now)
            }
            __syncwarp();
        }
#else
        rows = load_rows(nxt + stride); // indices of the tile after the next: in flight during the statistics, consumed at the next loop top
#endif
        // the fast order only needs the number of survivors per cell (stayers + arrivals, accumulate_cell_stats with arrive == nullptr),
        // so the cell the particle started in is not carried through the substep loop
        const unsigned sb = __ballot_sync(0xffffffffu, live), mb = 0u;
        const unsigned lb = __ballot_sync(0xffffffffu, lost);
        const int wm = __reduce_add_sync(0xffffffffu, moved);
        if (lane == 0) {
            if (wm) atomicAdd(&s_mov, wm);
            if (lb) atomicAdd(&s_lost, __popc(lb));
        }
        accumulate_cell_stats(subcell_mode, live, c, L0, L1, L2, sb, mb, lane, n_cells, ppc, level, sub_step, stay, (int *)nullptr,
                                                    cell_mask);
#ifdef PFEM2_MOVE_TRACE
        if (lane == 0 && g_trace_tile) {
            g_trace_tile[base >> 5] = make_uint4(trace_t1 - trace_t0, trace_t2 - trace_t1, trace_t3 - trace_t2, (unsigned)trace_now() - trace_t3);
            g_trace_warp[4 * trace_w + 3] += 1;
        }
#endif
    }
#ifdef PFEM2_MOVE_TRACE
    if (lane == 0 && g_trace_warp) g_trace_warp[4 * trace_w + 1] = trace_now();
#endif
    if (lane == 0) bulk_wait_group<0>();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_mov) atomicAdd(&ctr->movers, s_mov);
        if (s_lost) atomicAdd(&ctr->lost, s_lost);
    }
}

static __global__ void __launch_bounds__(kThreads) k_iota(unsigned *__restrict__ src, const Counters *ctr, int padded_to)
{
    // identity permutation of the live prefix; the pad up to a multiple of 32 names row 0 (a valid row: gathered, never used)
    const int n = ctr->count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < padded_to; i += gridDim.x * blockDim.x) src[i] = i < n ? (unsigned)i : 0u;
}

// nodal velocity, interleaved for the advect pass: V2[n] = (Vx[n], Vy[n])
// begin_ctr != nullptr: thread 0 also opens the advect (the former one-thread kernel k_begin_advect: counters of the pass, the
// emigrant counters and the re-seed cursor reset) -- one launch less per step
static __global__ void __launch_bounds__(kThreads)
k_pack_nodal(int node_lo, int node_hi, NodalVel vel, double2 *__restrict__ V2, Counters *begin_ctr, int capacity, int *rank_count,
             int n_rank_count, int *tail_cursor)
{
    if (begin_ctr && blockIdx.x == 0 && threadIdx.x == 0) {
        begin_ctr->lost = 0;
        begin_ctr->movers = 0;
        begin_ctr->added = 0;
        begin_ctr->capacity = capacity;
        begin_ctr->n_old = begin_ctr->count;
        begin_ctr->n_warps = (begin_ctr->count + 31) >> 5;
        for (int k = 0; k < n_rank_count; ++k) rank_count[k] = 0;
        if (tail_cursor) { // [0]: appended re-seeds; [kTileCursor0 + 32 i]: tile cursor of the i-th move launch of this advect
            *tail_cursor = 0;
            for (int k = 0; k < 32; ++k) tail_cursor[kTileCursor0 + 32 * k] = 0;
        }
    }
    const double *Vx, *Vy;
    vel.resolve(Vx, Vy);
    const int i = node_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < node_hi) V2[i] = make_double2(Vx[i], Vy[i]);
}

} // namespace pfem2
