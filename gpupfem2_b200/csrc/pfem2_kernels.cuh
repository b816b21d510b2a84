// pfem2_kernels.cuh -- the particle-step kernels (sm_100a).
//
// Data layout in HBM: particles are one array of 64-byte records (four 16-byte fields, accessed
// through strided field views), kept PHYSICALLY SORTED BY OWNING CELL after every advect, so cell c owns the segment
// [cell_start[c], cell_start[c+1]).  Mesh data the path reads is repacked once into one 64-byte
// CellGeom record per cell.  All particle counts live in device memory (Counters); kernels are
// grid-stride and read the live count themselves, so a step issues no device->host copy.
#pragma once

#include "pfem2_device.cuh"
#include "pfem2_sort.cuh"
#include "pfem2_tma.cuh"

#include <cuda.h> // CUtensorMap (type only; the encoder is resolved at run time through cudaGetDriverEntryPoint)

namespace pfem2 {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------
// mesh repack: CellGeom[c] = { invJacobi[c], vertices[cells[c].z], cells[c] }
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_build_geom(int n_cells, const double2 *__restrict__ vertices, const unsigned *__restrict__ cells,
             const double *__restrict__ inv_jacobi, CellGeom *__restrict__ geom)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    CellGeom g;
    g.n0 = cells[3 * (size_t)c];
    g.n1 = cells[3 * (size_t)c + 1];
    g.n2 = cells[3 * (size_t)c + 2];
    g.pad = 0;
    const double2 v3 = vertices[g.n2];
    g.v3x = v3.x;
    g.v3y = v3.y;
    g.j0 = inv_jacobi[4 * (size_t)c];
    g.j1 = inv_jacobi[4 * (size_t)c + 1];
    g.j2 = inv_jacobi[4 * (size_t)c + 2];
    g.j3 = inv_jacobi[4 * (size_t)c + 3];
    geom[c] = g;
}

// kCalculateInvJacobi + Matrix2x2::inverse (mesh_2d.cu:21-34, cuda_math.cuh:124-141) as compiled:
//   det = fma(d0, d3, -(d1*d2)) ; inv = 1/det ; { d3*inv, d1*(-inv), d2*(-inv), d0*inv }
__global__ void __launch_bounds__(kThreads)
k_inv_jacobi(int n_cells, const double2 *__restrict__ vertices, const unsigned *__restrict__ cells, double *__restrict__ out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const double2 a = vertices[cells[3 * (size_t)c]], b = vertices[cells[3 * (size_t)c + 1]], z = vertices[cells[3 * (size_t)c + 2]];
    const double d0 = __dsub_rn(a.x, z.x), d1 = __dsub_rn(a.y, z.y);
    const double d2 = __dsub_rn(b.x, z.x), d3 = __dsub_rn(b.y, z.y);
    const double det = __fma_rn(d0, d3, -__dmul_rn(d1, d2));
    const double inv = __ddiv_rn(1.0, det);
    out[4 * (size_t)c + 0] = __dmul_rn(d3, inv);
    out[4 * (size_t)c + 1] = __dmul_rn(d1, -inv);
    out[4 * (size_t)c + 2] = __dmul_rn(d2, -inv);
    out[4 * (size_t)c + 3] = __dmul_rn(d0, inv);
}

// ---------------------------------------------------------------------------------------------
// seeding: kSeedParticlesIntoCell (particle_handler_2d.cu:35-52).  Slot = cell * ppc + sub-cell
// (deterministic; the reference hands out slot blocks by atomicAdd), so the array starts sorted.
// One thread per particle: coalesced SoA stores.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_seed(int n_cells, int own_lo, int own_hi, int ppc, const double2 *__restrict__ vertices, const CellGeom *__restrict__ geom,
       const double *__restrict__ centers, ParticleSoA p, int *__restrict__ cell_start, Counters *ctr)
{
    const long long total = (long long)(own_hi - own_lo) * ppc;
    // cells outside the owned range [own_lo, own_hi) get empty segments
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= n_cells; c += gridDim.x * blockDim.x)
        if (c < own_lo || c >= own_hi) cell_start[c] = c < own_lo ? 0 : (int)total;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = own_lo + (int)(i / ppc), s = (int)(i - (long long)(c - own_lo) * ppc);
        const uint4 nn = __ldg(reinterpret_cast<const uint4 *>(&geom[c].n0));
        const double2 v0 = __ldg(&vertices[nn.x]), v1 = __ldg(&vertices[nn.y]), v2 = __ldg(&vertices[nn.z]);
        const double L0 = __ldg(&centers[3 * s]), L1 = __ldg(&centers[3 * s + 1]), L2 = __ldg(&centers[3 * s + 2]);
        p.pos[i] = make_double2(to_global1(L0, L1, L2, v0.x, v1.x, v2.x), to_global1(L0, L1, L2, v0.y, v1.y, v2.y));
        p.lab[i] = make_double2(L0, L1);
        st_tail(p.tail + i, L2, (unsigned)c, (unsigned)i);
        p.vel[i] = make_double2(0.0, 0.0);
        if (s == 0) cell_start[c] = (int)i;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctr->count = (int)total;
        ctr->live = (int)total;
        ctr->added = 0;
        ctr->lost = 0;
        ctr->movers = 0;
    }
}

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
template <int SUBCELL_MODE, bool MASK64>
__device__ __forceinline__ void accumulate_cell_stats(bool live, unsigned c, double L0, double L1, double L2, unsigned sb, unsigned mb,
                                                      int lane, int n_cells, int ppc, int level, double sub_step,
                                                      int *__restrict__ stay, int *__restrict__ arrive,
                                                      unsigned long long *__restrict__ cell_mask)
{
    unsigned fc = 0xffffffffu;
    unsigned long long bit = 0;
    if (live) {
        const int sub = SUBCELL_MODE == 0 ? subcell_index(L0, L1, L2, level, sub_step) : subcell_index_clamped(L0, L1, L2, level, sub_step);
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
    if (MASK64) hi = __reduce_or_sync(peers, (unsigned)(gbit >> 32));
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

// ---------------------------------------------------------------------------------------------
// advect + locate, all S substeps fused in one pass over the particles
//   kAdvectParticles :54-70, kCheckParticleInCell :117-131, kCheckParticleInNeighbors :133-162.
// The nodal field is frozen inside advectParticles and particles do not interact, so the S substeps
// need no global barrier between them.  Positions, local coordinates and the new cell are
// written in place.  A particle with no accepting cell in own ∪ one-ring is marked lost.
// Per warp of 32 consecutive particles the kernel also emits
//   stay_bits[w]    ballot of particles that end in the cell they started in (they keep their array order),
//   warp_movers[w]  number of particles that changed cell (sorted separately by new cell),
// and accumulates per cell, with one atomic per distinct cell per warp (match_any aggregation):
//   stay[c], arrive[c]  survivors that stayed in / moved into cell c,
//   cell_mask[c]        sub-cell occupancy bits, flat unclamped index like kCountParticlesInSubcells (:173-181).
// ---------------------------------------------------------------------------------------------
template <int SUBCELL_MODE, bool WALK, bool MASK64, int NSUB>
__global__ void __launch_bounds__(kThreads, 4)
k_advect_locate(ParticleSoA p, const CellGeom *__restrict__ geom, const int4 *__restrict__ edge_nbr,
                const int *__restrict__ nbr_off, const int *__restrict__ nbr_idx, NodalVel vel, double h, int substeps,
                int n_cells, int ppc, int level, double sub_step, Counters *ctr, unsigned *__restrict__ stay_bits,
                int *__restrict__ warp_movers, int *__restrict__ stay, int *__restrict__ arrive,
                unsigned long long *__restrict__ cell_mask, int do_count, const double *__restrict__ dvx,
                const double *__restrict__ dvy)
{
    const double *__restrict__ Vx, *__restrict__ Vy;
    {
        const double *a, *b;
        vel.resolve(a, b);
        Vx = a;
        Vy = b;
    }
    const int n = ctr->count;
    const int lane = threadIdx.x & 31;
    __shared__ int s_mov, s_lost;
    if (threadIdx.x == 0) s_mov = s_lost = 0;
    __syncthreads();
    // Particle records are streamed with cache-streaming (evict-first) 128-bit loads / stores so that they do not push
    // the re-used cell records and nodal velocities out of L1.
    const int stride = gridDim.x * blockDim.x;
    // warp-uniform loop (the aggregation below uses full-mask warp intrinsics); base is a multiple of 32
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += stride) {
        const int i = base + lane;
        const bool valid = i < n;
        unsigned c0 = 0, c = 0;
        double L0 = 0, L1 = 0, L2 = 0;
        int moved = 0; // substeps in which the particle left its cell
        bool lost = false;
        if (valid) {
            int4 *rec = reinterpret_cast<int4 *>(p.records() + i);
            const int4 r0 = __ldcs(rec), r1 = __ldcs(rec + 1), r2 = __ldcs(rec + 2);
            const unsigned id = (unsigned)r2.w;
            c0 = c = (unsigned)r2.z;
            double x = __hiloint2double(r0.y, r0.x);
            double y = __hiloint2double(r0.w, r0.z);
            L0 = __hiloint2double(r1.y, r1.x);
            L1 = __hiloint2double(r1.w, r1.z);
            L2 = __hiloint2double(r2.y, r2.x);
            // cell record and its six nodal velocities stay in registers while the particle stays in the cell
            CellGeom g = load_geom(geom, c);
            int4 e = __ldg(edge_nbr + c); // prefetched with the cell record so the first walk hop has no extra dependent load
            double ax0 = __ldg(Vx + g.n0), ax1 = __ldg(Vx + g.n1), ax2 = __ldg(Vx + g.n2);
            double ay0 = __ldg(Vy + g.n0), ay1 = __ldg(Vy + g.n1), ay2 = __ldg(Vy + g.n2);
            if (dvx) { // pending velocity correction, with the cell / local position the particle had at the correct call
                const int4 r3 = __ldcs(rec + 3);
                const double vx = __dadd_rn(__hiloint2double(r3.y, r3.x),
                                            interp3(L0, L1, L2, __ldg(dvx + g.n0), __ldg(dvx + g.n1), __ldg(dvx + g.n2)));
                const double vy = __dadd_rn(__hiloint2double(r3.w, r3.z),
                                            interp3(L0, L1, L2, __ldg(dvy + g.n0), __ldg(dvy + g.n1), __ldg(dvy + g.n2)));
                __stcs(rec + 3, make_int4(__double2loint(vx), __double2hiint(vx), __double2loint(vy), __double2hiint(vy)));
            }
            const int nsub = NSUB > 0 ? NSUB : substeps;
            // kept rolled on purpose: unrolling lets stayers run ahead into the next substep's code while the movers of the
            // warp are still in the walk, which costs more issue slots than it saves (measured 14.0 vs 10.8 ms)
#pragma unroll 1
            for (int s = 0; s < nsub; ++s) {
                // kAdvectParticles: velocity from the STORED local position and cell
                const double ux = interp3(L0, L1, L2, ax0, ax1, ax2);
                const double uy = interp3(L0, L1, L2, ay0, ay1, ay2);
                x = __fma_rn(ux, h, x);
                y = __fma_rn(uy, h, y);
                // own cell first (wins even if a neighbour would also accept); L is overwritten: every path below
                // either replaces it again or never reads it (a lost particle stores no local position)
                to_local(g, x, y, L0, L1, L2);
                if (inside_unit(L0, L1, L2)) continue;
                ++moved;
                if (!locate_mover<WALK>(geom, edge_nbr, nbr_off, nbr_idx, g, e, c, x, y, L0, L1, L2)) {
                    lost = true;
                    break;
                }
                if (s + 1 < nsub) {
                    ax0 = __ldg(Vx + g.n0); ax1 = __ldg(Vx + g.n1); ax2 = __ldg(Vx + g.n2);
                    ay0 = __ldg(Vy + g.n0); ay1 = __ldg(Vy + g.n1); ay2 = __ldg(Vy + g.n2);
                }
            }
            __stcs(rec, make_int4(__double2loint(x), __double2hiint(x), __double2loint(y), __double2hiint(y)));
            if (lost) {
                st_cell(p.tail + i, kLostCell);
            } else {
                __stcs(rec + 1, make_int4(__double2loint(L0), __double2hiint(L0), __double2loint(L1), __double2hiint(L1)));
                __stcs(rec + 2, make_int4(__double2loint(L2), __double2hiint(L2), (int)c, (int)id));
            }
        }
        const bool live = valid && !lost;
        const bool stays = live && c == c0;
        const unsigned sb = __ballot_sync(0xffffffffu, stays);
        const unsigned mb = __ballot_sync(0xffffffffu, live && !stays);
        const unsigned lb = __ballot_sync(0xffffffffu, lost);
        const int wm = __reduce_add_sync(0xffffffffu, moved);
        if (lane == 0) { // the two statistics counters: one shared-memory atomic per warp and iteration
            if (wm) atomicAdd(&s_mov, wm);
            if (lb) atomicAdd(&s_lost, __popc(lb));
            if (stay_bits) { // only the stable-order path consumes these
                stay_bits[base >> 5] = sb;
                warp_movers[base >> 5] = __popc(mb);
            }
        }
        if (do_count)
            accumulate_cell_stats<SUBCELL_MODE, MASK64>(live, c, L0, L1, L2, sb, mb, lane, n_cells, ppc, level, sub_step, stay, arrive,
                                                        cell_mask);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_mov) atomicAdd(&ctr->movers, s_mov);
        if (s_lost) atomicAdd(&ctr->lost, s_lost);
    }
}

// destination rank of a cell: bounds[r] <= cell < bounds[r + 1]   (n_ranks <= 64)
__device__ __forceinline__ int rank_of_cell(unsigned c, const int *__restrict__ bounds, int n_ranks)
{
    int r = 0;
    while (r + 1 < n_ranks && (int)c >= bounds[r + 1]) ++r;
    return r;
}

// ---------------------------------------------------------------------------------------------
// advect + locate, TMA-tiled variant (default).  Same arithmetic and the same outputs as k_advect_locate, different
// data movement: the one-lane-per-record 128-bit global loads / stores of k_advect_locate touch 16 cache lines per warp
// instruction (64-byte record stride), and with the gathered cell / nodal loads of the walk on top the L1 data pipe
// -- not HBM, not instruction issue -- was the busiest unit of that kernel (l1tex__data_pipe_lsu_wavefronts 76 %,
// profiles/r01_adv2_summary.md).  Here every warp owns two 2 KB tiles of shared memory and lets the copy engine move
// whole 32-record tiles global <-> shared (cp.async.bulk.tensor, 64-byte swizzle so that "lane r reads record r" is
// bank-conflict free): the LSU sees four shared loads and four shared stores per lane and iteration and nothing else
// for the particle state; the next tile lands while the current one is computed (ping-pong, one mbarrier per tile).
// Nodal velocities come interleaved (double2 per node: one 128-bit gather per node instead of two 64-bit ones).
// ---------------------------------------------------------------------------------------------
constexpr int kAdvTileBytes = 32 * (int)sizeof(ParticleRec); // one warp tile
constexpr size_t advect_tma_smem_bytes(int threads) { return (size_t)(threads / 32) * (2 * kAdvTileBytes + 2 * sizeof(uint64_t)) + 1024; }

// Block size / resident blocks per SM of the TMA-tiled move pass, compile-time.  Swept on B200 (channel16m, ms per pass):
// 256x4 8.42, 128x8 8.38, 64x16 8.45 (all 64 registers, 32 warps per SM); 192x5 8.67 (30 warps); 128x7 8.97 (72 registers, 28 warps);
// 128x6 9.14 and 256x3 9.24 (80 registers, no spills, 24 warps): the pass is latency-bound, resident warps beat registers.
#ifndef PFEM2_ADV_THREADS
#define PFEM2_ADV_THREADS 256
#define PFEM2_ADV_MINB 4
#endif
constexpr int kAdvThreads = PFEM2_ADV_THREADS, kAdvBlocksPerSM = PFEM2_ADV_MINB;
// FAST (the fast order, i.e. not pfem2_options.stable_order): only the number of survivors per cell is needed (stayers + arrivals are
// summed, accumulate_cell_stats with arrive == nullptr, and no stay bits are written), so the cell a particle started in is not carried
// through the substep loop -- one register less at the 64-register cap: ptxas then allocates the S = 3 form without a single spill
// (12 / 20 spilled bytes before; on this LSU-bound kernel 16 more spilled bytes had cost 0.47 ms, profiles/r01d_summary.md §3).
template <int SUBCELL_MODE, bool WALK, bool MASK64, int NSUB, bool FAST>
__global__ void __launch_bounds__(kAdvThreads, kAdvBlocksPerSM)
k_advect_locate_tma(const __grid_constant__ CUtensorMap tmap, const CellGeom *__restrict__ geom, const int4 *__restrict__ edge_nbr,
                    const int *__restrict__ nbr_off, const int *__restrict__ nbr_idx, const double2 *__restrict__ V2, double h,
                    int substeps, int n_cells, int ppc, int level, double sub_step, Counters *ctr, unsigned *__restrict__ stay_bits,
                    int *__restrict__ warp_movers, int *__restrict__ stay, int *__restrict__ arrive,
                    unsigned long long *__restrict__ cell_mask, int do_count, const double2 *__restrict__ dV2, int own_lo, int own_hi,
                    const int *__restrict__ rank_bounds, int n_ranks, int *__restrict__ rank_count, unsigned *__restrict__ emig_idx,
                    const int *__restrict__ chunk_start, int chunk_lo, int chunk_hi)
{
    // chunk_start != nullptr: only the particles of the cells [chunk_lo, chunk_hi) are moved (chunk_start = the segment table
    // of the sorted array): pfem2_step_host launches the pass chunk by chunk while the nodal field is still arriving over PCIe.
    // Multi-GPU (emig_idx != nullptr): a particle whose new cell lies outside the owned range [own_lo, own_hi) is an
    // emigrant.  It is written back like everybody else (the pack kernel removes it), but it is kept out of the per-cell
    // statistics, its array index is appended to emig_idx and rank_count[destination] / rank_count[n_ranks] (total) are
    // bumped, so that neither a counting nor a search pass over the whole array is needed afterwards.
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
        const uint32_t sa0 = buf + my0, sa1 = sa0 ^ 16u, sa2 = sa0 ^ 32u, sa3 = sa0 ^ 48u;
        const int base = tile << 5;
        const int i = base + lane;
        const bool valid = i >= p_lo && i < n;
        unsigned c0 = 0, c = 0;
        double L0 = 0, L1 = 0, L2 = 0;
        int moved = 0; // substeps in which the particle left its cell
        bool lost = false;
        if (valid) {
            const int4 r0 = lds128(sa0), r1 = lds128(sa1), r2 = lds128(sa2);
            c = (unsigned)r2.z;
            if (!FAST) c0 = c;
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
        // hand the tile back: generic-proxy writes -> visible to the async proxy -> one lane issues the store
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_tile_2d(&tmap, 0, base, buf);
            bulk_commit_group();
        }
        bool live = valid && !lost;
        if (emig_idx && live && ((int)c < own_lo || (int)c >= own_hi)) { // rare: a handful per tile row at a strip interface
            const int slot = atomicAdd(rank_count + n_ranks, 1);
            emig_idx[slot] = (unsigned)i;
            atomicAdd(rank_count + rank_of_cell(c, rank_bounds, n_ranks), 1);
            live = false;
        }
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
            accumulate_cell_stats<SUBCELL_MODE, MASK64>(live, c, L0, L1, L2, sb, mb, lane, n_cells, ppc, level, sub_step, stay, arrive,
                                                        cell_mask);
    }
    if (lane == 0) bulk_wait_group<0>(); // the last stores still read this block's shared memory
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_mov) atomicAdd(&ctr->movers, s_mov);
        if (s_lost) atomicAdd(&ctr->lost, s_lost);
    }
}

// nodal velocity, interleaved for the advect pass: V2[n] = (Vx[n], Vy[n])
__global__ void __launch_bounds__(kThreads) k_pack_nodal(int node_lo, int node_hi, NodalVel vel, double2 *__restrict__ V2)
{
    const double *Vx, *Vy;
    vel.resolve(Vx, Vy);
    const int i = node_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < node_hi) V2[i] = make_double2(Vx[i], Vy[i]);
}

// smallest / largest node id of the cells [cell_lo, cell_hi): out[0] = min (start INT_MAX), out[1] = max (start -1)
__global__ void __launch_bounds__(kThreads) k_node_minmax(int cell_lo, int cell_hi, const CellGeom *__restrict__ geom, int *out)
{
    const int c = cell_lo + blockIdx.x * blockDim.x + threadIdx.x;
    int lo = 0x7fffffff, hi = -1;
    if (c < cell_hi) {
        const unsigned a = geom[c].n0, b = geom[c].n1, d = geom[c].n2;
        lo = (int)min(a, min(b, d));
        hi = (int)max(a, max(b, d));
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((threadIdx.x & 31) == 0 && hi >= 0) {
        atomicMin(out, lo);
        atomicMax(out + 1, hi);
    }
}

// ---- pfem2_step_host pipeline plan (once per handle): how far the cell numbering couples distant cells, and which
// node ranges a chunk of cells depends on ----
// band[0] = max |neighbour - cell| over the one-ring lists: a particle changes its cell index by at most that per substep
__global__ void __launch_bounds__(kThreads) k_band_width(int n_cells, const int *__restrict__ nbr_off, const int *__restrict__ nbr_idx, int *band)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    int w = 0;
    if (c < n_cells)
        for (int k = __ldg(nbr_off + c); k < __ldg(nbr_off + c + 1); ++k) w = max(w, abs(__ldg(nbr_idx + k) - c));
    w = __reduce_max_sync(0xffffffffu, w);
    if ((threadIdx.x & 31) == 0 && w > 0) atomicMax(band, w);
}
// For the K chunks [cb[j], cb[j+1]) of the cell range and a reach of `ext` cells (substeps x band width):
//   up_need[j]  = 1 + the largest node id of any cell a particle of chunk j can visit (cells [cb[j]-ext, cb[j+1]+ext))
//   dn_ready[j] = the smallest node id of any cell behind chunk j (cells >= cb[j+1]): nodes below it are complete once the
//                 chunks 0..j have been projected
__global__ void __launch_bounds__(kThreads)
k_chunk_node_ranges(int n_cells, const CellGeom *__restrict__ geom, int K, const int *__restrict__ cb, int ext, int *__restrict__ up_need,
                    int *__restrict__ dn_ready)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = c < n_cells;
    int mn = 0x7fffffff, mx = -1;
    if (valid) {
        const uint4 nn = __ldg(reinterpret_cast<const uint4 *>(&geom[c].n0));
        mn = (int)min(nn.x, min(nn.y, nn.z));
        mx = (int)max(nn.x, max(nn.y, nn.z));
    }
    for (int j = 0; j < K; ++j) {
        const long long lo = (long long)__ldg(cb + j) - ext, hi = (long long)__ldg(cb + j + 1) + ext;
        const int a = __reduce_max_sync(0xffffffffu, (valid && c >= lo && c < hi) ? mx + 1 : 0);
        const int b = __reduce_min_sync(0xffffffffu, (valid && c >= __ldg(cb + j + 1)) ? mn : 0x7fffffff);
        if ((threadIdx.x & 31) == 0) {
            if (a > 0) atomicMax(up_need + j, a);
            if (b != 0x7fffffff) atomicMin(dn_ready + j, b);
        }
    }
}

// movers -> (new cell, array index) pairs in array order, at the positions the scan of warp_movers assigns
__global__ void __launch_bounds__(kThreads)
k_emit_movers(ParticleSoA p, Counters *ctr, const unsigned *__restrict__ stay_bits,
              const int *__restrict__ warp_mover_base, unsigned *__restrict__ keys, unsigned *__restrict__ vals)
{
    const int n = ctr->count;
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        const unsigned sb = __ldg(stay_bits + (base >> 5));
        unsigned c = kLostCell;
        if (i < n && !((sb >> lane) & 1u)) c = ld_cell(p.tail + i);
        const bool mover = c != kLostCell;
        const unsigned mb = __ballot_sync(0xffffffffu, mover);
        if (mover) {
            const int pos = __ldg(warp_mover_base + (base >> 5)) + __popc(mb & ((1u << lane) - 1));
            keys[pos] = c;
            vals[pos] = (unsigned)i;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ctr->n_movers = warp_mover_base[ctr->n_warps]; // scan total
}

// upload(): an arbitrary (unsorted) array is handled as "everybody is a mover"
__global__ void __launch_bounds__(kThreads)
k_all_movers(ParticleSoA p, int n_cells, Counters *ctr, unsigned *__restrict__ keys, unsigned *__restrict__ vals,
             int *__restrict__ arrive, int *__restrict__ n_movers)
{
    const int n = ctr->count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned c = ld_cell(p.tail + i);
        if (c >= (unsigned)n_cells) c = (unsigned)n_cells; // lost / handed-over particles sort behind every cell and are dropped
        keys[i] = c;
        vals[i] = (unsigned)i;
        if (arrive && c < (unsigned)n_cells) atomicAdd(arrive + c, 1);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_movers = n;
}

// ---------------------------------------------------------------------------------------------
// distribution check, planning part: kCountParticlesToBeAdded (:183-195)
//   packed[c] = (stay + arrive + missing)(c) | arrive(c) << 32   (one 64-bit scan yields both prefix sums)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_plan_cells(int n_cells, int own_lo, int own_hi, int ppc, int reseed, const int *__restrict__ stay, const int *__restrict__ arrive,
             const unsigned long long *__restrict__ cell_mask, unsigned long long *__restrict__ packed, Counters *ctr)
{
    // only the owned cell range [own_lo, own_hi) can hold particles (the whole mesh on a single GPU)
    const int c = own_lo + blockIdx.x * blockDim.x + threadIdx.x;
    int missing = 0;
    (void)n_cells;
    if (c < own_hi) {
        const unsigned long long full = ppc >= 64 ? ~0ull : ((1ull << ppc) - 1ull);
        missing = (reseed && c >= own_lo && c < own_hi) ? ppc - __popcll(cell_mask[c] & full) : 0; // only owned cells are re-seeded
        const unsigned a = (unsigned)arrive[c];
        packed[c] = (unsigned long long)((unsigned)stay[c] + a + (unsigned)missing) | ((unsigned long long)a << 32);
    }
    // total number of re-seeded particles
    const unsigned any = __ballot_sync(0xffffffffu, missing != 0);
    if (any) {
        int m = missing;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m += __shfl_xor_sync(0xffffffffu, m, d);
        if ((threadIdx.x & 31) == 0) atomicAdd(&ctr->added, m);
    }
}

__global__ void k_plan_finish(int own_hi, const unsigned long long *__restrict__ packed_start, Counters *ctr)
{
    const unsigned long long t = packed_start[own_hi];
    const long long total = (long long)(unsigned)(t & 0xffffffffull);
    ctr->live = (int)total - ctr->added;
    if (total > ctr->capacity) {
        ctr->overflow = 1;
        ctr->count = 0; // nothing valid to iterate over; the host reports PFEM2_ECAPACITY
    } else {
        ctr->count = (int)total;
    }
}

__device__ __forceinline__ void copy_particle(const ParticleSoA &src, int s, const ParticleSoA &dst, int d)
{
    const double2 a = src.pos[s], b = src.lab[s], v = src.vel[s];
    const int4 t = *reinterpret_cast<const int4 *>(src.tail + s);
    dst.pos[d] = a;
    dst.lab[d] = b;
    *reinterpret_cast<int4 *>(dst.tail + d) = t;
    dst.vel[d] = v;
}

// stayers keep their relative order: destination = new segment start + number of stayers before it in its old
// segment (popcount over the stay bits from the old segment start).  n_old = array length before the advect.
__global__ void __launch_bounds__(kThreads)
k_scatter_stayers(ParticleSoA src, ParticleSoA dst, const int *__restrict__ n_old_ptr,
                  const unsigned *__restrict__ stay_bits, const int *__restrict__ old_start,
                  const unsigned long long *__restrict__ packed_start, const Counters *ctr)
{
    if (ctr->overflow) return;
    const int n = *n_old_ptr;
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x) {
        const unsigned sb = __ldg(stay_bits + (base >> 5));
        if (!((sb >> lane) & 1u)) continue;
        const int i = base + lane;
        const unsigned c = ld_cell(src.tail + i);
        const int s0 = __ldg(old_start + c);
        int rank;
        if (s0 >= base) {
            rank = __popc(sb & ((1u << lane) - 1) & ~((1u << (s0 - base)) - 1));
        } else {
            rank = __popc(sb & ((1u << lane) - 1));
            int w = s0 >> 5;
            rank += __popc(__ldg(stay_bits + w) & ~((1u << (s0 & 31)) - 1));
            for (++w; w < (base >> 5); ++w) rank += __popc(__ldg(stay_bits + w));
        }
        const int d = (int)(unsigned)(packed_start[c] & 0xffffffffull) + rank;
        copy_particle(src, i, dst, d);
    }
}

// Fast reorder (default): one counting-sort scatter of ALL survivors in array order.  Lanes of a warp that go to
// the same cell form a group (match_any); the group leader reserves a contiguous run behind the cell's cursor with
// one atomicAdd, so each group writes one contiguous run per array.  The order of runs inside a cell depends on the
// order the atomics retire (owner cells and the particle SET are exact; only the slot order within a cell and hence
// the last bits of the projection sums can differ between runs).  pfem2_options.stable_order selects the
// deterministic stayer / sorted-mover path below instead.
// Register-staged variant (default: measured 10.5 ms vs 11.5 ms for the TMA variant on channel16m): four particles per
// thread; all loads, then all atomics, are in flight together.
__global__ void __launch_bounds__(kThreads)
k_scatter_all_regs(ParticleSoA src, ParticleSoA dst, const int *__restrict__ n_old_ptr, int *__restrict__ cursor, const Counters *ctr)
{
    if (ctr->overflow) return;
    const int n = *n_old_ptr;
    const int lane = threadIdx.x & 31;
    constexpr int U = 6; // particles per thread per iteration (measured: 2 -> 11.9 ms, 4 -> 10.7, 6 -> 10.3 on channel16m)
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp_global * (32 * U); base < n; base += warps_total * (32 * U)) {
        double2 a[U], b[U], v[U];
        int4 t[U];
        unsigned c[U], peers[U];
        int run[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * 32 + lane;
            c[u] = kLostCell;
            if (i < n) {
                t[u] = *reinterpret_cast<const int4 *>(src.tail + i);
                c[u] = (unsigned)t[u].z;
                a[u] = src.pos[i];
                b[u] = src.lab[i];
                v[u] = src.vel[i];
            }
        }
        // cursor[c] starts at the new segment start of cell c: one atomicAdd per (warp, cell) group reserves a run
#pragma unroll
        for (int u = 0; u < U; ++u) {
            peers[u] = __match_any_sync(0xffffffffu, c[u]);
            run[u] = 0;
            if (c[u] != kLostCell && (peers[u] & ((1u << lane) - 1)) == 0) run[u] = atomicAdd(cursor + c[u], __popc(peers[u]));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int leader = __ffs(peers[u]) - 1;
            const int r = __shfl_sync(0xffffffffu, run[u], leader);
            if (c[u] == kLostCell) continue;
            const int d = r + __popc(peers[u] & ((1u << lane) - 1));
            dst.pos[d] = a[u];
            dst.lab[d] = b[u];
            *reinterpret_cast<int4 *>(dst.tail + d) = t[u];
            dst.vel[d] = v[u];
        }
    }
}

// Quad-cooperative variant (default): FOUR lanes per record, lane (4 q + f) moves the 16-byte field f of record q of
// the warp's 8-record group.  A warp-wide 128-bit load then covers 512 contiguous bytes (4 L1 wavefronts for 8 records)
// and the four lanes of a quad store one whole 64-byte record (1 wavefront), where the one-lane-per-record form needs
// 16 wavefronts per load instruction and 4 per stored record: the L1 data pipe, not HBM, bounded that form
// (l1tex__data_pipe_lsu_wavefronts 67 % at 5.0 TB/s, profiles/r01_final_summary.md).  U groups are in flight per warp.
template <int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_scatter_all_quads(ParticleSoA src, ParticleSoA dst, const int *__restrict__ n_old_ptr, int *__restrict__ cursor, const Counters *ctr)
{
    if (ctr->overflow) return;
    const int n = *n_old_ptr;
    const int lane = threadIdx.x & 31;
    const int f = lane & 3, q = lane >> 2;
    const unsigned lt = (1u << lane) - 1;
    // U 8-record groups per warp and iteration: all loads, then all atomics, are in flight together
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long warps_total = ((long long)gridDim.x * blockDim.x) >> 5;
    const int4 *__restrict__ in = reinterpret_cast<const int4 *>(src.records());
    int4 *__restrict__ out = reinterpret_cast<int4 *>(dst.records());
    for (long long base = warp_global * (8 * U); base < n; base += warps_total * (8 * U)) {
        int4 v[U];
        unsigned peers[U]; // lanes of the records that go to the same cell (whole quads); 0 = lost / beyond the end
        int run[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = base + u * 8 + q;
            v[u] = make_int4(0, 0, (int)kLostCell, 0);
            if (r < n) v[u] = __ldcs(in + r * 4 + f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned c = (unsigned)__shfl_sync(0xffffffffu, v[u].z, (lane & ~3) | 2); // the record's cell sits in field 2 (tail)
            peers[u] = __match_any_sync(0xffffffffu, c);
            run[u] = 0;
            if (c == kLostCell) peers[u] = 0;
            else if ((peers[u] & lt) == 0) run[u] = atomicAdd(cursor + c, __popc(peers[u]) >> 2);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r0 = __shfl_sync(0xffffffffu, run[u], peers[u] ? __ffs(peers[u]) - 1 : 0);
            if (peers[u] == 0) continue;
            const long long d = r0 + (__popc(peers[u] & lt) >> 2);
            out[d * 4 + f] = v[u];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Trailing projection (pfem2_options.fuse_project = 1, EXPERIMENTAL, off by default): the re-sort scatter and the projection's cell pass run
// CONCURRENTLY, the second one trailing the first so closely that it reads the freshly written records from L2 instead
// of HBM (-64 B per particle-step of DRAM traffic).
//   producer  k_scatter_quads_ordered   the quad scatter with a block-contiguous schedule: in its k-th iteration block b
//             moves the records of "slab" k * gridDim + b (8 warps x U groups x 8 records) and counts in; the block that
//             completes iteration k raises prog[0] to k + 1 (release): all slabs below prog[0] * gridDim are done.
//   consumer  k_reseed_project_trailing  persistent blocks claim chunks of kTrailCells cells in ascending order, wait
//             until every source record that can still land in the chunk has been moved (a particle's cell index changes by
//             at most `reach` = band width of the one-ring lists x substeps, so the records of the old cells below
//             chunk_end + reach suffice), then re-seed the chunk's cells and reduce their segments to the nine sums.
// Neither kernel waits for the other's blocks to be scheduled: the producer never waits at all and its grid is sized to
// be fully resident, so the consumer's spin always ends; a watchdog turns a would-be hang into an error flag.
// Measured on channel16m (profiles/r01b_summary.md §8): producer 7.6 ms + consumer tail 1.3 ms = 9.1 ms against 6.1 + 2.8 ms for
// the separate passes -- break-even, because the ordered producer pays for its publish fence (6.5 vs 5.5 ms alone), is slowed
// further by the co-resident consumer, and the consumer (one 256-thread block per SM next to two producer blocks) cannot quite
// keep up.  Kept opt-in as the starting point for the persistent fused kernel of DESIGN.md §10.
// ---------------------------------------------------------------------------------------------
constexpr int kTrailCells = 256; // cells per consumer chunk (one thread per cell in the re-seed part)

template <int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_scatter_quads_ordered(ParticleSoA src, ParticleSoA dst, const int *__restrict__ n_old_ptr, int *__restrict__ cursor, const Counters *ctr,
                        int *prog)
{
    if (ctr->overflow) return;
    const int n = *n_old_ptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = lane & 3, q = lane >> 2;
    const unsigned lt = (1u << lane) - 1;
    constexpr int kSlab = (kThreads / 32) * 8 * U; // records per block iteration
    const int4 *__restrict__ in = reinterpret_cast<const int4 *>(src.records());
    int4 *__restrict__ out = reinterpret_cast<int4 *>(dst.records());
    // every block runs the same number of iterations (possibly with an empty last slab) so that "iteration k complete" is
    // simply "gridDim blocks have counted in"
    const long long n_slabs = ((long long)n + kSlab - 1) / kSlab;
    const int iters = (int)((n_slabs + gridDim.x - 1) / gridDim.x);
    int *iter_cnt = prog + 1; // prog[0] = completed iterations (monotone), prog[1 + k] = blocks that finished iteration k
    for (int k = 0; k < iters; ++k) {
        const long long slab = (long long)k * gridDim.x + blockIdx.x;
        const long long base = slab * kSlab + (long long)warp * (8 * U);
        int4 v[U];
        unsigned peers[U];
        int run[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = base + u * 8 + q;
            v[u] = make_int4(0, 0, (int)kLostCell, 0);
            if (r < n) v[u] = __ldcs(in + r * 4 + f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned c = (unsigned)__shfl_sync(0xffffffffu, v[u].z, (lane & ~3) | 2);
            peers[u] = __match_any_sync(0xffffffffu, c);
            run[u] = 0;
            if (c == kLostCell) peers[u] = 0;
            else if ((peers[u] & lt) == 0) run[u] = atomicAdd(cursor + c, __popc(peers[u]) >> 2);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r0 = __shfl_sync(0xffffffffu, run[u], peers[u] ? __ffs(peers[u]) - 1 : 0);
            if (peers[u] == 0) continue;
            const long long d = r0 + (__popc(peers[u] & lt) >> 2);
            out[d * 4 + f] = v[u];
        }
        // publish: the barrier orders the block's stores before thread 0's fence (cumulativity), one fence per slab.
        // (Measured alternatives: a fence in every thread 10.6 ms; per-warp counting with the fence deferred to the next
        // iteration 7.8 ms; this form 6.5 ms; the unordered k_scatter_all_quads 5.5 ms.)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(iter_cnt + k, 1) == (int)gridDim.x - 1) atomicMax(prog, k + 1);
        }
    }
}

template <int G>
__global__ void __launch_bounds__(kThreads, 4)
k_reseed_project_trailing(int n_cells, int ppc, int reach, int slab_records, int n_producers, const int *prog, int *next_chunk,
                          const int *__restrict__ old_start, const double2 *__restrict__ vertices, const CellGeom *__restrict__ geom,
                          const double *__restrict__ centers, NodalVel vel, const unsigned long long *__restrict__ cell_mask,
                          const int *__restrict__ stay, const int *__restrict__ arrive, const unsigned long long *__restrict__ packed_start,
                          ParticleSoA dst, int *__restrict__ cell_start, double *__restrict__ partial, Counters *ctr)
{
    __shared__ int s_chunk;
    const double *Vx, *Vy;
    vel.resolve(Vx, Vy);
    const int n_chunks = (n_cells + kTrailCells - 1) / kTrailCells;
    const bool dead = ctr->overflow != 0; // the plan did not fit: only the segment table is written
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(next_chunk, 1);
        __syncthreads();
        const int chunk = s_chunk;
        if (chunk >= n_chunks) break;
        const int c0 = chunk * kTrailCells, c1 = min(c0 + kTrailCells, n_cells);
        if (!dead) {
            // wait until all source records of the old cells [0, c1 + reach) have been moved
            const long long need_rec = __ldg(old_start + min((long long)c1 + reach, (long long)n_cells));
            const long long need_slab = (need_rec + slab_records - 1) / slab_records; // slabs [0, need_slab) must be done
            const int need_iter = (int)((need_slab + n_producers - 1) / n_producers);  // producer iterations that cover them
            if (threadIdx.x == 0) {
                const long long t0 = clock64();
                for (;;) {
                    int done;
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(done) : "l"(prog) : "memory");
                    if (done >= need_iter) break;
                    if (clock64() - t0 > (1ll << 32)) { // ~2 s without progress: fail loudly instead of hanging
                        atomicExch(&ctr->overflow, 2);
                        break;
                    }
                    __nanosleep(1000);
                }
            }
            __syncthreads();
        }
        // ---- re-seed (k_reseed) + segment table, one thread per cell ----
        {
            const int c = c0 + threadIdx.x;
            if (c1 == n_cells && threadIdx.x == 0) cell_start[n_cells] = (int)(unsigned)(packed_start[n_cells] & 0xffffffffull);
            if (c < c1) {
                const int start = (int)(unsigned)(packed_start[c] & 0xffffffffull);
                cell_start[c] = start;
                if (!dead) {
                    const int live = stay[c] + arrive[c];
                    const int missing = (int)(unsigned)(packed_start[c + 1] & 0xffffffffull) - start - live;
                    if (missing > 0) {
                        const uint4 nn = __ldg(reinterpret_cast<const uint4 *>(&geom[c].n0));
                        const double2 v0 = __ldg(&vertices[nn.x]), v1 = __ldg(&vertices[nn.y]), v2 = __ldg(&vertices[nn.z]);
                        const double ax0 = __ldg(Vx + nn.x), ax1 = __ldg(Vx + nn.y), ax2 = __ldg(Vx + nn.z);
                        const double ay0 = __ldg(Vy + nn.x), ay1 = __ldg(Vy + nn.y), ay2 = __ldg(Vy + nn.z);
                        const unsigned long long mask = cell_mask[c];
                        int d = start + live;
                        for (int s = 0; s < ppc; ++s) {
                            if ((mask >> s) & 1ull) continue;
                            const double L0 = __ldg(&centers[3 * s]), L1 = __ldg(&centers[3 * s + 1]), L2 = __ldg(&centers[3 * s + 2]);
                            dst.pos[d] = make_double2(to_global1(L0, L1, L2, v0.x, v1.x, v2.x), to_global1(L0, L1, L2, v0.y, v1.y, v2.y));
                            dst.lab[d] = make_double2(L0, L1);
                            st_tail(dst.tail + d, L2, (unsigned)c, (unsigned)d);
                            dst.vel[d] = make_double2(interp3(L0, L1, L2, ax0, ax1, ax2), interp3(L0, L1, L2, ay0, ay1, ay2));
                            ++d;
                        }
                    }
                }
            }
        }
        if (dead) continue;
        __threadfence_block();
        __syncthreads();
        // ---- project (k_project_cells<G>): G lanes per cell, same summation order, records read through L2 (ld.cg) ----
        const int lane = threadIdx.x & (G - 1);
        for (int cw = c0 + (threadIdx.x / G); cw - (threadIdx.x / G) < c1; cw += kThreads / G) {
            const int c = cw;
            const bool valid = c < c1;
            const int b = valid ? (int)(unsigned)(packed_start[c] & 0xffffffffull) : 0;
            const int e = valid ? (int)(unsigned)(packed_start[c + 1] & 0xffffffffull) : 0;
            double acc[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[k] = 0.0;
            int i = b + lane;
            for (; i + 3 * G < e; i += 4 * G) {
                double2 l4[4], v4[4];
                double z4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    l4[u] = __ldcg(&dst.lab[i + u * G]);
                    v4[u] = __ldcg(&dst.vel[i + u * G]);
                    z4[u] = __ldcg(&dst.tail[i + u * G].l2);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double Lu[3] = {l4[u].x, l4[u].y, z4[u]};
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        acc[3 * k + 0] = __dadd_rn(acc[3 * k + 0], __dmul_rn(Lu[k], v4[u].x));
                        acc[3 * k + 1] = __dadd_rn(acc[3 * k + 1], __dmul_rn(Lu[k], v4[u].y));
                        acc[3 * k + 2] = __dadd_rn(acc[3 * k + 2], Lu[k]);
                    }
                }
            }
            for (; i < e; i += G) {
                const double2 la = __ldcg(&dst.lab[i]);
                const double2 va = __ldcg(&dst.vel[i]);
                const double La[3] = {la.x, la.y, __ldcg(&dst.tail[i].l2)};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    acc[3 * k + 0] = __dadd_rn(acc[3 * k + 0], __dmul_rn(La[k], va.x));
                    acc[3 * k + 1] = __dadd_rn(acc[3 * k + 1], __dmul_rn(La[k], va.y));
                    acc[3 * k + 2] = __dadd_rn(acc[3 * k + 2], La[k]);
                }
            }
#pragma unroll
            for (int d = G / 2; d > 0; d >>= 1) {
#pragma unroll
                for (int k = 0; k < 9; ++k) acc[k] = __dadd_rn(acc[k], __shfl_xor_sync(0xffffffffu, acc[k], d, G));
            }
            if (valid) {
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    if (lane == (k % G)) partial[9 * (size_t)c + k] = acc[k];
            }
        }
    }
}

// One warp-tile of 32 particle records staged in shared memory by the copy engine (one bulk copy of 2 KB).
struct __align__(128) ScatterStage {
    int4 rec[32][4]; // 32 whole particle records: [.][0] pos, [1] lab, [2] tail, [3] vel
};
static_assert(sizeof(ScatterStage) == 2048, "one scatter stage is 2 KB");

// TMA-staged variant (pfem2_options.scatter_tma = 1).  Each warp runs its own S-stage TMA pipeline over its tiles (no block-level synchronisation): lane 0 issues
// cp.async.bulk copies of the next tiles and arms the stage's mbarrier; the warp waits on the phase parity, reads
// the cell keys from shared memory, reserves runs with one atomic per (warp, cell) group and streams the records
// shared -> global with 128-bit stores.  Registers hold no particle data across the atomic's latency, so many
// warps fit per SM and several tiles per warp are in flight.
template <int S>
__global__ void __launch_bounds__(kThreads)
k_scatter_all_tma(ParticleSoA src, ParticleSoA dst, const int *__restrict__ n_old_ptr, int *__restrict__ cursor, const Counters *ctr)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (ctr->overflow) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps_per_block = blockDim.x >> 5;
    ScatterStage *stage = reinterpret_cast<ScatterStage *>(smem_raw) + warp * S;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)warps_per_block * S * sizeof(ScatterStage)) + warp * S;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < S; ++k) mbar_init(bars + k, 1);
        fence_barrier_init();
    }
    __syncwarp();
    const int n = *n_old_ptr;
    const int tiles = (n + 31) >> 5;
    const int warp_global = blockIdx.x * warps_per_block + warp;
    const int warps_total = gridDim.x * warps_per_block;
    auto issue = [&](int tile, int k) {
        if (lane == 0) {
            const int base = tile << 5;
            const uint32_t bytes = (uint32_t)min(32, n - base) * (uint32_t)sizeof(ParticleRec);
            mbar_arrive_expect_tx(bars + k, bytes);
            bulk_g2s(stage[k].rec, src.records() + base, bytes, bars + k);
        }
    };
#pragma unroll
    for (int k = 0; k < S - 1; ++k) {
        const int tile = warp_global + k * warps_total;
        if (tile < tiles) issue(tile, k);
    }
    int it = 0;
    for (int tile = warp_global; tile < tiles; tile += warps_total, ++it) {
        const int k = it % S;
        {   // refill the stage consumed in the previous iteration with the tile S-1 ahead
            const int nxt = tile + (S - 1) * warps_total;
            if (nxt < tiles) issue(nxt, (it + S - 1) % S);
        }
        mbar_wait(bars + k, (uint32_t)((it / S) & 1));
        const int i = (tile << 5) + lane;
        const unsigned c = i < n ? (unsigned)stage[k].rec[lane][2].z : kLostCell;
        // cursor[c] starts at the new segment start of cell c: one atomicAdd per (warp, cell) group reserves a run
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        int run = 0;
        if (c != kLostCell && (peers & ((1u << lane) - 1)) == 0) run = atomicAdd(cursor + c, __popc(peers));
        run = __shfl_sync(0xffffffffu, run, __ffs(peers) - 1);
        if (c != kLostCell) {
            const int d = run + __popc(peers & ((1u << lane) - 1));
            int4 *out = reinterpret_cast<int4 *>(dst.records() + d);
            out[0] = stage[k].rec[lane][0];
            out[1] = stage[k].rec[lane][1];
            out[2] = stage[k].rec[lane][2];
            out[3] = stage[k].rec[lane][3];
        }
        __syncwarp(); // every lane has read stage k before it is refilled
    }
}
template <int S> constexpr size_t scatter_smem_bytes(int threads) { return (size_t)(threads / 32) * S * (sizeof(ScatterStage) + sizeof(uint64_t)); }

// cursor[c] = new segment start of cell c (low half of the scanned plan word), for the fast scatter
__global__ void __launch_bounds__(kThreads)
k_init_cursor(int own_lo, int own_hi, const unsigned long long *__restrict__ packed_start, int *__restrict__ cursor)
{
    const int c = own_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (c < own_hi) cursor[c] = (int)(unsigned)(packed_start[c] & 0xffffffffull);
}

// movers, sorted by new cell (stable: array order within a cell), go right behind the cell's stayers
__global__ void __launch_bounds__(kThreads)
k_scatter_movers(ParticleSoA src, ParticleSoA dst, int n_cells, const int *__restrict__ n_movers,
                 const unsigned *__restrict__ keys_sorted, const unsigned *__restrict__ vals_sorted, const int *__restrict__ stay,
                 const unsigned long long *__restrict__ packed_start, const Counters *ctr)
{
    if (ctr->overflow) return;
    const int m = *n_movers;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        const unsigned c = keys_sorted[j];
        if (c >= (unsigned)n_cells) continue; // lost
        const unsigned s = vals_sorted[j];
        const unsigned long long ps = packed_start[c];
        const int d = (int)(unsigned)(ps & 0xffffffffull) + __ldg(stay + c) + (j - (int)(unsigned)(ps >> 32));
        copy_particle(src, (int)s, dst, d);
    }
}

// kAddParticlesToCell (:197-236): one new particle at the centre of every empty sub-cell, velocity
// interpolated from the current nodal field; written right behind the cell's survivors.  Also
// materialises the segment table cell_start[].
__global__ void __launch_bounds__(kThreads)
k_reseed(int own_lo, int own_hi, int ppc, const double2 *__restrict__ vertices, const CellGeom *__restrict__ geom,
         const double *__restrict__ centers, NodalVel vel, const unsigned long long *__restrict__ cell_mask,
         const int *__restrict__ stay, const int *__restrict__ arrive, const unsigned long long *__restrict__ packed_start,
         ParticleSoA dst, int *__restrict__ cell_start, const Counters *ctr)
{
    const int c = own_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (c > own_hi) return;
    const int start = (int)(unsigned)(packed_start[c] & 0xffffffffull);
    cell_start[c] = start;
    if (c == own_hi || ctr->overflow) return;
    const int live = stay[c] + arrive[c];
    const int missing = (int)(unsigned)(packed_start[c + 1] & 0xffffffffull) - start - live;
    if (missing <= 0) return;
    const double *Vx, *Vy;
    vel.resolve(Vx, Vy);
    const uint4 nn = __ldg(reinterpret_cast<const uint4 *>(&geom[c].n0));
    const double2 v0 = __ldg(&vertices[nn.x]), v1 = __ldg(&vertices[nn.y]), v2 = __ldg(&vertices[nn.z]);
    const double ax0 = __ldg(Vx + nn.x), ax1 = __ldg(Vx + nn.y), ax2 = __ldg(Vx + nn.z);
    const double ay0 = __ldg(Vy + nn.x), ay1 = __ldg(Vy + nn.y), ay2 = __ldg(Vy + nn.z);
    const unsigned long long mask = cell_mask[c];
    int d = start + live;
    for (int s = 0; s < ppc; ++s) {
        if ((mask >> s) & 1ull) continue;
        const double L0 = __ldg(&centers[3 * s]), L1 = __ldg(&centers[3 * s + 1]), L2 = __ldg(&centers[3 * s + 2]);
        dst.pos[d] = make_double2(to_global1(L0, L1, L2, v0.x, v1.x, v2.x), to_global1(L0, L1, L2, v0.y, v1.y, v2.y));
        dst.lab[d] = make_double2(L0, L1);
        st_tail(dst.tail + d, L2, (unsigned)c, (unsigned)d);
        dst.vel[d] = make_double2(interp3(L0, L1, L2, ax0, ax1, ax2), interp3(L0, L1, L2, ay0, ay1, ay2));
        ++d;
    }
}

// ---------------------------------------------------------------------------------------------
// multi-GPU (strip partition): the move pass runs without statistics; after particles that left the
// owned cell range have been handed to their new owner and the immigrants appended, one pass over all live
// particles accumulates the per-cell counts and occupancy masks (everybody counts as "arrived").
// ---------------------------------------------------------------------------------------------
template <int SUBCELL_MODE, bool MASK64>
__global__ void __launch_bounds__(kThreads)
k_count_all(ParticleSoA p, const Counters *ctr, int n_cells, int ppc, int level, double sub_step, int *__restrict__ stay,
            int *__restrict__ arrive, unsigned long long *__restrict__ cell_mask)
{
    const int n = ctr->count;
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        bool live = false;
        unsigned c = 0;
        double L0 = 0, L1 = 0, L2 = 0;
        if (i < n) {
            const ParticleTail tl = ld_tail(p.tail + i);
            c = tl.cell;
            live = c != kLostCell;
            if (live) {
                const double2 lab = p.lab[i];
                L0 = lab.x;
                L1 = lab.y;
                L2 = tl.l2;
            }
        }
        const unsigned mb = __ballot_sync(0xffffffffu, live);
        accumulate_cell_stats<SUBCELL_MODE, MASK64>(live, c, L0, L1, L2, 0u, mb, lane, n_cells, ppc, level, sub_step, stay, arrive, cell_mask);
    }
}

// emigrants = live particles whose cell is outside [own_lo, own_hi).  Pass 1 counts them per destination rank.
__global__ void __launch_bounds__(kThreads)
k_emigrant_count(ParticleSoA p, const Counters *ctr, int own_lo, int own_hi, const int *__restrict__ bounds, int n_ranks,
                 int *__restrict__ rank_count)
{
    const int n = ctr->count;
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        int r = -1;
        if (i < n) {
            const unsigned c = ld_cell(p.tail + i);
            if (c != kLostCell && ((int)c < own_lo || (int)c >= own_hi)) r = rank_of_cell(c, bounds, n_ranks);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, r);
        if (r >= 0 && (peers & ((1u << lane) - 1)) == 0) atomicAdd(rank_count + r, __popc(peers));
    }
}

// Pass 2 packs them as 64-byte records {pos, lab, tail, vel} grouped by destination rank (rank_cursor starts at the
// exclusive prefix of the counts) and removes them from the local array (cell = lost).
__global__ void __launch_bounds__(kThreads)
k_emigrant_pack(ParticleSoA p, const Counters *ctr, int own_lo, int own_hi, const int *__restrict__ bounds, int n_ranks,
                int *__restrict__ rank_cursor, int4 *__restrict__ out)
{
    const int n = ctr->count;
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + lane;
        int r = -1;
        if (i < n) {
            const unsigned c = ld_cell(p.tail + i);
            if (c != kLostCell && ((int)c < own_lo || (int)c >= own_hi)) r = rank_of_cell(c, bounds, n_ranks);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, r);
        if (r < 0) continue;
        const int leader = __ffs(peers) - 1;
        int slot = 0;
        if (lane == leader) slot = atomicAdd(rank_cursor + r, __popc(peers));
        slot = __shfl_sync(peers, slot, leader) + __popc(peers & ((1u << lane) - 1));
        int4 *rec = out + 4 * (size_t)slot;
        rec[0] = *reinterpret_cast<const int4 *>(p.pos + i);
        rec[1] = *reinterpret_cast<const int4 *>(p.lab + i);
        rec[2] = *reinterpret_cast<const int4 *>(p.tail + i);
        rec[3] = *reinterpret_cast<const int4 *>(p.vel + i);
        st_cell(p.tail + i, kLostCell);
    }
}

// Fused multi-GPU path: the move pass left the array indices of the emigrants in emig_idx (k_advect_locate_tma); pack
// them grouped by destination rank (rank_cursor starts at the exclusive prefix of the counts) and remove them locally.
__global__ void __launch_bounds__(kThreads)
k_emigrant_pack_list(ParticleSoA p, const unsigned *__restrict__ emig_idx, int n_emig, const int *__restrict__ bounds, int n_ranks,
                     int *__restrict__ rank_cursor, int4 *__restrict__ out)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_emig; j += gridDim.x * blockDim.x) {
        const unsigned i = emig_idx[j];
        const int4 t = *reinterpret_cast<const int4 *>(p.tail + i);
        const int slot = atomicAdd(rank_cursor + rank_of_cell((unsigned)t.z, bounds, n_ranks), 1);
        int4 *rec = out + 4 * (size_t)slot;
        rec[0] = *reinterpret_cast<const int4 *>(p.pos + i);
        rec[1] = *reinterpret_cast<const int4 *>(p.lab + i);
        rec[2] = t;
        rec[3] = *reinterpret_cast<const int4 *>(p.vel + i);
        st_cell(p.tail + i, kLostCell);
    }
}

// per-cell statistics of the m records behind the current array end (the immigrants just appended): all "arrived"
template <int SUBCELL_MODE, bool MASK64>
__global__ void __launch_bounds__(kThreads)
k_count_appended(ParticleSoA p, const Counters *ctr, int m, int n_cells, int ppc, int level, double sub_step, int *__restrict__ stay,
                 int *__restrict__ arrive, unsigned long long *__restrict__ cell_mask)
{
    const int n0 = ctr->count;
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < m; base += gridDim.x * blockDim.x) {
        const int j = base + lane;
        bool live = false;
        unsigned c = 0;
        double L0 = 0, L1 = 0, L2 = 0;
        if (j < m) {
            const ParticleTail tl = ld_tail(p.tail + (n0 + j));
            c = tl.cell;
            live = c != kLostCell;
            const double2 lab = p.lab[n0 + j];
            L0 = lab.x;
            L1 = lab.y;
            L2 = tl.l2;
        }
        const unsigned mb = __ballot_sync(0xffffffffu, live);
        accumulate_cell_stats<SUBCELL_MODE, MASK64>(live, c, L0, L1, L2, 0u, mb, lane, n_cells, ppc, level, sub_step, stay, arrive, cell_mask);
    }
}

// immigrants: 64-byte records appended behind the current array
__global__ void __launch_bounds__(kThreads)
k_immigrant_append(ParticleSoA p, Counters *ctr, const int4 *__restrict__ in, int m)
{
    const int n = ctr->count;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        const int4 *rec = in + 4 * (size_t)j;
        *reinterpret_cast<int4 *>(p.pos + (n + j)) = rec[0];
        *reinterpret_cast<int4 *>(p.lab + (n + j)) = rec[1];
        *reinterpret_cast<int4 *>(p.tail + (n + j)) = rec[2];
        *reinterpret_cast<int4 *>(p.vel + (n + j)) = rec[3];
    }
}
__global__ void k_add_count(Counters *ctr, int m)
{
    ctr->count += m;
    ctr->n_old = ctr->count;
    ctr->n_warps = (ctr->count + 31) >> 5;
}

// ---------------------------------------------------------------------------------------------
// Neighbour protocol of the strip partition: no host round trip between the move pass and the re-sort.
// A migration buffer is  [64-byte header | cap 64-byte records]  of fixed capacity, so that the NCCL transfer needs no
// size negotiation; the number of valid records travels in the header and is only ever read on the device.
// ---------------------------------------------------------------------------------------------
struct __align__(16) MigrationHeader {
    int count;                   // records the sender wanted to pack (> cap: overflow, only cap of them are present)
    int flags;                   // 1: the sender ran out of capacity; 2: an emigrant's destination was not an adjacent strip
    int pad0[2];
    unsigned long long spill[4]; // right-going only: occupancy words of the cells own_hi .. own_hi + 3 of the sender (SURVEY N4:
                                 // the flat sub-cell index of a particle in the tolerance band lands in the next cell's word)
    int pad1[4];
};
static_assert(sizeof(MigrationHeader) == sizeof(ParticleRec), "the header occupies exactly one record slot");
constexpr int kOverflowMigration = 4; // Counters.overflow bit: migration buffer too small / non-adjacent destination

// the move pass (k_advect_locate_tma) left the array indices of the emigrants in emig_idx and their number in
// rank_count[n_ranks]; pack them for the left / right neighbour and remove them locally.  Headers are zeroed by the caller.
__global__ void __launch_bounds__(kThreads)
k_emigrant_pack_nbr(ParticleSoA p, const unsigned *__restrict__ emig_idx, const int *__restrict__ rank_count, int n_ranks,
                    const int *__restrict__ bounds, int rank, int4 *out_left, int4 *out_right, int cap, Counters *ctr,
                    const unsigned long long *__restrict__ cell_mask, int own_hi, int n_cells)
{
    const int n_emig = rank_count[n_ranks];
    if (blockIdx.x == 0 && threadIdx.x < 4 && out_right) {
        const int c = own_hi + (int)threadIdx.x;
        reinterpret_cast<MigrationHeader *>(out_right)->spill[threadIdx.x] = c < n_cells ? cell_mask[c] : 0ull;
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_emig; j += gridDim.x * blockDim.x) {
        const unsigned i = emig_idx[j];
        const int4 t = *reinterpret_cast<const int4 *>(p.tail + i);
        const int dest = rank_of_cell((unsigned)t.z, bounds, n_ranks);
        int4 *out = dest == rank - 1 ? out_left : (dest == rank + 1 ? out_right : nullptr);
        st_cell(p.tail + i, kLostCell);
        if (!out) { // more than one strip away in one step (or no such neighbour): the strips are too thin for this time step
            atomicOr(&ctr->overflow, kOverflowMigration);
            continue;
        }
        MigrationHeader *hd = reinterpret_cast<MigrationHeader *>(out);
        const int slot = atomicAdd(&hd->count, 1);
        if (slot >= cap) {
            atomicOr(&hd->flags, 1);
            atomicOr(&ctr->overflow, kOverflowMigration);
            continue;
        }
        int4 *rec = out + 4 * ((size_t)slot + 1);
        rec[0] = *reinterpret_cast<const int4 *>(p.pos + i);
        rec[1] = *reinterpret_cast<const int4 *>(p.lab + i);
        rec[2] = t;
        rec[3] = *reinterpret_cast<const int4 *>(p.vel + i);
    }
}

__device__ __forceinline__ int migration_count(const int4 *buf, int cap)
{
    return min(max(reinterpret_cast<const MigrationHeader *>(buf)->count, 0), cap);
}

// immigrants of one received migration buffer, appended behind the current array (count taken from the header)
__global__ void __launch_bounds__(kThreads)
k_immigrant_append_dev(ParticleSoA p, const Counters *ctr, const int4 *__restrict__ buf, int cap)
{
    const int m = migration_count(buf, cap), n = ctr->count;
    if ((long long)n + m > ctr->capacity) return; // k_add_count_dev raises the overflow flag
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        const int4 *rec = buf + 4 * ((size_t)j + 1);
        *reinterpret_cast<int4 *>(p.pos + (n + j)) = rec[0];
        *reinterpret_cast<int4 *>(p.lab + (n + j)) = rec[1];
        *reinterpret_cast<int4 *>(p.tail + (n + j)) = rec[2];
        *reinterpret_cast<int4 *>(p.vel + (n + j)) = rec[3];
    }
}

// per-cell statistics of those records (all "arrived"), like k_count_appended
template <int SUBCELL_MODE, bool MASK64>
__global__ void __launch_bounds__(kThreads)
k_count_appended_dev(ParticleSoA p, const Counters *ctr, const int4 *__restrict__ buf, int cap, int n_cells, int ppc, int level,
                     double sub_step, int *__restrict__ stay, int *__restrict__ arrive, unsigned long long *__restrict__ cell_mask)
{
    const int m = migration_count(buf, cap), n0 = ctr->count;
    if ((long long)n0 + m > ctr->capacity) return;
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < m; base += gridDim.x * blockDim.x) {
        const int j = base + lane;
        bool live = false;
        unsigned c = 0;
        double L0 = 0, L1 = 0, L2 = 0;
        if (j < m) {
            const ParticleTail tl = ld_tail(p.tail + (n0 + j));
            c = tl.cell;
            live = c != kLostCell;
            const double2 lab = p.lab[n0 + j];
            L0 = lab.x;
            L1 = lab.y;
            L2 = tl.l2;
        }
        const unsigned mb = __ballot_sync(0xffffffffu, live);
        accumulate_cell_stats<SUBCELL_MODE, MASK64>(live, c, L0, L1, L2, 0u, mb, lane, n_cells, ppc, level, sub_step, stay, arrive, cell_mask);
    }
}

// one thread: the array grows by the appended records; the left neighbour's spilled occupancy bits join this strip's first cells
__global__ void k_add_count_dev(Counters *ctr, const int4 *__restrict__ buf, int cap, unsigned long long *__restrict__ cell_mask,
                                int own_lo, int own_hi, int from_left)
{
    const MigrationHeader *hd = reinterpret_cast<const MigrationHeader *>(buf);
    int m = migration_count(buf, cap);
    if (hd->count > cap || hd->count < 0 || hd->flags) ctr->overflow |= kOverflowMigration;
    if ((long long)ctr->count + m > ctr->capacity) {
        ctr->overflow |= 1;
        m = 0;
    }
    ctr->count += m;
    ctr->n_old = ctr->count;
    ctr->n_warps = (ctr->count + 31) >> 5;
    if (from_left)
        for (int k = 0; k < 4; ++k)
            if (own_lo + k < own_hi && hd->spill[k]) cell_mask[own_lo + k] |= hd->spill[k];
}

// ---------------------------------------------------------------------------------------------
// P2P transport of the neighbour protocol: NVLink peer memory instead of NCCL, nothing on the host in the loop.
// Every strip allocates one INBOX per neighbour and hands its CUDA IPC handle to that neighbour, which maps it and
// from then on stores into it directly:
//     [ P2PInboxHead 64 B ][ parity 0: header | cap records ][ parity 1: header | cap records ][ halo 0 ][ halo 1 ]
// The pack kernel of the sender writes the 64-byte records of its emigrants straight into the receiver's HBM (plain
// 128-bit stores over NVLink), a one-thread publish kernel then writes the header (count, flags, spill words), fences
// (system scope) and releases the sequence number into the head; the receiver's stream waits on that flag with a
// one-thread acquire spin (bounded by a watchdog: a dead peer raises the overflow flag instead of hanging the GPU) and
// appends from its own memory.  The projection halo uses the same inbox: interface-node accumulators are stored into
// the neighbour's halo block and added there.  Blocks alternate with the parity of the sequence number: a sender can be
// at most one delivery ahead of what the receiver has consumed (it waited for the receiver's delivery of the same step).
// ---------------------------------------------------------------------------------------------
constexpr unsigned kP2PMagic = 0x50324232u;
constexpr int kOverflowP2PTimeout = 8; // Counters.overflow bit: a neighbour's delivery did not arrive in time
struct __align__(64) P2PInboxHead {
    unsigned magic;
    unsigned flag_mig;    // sequence number of the last complete migration delivery (written by the neighbour)
    unsigned flag_halo;   // ... of the last complete halo delivery
    int capacity_records;
    int n_halo_nodes;
    int pad[11];
};
static_assert(sizeof(P2PInboxHead) == 64, "inbox head is one record slot");
__host__ __device__ inline size_t p2p_block_bytes(int cap) { return ((size_t)cap + 1) * sizeof(ParticleRec); }
__host__ __device__ inline size_t p2p_block_offset(int cap, int parity) { return sizeof(P2PInboxHead) + (size_t)parity * p2p_block_bytes(cap); }
__host__ __device__ inline size_t p2p_halo_offset(int cap, int n_nodes, int parity)
{
    return sizeof(P2PInboxHead) + 2 * p2p_block_bytes(cap) + (size_t)parity * (size_t)n_nodes * 3 * sizeof(double);
}
__host__ __device__ inline size_t p2p_inbox_bytes(int cap, int n_nodes) { return p2p_halo_offset(cap, n_nodes, 2); }

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// emigrants listed by the move pass -> records stored directly into the neighbours' inbox blocks (peer memory); slots come
// from LOCAL cursors (no atomics over NVLink).  rec_left / rec_right point at the first record of the peer's block.
__global__ void __launch_bounds__(kThreads)
k_emigrant_pack_p2p(ParticleSoA p, const unsigned *__restrict__ emig_idx, const int *__restrict__ rank_count, int n_ranks,
                    const int *__restrict__ bounds, int rank, int4 *rec_left, int4 *rec_right, int cap, Counters *ctr, int *cursors)
{
    const int n_emig = rank_count[n_ranks];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_emig; j += gridDim.x * blockDim.x) {
        const unsigned i = emig_idx[j];
        const int4 t = *reinterpret_cast<const int4 *>(p.tail + i);
        const int dest = rank_of_cell((unsigned)t.z, bounds, n_ranks);
        const int side = dest == rank - 1 ? 0 : (dest == rank + 1 ? 1 : -1);
        int4 *out = side == 0 ? rec_left : (side == 1 ? rec_right : nullptr);
        st_cell(p.tail + i, kLostCell);
        if (!out) {
            atomicOr(&ctr->overflow, kOverflowMigration);
            continue;
        }
        const int slot = atomicAdd(cursors + side, 1);
        if (slot >= cap) {
            atomicOr(&ctr->overflow, kOverflowMigration);
            continue;
        }
        int4 *rec = out + 4 * (size_t)slot;
        rec[0] = *reinterpret_cast<const int4 *>(p.pos + i);
        rec[1] = *reinterpret_cast<const int4 *>(p.lab + i);
        rec[2] = t;
        rec[3] = *reinterpret_cast<const int4 *>(p.vel + i);
    }
}

// one thread, after the pack kernel has completed (its peer stores are performed): headers, fence, then the flags
__global__ void k_p2p_publish_migration(MigrationHeader *hdr_left, unsigned *flag_left, MigrationHeader *hdr_right, unsigned *flag_right,
                                        int *cursors, int cap, const unsigned long long *__restrict__ cell_mask, int own_hi,
                                        int n_cells, unsigned seq)
{
    if (hdr_left) {
        const int n = cursors[0];
        hdr_left->count = n;
        hdr_left->flags = n > cap ? 1 : 0;
        for (int k = 0; k < 4; ++k) hdr_left->spill[k] = 0ull;
    }
    if (hdr_right) {
        const int n = cursors[1];
        hdr_right->count = n;
        hdr_right->flags = n > cap ? 1 : 0;
        for (int k = 0; k < 4; ++k) hdr_right->spill[k] = own_hi + k < n_cells ? cell_mask[own_hi + k] : 0ull;
    }
    cursors[2] = cursors[0] + cursors[1]; // what this strip handed over (read back on demand)
    cursors[0] = cursors[1] = 0;
    __threadfence_system();
    if (flag_left) st_release_sys(flag_left, seq);
    if (flag_right) st_release_sys(flag_right, seq);
}

__global__ void k_p2p_publish_flag(unsigned *flag_left, unsigned *flag_right, unsigned seq)
{
    __threadfence_system();
    if (flag_left) st_release_sys(flag_left, seq);
    if (flag_right) st_release_sys(flag_right, seq);
}

// one thread: the stream continues once both neighbours have delivered sequence number `seq` (or the watchdog fires)
__global__ void k_p2p_wait(const unsigned *flag_a, const unsigned *flag_b, unsigned seq, Counters *ctr, unsigned long long timeout_ns)
{
    const unsigned long long t0 = global_timer_ns();
    const unsigned *flags[2] = {flag_a, flag_b};
    for (int k = 0; k < 2; ++k) {
        if (!flags[k]) continue;
        while ((int)(ld_acquire_sys(flags[k]) - seq) < 0) {
            if (global_timer_ns() - t0 > timeout_ns) {
                atomicOr(&ctr->overflow, kOverflowP2PTimeout);
                return;
            }
            __nanosleep(200);
        }
    }
}

// projection halo: my accumulators of the interface nodes -> the neighbour's halo block (peer memory)
__global__ void __launch_bounds__(kThreads)
k_halo_send(const double *__restrict__ acc3, const int *__restrict__ idx, int n, double *peer_halo)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *a = acc3 + 3 * (size_t)idx[i];
    peer_halo[3 * (size_t)i] = a[0];
    peer_halo[3 * (size_t)i + 1] = a[1];
    peer_halo[3 * (size_t)i + 2] = a[2];
}

// ... and the neighbour's contribution added to mine (two contributions per shared node: a + b == b + a bit for bit)
__global__ void __launch_bounds__(kThreads)
k_halo_add(double *__restrict__ acc3, const int *__restrict__ idx, int n, const double *__restrict__ halo)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double *a = acc3 + 3 * (size_t)idx[i];
    a[0] = __dadd_rn(a[0], halo[3 * (size_t)i]);
    a[1] = __dadd_rn(a[1], halo[3 * (size_t)i + 1]);
    a[2] = __dadd_rn(a[2], halo[3 * (size_t)i + 2]);
}

__global__ void k_p2p_init_head(P2PInboxHead *hd, int cap, int n_nodes)
{
    hd->magic = kP2PMagic;
    hd->flag_mig = 0;
    hd->flag_halo = 0;
    hd->capacity_records = cap;
    hd->n_halo_nodes = n_nodes;
}

// projection, multi-GPU flavour: per-node accumulators {sum L v_x, sum L v_y, sum L} without the division, so that the
// contributions of the strips sharing an interface node can be added before kFinalizeVelocityProjection's division
__global__ void __launch_bounds__(kThreads)
k_project_nodes_acc(int n_list, const int *__restrict__ node_list, const int *__restrict__ node_off,
                    const int *__restrict__ node_inc, const double *__restrict__ partial, double *__restrict__ acc3)
{
    // node_list: the nodes of the owned cells (nullptr = all nodes)
    const int q0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (q0 >= n_list) return;
    const int i = node_list ? node_list[q0] : q0;
    double sx = 0.0, sy = 0.0, sw = 0.0;
    const int e = __ldg(node_off + i + 1);
    for (int q = __ldg(node_off + i); q < e; ++q) {
        const double *a = partial + 3 * (size_t)__ldg(node_inc + q);
        sx = __dadd_rn(sx, a[0]);
        sy = __dadd_rn(sy, a[1]);
        sw = __dadd_rn(sw, a[2]);
    }
    acc3[3 * (size_t)i] = sx;
    acc3[3 * (size_t)i + 1] = sy;
    acc3[3 * (size_t)i + 2] = sw;
}

__global__ void __launch_bounds__(kThreads)
k_project_finalize(int n_list, const int *__restrict__ node_list, const double *__restrict__ acc3, double *__restrict__ vx,
                   double *__restrict__ vy)
{
    const int q0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (q0 >= n_list) return;
    const int i = node_list ? node_list[q0] : q0;
    const double sw = acc3[3 * (size_t)i + 2];
    vx[i] = __ddiv_rn(acc3[3 * (size_t)i], sw);
    vy[i] = __ddiv_rn(acc3[3 * (size_t)i + 1], sw);
}

// ---------------------------------------------------------------------------------------------
// projection: kProjectParticleVelocityOntoGrid (:90-107) + kFinalizeVelocityProjection (:109-115)
// as a sorted segmented reduction instead of 9 fp64 atomics per particle.
//   pass 1: G lanes per cell reduce the cell's contiguous segment to 9 partial sums
//           partial[(3c + i)*3 + {0,1,2}] = sum_p { L_i vx, L_i vy, L_i }
//   pass 2: one thread per node sums the partials of its incident (cell, i) pairs in ascending order
//           and divides (IEEE).  No atomics, bit-reproducible for a given particle order.
// ---------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(kThreads)
k_project_cells(int c_lo, int n_cells, ParticleSoA p, const int *__restrict__ cell_start, double *__restrict__ partial)
{
    // cells [c_lo, n_cells): the owned range (partials of cells that can never hold particles stay zero)
    const int lane = threadIdx.x & (G - 1);
    constexpr int groups_per_warp = 32 / G, groups_per_block = kThreads / G;
    // the loop bound is warp-uniform (the full-mask shuffles below need all 32 lanes), cells are guarded inside
    for (int cw = c_lo + blockIdx.x * groups_per_block + (threadIdx.x >> 5) * groups_per_warp; cw < n_cells;
         cw += gridDim.x * groups_per_block) {
        const int c = cw + ((threadIdx.x & 31) / G);
        const bool valid = c < n_cells;
        const int b = valid ? __ldg(cell_start + c) : 0, e = valid ? __ldg(cell_start + c + 1) : 0;
        double acc[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] = 0.0;
        // each lane walks the segment with stride G; four (then two) particles in flight per lane
        int i = b + lane;
        for (; i + 3 * G < e; i += 4 * G) {
            double2 l4[4], v4[4];
            double z4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                l4[u] = p.lab[i + u * G];
                v4[u] = p.vel[i + u * G];
                z4[u] = p.tail[i + u * G].l2;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double Lu[3] = {l4[u].x, l4[u].y, z4[u]};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    acc[3 * k + 0] = __dadd_rn(acc[3 * k + 0], __dmul_rn(Lu[k], v4[u].x));
                    acc[3 * k + 1] = __dadd_rn(acc[3 * k + 1], __dmul_rn(Lu[k], v4[u].y));
                    acc[3 * k + 2] = __dadd_rn(acc[3 * k + 2], Lu[k]);
                }
            }
        }
        for (; i + G < e; i += 2 * G) {
            const double2 la = p.lab[i], lb = p.lab[i + G];
            const double2 va = p.vel[i], vb = p.vel[i + G];
            const double za = p.tail[i].l2, zb = p.tail[i + G].l2;
            const double La[3] = {la.x, la.y, za}, Lb[3] = {lb.x, lb.y, zb};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                acc[3 * k + 0] = __dadd_rn(acc[3 * k + 0], __dmul_rn(La[k], va.x)); // t = L_i * v (plain mul), then add
                acc[3 * k + 1] = __dadd_rn(acc[3 * k + 1], __dmul_rn(La[k], va.y));
                acc[3 * k + 2] = __dadd_rn(acc[3 * k + 2], La[k]);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                acc[3 * k + 0] = __dadd_rn(acc[3 * k + 0], __dmul_rn(Lb[k], vb.x));
                acc[3 * k + 1] = __dadd_rn(acc[3 * k + 1], __dmul_rn(Lb[k], vb.y));
                acc[3 * k + 2] = __dadd_rn(acc[3 * k + 2], Lb[k]);
            }
        }
        if (i < e) {
            const double2 la = p.lab[i];
            const double2 va = p.vel[i];
            const double La[3] = {la.x, la.y, p.tail[i].l2};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                acc[3 * k + 0] = __dadd_rn(acc[3 * k + 0], __dmul_rn(La[k], va.x));
                acc[3 * k + 1] = __dadd_rn(acc[3 * k + 1], __dmul_rn(La[k], va.y));
                acc[3 * k + 2] = __dadd_rn(acc[3 * k + 2], La[k]);
            }
        }
#pragma unroll
        for (int d = G / 2; d > 0; d >>= 1) {
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[k] = __dadd_rn(acc[k], __shfl_xor_sync(0xffffffffu, acc[k], d, G));
        }
        if (valid) {
            // after the butterfly every lane of the group holds all nine sums; spread the stores over the lanes
#pragma unroll
            for (int k = 0; k < 9; ++k)
                if (lane == (k % G)) partial[9 * (size_t)c + k] = acc[k];
        }
    }
}

__global__ void __launch_bounds__(kThreads)
k_project_nodes(int node_lo, int n_nodes, const int *__restrict__ node_off, const int *__restrict__ node_inc,
                const double *__restrict__ partial, double *vx_arg, double *vy_arg, double *const *table, double *cx_arg = nullptr,
                double *cy_arg = nullptr, double *const *table_copy = nullptr)
{
    // nodes [node_lo, n_nodes).  Optional second destination (pfem2_project_dual): the cases copy the projected field into their
    // "old" solution right after the call (copy_d2d, cases/Cylinder2D/main.cu:804-805); written here it costs no extra pass
    double *Vx = table ? table[0] : vx_arg;
    double *Vy = table ? table[1] : vy_arg;
    double *Cx = table_copy ? table_copy[0] : cx_arg;
    double *Cy = table_copy ? table_copy[1] : cy_arg;
    const int i = node_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    double sx = 0.0, sy = 0.0, sw = 0.0;
    const int e = __ldg(node_off + i + 1);
    for (int q = __ldg(node_off + i); q < e; ++q) {
        const double *a = partial + 3 * (size_t)__ldg(node_inc + q);
        sx = __dadd_rn(sx, a[0]);
        sy = __dadd_rn(sy, a[1]);
        sw = __dadd_rn(sw, a[2]);
    }
    const double qx = __ddiv_rn(sx, sw), qy = __ddiv_rn(sy, sw);
    Vx[i] = qx;
    Vy[i] = qy;
    if (Cx) {
        Cx[i] = qx;
        Cy[i] = qy;
    }
}

// ---------------------------------------------------------------------------------------------
// correction: kCorrectParticleVelocity (:72-88); Vold == nullptr -> initParticleVelocity (:322-326)
//   d_i = V_i - Vold_i (plain sub) ; inc = fma chain from 0 ; v = v + inc (plain add)
// ---------------------------------------------------------------------------------------------
template <bool HAS_OLD>
__global__ void __launch_bounds__(kThreads)
k_correct(ParticleSoA p, const CellGeom *__restrict__ geom, NodalVel vel, NodalVel vel_old, const Counters *ctr)
{
    const double *__restrict__ Vx, *__restrict__ Vy, *__restrict__ Ox = nullptr, *__restrict__ Oy = nullptr;
    {
        const double *a, *b;
        vel.resolve(a, b);
        Vx = a;
        Vy = b;
        if (HAS_OLD) {
            vel_old.resolve(a, b);
            Ox = a;
            Oy = b;
        }
    }
    const int n = ctr->count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double2 lab = p.lab[i];
        const ParticleTail tl = ld_tail(p.tail + i);
        const double2 vel = p.vel[i];
        const unsigned c = tl.cell;
        const uint4 nn = __ldg(reinterpret_cast<const uint4 *>(&geom[c].n0));
        const double L0 = lab.x, L1 = lab.y, L2 = tl.l2;
        double dx0 = __ldg(Vx + nn.x), dx1 = __ldg(Vx + nn.y), dx2 = __ldg(Vx + nn.z);
        double dy0 = __ldg(Vy + nn.x), dy1 = __ldg(Vy + nn.y), dy2 = __ldg(Vy + nn.z);
        if (HAS_OLD) {
            dx0 = __dsub_rn(dx0, __ldg(Ox + nn.x));
            dx1 = __dsub_rn(dx1, __ldg(Ox + nn.y));
            dx2 = __dsub_rn(dx2, __ldg(Ox + nn.z));
            dy0 = __dsub_rn(dy0, __ldg(Oy + nn.x));
            dy1 = __dsub_rn(dy1, __ldg(Oy + nn.y));
            dy2 = __dsub_rn(dy2, __ldg(Oy + nn.z));
        }
        p.vel[i] = make_double2(__dadd_rn(vel.x, interp3(L0, L1, L2, dx0, dx1, dx2)),
                                __dadd_rn(vel.y, interp3(L0, L1, L2, dy0, dy1, dy2)));
    }
}

// Deferred correction: correctParticleVelocity only snapshots the nodal increment
// dV = V - Vold (N values, plain subtraction exactly as kCorrectParticleVelocity computes it per use); the particle
// update v += sum_i L_i dV_i is folded into the next advect pass, which reads the particle anyway.  The increment
// is evaluated with the particle's cell and local position at the time of the correct call (nothing moves between
// the two), so the bits are those of the eager kernel.  Any reader of particle velocities flushes first.
__global__ void __launch_bounds__(kThreads)
k_snapshot_dv(int node_lo, int node_hi, NodalVel vel, NodalVel vel_old, int has_old, double *__restrict__ dvx, double *__restrict__ dvy,
              double2 *__restrict__ dv2)
{
    // [node_lo, node_hi): the nodes of the owned cells (multi-GPU: a strip's share of the mesh; otherwise all nodes)
    const double *Vx, *Vy, *Ox = nullptr, *Oy = nullptr;
    vel.resolve(Vx, Vy);
    if (has_old) vel_old.resolve(Ox, Oy);
    const int i = node_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= node_hi) return;
    const double dx = has_old ? __dsub_rn(Vx[i], Ox[i]) : Vx[i];
    const double dy = has_old ? __dsub_rn(Vy[i], Oy[i]) : Vy[i];
    dvx[i] = dx;
    dvy[i] = dy;
    dv2[i] = make_double2(dx, dy); // interleaved copy for the TMA-tiled advect pass
}

// ---------------------------------------------------------------------------------------------
// getParticles(): materialise the reference's 96-byte AoS Particle2D records (particle_2d.cuh:51-57;
// ID@0 position@16 localPosition@32 velocity@64 cellID@80, 12 B tail pad).  6 x 16-byte stores each.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_export_aos(ParticleSoA p, const Counters *ctr, uint4 *__restrict__ out)
{
    const int n = ctr->count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint4 *rec = out + 6 * (size_t)i;
        const double2 pos = p.pos[i], lab = p.lab[i], vel = p.vel[i];
        const ParticleTail tl = ld_tail(p.tail + i);
        rec[0] = make_uint4(tl.id, 0u, 0u, 0u);
        reinterpret_cast<double2 *>(rec)[1] = pos;
        reinterpret_cast<double2 *>(rec)[2] = lab;
        reinterpret_cast<double2 *>(rec)[3] = make_double2(tl.l2, 0.0);
        reinterpret_cast<double2 *>(rec)[4] = vel;
        rec[5] = make_uint4(tl.cell, 0u, 0u, 0u);
    }
}

// ---------------------------------------------------------------------------------------------
// node -> (cell, local vertex) incidence keys, for the projection's gather pass and the one-ring builder
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_incidence_keys(int n_cells, const unsigned *__restrict__ cells, unsigned *__restrict__ keys, unsigned *__restrict__ vals,
                 int *__restrict__ node_count)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x; // q = 3c + i
    if (q >= 3 * n_cells) return;
    const unsigned node = cells[q];
    keys[q] = node;
    vals[q] = (unsigned)q;
    atomicAdd(node_count + node, 1);
}

// Mesh2D::fillCellNeighborIndices (mesh_2d.cu:107-139) in O(C): the one-ring of cell c is the union of
// the cells incident to its three nodes, minus c, ascending.  Pass 1 (indices == nullptr) counts.
constexpr int kMaxRing = 96;
__global__ void __launch_bounds__(128)
k_one_ring(int n_cells, const unsigned *__restrict__ cells, const int *__restrict__ node_off,
           const unsigned *__restrict__ node_inc, int *__restrict__ counts, const int *__restrict__ offsets,
           int *__restrict__ indices, int *__restrict__ error)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    int buf[kMaxRing];
    int m = 0;
    for (int k = 0; k < 3; ++k) {
        const unsigned node = cells[3 * (size_t)c + k];
        const int e = node_off[node + 1];
        for (int q = node_off[node]; q < e; ++q) {
            const int other = (int)(node_inc[q] / 3u);
            if (other == c) continue;
            // sorted insert without duplicates
            int pos = m;
            bool dup = false;
            for (int t = 0; t < m; ++t) {
                if (buf[t] == other) { dup = true; break; }
                if (buf[t] > other) { pos = t; break; }
            }
            if (dup) continue;
            if (m >= kMaxRing) { *error = 1; continue; }
            for (int t = m; t > pos; --t) buf[t] = buf[t - 1];
            buf[pos] = other;
            ++m;
        }
    }
    if (indices) {
        const int o = offsets[c];
        for (int t = 0; t < m; ++t) indices[o + t] = buf[t];
    } else {
        counts[c] = m;
    }
}

// ---------------------------------------------------------------------------------------------
// locate acceleration data, built once at create():
//   edge_nbr[c] = cells across the edges opposite to local vertices 0,1,2 (-1 on the domain boundary),
//   CellGeom.pad = strict-interior margin of cell c as a float (see locate_mover and DESIGN.md):
//       margin_c = 12 * 2e-6 * (longest edge in the mesh) / (smallest height of c), at least 1e-5.
// A cell T' accepts a point p (all barycentrics >= -tol) only if dist(p, T') <= 6 tol diam(T'); a point whose
// barycentrics in T all exceed margin_T is farther than that from every other cell of a non-overlapping mesh.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_cell_metrics(int n_cells, const double2 *__restrict__ vertices, const CellGeom *__restrict__ geom, double *__restrict__ hmin,
               unsigned long long *__restrict__ dmax_bits)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    double longest = 0.0;
    if (c < n_cells) {
        const CellGeom g = geom[c];
        const double2 a = vertices[g.n0], b = vertices[g.n1], z = vertices[g.n2];
        const double e0 = hypot(b.x - z.x, b.y - z.y), e1 = hypot(a.x - z.x, a.y - z.y), e2 = hypot(a.x - b.x, a.y - b.y);
        longest = fmax(e0, fmax(e1, e2));
        const double area2 = fabs((a.x - z.x) * (b.y - z.y) - (a.y - z.y) * (b.x - z.x));
        hmin[c] = area2 / longest;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) longest = fmax(longest, __shfl_xor_sync(0xffffffffu, longest, d));
    if ((threadIdx.x & 31) == 0) atomicMax(dmax_bits, (unsigned long long)__double_as_longlong(longest));
}

__global__ void __launch_bounds__(kThreads)
k_build_locate_data(int n_cells, CellGeom *__restrict__ geom, const int *__restrict__ nbr_off, const int *__restrict__ nbr_idx,
                    const double *__restrict__ hmin, const unsigned long long *__restrict__ dmax_bits, int4 *__restrict__ edge_nbr)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const unsigned a0 = geom[c].n0, a1 = geom[c].n1, a2 = geom[c].n2;
    int4 e = make_int4(-1, -1, -1, 0);
    for (int k = nbr_off[c]; k < nbr_off[c + 1]; ++k) {
        const int nb = nbr_idx[k];
        const unsigned b0 = geom[nb].n0, b1 = geom[nb].n1, b2 = geom[nb].n2;
        const bool s0 = a0 == b0 || a0 == b1 || a0 == b2;
        const bool s1 = a1 == b0 || a1 == b1 || a1 == b2;
        const bool s2 = a2 == b0 || a2 == b1 || a2 == b2;
        if (s1 && s2 && !s0 && e.x < 0) e.x = nb; // shares the edge opposite to vertex 0
        if (s0 && s2 && !s1 && e.y < 0) e.y = nb;
        if (s0 && s1 && !s2 && e.z < 0) e.z = nb;
    }
    edge_nbr[c] = e;
    const double dmax = __longlong_as_double((long long)*dmax_bits);
    double m = 12.0 * 2e-6 * dmax / hmin[c];
    if (!(m >= 1e-5)) m = 1e-5;
    if (!(m < 0.3)) m = 2.0; // degenerate cell: never take the fast path into it
    geom[c].pad = __float_as_uint(__double2float_ru(m * 1.0001));
}

// nodes touched by the owned cells -> compact list (multi-GPU: per-node work only for these)
__global__ void __launch_bounds__(kThreads)
k_mark_nodes(int c_lo, int c_hi, const unsigned *__restrict__ cells, int *__restrict__ flag)
{
    const int c = c_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_hi) return;
    flag[cells[3 * (size_t)c]] = 1;
    flag[cells[3 * (size_t)c + 1]] = 1;
    flag[cells[3 * (size_t)c + 2]] = 1;
}
__global__ void __launch_bounds__(kThreads)
k_compact_nodes(int n_nodes, const int *__restrict__ flag, const int *__restrict__ pos, int *__restrict__ list)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes && flag[i]) list[pos[i]] = i;
}

__global__ void k_set_counters(Counters *ctr, int count, int capacity)
{
    ctr->count = count;
    ctr->live = count;
    ctr->added = 0;
    ctr->lost = 0;
    ctr->movers = 0;
    ctr->overflow = 0;
    ctr->capacity = capacity;
    ctr->n_old = count;
    ctr->n_warps = (count + 31) >> 5;
    ctr->n_movers = 0;
}

__global__ void k_begin_advect(Counters *ctr, int capacity)
{
    ctr->lost = 0;
    ctr->movers = 0;
    ctr->added = 0;
    ctr->capacity = capacity;
    ctr->n_old = ctr->count;
    ctr->n_warps = (ctr->count + 31) >> 5;
}

} // namespace pfem2
