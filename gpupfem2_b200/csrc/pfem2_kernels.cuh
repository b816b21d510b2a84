// pfem2_kernels.cuh -- the particle-step kernels (sm_100a).
//
// Data layout in HBM: particles are nine SoA arrays (64 B of state per particle),
// kept PHYSICALLY SORTED BY OWNING CELL after every advect, so cell c owns the contiguous segment
// [cell_start[c], cell_start[c+1]).  Mesh data the path reads is repacked once into one 64-byte
// CellGeom record per cell.  All particle counts live in device memory (Counters); kernels are
// grid-stride and read the live count themselves, so a step issues no device->host copy.
#pragma once

#include "pfem2_device.cuh"
#include "pfem2_sort.cuh"

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
k_seed(int n_cells, int ppc, const double2 *__restrict__ vertices, const CellGeom *__restrict__ geom,
       const double *__restrict__ centers, ParticleSoA p, int *__restrict__ cell_start, Counters *ctr)
{
    const long long total = (long long)n_cells * ppc;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i / ppc), s = (int)(i - (long long)c * ppc);
        const uint4 nn = __ldg(reinterpret_cast<const uint4 *>(&geom[c].n0));
        const double2 v0 = __ldg(&vertices[nn.x]), v1 = __ldg(&vertices[nn.y]), v2 = __ldg(&vertices[nn.z]);
        const double L0 = __ldg(&centers[3 * s]), L1 = __ldg(&centers[3 * s + 1]), L2 = __ldg(&centers[3 * s + 2]);
        p.x[i] = to_global1(L0, L1, L2, v0.x, v1.x, v2.x);
        p.y[i] = to_global1(L0, L1, L2, v0.y, v1.y, v2.y);
        p.l0[i] = L0;
        p.l1[i] = L1;
        p.l2[i] = L2;
        p.vx[i] = 0.0;
        p.vy[i] = 0.0;
        p.cell[i] = (unsigned)c;
        p.id[i] = (unsigned)i;
        if (s == 0) cell_start[c] = (int)i;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cell_start[n_cells] = (int)total;
        ctr->count = (int)total;
        ctr->live = (int)total;
        ctr->added = 0;
        ctr->lost = 0;
        ctr->movers = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// advect + locate, all S substeps fused in one pass over the particles
//   kAdvectParticles :54-70, kCheckParticleInCell :117-131, kCheckParticleInNeighbors :133-162.
// The nodal field is frozen inside advectParticles and particles do not interact, so the S substeps
// need no global barrier between them.  A particle with no accepting cell in
// own ∪ one-ring is marked lost (cell = kLostCell) and dropped by the following sort.
// Also accumulates, per surviving particle: the live count of its cell, and its sub-cell bit in the
// cell occupancy mask with the reference's flat, unclamped index (kCountParticlesInSubcells :173-181).
// ---------------------------------------------------------------------------------------------
template <int SUBCELL_MODE>
__global__ void __launch_bounds__(kThreads)
k_advect_locate(ParticleSoA p, const CellGeom *__restrict__ geom, const int *__restrict__ nbr_off,
                const int *__restrict__ nbr_idx, NodalVel vel, double h, int substeps, int n_cells, int ppc, int level,
                double sub_step, Counters *ctr, unsigned *__restrict__ keys, unsigned *__restrict__ vals,
                int *__restrict__ cell_count, unsigned long long *__restrict__ cell_mask)
{
    const double *__restrict__ Vx, *__restrict__ Vy;
    {
        const double *a, *b;
        vel.resolve(a, b);
        Vx = a;
        Vy = b;
    }
    const int n = ctr->count;
    int my_movers = 0, my_lost = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned c = p.cell[i];
        double x = p.x[i], y = p.y[i];
        double L0 = p.l0[i], L1 = p.l1[i], L2 = p.l2[i];
        bool lost = false;
        for (int s = 0; s < substeps; ++s) {
            const CellGeom g = load_geom(geom, c);
            // kAdvectParticles: velocity from the STORED local position and cell
            const double ux = interp3(L0, L1, L2, __ldg(Vx + g.n0), __ldg(Vx + g.n1), __ldg(Vx + g.n2));
            const double uy = interp3(L0, L1, L2, __ldg(Vy + g.n0), __ldg(Vy + g.n1), __ldg(Vy + g.n2));
            x = __fma_rn(ux, h, x);
            y = __fma_rn(uy, h, y);
            // own cell first (wins even if a neighbour would also accept)
            double a0, a1, a2;
            to_local(g, x, y, a0, a1, a2);
            if (inside_unit(a0, a1, a2)) {
                L0 = a0; L1 = a1; L2 = a2;
                continue;
            }
            ++my_movers;
            // ascending one-ring, first accepting cell wins
            bool found = false;
            const int k1 = __ldg(nbr_off + c + 1);
            for (int k = __ldg(nbr_off + c); k < k1; ++k) {
                const unsigned nb = (unsigned)__ldg(nbr_idx + k);
                const CellGeom gn = load_geom(geom, nb);
                to_local(gn, x, y, a0, a1, a2);
                if (inside_unit(a0, a1, a2)) {
                    c = nb;
                    L0 = a0; L1 = a1; L2 = a2;
                    found = true;
                    break;
                }
            }
            if (!found) {
                lost = true;
                break;
            }
        }
        p.x[i] = x;
        p.y[i] = y;
        vals[i] = (unsigned)i;
        if (lost) {
            p.cell[i] = kLostCell;
            keys[i] = (unsigned)n_cells; // sorts behind every live particle
            ++my_lost;
        } else {
            p.l0[i] = L0;
            p.l1[i] = L1;
            p.l2[i] = L2;
            p.cell[i] = c;
            keys[i] = c;
            atomicAdd(cell_count + c, 1);
            const int sub = SUBCELL_MODE == 0 ? subcell_index(L0, L1, L2, level, sub_step)
                                              : subcell_index_clamped(L0, L1, L2, level, sub_step);
            // reference: ++hist[cellID * ppc + sub] in unsigned arithmetic, unchecked
            const unsigned long long flat = (unsigned long long)(unsigned)(c * (unsigned)ppc + (unsigned)sub);
            if (flat < (unsigned long long)n_cells * ppc) {
                const unsigned fc = (unsigned)(flat / (unsigned)ppc);
                atomicOr(cell_mask + fc, 1ull << (unsigned)(flat - (unsigned long long)fc * ppc));
            }
        }
    }
    // block-level reduction of the two statistics counters
    __shared__ int s_mov, s_lost;
    if (threadIdx.x == 0) s_mov = s_lost = 0;
    __syncthreads();
    if (my_movers) atomicAdd(&s_mov, my_movers);
    if (my_lost) atomicAdd(&s_lost, my_lost);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_mov) atomicAdd(&ctr->movers, s_mov);
        if (s_lost) atomicAdd(&ctr->lost, s_lost);
    }
}

// keys / counts of an existing (possibly unsorted) array, used by upload(): no motion, no re-seeding
__global__ void __launch_bounds__(kThreads)
k_keys_from_cells(ParticleSoA p, int n_cells, Counters *ctr, unsigned *__restrict__ keys, unsigned *__restrict__ vals,
                  int *__restrict__ cell_count)
{
    const int n = ctr->count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned c = p.cell[i];
        vals[i] = (unsigned)i;
        if (c < (unsigned)n_cells) {
            keys[i] = c;
            atomicAdd(cell_count + c, 1);
        } else {
            keys[i] = (unsigned)n_cells;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// distribution check, planning part: kCountParticlesToBeAdded (:183-195)
//   packed[c] = live(c) | missing(c) << 32   (one 64-bit scan yields both prefix sums)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_plan_cells(int n_cells, int ppc, int reseed, const int *__restrict__ cell_count,
             const unsigned long long *__restrict__ cell_mask, unsigned long long *__restrict__ packed)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const unsigned long long full = ppc >= 64 ? ~0ull : ((1ull << ppc) - 1ull);
    const int missing = reseed ? ppc - __popcll(cell_mask[c] & full) : 0;
    packed[c] = (unsigned long long)(unsigned)cell_count[c] | ((unsigned long long)(unsigned)missing << 32);
}

__global__ void k_plan_finish(int n_cells, const unsigned long long *__restrict__ packed_start, Counters *ctr)
{
    const unsigned long long t = packed_start[n_cells];
    const int live = (int)(unsigned)(t & 0xffffffffull), added = (int)(unsigned)(t >> 32);
    ctr->live = live;
    ctr->added = added;
    const long long total = (long long)live + added;
    if (total > ctr->capacity) {
        ctr->overflow = 1;
        ctr->count = 0; // nothing valid to iterate over; the host reports PFEM2_ECAPACITY
    } else {
        ctr->count = (int)total;
    }
}

// gather the survivors into cell order: sorted position j (within the live particles) -> destination
// j + (number of re-seeded particles in all earlier cells).  src order within a cell is the stable
// sort order, i.e. the previous array order.
__global__ void __launch_bounds__(kThreads)
k_gather_sorted(ParticleSoA src, ParticleSoA dst, const unsigned *__restrict__ keys_sorted,
                const unsigned *__restrict__ vals_sorted, const unsigned long long *__restrict__ packed_start,
                const Counters *ctr)
{
    if (ctr->overflow) return;
    const int live = ctr->live;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < live; j += gridDim.x * blockDim.x) {
        const unsigned c = keys_sorted[j];
        const unsigned s = vals_sorted[j];
        const int d = j + (int)(unsigned)(packed_start[c] >> 32);
        dst.x[d] = src.x[s];
        dst.y[d] = src.y[s];
        dst.l0[d] = src.l0[s];
        dst.l1[d] = src.l1[s];
        dst.l2[d] = src.l2[s];
        dst.vx[d] = src.vx[s];
        dst.vy[d] = src.vy[s];
        dst.cell[d] = c;
        dst.id[d] = src.id[s];
    }
}

// kAddParticlesToCell (:197-236): one new particle at the centre of every empty sub-cell, velocity
// interpolated from the current nodal field; written right behind the cell's survivors.  Also
// materialises the segment table cell_start[].
__global__ void __launch_bounds__(kThreads)
k_reseed(int n_cells, int ppc, const double2 *__restrict__ vertices, const CellGeom *__restrict__ geom,
         const double *__restrict__ centers, NodalVel vel, const unsigned long long *__restrict__ cell_mask,
         const unsigned long long *__restrict__ packed_start, ParticleSoA dst, int *__restrict__ cell_start,
         const Counters *ctr)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_cells) return;
    const unsigned long long ps = packed_start[c];
    const int start = (int)(unsigned)(ps & 0xffffffffull) + (int)(unsigned)(ps >> 32);
    cell_start[c] = start;
    if (c == n_cells || ctr->overflow) return;
    const unsigned long long pn = packed_start[c + 1];
    const int missing = (int)(unsigned)(pn >> 32) - (int)(unsigned)(ps >> 32);
    if (missing == 0) return;
    const int live = (int)(unsigned)(pn & 0xffffffffull) - (int)(unsigned)(ps & 0xffffffffull);
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
        dst.x[d] = to_global1(L0, L1, L2, v0.x, v1.x, v2.x);
        dst.y[d] = to_global1(L0, L1, L2, v0.y, v1.y, v2.y);
        dst.l0[d] = L0;
        dst.l1[d] = L1;
        dst.l2[d] = L2;
        dst.vx[d] = interp3(L0, L1, L2, ax0, ax1, ax2);
        dst.vy[d] = interp3(L0, L1, L2, ay0, ay1, ay2);
        dst.cell[d] = (unsigned)c;
        dst.id[d] = (unsigned)d;
        ++d;
    }
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
k_project_cells(int n_cells, ParticleSoA p, const int *__restrict__ cell_start, double *__restrict__ partial)
{
    const int lane = threadIdx.x & (G - 1);
    constexpr int groups_per_warp = 32 / G, groups_per_block = kThreads / G;
    // the loop bound is warp-uniform (the full-mask shuffles below need all 32 lanes), cells are guarded inside
    for (int cw = blockIdx.x * groups_per_block + (threadIdx.x >> 5) * groups_per_warp; cw < n_cells;
         cw += gridDim.x * groups_per_block) {
        const int c = cw + ((threadIdx.x & 31) / G);
        const bool valid = c < n_cells;
        const int b = valid ? __ldg(cell_start + c) : 0, e = valid ? __ldg(cell_start + c + 1) : 0;
        double acc[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] = 0.0;
        for (int i = b + lane; i < e; i += G) {
            const double L[3] = {p.l0[i], p.l1[i], p.l2[i]};
            const double vx = p.vx[i], vy = p.vy[i];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                acc[3 * k + 0] = __dadd_rn(acc[3 * k + 0], __dmul_rn(L[k], vx)); // t = L_i * v (plain mul), then add
                acc[3 * k + 1] = __dadd_rn(acc[3 * k + 1], __dmul_rn(L[k], vy));
                acc[3 * k + 2] = __dadd_rn(acc[3 * k + 2], L[k]);
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
k_project_nodes(int n_nodes, const int *__restrict__ node_off, const int *__restrict__ node_inc,
                const double *__restrict__ partial, double *vx_arg, double *vy_arg, double *const *table)
{
    double *Vx = table ? table[0] : vx_arg;
    double *Vy = table ? table[1] : vy_arg;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    double sx = 0.0, sy = 0.0, sw = 0.0;
    const int e = __ldg(node_off + i + 1);
    for (int q = __ldg(node_off + i); q < e; ++q) {
        const double *a = partial + 3 * (size_t)__ldg(node_inc + q);
        sx = __dadd_rn(sx, a[0]);
        sy = __dadd_rn(sy, a[1]);
        sw = __dadd_rn(sw, a[2]);
    }
    Vx[i] = __ddiv_rn(sx, sw);
    Vy[i] = __ddiv_rn(sy, sw);
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
        const unsigned c = p.cell[i];
        const uint4 nn = __ldg(reinterpret_cast<const uint4 *>(&geom[c].n0));
        const double L0 = p.l0[i], L1 = p.l1[i], L2 = p.l2[i];
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
        p.vx[i] = __dadd_rn(p.vx[i], interp3(L0, L1, L2, dx0, dx1, dx2));
        p.vy[i] = __dadd_rn(p.vy[i], interp3(L0, L1, L2, dy0, dy1, dy2));
    }
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
        const double x = p.x[i], y = p.y[i], l0 = p.l0[i], l1 = p.l1[i], l2 = p.l2[i], vx = p.vx[i], vy = p.vy[i];
        rec[0] = make_uint4(p.id[i], 0u, 0u, 0u);
        reinterpret_cast<double2 *>(rec)[1] = make_double2(x, y);
        reinterpret_cast<double2 *>(rec)[2] = make_double2(l0, l1);
        reinterpret_cast<double2 *>(rec)[3] = make_double2(l2, 0.0);
        reinterpret_cast<double2 *>(rec)[4] = make_double2(vx, vy);
        rec[5] = make_uint4(p.cell[i], 0u, 0u, 0u);
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

__global__ void k_set_counters(Counters *ctr, int count, int capacity)
{
    ctr->count = count;
    ctr->live = count;
    ctr->added = 0;
    ctr->lost = 0;
    ctr->movers = 0;
    ctr->overflow = 0;
    ctr->capacity = capacity;
}

__global__ void k_begin_advect(Counters *ctr, int capacity)
{
    ctr->lost = 0;
    ctr->movers = 0;
    ctr->added = 0;
    ctr->capacity = capacity;
}

} // namespace pfem2
