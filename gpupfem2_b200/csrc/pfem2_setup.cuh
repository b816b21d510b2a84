// pfem2_setup.cuh -- one-time kernels: mesh repack, inverse Jacobians, O(C) one-ring builder, locate data, seeding, counters,
// AoS export.
#pragma once

#include "pfem2_common.cuh"

namespace pfem2 {

// ---------------------------------------------------------------------------------------------
// mesh repack: CellGeom[c] = { invJacobi[c], vertices[cells[c].z], cells[c] }
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(kThreads)
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
static __global__ void __launch_bounds__(kThreads)
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
static __global__ void __launch_bounds__(kThreads)
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
// getParticles(): materialise the reference's 96-byte AoS Particle2D records (particle_2d.cuh:51-57;
// ID@0 position@16 localPosition@32 velocity@64 cellID@80, 12 B tail pad).  6 x 16-byte stores each.
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(kThreads) k_export_aos(ParticleSoA p, const Counters *ctr, uint4 *__restrict__ out)
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
static __global__ void __launch_bounds__(kThreads)
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
static __global__ void __launch_bounds__(128)
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
static __global__ void __launch_bounds__(kThreads)
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

static __global__ void __launch_bounds__(kThreads)
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
static __global__ void __launch_bounds__(kThreads)
k_mark_nodes(int c_lo, int c_hi, const unsigned *__restrict__ cells, int *__restrict__ flag)
{
    const int c = c_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c_hi) return;
    flag[cells[3 * (size_t)c]] = 1;
    flag[cells[3 * (size_t)c + 1]] = 1;
    flag[cells[3 * (size_t)c + 2]] = 1;
}
static __global__ void __launch_bounds__(kThreads)
k_compact_nodes(int n_nodes, const int *__restrict__ flag, const int *__restrict__ pos, int *__restrict__ list)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes && flag[i]) list[pos[i]] = i;
}

static __global__ void k_set_counters(Counters *ctr, int count, int capacity)
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

static __global__ void k_begin_advect(Counters *ctr, int capacity)
{
    ctr->lost = 0;
    ctr->movers = 0;
    ctr->added = 0;
    ctr->capacity = capacity;
    ctr->n_old = ctr->count;
    ctr->n_warps = (ctr->count + 31) >> 5;
}

// smallest / largest node id of the cells [cell_lo, cell_hi): out[0] = min (start INT_MAX), out[1] = max (start -1)
static __global__ void __launch_bounds__(kThreads) k_node_minmax(int cell_lo, int cell_hi, const CellGeom *__restrict__ geom, int *out)
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
static __global__ void __launch_bounds__(kThreads) k_band_width(int n_cells, const int *__restrict__ nbr_off, const int *__restrict__ nbr_idx, int *band)
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
//   dn_ready[j] = the smallest node id of any cell behind chunk j (cells [cb[j+1], cb[K])): nodes below it are complete once the
//                 chunks 0..j have been projected
static __global__ void __launch_bounds__(kThreads)
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
        const int b = __reduce_min_sync(0xffffffffu, (valid && c >= __ldg(cb + j + 1) && c < __ldg(cb + K)) ? mn : 0x7fffffff);
        if ((threadIdx.x & 31) == 0) {
            if (a > 0) atomicMax(up_need + j, a);
            if (b != 0x7fffffff) atomicMin(dn_ready + j, b);
        }
    }
}

} // namespace pfem2
