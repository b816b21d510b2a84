/*
 * pfem2_oracle.cpp -- CPU oracle (C++17 / OpenMP) for the PFEM-2 particle step.
 *
 * TEST INFRASTRUCTURE ONLY (see pfem2_oracle.h).  Every function restates one piece of the
 * reference's CUDA path and cites it.  The reference is compiled with nvcc's default -fmad=true, so
 * the fused-multiply-add contraction nvcc 12.9 chose for sm_100a is part of its arithmetic; it is
 * written out here with std::fma and this file must be compiled with -ffp-contract=off.
 * The contraction pattern was read from `cuobjdump -sass` of the reference objects
 * (oracle/dump_ref_sass.sh re-derives it).
 *
 * Build: g++ -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp -shared -fPIC pfem2_oracle.cpp
 */
#include "pfem2_oracle.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr double DOUBLE_MIN = 2e-6;              // src/common/constants.h:5
constexpr double ONE_THIRD = 0.3333333333333333; // src/common/constants.h:6
constexpr unsigned LOST = 0xFFFFFFFFu;

// CUDA F2I.F64.TRUNC semantics (NaN -> 0, saturating), used by `int i = <double>` in device code.
inline int cvt_rzi(double v)
{
    if (std::isnan(v)) return 0;
    if (v >= 2147483647.0) return INT_MAX;
    if (v <= -2147483648.0) return INT_MIN;
    return (int)v;
}

// GEOMETRY::transformGlobalToLocal, src/geometry.cuh:14-23; SASS of Particle2D::isInsideCell:
//   dx = px - v3x ; dy = py - v3y ; Lx = fma(dx, J0, dy*J2) ; Ly = fma(dx, J1, dy*J3) ; Lz = (1 - Lx) - Ly
inline void to_local(const double *J, double v3x, double v3y, double px, double py, double &Lx, double &Ly, double &Lz)
{
    const double dx = px - v3x;
    const double dy = py - v3y;
    Lx = std::fma(dx, J[0], dy * J[2]);
    Ly = std::fma(dx, J[1], dy * J[3]);
    Lz = (1.0 - Lx) - Ly;
}

// GEOMETRY::isPointInsideUnitTriangle, src/geometry.cuh:37-46 (NaN compares false -> counts as inside)
inline bool inside(double Lx, double Ly, double Lz)
{
    if (Lx > 1.0 + DOUBLE_MIN || Lx < -DOUBLE_MIN) return false;
    if (Ly > 1.0 + DOUBLE_MIN || Ly < -DOUBLE_MIN) return false;
    if (Lz > 1.0 + DOUBLE_MIN || Lz < -DOUBLE_MIN) return false;
    return true;
}

// determineSubcell, src/particles/particle_handler_2d.cu:10-33; SASS:
//   i = trunc((1 - Ly) * n) ; j = trunc(Lx * n) ; res = (i >= 1 ? i*i : 0) + 2*j
//   if (j != i) { z = fma(j+1, -step, 1) + fma(i+1, step, -1) ; if (Lz < z) ++res }
inline int subcell(double Lx, double Ly, double Lz, int n, double step)
{
    const int i = cvt_rzi((1.0 - Ly) * (double)n);
    const int j = cvt_rzi(Lx * (double)n);
    int res = (i >= 1) ? (int)((unsigned)i * (unsigned)i) : 0;
    res = (int)((unsigned)res + 2u * (unsigned)j);
    if (j != i) {
        const double z = std::fma((double)(j + 1), -step, 1.0) + std::fma((double)(i + 1), step, -1.0);
        if (Lz < z) ++res;
    }
    return res;
}

// clamped variant ("fixed" mode, not reference behaviour): row/column forced into the triangle
inline int subcell_clamped(double Lx, double Ly, double Lz, int n, double step)
{
    int i = cvt_rzi((1.0 - Ly) * (double)n);
    int j = cvt_rzi(Lx * (double)n);
    i = std::min(std::max(i, 0), n - 1);
    j = std::min(std::max(j, 0), i);
    int res = i * i + 2 * j;
    if (j != i) {
        const double z = std::fma((double)(j + 1), -step, 1.0) + std::fma((double)(i + 1), step, -1.0);
        if (Lz < z) ++res;
    }
    return res;
}

} // namespace

struct orc_handle {
    int n_nodes = 0, n_cells = 0;
    std::vector<double> vert;      // 2 per node
    std::vector<unsigned> cells;   // 3 per cell
    std::vector<double> invJ;      // 4 per cell
    std::vector<int> nbr_off, nbr_idx;
    int level = 1, ppc = 1, mode = 0;
    double step = 1.0;
    std::vector<double> centers;   // 3 per sub-cell

    // particle state (SoA)
    std::vector<double> x, y, l0, l1, l2, vx, vy;
    std::vector<unsigned> cell, id;

    // scratch
    std::vector<int> hist;
    std::vector<int> node_off, node_inc; // node -> (cell*3 + local) incidences, ascending
    int lost[16] = {0};
    int added = 0;

    int count() const { return (int)x.size(); }
};

extern "C" {

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// kCalculateInvJacobi + Matrix2x2::inverse, src/mesh_2d.cu:21-34, src/common/cuda_math.cuh:124-141; SASS:
//   det = fma(d0, d3, -(d1*d2)) ; inv = 1/det (IEEE) ; {d3*inv, -(d1*inv), -(d2*inv), d0*inv}
void orc_inv_jacobi(int n_cells, const double *v, const unsigned *cells, double *out)
{
#pragma omp parallel for schedule(static)
    for (int c = 0; c < n_cells; ++c) {
        const unsigned a = cells[3 * c], b = cells[3 * c + 1], z = cells[3 * c + 2];
        const double d0 = v[2 * a] - v[2 * z], d1 = v[2 * a + 1] - v[2 * z + 1];
        const double d2 = v[2 * b] - v[2 * z], d3 = v[2 * b + 1] - v[2 * z + 1];
        const double det = std::fma(d0, d3, -(d1 * d2));
        const double inv = 1.0 / det;
        out[4 * c + 0] = d3 * inv;
        out[4 * c + 1] = d1 * -inv;
        out[4 * c + 2] = d2 * -inv;
        out[4 * c + 3] = d0 * inv;
    }
}

// Mesh2D::fillCellNeighborIndices, src/mesh_2d.cu:107-139: cells sharing at least one vertex,
// each list ascending (std::set order), self excluded.  Built through a vertex->cell CSR.
void orc_one_ring(int n_nodes, int n_cells, const unsigned *cells, int *offsets, int *indices)
{
    std::vector<int> voff(n_nodes + 1, 0);
    for (int c = 0; c < n_cells; ++c)
        for (int k = 0; k < 3; ++k) ++voff[cells[3 * c + k] + 1];
    for (int n = 0; n < n_nodes; ++n) voff[n + 1] += voff[n];
    std::vector<int> vcell(voff[n_nodes]);
    {
        std::vector<int> cur(voff.begin(), voff.end() - 1);
        for (int c = 0; c < n_cells; ++c)
            for (int k = 0; k < 3; ++k) vcell[cur[cells[3 * c + k]]++] = c;
    }
    auto ring = [&](int c, int *buf) {
        int m = 0;
        for (int k = 0; k < 3; ++k) {
            const unsigned n = cells[3 * c + k];
            for (int p = voff[n]; p < voff[n + 1]; ++p)
                if (vcell[p] != c) buf[m++] = vcell[p];
        }
        std::sort(buf, buf + m);
        return (int)(std::unique(buf, buf + m) - buf);
    };
    int maxdeg = 0;
    for (int n = 0; n < n_nodes; ++n) maxdeg = std::max(maxdeg, voff[n + 1] - voff[n]);
    offsets[0] = 0;
#pragma omp parallel
    {
        std::vector<int> buf(3 * maxdeg + 3);
#pragma omp for schedule(static)
        for (int c = 0; c < n_cells; ++c) offsets[c + 1] = ring(c, buf.data());
    }
    for (int c = 0; c < n_cells; ++c) offsets[c + 1] += offsets[c];
    if (!indices) return;
#pragma omp parallel
    {
        std::vector<int> buf(3 * maxdeg + 3);
#pragma omp for schedule(static)
        for (int c = 0; c < n_cells; ++c) {
            const int m = ring(c, buf.data());
            std::copy(buf.begin(), buf.begin() + m, indices + offsets[c]);
        }
    }
}

// ParticleHandler2D::ParticleHandler2D, src/particles/particle_handler_2d.cu:238-294
orc_handle *orc_create(int n_nodes, int n_cells, const double *vertices, const unsigned *cells,
                       const double *inv_jacobi, const int *nbr_offsets, const int *nbr_indices,
                       int cell_division_level, int max_level, int subcell_mode)
{
    orc_handle *h = new orc_handle;
    h->n_nodes = n_nodes;
    h->n_cells = n_cells;
    h->vert.assign(vertices, vertices + 2 * (size_t)n_nodes);
    h->cells.assign(cells, cells + 3 * (size_t)n_cells);
    h->invJ.resize(4 * (size_t)n_cells);
    if (inv_jacobi)
        std::copy(inv_jacobi, inv_jacobi + 4 * (size_t)n_cells, h->invJ.begin());
    else
        orc_inv_jacobi(n_cells, vertices, cells, h->invJ.data());
    h->nbr_off.assign(nbr_offsets, nbr_offsets + n_cells + 1);
    h->nbr_idx.assign(nbr_indices, nbr_indices + nbr_offsets[n_cells]);

    if (max_level < 1) max_level = 4; // CONSTANTS::MAX_CELL_DIVISION_LEVEL
    const int n = std::max(std::min(cell_division_level, max_level), 1); // :241
    h->level = n;
    h->ppc = n * n;       // :242
    h->step = 1.0 / n;    // :243
    h->mode = subcell_mode;

    // sub-cell centres, :248-274 (host arithmetic, no contraction)
    h->centers.resize(3 * (size_t)h->ppc);
    int num = -1;
    const double dx = 1.0 / n;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 2 * i + 1; ++j) {
            const double xmin = (j / 2) * dx;
            const double xmax = xmin + dx;
            const double ymin = (n - 1 - i) * dx;
            const double ymax = ymin + dx;
            double v0[3] = {xmin, ymax, 1.0 - xmin - ymax};
            double v1[3];
            v1[0] = (j % 2 == 0) ? xmin : xmax;
            v1[1] = (j % 2 == 0) ? ymin : ymax;
            v1[2] = 1.0 - v1[0] - v1[1];
            double v2[3] = {xmax, ymin, 1.0 - xmax - ymin};
            ++num;
            for (int k = 0; k < 3; ++k) h->centers[3 * num + k] = ((v0[k] + v1[k]) + v2[k]) * ONE_THIRD;
        }

    // node -> (cell, local index) incidences, ascending in cell (used by the projection only)
    h->node_off.assign(n_nodes + 1, 0);
    for (int c = 0; c < n_cells; ++c)
        for (int k = 0; k < 3; ++k) ++h->node_off[cells[3 * c + k] + 1];
    for (int i = 0; i < n_nodes; ++i) h->node_off[i + 1] += h->node_off[i];
    h->node_inc.resize(h->node_off[n_nodes]);
    {
        std::vector<int> cur(h->node_off.begin(), h->node_off.end() - 1);
        for (int c = 0; c < n_cells; ++c)
            for (int k = 0; k < 3; ++k) h->node_inc[cur[cells[3 * c + k]]++] = 3 * c + k;
    }
    return h;
}

void orc_destroy(orc_handle *h) { delete h; }

int orc_particles_per_cell(const orc_handle *h) { return h->ppc; }

void orc_subcell_centers(const orc_handle *h, double *out)
{
    std::copy(h->centers.begin(), h->centers.end(), out);
}

int orc_count(const orc_handle *h) { return h->count(); }

void orc_last_stats(const orc_handle *h, int *lost, int *added)
{
    if (lost) std::memcpy(lost, h->lost, sizeof(h->lost));
    if (added) *added = h->added;
}

static void resize_all(orc_handle *h, size_t n)
{
    h->x.resize(n); h->y.resize(n); h->l0.resize(n); h->l1.resize(n); h->l2.resize(n);
    h->vx.resize(n); h->vy.resize(n); h->cell.resize(n); h->id.resize(n);
}

// GEOMETRY::transformLocalToGlobal, src/geometry.cuh:6-8; SASS of kSeedParticlesIntoCell:
//   p = fma(Lz, v2, fma(Lx, v0, Ly*v1)) per component
static inline void to_global(const orc_handle *h, int c, const double *L, double &px, double &py)
{
    const unsigned a = h->cells[3 * c], b = h->cells[3 * c + 1], z = h->cells[3 * c + 2];
    const double *v = h->vert.data();
    px = std::fma(L[2], v[2 * z], std::fma(L[0], v[2 * a], L[1] * v[2 * b]));
    py = std::fma(L[2], v[2 * z + 1], std::fma(L[0], v[2 * a + 1], L[1] * v[2 * b + 1]));
}

// seedParticles + kSeedParticlesIntoCell, :304-320, :35-52.  Slot blocks are handed out in cell
// order here (the reference's atomicAdd order is scheduling dependent).
int orc_seed(orc_handle *h)
{
    const int C = h->n_cells, ppc = h->ppc;
    resize_all(h, (size_t)C * ppc);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; ++c)
        for (int s = 0; s < ppc; ++s) {
            const size_t p = (size_t)c * ppc + s;
            const double *L = &h->centers[3 * s];
            to_global(h, c, L, h->x[p], h->y[p]);
            h->l0[p] = L[0]; h->l1[p] = L[1]; h->l2[p] = L[2];
            h->vx[p] = 0.0; h->vy[p] = 0.0;
            h->cell[p] = (unsigned)c;
            h->id[p] = (unsigned)p;
        }
    return h->count();
}

// u = fma(L2, V[n2], fma(L1, V[n1], fma(L0, V[n0], 0)))   (SASS of kAdvectParticles / kAddParticlesToCell)
static inline double interp(const double *V, const unsigned *tri, double L0, double L1, double L2)
{
    double u = std::fma(L0, V[tri[0]], 0.0);
    u = std::fma(L1, V[tri[1]], u);
    u = std::fma(L2, V[tri[2]], u);
    return u;
}

// kCorrectParticleVelocity, :72-88 (old == nullptr: initParticleVelocity, :322-326)
void orc_correct(orc_handle *h, const double *vx, const double *vy, const double *vx_old, const double *vy_old)
{
    const int n = h->count();
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n; ++p) {
        const unsigned *tri = &h->cells[3 * (size_t)h->cell[p]];
        double ix = 0.0, iy = 0.0;
        const double L[3] = {h->l0[p], h->l1[p], h->l2[p]};
        for (int k = 0; k < 3; ++k) {
            const double dx = vx_old ? vx[tri[k]] - vx_old[tri[k]] : vx[tri[k]];
            const double dy = vy_old ? vy[tri[k]] - vy_old[tri[k]] : vy[tri[k]];
            ix = std::fma(L[k], dx, ix);
            iy = std::fma(L[k], dy, iy);
        }
        h->vx[p] = h->vx[p] + ix;
        h->vy[p] = h->vy[p] + iy;
    }
}

void orc_init_velocity(orc_handle *h, const double *vx, const double *vy)
{
    orc_correct(h, vx, vy, nullptr, nullptr);
}

// advectParticles, :328-342 = S x (kAdvectParticles :54-70 ; sortParticlesInCells :363-395) ;
// checkParticleDistribution :397-423
static int check_distribution(orc_handle *h, const double *vx, const double *vy, int own_lo, int own_hi);

// the S substeps of advectParticles (:333-338) without the distribution check; multi-rank tests call the two halves
int orc_move(orc_handle *h, const double *vx, const double *vy, double dt, int substeps)
{
    const double hstep = dt / substeps; // :330, host double
    std::memset(h->lost, 0, sizeof(h->lost));
    for (int s = 0; s < substeps; ++s) {
        const int n = h->count();
        int lost = 0;
#pragma omp parallel for schedule(static) reduction(+ : lost)
        for (int p = 0; p < n; ++p) {
            unsigned c = h->cell[p];
            const unsigned *tri = &h->cells[3 * (size_t)c];
            // kAdvectParticles: uses the STORED local position and cell
            const double ux = interp(vx, tri, h->l0[p], h->l1[p], h->l2[p]);
            const double uy = interp(vy, tri, h->l0[p], h->l1[p], h->l2[p]);
            const double px = std::fma(ux, hstep, h->x[p]);
            const double py = std::fma(uy, hstep, h->y[p]);
            h->x[p] = px;
            h->y[p] = py;
            // kCheckParticleInCell :117-131 (own cell wins)
            double Lx, Ly, Lz;
            to_local(&h->invJ[4 * (size_t)c], h->vert[2 * tri[2]], h->vert[2 * tri[2] + 1], px, py, Lx, Ly, Lz);
            bool found = inside(Lx, Ly, Lz);
            if (!found) {
                // kCheckParticleInNeighbors :133-162 (ascending one-ring, first hit wins)
                for (int k = h->nbr_off[c]; k < h->nbr_off[c + 1]; ++k) {
                    const int nb = h->nbr_idx[k];
                    const unsigned v3 = h->cells[3 * (size_t)nb + 2];
                    to_local(&h->invJ[4 * (size_t)nb], h->vert[2 * v3], h->vert[2 * v3 + 1], px, py, Lx, Ly, Lz);
                    if (inside(Lx, Ly, Lz)) {
                        c = (unsigned)nb;
                        found = true;
                        break;
                    }
                }
            }
            if (found) {
                h->cell[p] = c;
                h->l0[p] = Lx; h->l1[p] = Ly; h->l2[p] = Lz;
            } else {
                h->cell[p] = LOST;
                ++lost;
            }
        }
        if (s < 16) h->lost[s] = lost;
        if (lost) {
            // kDeleteParticles :164-171 -- set semantics; survivors keep their order here
            // (holes are filled with live particles taken from the tail: O(lost) moves, order unspecified like the reference)
            const size_t keep = (size_t)n - (size_t)lost;
            size_t t = (size_t)n; // one past the candidate tail element
            for (size_t hole = 0; hole < keep; ++hole) {
                if (h->cell[hole] != LOST) continue;
                do { --t; } while (h->cell[t] == LOST);
                h->x[hole] = h->x[t]; h->y[hole] = h->y[t];
                h->l0[hole] = h->l0[t]; h->l1[hole] = h->l1[t]; h->l2[hole] = h->l2[t];
                h->vx[hole] = h->vx[t]; h->vy[hole] = h->vy[t];
                h->cell[hole] = h->cell[t]; h->id[hole] = h->id[t];
            }
            resize_all(h, keep);
        }
    }
    return h->count();
}

int orc_advect(orc_handle *h, const double *vx, const double *vy, double dt, int substeps)
{
    orc_move(h, vx, vy, dt, substeps);
    return check_distribution(h, vx, vy, 0, h->n_cells);
}

// re-seeding restricted to the owned cell range [own_lo, own_hi) (the reference is single-GPU: whole mesh)
int orc_check_distribution(orc_handle *h, const double *vx, const double *vy, int own_lo, int own_hi)
{
    return check_distribution(h, vx, vy, own_lo, own_hi);
}

static int check_distribution(orc_handle *h, const double *vx, const double *vy, int own_lo, int own_hi)
{
    // checkParticleDistribution :397-423
    const int C = h->n_cells, ppc = h->ppc;
    const long long hist_size = (long long)C * ppc;
    h->hist.assign((size_t)hist_size, 0);
    const int n = h->count();
#pragma omp parallel for schedule(static)
    for (int p = 0; p < n; ++p) { // kCountParticlesInSubcells :173-181
        const int sub = h->mode == 0 ? subcell(h->l0[p], h->l1[p], h->l2[p], h->level, h->step)
                                     : subcell_clamped(h->l0[p], h->l1[p], h->l2[p], h->level, h->step);
        // reference: unchecked flat index (unsigned arithmetic); out-of-range writes hit no counter
        const long long k = (long long)(unsigned)(h->cell[p] * (unsigned)ppc + (unsigned)sub);
        if (k < hist_size) {
#pragma omp atomic
            ++h->hist[(size_t)k];
        }
    }
    // kCountParticlesToBeAdded :183-195 ; kAddParticlesToCell :197-236 -- new particles are appended in cell order
    std::vector<int> add_start((size_t)(own_hi - own_lo) + 1, 0);
#pragma omp parallel for schedule(static)
    for (int c = own_lo; c < own_hi; ++c) {
        int m = 0;
        for (int s = 0; s < ppc; ++s) m += h->hist[(size_t)c * ppc + s] == 0;
        add_start[(size_t)(c - own_lo) + 1] = m;
    }
    for (size_t k = 1; k < add_start.size(); ++k) add_start[k] += add_start[k - 1];
    const int added = add_start.back();
    const size_t old_n = (size_t)n;
    resize_all(h, old_n + (size_t)added);
#pragma omp parallel for schedule(static)
    for (int c = own_lo; c < own_hi; ++c) {
        const unsigned *tri = &h->cells[3 * (size_t)c];
        size_t w = old_n + (size_t)add_start[(size_t)(c - own_lo)];
        for (int s = 0; s < ppc; ++s) {
            if (h->hist[(size_t)c * ppc + s] != 0) continue;
            const double *L = &h->centers[3 * s];
            to_global(h, c, L, h->x[w], h->y[w]);
            h->l0[w] = L[0]; h->l1[w] = L[1]; h->l2[w] = L[2];
            h->vx[w] = interp(vx, tri, L[0], L[1], L[2]); // :223-227
            h->vy[w] = interp(vy, tri, L[0], L[1], L[2]);
            h->cell[w] = (unsigned)c;
            h->id[w] = (unsigned)w;
            ++w;
        }
    }
    h->added = added;
    return h->count();
}

// projectVelocityOntoGrid :350-361 = kProjectParticleVelocityOntoGrid :90-107 (t = L_i * v, then sum; the
// reference sums with fp64 atomics in scheduling order) ; kFinalizeVelocityProjection :109-115 (IEEE division).
// Summation order here: per cell in particle order, then per node over incident cells ascending.
static void project_acc(orc_handle *h, double *acc3);

void orc_project(orc_handle *h, double *vx, double *vy)
{
    const int N = h->n_nodes;
    std::vector<double> acc(3 * (size_t)N);
    project_acc(h, acc.data());
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        vx[i] = acc[3 * (size_t)i] / acc[3 * (size_t)i + 2];
        vy[i] = acc[3 * (size_t)i + 1] / acc[3 * (size_t)i + 2];
    }
}

// per-node accumulators {sum L v_x, sum L v_y, sum L} without the division (multi-rank tests add the strips' parts)
void orc_project_accumulate(orc_handle *h, double *acc3) { project_acc(h, acc3); }

static void project_acc(orc_handle *h, double *acc3)
{
    const int C = h->n_cells, N = h->n_nodes, n = h->count();
    std::vector<int> start(C + 1, 0);
    for (int p = 0; p < n; ++p) ++start[h->cell[p] + 1];
    for (int c = 0; c < C; ++c) start[c + 1] += start[c];
    std::vector<int> order(n);
    {
        std::vector<int> cur(start.begin(), start.end() - 1);
        for (int p = 0; p < n; ++p) order[cur[h->cell[p]]++] = p;
    }
    std::vector<double> part(9 * (size_t)C);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; ++c) {
        double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int q = start[c]; q < start[c + 1]; ++q) {
            const int p = order[q];
            const double L[3] = {h->l0[p], h->l1[p], h->l2[p]};
            for (int k = 0; k < 3; ++k) {
                acc[3 * k + 0] += L[k] * h->vx[p];
                acc[3 * k + 1] += L[k] * h->vy[p];
                acc[3 * k + 2] += L[k];
            }
        }
        std::copy(acc, acc + 9, &part[9 * (size_t)c]);
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        double sx = 0.0, sy = 0.0, sw = 0.0;
        for (int q = h->node_off[i]; q < h->node_off[i + 1]; ++q) {
            const double *a = &part[3 * (size_t)h->node_inc[q]];
            sx += a[0]; sy += a[1]; sw += a[2];
        }
        acc3[3 * (size_t)i] = sx;
        acc3[3 * (size_t)i + 1] = sy;
        acc3[3 * (size_t)i + 2] = sw;
    }
}

void orc_download(const orc_handle *h, double *x, double *y, double *l0, double *l1, double *l2,
                  double *vx, double *vy, unsigned *cell, unsigned *id)
{
    auto cp = [](auto &src, auto *dst) { if (dst) std::copy(src.begin(), src.end(), dst); };
    cp(h->x, x); cp(h->y, y); cp(h->l0, l0); cp(h->l1, l1); cp(h->l2, l2);
    cp(h->vx, vx); cp(h->vy, vy); cp(h->cell, cell); cp(h->id, id);
}

void orc_upload(orc_handle *h, int n, const double *x, const double *y, const double *l0, const double *l1,
                const double *l2, const double *vx, const double *vy, const unsigned *cell, const unsigned *id)
{
    resize_all(h, (size_t)n);
    auto cp = [n](const auto *src, auto &dst) { if (src) std::copy(src, src + n, dst.begin()); };
    cp(x, h->x); cp(y, h->y); cp(l0, h->l0); cp(l1, h->l1); cp(l2, h->l2);
    cp(vx, h->vx); cp(vy, h->vy); cp(cell, h->cell); cp(id, h->id);
}

void orc_to_local(const double *J, const double *v3, double px, double py, double *L3)
{
    to_local(J, v3[0], v3[1], px, py, L3[0], L3[1], L3[2]);
}

int orc_inside(const double *L3) { return inside(L3[0], L3[1], L3[2]) ? 1 : 0; }

int orc_subcell(const double *L3, int level) { return subcell(L3[0], L3[1], L3[2], level, 1.0 / level); }

} // extern "C"
