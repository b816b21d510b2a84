// pfem2_project.cuh -- projection (sorted segmented reduction, no atomics) and velocity correction.
#pragma once

#include "pfem2_common.cuh"

namespace pfem2 {

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

// ---------------------------------------------------------------------------------------------
// projection through the permutation: k_project_cells with records[src[j]] instead of records[j].  Same sums, same order inside a
// segment, G lanes per cell.
// ---------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(kThreads)
k_project_cells_lazy(int c_lo, int n_cells, ParticleSoA p, const unsigned *__restrict__ src, const int *__restrict__ cell_start,
                     double *__restrict__ partial)
{
    const int lane = threadIdx.x & (G - 1);
    constexpr int groups_per_warp = 32 / G, groups_per_block = kThreads / G;
    for (int cw = c_lo + blockIdx.x * groups_per_block + (threadIdx.x >> 5) * groups_per_warp; cw < n_cells;
         cw += gridDim.x * groups_per_block) {
        const int c = cw + ((threadIdx.x & 31) / G);
        const bool valid = c < n_cells;
        const int b = valid ? __ldg(cell_start + c) : 0, e = valid ? __ldg(cell_start + c + 1) : 0;
        double acc[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] = 0.0;
        int i = b + lane;
        for (; i + 3 * G < e; i += 4 * G) { // four (then two) particles in flight per lane, as in k_project_cells
            long long r4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) r4[u] = __ldg(src + i + u * G);
            double2 l4[4], v4[4];
            double z4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                l4[u] = p.lab[r4[u]];
                v4[u] = p.vel[r4[u]];
                z4[u] = p.tail[r4[u]].l2;
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
            const long long ra = __ldg(src + i), rb = __ldg(src + i + G);
            const double2 la = p.lab[ra], lb = p.lab[rb];
            const double2 va = p.vel[ra], vb = p.vel[rb];
            const double za = p.tail[ra].l2, zb = p.tail[rb].l2;
            const double La[3] = {la.x, la.y, za}, Lb[3] = {lb.x, lb.y, zb};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                acc[3 * k + 0] = __dadd_rn(acc[3 * k + 0], __dmul_rn(La[k], va.x));
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
            const long long ra = __ldg(src + i);
            const double2 la = p.lab[ra];
            const double2 va = p.vel[ra];
            const double La[3] = {la.x, la.y, p.tail[ra].l2};
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


static __global__ void __launch_bounds__(kThreads)
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


// ---- strips (pfem2_step_host_p2p): the same two node passes over a node id RANGE, restricted to the nodes an owned cell touches ----
__device__ __forceinline__ bool node_accumulate_owned(int i, int own_lo, int own_hi, const int *__restrict__ node_off,
                                                      const int *__restrict__ node_inc, const double *__restrict__ partial, double &sx,
                                                      double &sy, double &sw)
{
    bool owned = false;
    sx = sy = sw = 0.0;
    const int e = __ldg(node_off + i + 1);
    for (int q = __ldg(node_off + i); q < e; ++q) {
        const int inc = __ldg(node_inc + q);
        const int cell = inc / 3;
        owned |= cell >= own_lo && cell < own_hi;
        const double *a = partial + 3 * (size_t)inc;
        sx = __dadd_rn(sx, a[0]);
        sy = __dadd_rn(sy, a[1]);
        sw = __dadd_rn(sw, a[2]);
    }
    return owned;
}

// acc3[i] = the three sums of node i (no division: the neighbour strip's share of an interface node is added first)
static __global__ void __launch_bounds__(kThreads)
k_project_nodes_acc_range(int node_lo, int node_hi, int own_lo, int own_hi, const int *__restrict__ node_off, const int *__restrict__ node_inc,
                          const double *__restrict__ partial, double *__restrict__ acc3)
{
    const int i = node_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= node_hi) return;
    double sx, sy, sw;
    if (!node_accumulate_owned(i, own_lo, own_hi, node_off, node_inc, partial, sx, sy, sw)) return;
    acc3[3 * (size_t)i] = sx;
    acc3[3 * (size_t)i + 1] = sy;
    acc3[3 * (size_t)i + 2] = sw;
}

// kFinalizeVelocityProjection over the owned nodes of [node_lo, node_hi)
static __global__ void __launch_bounds__(kThreads)
k_project_finalize_range(int node_lo, int node_hi, int own_lo, int own_hi, const int *__restrict__ node_off, const int *__restrict__ node_inc,
                         const double *__restrict__ acc3, double *__restrict__ vx, double *__restrict__ vy)
{
    const int i = node_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= node_hi) return;
    bool owned = false;
    const int e = __ldg(node_off + i + 1);
    for (int q = __ldg(node_off + i); q < e && !owned; ++q) {
        const int cell = __ldg(node_inc + q) / 3;
        owned = cell >= own_lo && cell < own_hi;
    }
    if (!owned) return;
    const double sw = acc3[3 * (size_t)i + 2];
    vx[i] = __ddiv_rn(acc3[3 * (size_t)i], sw);
    vy[i] = __ddiv_rn(acc3[3 * (size_t)i + 1], sw);
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
static __global__ void __launch_bounds__(kThreads)
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

} // namespace pfem2
