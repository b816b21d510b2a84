// pfem2_device.cuh -- device-side building blocks of the B200 PFEM-2 particle step.
//
// Arithmetic contract: the reference is compiled with nvcc's default -fmad=true, so
// the FMA contraction nvcc chose is part of its results.  Every parity-relevant expression below is
// written with explicit __fma_rn/__dmul_rn/__dadd_rn/__dsub_rn intrinsics (which the compiler never
// re-associates or contracts) in exactly the operation order of the reference's sm_100a SASS.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace pfem2 {

constexpr unsigned kLostCell = 0xFFFFFFFFu;
constexpr int kMaxLevel = 8;                 // ppc <= 64 -> one 64-bit occupancy mask per cell
constexpr double kTolHi = 1.0 + 2e-6;        // 1.0 + CONSTANTS::DOUBLE_MIN (geometry.cuh:38), folded in double
constexpr double kTolLo = -2e-6;

// Per-cell record, one 64-byte line: everything the locate / interpolate steps need from the mesh.
// Built once at create() from the borrowed Mesh2D arrays (cells, vertices[cells.z], invJacobi).
struct __align__(16) CellGeom {
    double j0, j1, j2, j3; // Matrix2x2::data, row-major (cuda_math.cuh:180)
    double v3x, v3y;       // vertices[cells[c].z]
    unsigned n0, n1, n2;   // cells[c].{x,y,z}
    unsigned pad;          // float bits: strict-interior margin of the locate fast path (k_build_locate_data)
};
static_assert(sizeof(CellGeom) == 64, "CellGeom must be one 64-byte record");

// Particle storage: ONE array of 64-byte records (exactly two 32-byte DRAM sectors, 64-byte aligned), each made of
// four 16-byte fields, so every access is a 128-bit load / store:
//   +0  pos  = Particle2D::position            (x, y)
//   +16 lab  = Particle2D::localPosition.x, .y (first two barycentrics)
//   +32 tail = localPosition.z, cellID, ID
//   +48 vel  = Particle2D::velocity            (vx, vy)
// Why records and not four field arrays: the per-step re-sort scatters particles in short runs (3-4 records); with
// whole-sector records every store completes its sectors, which measured 3x the scatter bandwidth of the field-array
// layout on B200 (profiles/r01_scatter_pattern_microbench.md).  The kernels address the fields through strided views,
// so `p.pos[i]`, `p.tail + i` read like a structure of arrays.
struct __align__(16) ParticleTail {
    double l2;
    unsigned cell;
    unsigned id;
};
static_assert(sizeof(ParticleTail) == 16, "ParticleTail must be a 16-byte field");

struct __align__(64) ParticleRec {
    double2 pos;
    double2 lab;
    ParticleTail tail;
    double2 vel;
};
static_assert(sizeof(ParticleRec) == 64, "a particle is one 64-byte record");

template <class T> struct FieldView { // field of record i at base + 64 i
    char *base;
    __host__ __device__ __forceinline__ T &operator[](long long i) const { return *reinterpret_cast<T *>(base + i * (long long)sizeof(ParticleRec)); }
    __host__ __device__ __forceinline__ T *operator+(long long i) const { return reinterpret_cast<T *>(base + i * (long long)sizeof(ParticleRec)); }
};

struct ParticleSoA {
    FieldView<double2> pos;
    FieldView<double2> lab;
    FieldView<ParticleTail> tail;
    FieldView<double2> vel;
    __host__ __device__ ParticleRec *records() const { return reinterpret_cast<ParticleRec *>(pos.base); }
    __host__ void bind(ParticleRec *r)
    {
        char *b = reinterpret_cast<char *>(r);
        pos.base = b;
        lab.base = b + 16;
        tail.base = b + 32;
        vel.base = b + 48;
    }
};

__device__ __forceinline__ ParticleTail ld_tail(const ParticleTail *p)
{
    const int4 r = *reinterpret_cast<const int4 *>(p);
    ParticleTail t;
    t.l2 = __hiloint2double(r.y, r.x);
    t.cell = (unsigned)r.z;
    t.id = (unsigned)r.w;
    return t;
}
__device__ __forceinline__ void st_tail(ParticleTail *p, double l2, unsigned cell, unsigned id)
{
    *reinterpret_cast<int4 *>(p) = make_int4(__double2loint(l2), __double2hiint(l2), (int)cell, (int)id);
}
__device__ __forceinline__ unsigned ld_cell(const ParticleTail *p) { return reinterpret_cast<const unsigned *>(p)[2]; }
__device__ __forceinline__ void st_cell(ParticleTail *p, unsigned c) { reinterpret_cast<unsigned *>(p)[2] = c; }

// device-resident counters (one per handle); mirrors pfem2_stats
struct Counters {
    int count;    // live particles (valid prefix of the current SoA)
    int live;     // survivors of the advect in flight
    int added;    // re-seeded by the last advect
    int lost;     // deleted by the last advect
    int movers;   // particle-substeps that left their cell
    int overflow; // capacity exceeded
    int capacity;
    int n_old;    // array length before the advect in flight
    int n_warps;  // ceil(n_old / 32)
    int n_movers; // particles that changed cell in the advect in flight
    int pad[2];
};

__device__ __forceinline__ CellGeom load_geom(const CellGeom *__restrict__ g, unsigned c)
{
    const double2 *p = reinterpret_cast<const double2 *>(g + c);
    const double2 a = __ldg(p), b = __ldg(p + 1), v = __ldg(p + 2);
    const uint4 n = __ldg(reinterpret_cast<const uint4 *>(p + 3));
    CellGeom r;
    r.j0 = a.x; r.j1 = a.y; r.j2 = b.x; r.j3 = b.y;
    r.v3x = v.x; r.v3y = v.y;
    r.n0 = n.x; r.n1 = n.y; r.n2 = n.z; r.pad = n.w;
    return r;
}

// GEOMETRY::transformGlobalToLocal (geometry.cuh:14-23) as compiled into Particle2D::isInsideCell:
//   dx = px - v3x ; dy = py - v3y ; Lx = fma(dx, J0, dy*J2) ; Ly = fma(dx, J1, dy*J3) ; Lz = (1 - Lx) - Ly
__device__ __forceinline__ void to_local(const CellGeom &g, double px, double py, double &L0, double &L1, double &L2)
{
    const double dx = __dsub_rn(px, g.v3x);
    const double dy = __dsub_rn(py, g.v3y);
    L0 = __fma_rn(dx, g.j0, __dmul_rn(dy, g.j2));
    L1 = __fma_rn(dx, g.j1, __dmul_rn(dy, g.j3));
    L2 = __dsub_rn(__dsub_rn(1.0, L0), L1);
}

// GEOMETRY::isPointInsideUnitTriangle (geometry.cuh:37-46); NaN compares false everywhere -> inside.
// Written as two chains of three ordered fp64 compares on predicates (6 DSETP): the plain C++ form
// `(L0 > hi) | (L0 < lo) | ...` is turned into max3 / min3 by the compiler, and fp64 max / min are emulated with
// ~12 instructions each on sm_100a (45 instructions for this test in the hot loop of the advect pass).
__device__ __forceinline__ bool inside_unit(double L0, double L1, double L2)
{
    int out;
    asm("{\n\t"
        ".reg .pred ph, pl;\n\t"
        "setp.gt.f64 ph, %1, %4;\n\t"
        "setp.lt.f64 pl, %1, %5;\n\t"
        "setp.gt.or.f64 ph, %2, %4, ph;\n\t"
        "setp.lt.or.f64 pl, %2, %5, pl;\n\t"
        "setp.gt.or.f64 ph, %3, %4, ph;\n\t"
        "setp.lt.or.f64 pl, %3, %5, pl;\n\t"
        "or.pred ph, ph, pl;\n\t"
        "selp.s32 %0, 1, 0, ph;\n\t"
        "}"
        : "=r"(out)
        : "d"(L0), "d"(L1), "d"(L2), "d"(kTolHi), "d"(kTolLo));
    return out == 0;
}

// determineSubcell (particle_handler_2d.cu:10-33) as compiled:
//   i = trunc((1 - Ly) * n) ; j = trunc(Lx * n) ; res = (i >= 1 ? i*i : 0) + 2j
//   if (j != i) { z = fma(j+1, -step, 1) + fma(i+1, step, -1) ; if (Lz < z) ++res }
// No clamping: for particles in the tolerance band the index leaves [0, ppc).
__device__ __forceinline__ int subcell_index(double L0, double L1, double L2, int n, double step)
{
    const double dn = (double)n;
    const int i = __double2int_rz(__dmul_rn(__dsub_rn(1.0, L1), dn));
    const int j = __double2int_rz(__dmul_rn(L0, dn));
    int res = (i >= 1) ? (int)((unsigned)i * (unsigned)i) : 0;
    res = (int)((unsigned)res + 2u * (unsigned)j);
    if (j != i) {
        const double z = __dadd_rn(__fma_rn((double)(j + 1), -step, 1.0), __fma_rn((double)(i + 1), step, -1.0));
        if (L2 < z) ++res;
    }
    return res;
}

// "fixed" mode (pfem2_options.subcell_mode = 1): row / column clamped into the triangle
__device__ __forceinline__ int subcell_index_clamped(double L0, double L1, double L2, int n, double step)
{
    const double dn = (double)n;
    int i = __double2int_rz(__dmul_rn(__dsub_rn(1.0, L1), dn));
    int j = __double2int_rz(__dmul_rn(L0, dn));
    i = min(max(i, 0), n - 1);
    j = min(max(j, 0), i);
    int res = i * i + 2 * j;
    if (j != i) {
        const double z = __dadd_rn(__fma_rn((double)(j + 1), -step, 1.0), __fma_rn((double)(i + 1), step, -1.0));
        if (L2 < z) ++res;
    }
    return res;
}

// u = fma(L2, V2, fma(L1, V1, fma(L0, V0, 0)))  -- kAdvectParticles / kCorrectParticleVelocity / kAddParticlesToCell
__device__ __forceinline__ double interp3(double L0, double L1, double L2, double a0, double a1, double a2)
{
    double u = __fma_rn(L0, a0, 0.0);
    u = __fma_rn(L1, a1, u);
    u = __fma_rn(L2, a2, u);
    return u;
}

// GEOMETRY::transformLocalToGlobal (geometry.cuh:6-8) as compiled: p = fma(Lz, v2, fma(Lx, v0, Ly*v1))
__device__ __forceinline__ double to_global1(double L0, double L1, double L2, double a0, double a1, double a2)
{
    return __fma_rn(L2, a2, __fma_rn(L0, a0, __dmul_rn(L1, a1)));
}

// nodal velocity components either given directly or through the reference's device pointer table
struct NodalVel {
    const double *x;
    const double *y;
    double *const *table; // deviceVector<double*>::data, or nullptr
    __device__ __forceinline__ void resolve(const double *&px, const double *&py) const
    {
        if (table) {
            px = table[0];
            py = table[1];
        } else {
            px = x;
            py = y;
        }
    }
};

} // namespace pfem2
