// pfem2_resort.cuh -- distribution check (plan), re-sort by owning cell (physical scatter / stable radix path) and re-seeding.
#pragma once

#include "pfem2_common.cuh"
#include "pfem2_move.cuh" // accumulate_cell_stats

namespace pfem2 {

// movers -> (new cell, array index) pairs in array order, at the positions the scan of warp_movers assigns
static __global__ void __launch_bounds__(kThreads)
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
static __global__ void __launch_bounds__(kThreads)
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
static __global__ void __launch_bounds__(kThreads)
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

// Epilogue of the plan's scan (exclusive_scan_dev over packed[own_lo .. own_hi)): the low half of the prefix of cell c is its new
// segment start -- the cursor of the counting sort / rank pass when `cursor` is given --, the low half of the total closes the plan
// (live count, new count, overflow flag).  Formerly two kernels of their own behind the scan (k_init_cursor / k_plan_finish).
struct PlanEpilogue {
    int *cursor; // nullptr: the stable order has no cursors
    Counters *ctr;
    __device__ __forceinline__ void store(int i, unsigned long long x) const
    {
        if (cursor) cursor[i] = (int)(unsigned)(x & 0xffffffffull);
    }
    __device__ __forceinline__ void finish(unsigned long long t) const
    {
        const long long total = (long long)(unsigned)(t & 0xffffffffull);
        ctr->live = (int)total - ctr->added;
        if (total > ctr->capacity) {
            ctr->overflow |= 1; // (other bits: migration / P2P watchdog)
            ctr->count = 0; // nothing valid to iterate over; the host reports PFEM2_ECAPACITY
        } else {
            ctr->count = (int)total;
        }
    }
};

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
static __global__ void __launch_bounds__(kThreads)
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

// movers, sorted by new cell (stable: array order within a cell), go right behind the cell's stayers
static __global__ void __launch_bounds__(kThreads)
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

// The new particles of the 32 cells of a warp (missing > 0 in the lanes whose cell needs some), one LANE PER NEW PARTICLE:
// kAddParticlesToCell (:197-236), one particle at the centre of every empty sub-cell in ascending sub-cell order, velocity interpolated
// from the current nodal field.  The cell of lane o gets the rows d(o) .. d(o) + missing(o) - 1 of `rec`; with src_new != nullptr
// (lazy re-sort) their indices go to src_new[j(o) ..].  All 32 lanes must call.
// (Forms measured before this one, stress case = 1.35 new particles per cell and step: one thread looping over its cell's sub-cells and
// writing up to ppc 64-byte records on its own with an atomic per cell on the one cursor 0.83 ms; the needy cells of a warp served one
// after the other by the whole warp, one atomic per warp 0.32 ms; profiles/r03_summary.md §7.)
__device__ __forceinline__ void reseed_cells_of_warp(int c, int missing, int d, int j, unsigned long long mask, int ppc,
                                                     const double2 *__restrict__ vertices, const CellGeom *__restrict__ geom,
                                                     const double *__restrict__ centers, NodalVel vel, const ParticleSoA &rec,
                                                     unsigned *__restrict__ src_new)
{
    const int lane = threadIdx.x & 31;
    int incl = missing; // inclusive prefix sums of `missing` over the lanes: new particle k of the warp belongs to the lane o with excl(o) <= k < incl(o)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    const int excl = incl - missing;
    const double *Vx, *Vy;
    vel.resolve(Vx, Vy);
    const unsigned long long full = ppc >= 64 ? ~0ull : ((1ull << ppc) - 1ull);
    for (int k0 = 0; k0 < total; k0 += 32) {
        const int k = k0 + lane;
        int o = 0; // binary search over the lanes' inclusive sums (every lane takes part in the shuffles)
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
            if (__shfl_sync(0xffffffffu, incl, o + step - 1) <= k) o += step;
        o = min(o, 31);
        const int cc = __shfl_sync(0xffffffffu, c, o), d0 = __shfl_sync(0xffffffffu, d, o), j0 = __shfl_sync(0xffffffffu, j, o);
        const int r = k - __shfl_sync(0xffffffffu, excl, o); // rank among the cell's empty sub-cells
        const unsigned long long m = ((unsigned long long)__shfl_sync(0xffffffffu, (unsigned)(mask >> 32), o) << 32) |
                                     __shfl_sync(0xffffffffu, (unsigned)mask, o);
        if (k >= total) continue;
        const unsigned long long empty = ~m & full;
        const unsigned lo = (unsigned)empty, hi = (unsigned)(empty >> 32);
        const int nlo = __popc(lo);
        const int s = r < nlo ? (int)__fns(lo, 0, r + 1) : 32 + (int)__fns(hi, 0, r - nlo + 1); // the r-th empty sub-cell
        const uint4 nn = __ldg(reinterpret_cast<const uint4 *>(&geom[cc].n0));
        const double2 v0 = __ldg(&vertices[nn.x]), v1 = __ldg(&vertices[nn.y]), v2 = __ldg(&vertices[nn.z]);
        const double ax0 = __ldg(Vx + nn.x), ax1 = __ldg(Vx + nn.y), ax2 = __ldg(Vx + nn.z);
        const double ay0 = __ldg(Vy + nn.x), ay1 = __ldg(Vy + nn.y), ay2 = __ldg(Vy + nn.z);
        const double L0 = __ldg(&centers[3 * s]), L1 = __ldg(&centers[3 * s + 1]), L2 = __ldg(&centers[3 * s + 2]);
        const int row = d0 + r;
        rec.pos[row] = make_double2(to_global1(L0, L1, L2, v0.x, v1.x, v2.x), to_global1(L0, L1, L2, v0.y, v1.y, v2.y));
        rec.lab[row] = make_double2(L0, L1);
        st_tail(rec.tail + row, L2, (unsigned)cc, (unsigned)row);
        rec.vel[row] = make_double2(interp3(L0, L1, L2, ax0, ax1, ax2), interp3(L0, L1, L2, ay0, ay1, ay2));
        if (src_new) src_new[j0 + r] = (unsigned)row;
    }
}

// kAddParticlesToCell (:197-236): one new particle at the centre of every empty sub-cell, velocity
// interpolated from the current nodal field; written right behind the cell's survivors.  Also
// materialises the segment table cell_start[].
static __global__ void __launch_bounds__(kThreads)
k_reseed(int own_lo, int own_hi, int ppc, const double2 *__restrict__ vertices, const CellGeom *__restrict__ geom,
         const double *__restrict__ centers, NodalVel vel, const unsigned long long *__restrict__ cell_mask,
         const int *__restrict__ stay, const int *__restrict__ arrive, const unsigned long long *__restrict__ packed_start,
         ParticleSoA dst, int *__restrict__ cell_start, const Counters *ctr)
{
    const int c = own_lo + blockIdx.x * blockDim.x + threadIdx.x;
    int missing = 0, d = 0;
    unsigned long long mask = 0;
    if (c <= own_hi) {
        const int start = (int)(unsigned)(packed_start[c] & 0xffffffffull);
        cell_start[c] = start;
        if (c < own_hi && !ctr->overflow) {
            const int live = stay[c] + arrive[c];
            missing = max((int)(unsigned)(packed_start[c + 1] & 0xffffffffull) - start - live, 0);
            mask = missing ? cell_mask[c] : 0ull;
            d = start + live; // right behind the cell's survivors
        }
    }
    reseed_cells_of_warp(c, missing, d, 0, mask, ppc, vertices, geom, centers, vel, dst, nullptr);
}

// ---------------------------------------------------------------------------------------------
// rank pass: the counting sort's scatter, applied to 4-byte indices instead of 64-byte records.
// keys[i] = new cell of record i of the (dense) current buffer, i < n_old; cursor[c] starts at the new segment start of cell c
// (PlanEpilogue of the plan's scan).  One atomic per (warp, cell) group like k_scatter_all_quads; the order inside a cell is the order of atomic
// retirement, as in the fast order today.
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(kThreads)
k_rank(const unsigned *__restrict__ keys, const int *__restrict__ n_old_ptr, int *__restrict__ cursor, unsigned *__restrict__ src_new,
       const Counters *ctr)
{
    if (ctr->overflow) return;
    const int n = *n_old_ptr;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1;
    constexpr int U = 4; // keys per lane and iteration: all loads, then all atomics, are in flight together (as in k_scatter_all_regs)
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long warps_total = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = warp_global * (32 * U); base < n; base += warps_total * (32 * U)) {
        unsigned c[U], peers[U];
        int run[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = base + u * 32 + lane;
            c[u] = i < n ? __ldg(keys + i) : kLostCell;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            peers[u] = __match_any_sync(0xffffffffu, c[u]);
            run[u] = 0;
            if (c[u] != kLostCell && (peers[u] & lt) == 0) run[u] = atomicAdd(cursor + c[u], __popc(peers[u]));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = __shfl_sync(0xffffffffu, run[u], __ffs(peers[u]) - 1);
            if (c[u] != kLostCell) src_new[r + __popc(peers[u] & lt)] = (unsigned)(base + u * 32 + lane);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// re-seeding of the lazy re-sort: kAddParticlesToCell (:197-236) with the new particles APPENDED behind the dense array
// (records [n_old, n_old + added)); their indices fill the tail of the cell's range of src_new.  Also materialises cell_start[] and pads
// src_new to a multiple of 32 behind the last sorted position (cell == own_hi does that).
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(kThreads)
k_reseed_lazy(int own_lo, int own_hi, int ppc, const double2 *__restrict__ vertices, const CellGeom *__restrict__ geom,
              const double *__restrict__ centers, NodalVel vel, const unsigned long long *__restrict__ cell_mask, const int *__restrict__ stay,
              const unsigned long long *__restrict__ packed_start, ParticleSoA rec, const int *__restrict__ n_old_ptr, int *__restrict__ tail_cursor,
              unsigned *__restrict__ src_new, int *__restrict__ cell_start, const Counters *ctr)
{
    // one lane per cell finds out what its cell needs, then the whole warp serves the needy cells (reseed_cells_of_warp)
    const int c = own_lo + blockIdx.x * blockDim.x + threadIdx.x;
    int missing = 0, d = 0, j = 0;
    unsigned long long mask = 0;
    if (c <= own_hi) {
        const int start = (int)(unsigned)(packed_start[c] & 0xffffffffull);
        cell_start[c] = start;
        if (!ctr->overflow) { // (the plan found more live particles than the capacity: `start` may lie behind src_new[])
            if (c == own_hi) { // behind the last sorted position: pad to a whole tile with a valid row
                for (int k = start; k < ((start + 31) & ~31); ++k) src_new[k] = 0u;
            } else {
                const int live = stay[c]; // fast order: everybody was counted into stay[]
                missing = max((int)(unsigned)(packed_start[c + 1] & 0xffffffffull) - start - live, 0);
                if (missing) {
                    mask = cell_mask[c];
                    j = start + live;
                }
            }
        }
    }
    // rows behind the array for the whole warp with ONE atomic (an atomic per needy cell on the one cursor was what the kernel waited
    // for: one address takes ~0.4 atomics per ns, profiles/r03_summary.md §7)
    const int lane = threadIdx.x & 31;
    int incl = missing;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    int base = 0;
    if (lane == 31) base = *n_old_ptr + atomicAdd(tail_cursor, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if ((long long)base + total > ctr->capacity) { // (the plan only checked the number of live particles; the dense array also holds the lost ones)
        if (lane == 31) atomicExch(const_cast<int *>(&ctr->overflow), 1);
        return;
    }
    d = base + incl - missing;
    reseed_cells_of_warp(c, missing, d, j, mask, ppc, vertices, geom, centers, vel, rec, src_new);
}


// ---------------------------------------------------------------------------------------------
// back to the ordinary state: out[j] = in[src[j]] for the sorted positions j < count, four lanes per record (a warp-wide 128-bit
// store covers 512 contiguous bytes of the destination).
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(kThreads)
k_materialize(ParticleSoA in_, ParticleSoA out_, const unsigned *__restrict__ src, const Counters *ctr)
{
    const int n = ctr->count;
    const int lane = threadIdx.x & 31, f = lane & 3, q = lane >> 2;
    const int4 *__restrict__ in = reinterpret_cast<const int4 *>(in_.records());
    int4 *__restrict__ out = reinterpret_cast<int4 *>(out_.records());
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = warp * 8; base < n; base += warps * 8) {
        const long long j = base + q;
        if (j < n) out[j * 4 + f] = __ldcs(in + (long long)__ldg(src + j) * 4 + f);
    }
}


// ---------------------------------------------------------------------------------------------
// multi-GPU (strip partition): the move pass runs without statistics; after particles that left the
// owned cell range have been handed to their new owner and the immigrants appended, one pass over all live
// particles accumulates the per-cell counts and occupancy masks (everybody counts as "arrived").
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(kThreads)
k_count_all(ParticleSoA p, const Counters *ctr, int subcell_mode, int n_cells, int ppc, int level, double sub_step, int *__restrict__ stay,
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
        accumulate_cell_stats(subcell_mode, live, c, L0, L1, L2, 0u, mb, lane, n_cells, ppc, level, sub_step, stay, arrive, cell_mask);
    }
}


} // namespace pfem2
