// pfem2_lazy.cuh -- kernels of the lazy re-sort, pfem2_options.lazy_sort = 1 (EXPERIMENTAL, off by default).
// Part of libpfem2_b200.so since the end of round 1, compiled for sm_100a, NOT yet run on hardware (the GPU budget of the round was
// spent): the parity tests of the path are tests/test_gpu_lazy.py (opt-in: PFEM2_TEST_LAZY=1), the first thing to run in round 2
// (tools/lazy_check.sh).  What IS measured is the mechanism (profiles/r01d_summary.md §4: tile::gather4 moves 64-byte records
// from arbitrary rows into a warp's shared-memory tile at 31-38 G records/s with no LSU work) and the data-movement model of the
// whole step (tools/micro/lazy_resort_model.cu: 12.75 ms against 14.27 ms of record / index traffic at channel16m scale).
//
// Idea: the records move ONCE per step.  A step keeps a permutation src[] (sorted position j -> record index) instead of
// physically re-sorting the array:
//
//   advect   k_advect_locate_lazy   tile j of the SORTED order is gathered through src[] from buffer A (8 x gather4 per 32 records),
//                                   moved exactly like k_advect_locate_tma, and stored DENSE to buffer B at rows 32 j .. 32 j + 31
//                                   (B = "sorted by the cell of the previous step"); the new cell also goes to the dense key array.
//                                   Per-cell survivor counts and occupancy masks as today (accumulate_cell_stats).
//            k_plan_cells, scan, k_plan_finish, k_init_cursor                      unchanged: new segment table from the counts
//            k_rank                 slot = cursor[cell]++ (one atomic per (warp, cell) group) ; src_new[slot] = i    (4 + 4 bytes per particle
//                                   instead of the 128-byte-per-particle scatter)
//            k_reseed_lazy          new particles are appended behind the array; their indices fill the tail of their cell's src_new range
//   project  k_project_cells_lazy   the nine sums of a cell from records[src_new[j]], j in the cell's segment
//   anything else that wants the sorted array (getParticles, download, eager correct, multi-GPU, grow)
//            k_materialize          out[j] = in[src[j]] (four lanes per record), then the handle is in the ordinary state again
//
// Lost particles are never referenced by src_new, so the next move pass compacts them away for free.
#pragma once

#include "pfem2_kernels.cuh"

namespace pfem2 {

// four arbitrary rows of the [rows x 64 B] record tensor -> 256 contiguous bytes of shared memory (box {16, 1}: measured,
// a 4-row box raises an illegal instruction)
__device__ __forceinline__ void tma_gather4_rows(uint32_t smem_dst, const void *tmap, int4 rows, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                     smem_dst),
                 "l"(tmap), "r"(0), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w), "r"(bar)
                 : "memory");
}

__global__ void __launch_bounds__(kThreads) k_iota(unsigned *__restrict__ src, const Counters *ctr, int padded_to)
{
    // identity permutation of the live prefix; the pad up to a multiple of 32 names row 0 (a valid row: gathered, never used)
    const int n = ctr->count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < padded_to; i += gridDim.x * blockDim.x) src[i] = i < n ? (unsigned)i : 0u;
}

// ---------------------------------------------------------------------------------------------
// move pass of the lazy re-sort.  Same arithmetic, same statistics as k_advect_locate_tma (pfem2_kernels.cuh); differences:
//   * input tile = rows src[32 t .. 32 t + 31] of `gmap` (the current buffer), fetched with 8 gather4 operations (lanes 0..7, one each);
//   * output tile = rows 32 t .. of `tmap_out` (the OTHER buffer): the pass is not in place;
//   * keys[32 t + lane] = new cell (kLostCell for lost and padding lanes): what k_rank consumes;
//   * n_sorted = number of sorted positions (live particles of the previous step) comes from ctr->count; src[] is padded with a valid
//     row index up to a multiple of 32.
// The swizzle of `gmap` must be the one of `tmap_out` (64-byte): TMA swizzling is a function of the shared-memory address, so the four
// rows a gather4 drops at tile + 256 k land exactly where a 32-row tile load would have put rows 4 k .. 4 k + 3.   [to be verified on hardware;
// the micro-benchmark verified the linear layout only]
// ---------------------------------------------------------------------------------------------
// SWZ = false (PFEM2_LAZY_SWIZZLE=0): both maps are encoded without swizzle and lane r reads its record at r * 64 (4-way bank conflicts on
// the shared-memory side, correct by construction): the fallback should the hardware check above fail.
template <int SUBCELL_MODE, bool WALK, bool MASK64, int NSUB, bool SWZ>
__global__ void __launch_bounds__(kAdvThreads, kAdvBlocksPerSM)
k_advect_locate_lazy(const __grid_constant__ CUtensorMap gmap, const __grid_constant__ CUtensorMap tmap_out, const int4 *__restrict__ src,
                     unsigned *__restrict__ keys, const CellGeom *__restrict__ geom, const int4 *__restrict__ edge_nbr,
                     const int *__restrict__ nbr_off, const int *__restrict__ nbr_idx, const double2 *__restrict__ V2, double h, int substeps,
                     int n_cells, int ppc, int level, double sub_step, Counters *ctr, int *__restrict__ stay,
                     unsigned long long *__restrict__ cell_mask, const double2 *__restrict__ dV2, const int *__restrict__ chunk_start,
                     int chunk_lo, int chunk_hi)
{
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
        p_lo = min(__ldg(chunk_start + chunk_lo), n) & ~31;
        if (chunk_hi < n_cells) n = min(__ldg(chunk_start + chunk_hi), n) & ~31;
    }
    const int first = p_lo + ((blockIdx.x * warps_per_block + warp) << 5);
    const int stride = (int)gridDim.x * (warps_per_block * 32);
    const uint32_t my0 = (uint32_t)lane * 64 + (SWZ ? ((((uint32_t)lane >> 1) & 3) << 4) : 0u);
    // Fetch of a tile = 8 gather4 operations.  The eight lanes 0..7 each hold the four row indices of "their" quarter-kilobyte and issue
    // one gather4 (the hardware instruction is warp-uniform: ptxas wraps a divergent issue in an elect / R2UR loop over the active lanes,
    // so this costs the same 8 UTMALDG as one lane issuing all eight), but the indices arrive with ONE coalesced 128-byte load per tile
    // that is issued an iteration ahead -- in the one-lane form each 16-byte index load sat right in front of the gather4 that consumed it
    // (the asm statements are memory barriers to the compiler): eight DRAM latencies in a row on the warp's critical path.
    auto load_rows = [&](int base) { // indices of the tile at sorted position `base` (lanes 0..7), zeros otherwise
        int4 r = make_int4(0, 0, 0, 0);
        if (lane < 8 && base < n) r = __ldg(src + ((size_t)base >> 2) + lane);
        return r;
    };
    auto issue = [&](int4 r, uint32_t buf, uint32_t bar) { // all lanes call; expect_tx was armed by lane 0 before the __syncwarp
        if (lane < 8) tma_gather4_rows(buf + (uint32_t)lane * 256u, &gmap, r, bar);
    };
    int4 rows = load_rows(first);
    if (first < n) {
        if (lane == 0) mbar_arrive_expect_tx(bar0, kAdvTileBytes);
        __syncwarp();
        issue(rows, tile0, bar0);
    }
    rows = load_rows(first + stride); // for the fetch at the top of the first iteration
    uint32_t b = 0, par = 0;
    for (int base = first; base < n; base += stride) {
        const uint32_t buf = tile0 + b * kAdvTileBytes;
        const int nxt = base + stride;
        if (lane == 0) {
            bulk_wait_group_read<0>(); // the other buffer's store (previous iteration) has finished reading shared memory
            if (nxt < n) mbar_arrive_expect_tx(bar0 + (b ^ 1) * 8, kAdvTileBytes);
        }
        __syncwarp(); // lanes 1..7 write into that buffer too: behind lane 0's wait
        if (nxt < n) issue(rows, tile0 + (b ^ 1) * kAdvTileBytes, bar0 + (b ^ 1) * 8);
        mbar_wait(bar0 + b * 8, par);
        par ^= b;
        b ^= 1;
        const uint32_t sa0 = buf + my0, sa1 = sa0 ^ 16u, sa2 = sa0 ^ 32u, sa3 = sa0 ^ 48u;
        const int i = base + lane;
        const bool valid = i < n;
        unsigned c = 0;
        double L0 = 0, L1 = 0, L2 = 0;
        int moved = 0;
        bool lost = false;
        if (valid) {
            const int4 r0 = lds128(sa0), r1 = lds128(sa1), r2 = lds128(sa2);
            c = (unsigned)r2.z;
            double x = __hiloint2double(r0.y, r0.x);
            double y = __hiloint2double(r0.w, r0.z);
            L0 = __hiloint2double(r1.y, r1.x);
            L1 = __hiloint2double(r1.w, r1.z);
            L2 = __hiloint2double(r2.y, r2.x);
            CellGeom g = load_geom(geom, c);
            int4 e = __ldg(edge_nbr + c);
            double2 a0 = __ldg(V2 + g.n0), a1 = __ldg(V2 + g.n1), a2 = __ldg(V2 + g.n2);
            if (dV2) { // pending velocity correction (deferred correctParticleVelocity), exactly as in k_advect_locate_tma
                const int4 r3 = lds128(sa3);
                const double2 d0 = __ldg(dV2 + g.n0), d1 = __ldg(dV2 + g.n1), d2 = __ldg(dV2 + g.n2);
                const double vx = __dadd_rn(__hiloint2double(r3.y, r3.x), interp3(L0, L1, L2, d0.x, d1.x, d2.x));
                const double vy = __dadd_rn(__hiloint2double(r3.w, r3.z), interp3(L0, L1, L2, d0.y, d1.y, d2.y));
                sts128(sa3, make_int4(__double2loint(vx), __double2hiint(vx), __double2loint(vy), __double2hiint(vy)));
            }
            const int nsub = NSUB > 0 ? NSUB : substeps;
#pragma unroll 1
            for (int s = 0; s < nsub; ++s) {
                const double ux = interp3(L0, L1, L2, a0.x, a1.x, a2.x);
                const double uy = interp3(L0, L1, L2, a0.y, a1.y, a2.y);
                x = __fma_rn(ux, h, x);
                y = __fma_rn(uy, h, y);
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
        // the dense key array of the rank pass: one coalesced 128-byte store per tile (padding lanes and lost particles: kLostCell)
        keys[i] = (valid && !lost) ? c : kLostCell;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { // dense store into the OTHER buffer at the sorted position of the tile
            tma_store_tile_2d(&tmap_out, 0, base, buf);
            bulk_commit_group();
        }
        rows = load_rows(nxt + stride); // indices of the tile after the next: in flight during the statistics, consumed at the next loop top
        const bool live = valid && !lost;
        // the fast order only needs the number of survivors per cell (stayers + arrivals, accumulate_cell_stats with arrive == nullptr),
        // so the cell the particle started in is not carried through the substep loop
        const unsigned sb = __ballot_sync(0xffffffffu, live), mb = 0u;
        const unsigned lb = __ballot_sync(0xffffffffu, lost);
        const int wm = __reduce_add_sync(0xffffffffu, moved);
        if (lane == 0) {
            if (wm) atomicAdd(&s_mov, wm);
            if (lb) atomicAdd(&s_lost, __popc(lb));
        }
        accumulate_cell_stats<SUBCELL_MODE, MASK64>(live, c, L0, L1, L2, sb, mb, lane, n_cells, ppc, level, sub_step, stay, (int *)nullptr,
                                                    cell_mask);
    }
    if (lane == 0) bulk_wait_group<0>();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_mov) atomicAdd(&ctr->movers, s_mov);
        if (s_lost) atomicAdd(&ctr->lost, s_lost);
    }
}

// ---------------------------------------------------------------------------------------------
// rank pass: the counting sort's scatter, applied to 4-byte indices instead of 64-byte records.
// keys[i] = new cell of record i of the (dense) current buffer, i < n_old; cursor[c] starts at the new segment start of cell c
// (k_init_cursor).  One atomic per (warp, cell) group like k_scatter_all_quads; the order inside a cell is the order of atomic
// retirement, as in the fast order today.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
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
__global__ void __launch_bounds__(kThreads)
k_reseed_lazy(int own_lo, int own_hi, int ppc, const double2 *__restrict__ vertices, const CellGeom *__restrict__ geom,
              const double *__restrict__ centers, NodalVel vel, const unsigned long long *__restrict__ cell_mask, const int *__restrict__ stay,
              const unsigned long long *__restrict__ packed_start, ParticleSoA rec, const int *__restrict__ n_old_ptr, int *__restrict__ tail_cursor,
              unsigned *__restrict__ src_new, int *__restrict__ cell_start, const Counters *ctr)
{
    const int c = own_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (c > own_hi) return;
    const int start = (int)(unsigned)(packed_start[c] & 0xffffffffull);
    cell_start[c] = start;
    if (c == own_hi) { // behind the last sorted position: pad to a whole tile with a valid row
        for (int j = start; j < ((start + 31) & ~31); ++j) src_new[j] = 0u;
        return;
    }
    if (ctr->overflow) return;
    const int live = stay[c]; // fast order: everybody was counted into stay[]
    const int missing = (int)(unsigned)(packed_start[c + 1] & 0xffffffffull) - start - live;
    if (missing <= 0) return;
    const double *Vx, *Vy;
    vel.resolve(Vx, Vy);
    const uint4 nn = __ldg(reinterpret_cast<const uint4 *>(&geom[c].n0));
    const double2 v0 = __ldg(&vertices[nn.x]), v1 = __ldg(&vertices[nn.y]), v2 = __ldg(&vertices[nn.z]);
    const double ax0 = __ldg(Vx + nn.x), ax1 = __ldg(Vx + nn.y), ax2 = __ldg(Vx + nn.z);
    const double ay0 = __ldg(Vy + nn.x), ay1 = __ldg(Vy + nn.y), ay2 = __ldg(Vy + nn.z);
    const unsigned long long mask = cell_mask[c];
    int d = *n_old_ptr + atomicAdd(tail_cursor, missing); // a block of `missing` records behind the array
    if ((long long)d + missing > ctr->capacity) { // (the plan only checked the number of live particles; the dense array also holds the lost ones)
        atomicExch(const_cast<int *>(&ctr->overflow), 1);
        return;
    }
    int j = start + live;
    for (int s = 0; s < ppc; ++s) {
        if ((mask >> s) & 1ull) continue;
        const double L0 = __ldg(&centers[3 * s]), L1 = __ldg(&centers[3 * s + 1]), L2 = __ldg(&centers[3 * s + 2]);
        rec.pos[d] = make_double2(to_global1(L0, L1, L2, v0.x, v1.x, v2.x), to_global1(L0, L1, L2, v0.y, v1.y, v2.y));
        rec.lab[d] = make_double2(L0, L1);
        st_tail(rec.tail + d, L2, (unsigned)c, (unsigned)d);
        rec.vel[d] = make_double2(interp3(L0, L1, L2, ax0, ax1, ax2), interp3(L0, L1, L2, ay0, ay1, ay2));
        src_new[j++] = (unsigned)d;
        ++d;
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

// ---------------------------------------------------------------------------------------------
// back to the ordinary state: out[j] = in[src[j]] for the sorted positions j < count, four lanes per record (a warp-wide 128-bit
// store covers 512 contiguous bytes of the destination).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
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

} // namespace pfem2
