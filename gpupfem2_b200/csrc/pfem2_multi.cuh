// pfem2_multi.cuh -- multi-GPU kernels of the strip partition: emigrant packing, immigrant append, P2P (NVLink peer memory)
// publish / wait, projection halo.
#pragma once

#include "pfem2_common.cuh"
#include "pfem2_move.cuh"

namespace pfem2 {

// emigrants = live particles whose cell is outside [own_lo, own_hi).  Pass 1 counts them per destination rank.
static __global__ void __launch_bounds__(kThreads)
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
static __global__ void __launch_bounds__(kThreads)
k_emigrant_pack(ParticleSoA p, const Counters *ctr, int own_lo, int own_hi, const int *__restrict__ bounds, int n_ranks,
                int *__restrict__ rank_cursor, int4 *__restrict__ out, int cell_base)
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
        int4 t = *reinterpret_cast<const int4 *>(p.tail + i);
        t.z += cell_base; // records between strips carry GLOBAL cell ids
        rec[2] = t;
        rec[3] = *reinterpret_cast<const int4 *>(p.vel + i);
        st_cell(p.tail + i, kLostCell);
    }
}

// Fused multi-GPU path: the move pass left the array indices of the emigrants in emig_idx (k_advect_locate_tma); pack
// them grouped by destination rank (rank_cursor starts at the exclusive prefix of the counts) and remove them locally.
static __global__ void __launch_bounds__(kThreads)
k_emigrant_pack_list(ParticleSoA p, const unsigned *__restrict__ emig_idx, int n_emig, const int *__restrict__ bounds, int n_ranks,
                     int *__restrict__ rank_cursor, int4 *__restrict__ out, int cell_base)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_emig; j += gridDim.x * blockDim.x) {
        const unsigned i = emig_idx[j];
        int4 t = *reinterpret_cast<const int4 *>(p.tail + i);
        const int slot = atomicAdd(rank_cursor + rank_of_cell((unsigned)t.z, bounds, n_ranks), 1);
        int4 *rec = out + 4 * (size_t)slot;
        rec[0] = *reinterpret_cast<const int4 *>(p.pos + i);
        rec[1] = *reinterpret_cast<const int4 *>(p.lab + i);
        t.z += cell_base;
        rec[2] = t;
        rec[3] = *reinterpret_cast<const int4 *>(p.vel + i);
        st_cell(p.tail + i, kLostCell);
    }
}

// per-cell statistics of the m records behind the current array end (the immigrants just appended): all "arrived"
static __global__ void __launch_bounds__(kThreads)
k_count_appended(ParticleSoA p, const Counters *ctr, int m, int subcell_mode, int n_cells, int ppc, int level, double sub_step, int *__restrict__ stay,
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
        accumulate_cell_stats(subcell_mode, live, c, L0, L1, L2, 0u, mb, lane, n_cells, ppc, level, sub_step, stay, arrive, cell_mask);
    }
}

// immigrants: 64-byte records appended behind the current array
// (keys != nullptr, lazy re-sort: the appended rows also get their entry in the dense key array the rank pass reads)
static __global__ void __launch_bounds__(kThreads)
k_immigrant_append(ParticleSoA p, Counters *ctr, const int4 *__restrict__ in, int m, unsigned *__restrict__ keys, int cell_base)
{
    const int n = ctr->count;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        const int4 *rec = in + 4 * (size_t)j;
        int4 t = rec[2];
        t.z -= cell_base; // global -> this strip's numbering
        *reinterpret_cast<int4 *>(p.pos + (n + j)) = rec[0];
        *reinterpret_cast<int4 *>(p.lab + (n + j)) = rec[1];
        *reinterpret_cast<int4 *>(p.tail + (n + j)) = t;
        *reinterpret_cast<int4 *>(p.vel + (n + j)) = rec[3];
        if (keys) keys[n + j] = (unsigned)t.z;
    }
}
static __global__ void k_add_count(Counters *ctr, int m)
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

// the move pass (k_advect_locate_tma) left the array indices of the emigrants in emig_idx and their number in
// rank_count[n_ranks]; pack them for the left / right neighbour and remove them locally.  Headers are zeroed by the caller.
static __global__ void __launch_bounds__(kThreads)
k_emigrant_pack_nbr(ParticleSoA p, const unsigned *__restrict__ emig_idx, const int *__restrict__ rank_count, int n_ranks,
                    const int *__restrict__ bounds, int rank, int4 *out_left, int4 *out_right, int cap, Counters *ctr,
                    const unsigned long long *__restrict__ cell_mask, int own_hi, int n_cells, int cell_base)
{
    const int n_emig = rank_count[n_ranks];
    if (blockIdx.x == 0 && threadIdx.x < 4 && out_right) {
        const int c = own_hi + (int)threadIdx.x;
        reinterpret_cast<MigrationHeader *>(out_right)->spill[threadIdx.x] = c < n_cells ? cell_mask[c] : 0ull;
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_emig; j += gridDim.x * blockDim.x) {
        const unsigned i = emig_idx[j];
        int4 t = *reinterpret_cast<const int4 *>(p.tail + i);
        const int dest = rank_of_cell((unsigned)t.z, bounds, n_ranks);
        t.z += cell_base; // records between strips carry GLOBAL cell ids
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
static __global__ void __launch_bounds__(kThreads)
k_immigrant_append_dev(ParticleSoA p, const Counters *ctr, const int4 *__restrict__ buf, int cap, unsigned *__restrict__ keys,
                       int cell_base)
{
    const int m = migration_count(buf, cap), n = ctr->count;
    if ((long long)n + m > ctr->capacity) return; // k_add_count_dev raises the overflow flag
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
        const int4 *rec = buf + 4 * ((size_t)j + 1);
        int4 t = rec[2];
        t.z -= cell_base; // global -> this strip's numbering
        *reinterpret_cast<int4 *>(p.pos + (n + j)) = rec[0];
        *reinterpret_cast<int4 *>(p.lab + (n + j)) = rec[1];
        *reinterpret_cast<int4 *>(p.tail + (n + j)) = t;
        *reinterpret_cast<int4 *>(p.vel + (n + j)) = rec[3];
        if (keys) keys[n + j] = (unsigned)t.z; // lazy re-sort: dense key array of the rank pass
    }
}

// per-cell statistics of those records (all "arrived"), like k_count_appended
static __global__ void __launch_bounds__(kThreads)
k_count_appended_dev(ParticleSoA p, const Counters *ctr, const int4 *__restrict__ buf, int cap, int subcell_mode, int n_cells, int ppc, int level,
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
        accumulate_cell_stats(subcell_mode, live, c, L0, L1, L2, 0u, mb, lane, n_cells, ppc, level, sub_step, stay, arrive, cell_mask);
    }
}

// one thread: the array grows by the appended records; the left neighbour's spilled occupancy bits join this strip's first cells
static __global__ void k_add_count_dev(Counters *ctr, const int4 *__restrict__ buf, int cap, unsigned long long *__restrict__ cell_mask,
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
static __global__ void __launch_bounds__(kThreads)
k_emigrant_pack_p2p(ParticleSoA p, const unsigned *__restrict__ emig_idx, const int *__restrict__ rank_count, int n_ranks,
                    const int *__restrict__ bounds, int rank, int4 *rec_left, int4 *rec_right, int cap, Counters *ctr, int *cursors,
                    int cell_base)
{
    const int n_emig = rank_count[n_ranks];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_emig; j += gridDim.x * blockDim.x) {
        const unsigned i = emig_idx[j];
        int4 t = *reinterpret_cast<const int4 *>(p.tail + i);
        const int dest = rank_of_cell((unsigned)t.z, bounds, n_ranks);
        t.z += cell_base; // records between strips carry GLOBAL cell ids
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
static __global__ void k_p2p_publish_migration(MigrationHeader *hdr_left, unsigned *flag_left, MigrationHeader *hdr_right, unsigned *flag_right,
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

static __global__ void k_p2p_publish_flag(unsigned *flag_left, unsigned *flag_right, unsigned seq)
{
    __threadfence_system();
    if (flag_left) st_release_sys(flag_left, seq);
    if (flag_right) st_release_sys(flag_right, seq);
}

// one thread: the stream continues once both neighbours have delivered sequence number `seq` (or the watchdog fires)
static __global__ void k_p2p_wait(const unsigned *flag_a, const unsigned *flag_b, unsigned seq, Counters *ctr, unsigned long long timeout_ns)
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
static __global__ void __launch_bounds__(kThreads)
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
static __global__ void __launch_bounds__(kThreads)
k_halo_add(double *__restrict__ acc3, const int *__restrict__ idx, int n, const double *__restrict__ halo)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double *a = acc3 + 3 * (size_t)idx[i];
    a[0] = __dadd_rn(a[0], halo[3 * (size_t)i]);
    a[1] = __dadd_rn(a[1], halo[3 * (size_t)i + 1]);
    a[2] = __dadd_rn(a[2], halo[3 * (size_t)i + 2]);
}

static __global__ void k_p2p_init_head(P2PInboxHead *hd, int cap, int n_nodes)
{
    hd->magic = kP2PMagic;
    hd->flag_mig = 0;
    hd->flag_halo = 0;
    hd->capacity_records = cap;
    hd->n_halo_nodes = n_nodes;
}


// ---------------------------------------------------------------------------------------------
// Fused forms of the P2P step (round 2: 30 -> 17 launches per strip and step; at 8 GPUs a step is ~2 ms and every launch +
// drain costs ~5 us).  "Last block" pattern: every block counts in on a device word after a __threadfence(); the block that
// finds the count complete does the one-thread epilogue (headers, flags, counters) and resets the word.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool last_block_done(unsigned *done)
{
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(done, 1u);
        s_last = prev == gridDim.x - 1;
        if (s_last) *done = 0u;
    }
    __syncthreads();
    return s_last;
}

// k_emigrant_pack_p2p + k_p2p_publish_migration
static __global__ void __launch_bounds__(kThreads)
k_p2p_send(ParticleSoA p, const unsigned *__restrict__ emig_idx, const int *__restrict__ rank_count, int n_ranks,
           const int *__restrict__ bounds, int rank, MigrationHeader *hdr_left, unsigned *flag_left, MigrationHeader *hdr_right,
           unsigned *flag_right, int cap, Counters *ctr, int *cursors, int cell_base, const unsigned long long *__restrict__ cell_mask,
           int own_hi, int n_cells, unsigned seq, unsigned *done)
{
    int4 *rec_left = hdr_left ? reinterpret_cast<int4 *>(hdr_left + 1) : nullptr;
    int4 *rec_right = hdr_right ? reinterpret_cast<int4 *>(hdr_right + 1) : nullptr;
    const int n_emig = rank_count[n_ranks];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_emig; j += gridDim.x * blockDim.x) {
        const unsigned i = emig_idx[j];
        int4 t = *reinterpret_cast<const int4 *>(p.tail + i);
        const int dest = rank_of_cell((unsigned)t.z, bounds, n_ranks);
        t.z += cell_base; // records between strips carry GLOBAL cell ids
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
    __threadfence_system(); // this block's peer stores are performed before it counts in
    if (!last_block_done(done)) return;
    if (threadIdx.x == 0) {
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
}

// every block: wait until both neighbours have delivered sequence number `seq` (thread 0 spins, watchdog); false = timed out
__device__ __forceinline__ bool p2p_block_wait(const unsigned *flag_a, const unsigned *flag_b, unsigned seq, Counters *ctr,
                                               unsigned long long timeout_ns)
{
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        int ok = 1;
        const unsigned long long t0 = global_timer_ns();
        const unsigned *flags[2] = {flag_a, flag_b};
        for (int k = 0; k < 2 && ok; ++k) {
            if (!flags[k]) continue;
            while ((int)(ld_acquire_sys(flags[k]) - seq) < 0) {
                if (global_timer_ns() - t0 > timeout_ns) {
                    atomicOr(&ctr->overflow, kOverflowP2PTimeout);
                    ok = 0;
                    break;
                }
                __nanosleep(200);
            }
        }
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}

// k_p2p_wait + (k_immigrant_append_dev + k_count_appended_dev + k_add_count_dev) x 2: the immigrants of both inbox blocks are
// appended behind the array (left block first), counted into the per-cell statistics, and the last block grows the array
static __global__ void __launch_bounds__(kThreads)
k_p2p_receive(ParticleSoA p, Counters *ctr, const unsigned *flag_l, const unsigned *flag_r, unsigned seq, unsigned long long timeout_ns,
              const int4 *__restrict__ buf_l, const int4 *__restrict__ buf_r, int cap, unsigned *__restrict__ keys, int cell_base,
              int subcell_mode, int n_cells, int ppc, int level, double sub_step, int *__restrict__ stay, int *__restrict__ arrive,
              unsigned long long *__restrict__ cell_mask, int own_lo, int own_hi, unsigned *done)
{
    if (!p2p_block_wait(flag_l, flag_r, seq, ctr, timeout_ns)) return;
    const int ml = buf_l ? migration_count(buf_l, cap) : 0, mr = buf_r ? migration_count(buf_r, cap) : 0;
    const int m = ml + mr, n0 = ctr->count;
    const bool fits = (long long)n0 + m <= ctr->capacity;
    const int lane = threadIdx.x & 31;
    if (fits) {
        for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < m; base += gridDim.x * blockDim.x) {
            const int j = base + lane;
            bool live = false;
            unsigned c = 0;
            double L0 = 0, L1 = 0, L2 = 0;
            if (j < m) {
                const int4 *rec = j < ml ? buf_l + 4 * ((size_t)j + 1) : buf_r + 4 * ((size_t)(j - ml) + 1);
                const int4 a = rec[0], b = rec[1], v = rec[3];
                int4 t = rec[2];
                t.z -= cell_base; // global -> this strip's numbering
                *reinterpret_cast<int4 *>(p.pos + (n0 + j)) = a;
                *reinterpret_cast<int4 *>(p.lab + (n0 + j)) = b;
                *reinterpret_cast<int4 *>(p.tail + (n0 + j)) = t;
                *reinterpret_cast<int4 *>(p.vel + (n0 + j)) = v;
                if (keys) keys[n0 + j] = (unsigned)t.z; // lazy re-sort: dense key array of the rank pass
                c = (unsigned)t.z;
                live = c != kLostCell;
                L0 = __hiloint2double(b.y, b.x);
                L1 = __hiloint2double(b.w, b.z);
                L2 = __hiloint2double(t.y, t.x);
            }
            const unsigned mb = __ballot_sync(0xffffffffu, live);
            accumulate_cell_stats(subcell_mode, live, c, L0, L1, L2, 0u, mb, lane, n_cells, ppc, level, sub_step, stay, arrive, cell_mask);
        }
    }
    if (!last_block_done(done)) return;
    if (threadIdx.x == 0) {
        int add = m;
        for (int k = 0; k < 2; ++k) {
            const MigrationHeader *hd = reinterpret_cast<const MigrationHeader *>(k ? buf_r : buf_l);
            if (hd && (hd->count > cap || hd->count < 0 || hd->flags)) ctr->overflow |= kOverflowMigration;
        }
        if (!fits) {
            ctr->overflow |= 1;
            add = 0;
        }
        ctr->count += add;
        ctr->n_old = ctr->count;
        ctr->n_warps = (ctr->count + 31) >> 5;
        if (buf_l) { // the left neighbour's spilled occupancy bits join this strip's first cells
            const MigrationHeader *hd = reinterpret_cast<const MigrationHeader *>(buf_l);
            for (int k = 0; k < 4; ++k)
                if (own_lo + k < own_hi && hd->spill[k]) cell_mask[own_lo + k] |= hd->spill[k];
        }
    }
}

// projection halo, both sides in one launch: k_halo_send x 2 + k_p2p_publish_flag
static __global__ void __launch_bounds__(kThreads)
k_halo_send2(const double *__restrict__ acc3, const int *__restrict__ idx_l, int n_l, double *peer_l, unsigned *flag_l,
             const int *__restrict__ idx_r, int n_r, double *peer_r, unsigned *flag_r, unsigned seq, unsigned *done)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_l + n_r) {
        const bool left = i < n_l;
        const int k = left ? i : i - n_l;
        const double *a = acc3 + 3 * (size_t)(left ? idx_l[k] : idx_r[k]);
        double *o = (left ? peer_l : peer_r) + 3 * (size_t)k;
        o[0] = a[0];
        o[1] = a[1];
        o[2] = a[2];
    }
    __threadfence_system();
    if (!last_block_done(done)) return;
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (flag_l) st_release_sys(flag_l, seq);
        if (flag_r) st_release_sys(flag_r, seq);
    }
}

// k_p2p_wait + k_halo_add x 2
static __global__ void __launch_bounds__(kThreads)
k_halo_recv2(double *__restrict__ acc3, const unsigned *flag_l, const unsigned *flag_r, unsigned seq, Counters *ctr,
             unsigned long long timeout_ns, const int *__restrict__ idx_l, int n_l, const double *__restrict__ halo_l,
             const int *__restrict__ idx_r, int n_r, const double *__restrict__ halo_r)
{
    if (!p2p_block_wait(flag_l, flag_r, seq, ctr, timeout_ns)) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_l + n_r) return;
    const bool left = i < n_l;
    const int k = left ? i : i - n_l;
    double *a = acc3 + 3 * (size_t)(left ? idx_l[k] : idx_r[k]);
    const double *hsrc = (left ? halo_l : halo_r) + 3 * (size_t)k;
    a[0] = __dadd_rn(a[0], hsrc[0]);
    a[1] = __dadd_rn(a[1], hsrc[1]);
    a[2] = __dadd_rn(a[2], hsrc[2]);
}

// projection, multi-GPU flavour: per-node accumulators {sum L v_x, sum L v_y, sum L} without the division, so that the
// contributions of the strips sharing an interface node can be added before kFinalizeVelocityProjection's division
static __global__ void __launch_bounds__(kThreads)
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

static __global__ void __launch_bounds__(kThreads)
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

} // namespace pfem2
