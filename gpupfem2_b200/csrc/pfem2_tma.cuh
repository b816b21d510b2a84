// pfem2_tma.cuh -- bulk asynchronous copies (TMA, cp.async.bulk / UBLKCP) and mbarrier helpers for sm_100a.
//
// The particle kernels stream their inputs through shared memory: one elected lane per warp issues
// `cp.async.bulk.shared::cluster.global` copies of the warp's next tile (contiguous 16-byte records) and arms an
// mbarrier with the expected byte count; the copy engine lands the bytes without occupying registers or LSU
// slots, and the warp waits on the barrier's phase parity only when it actually needs the tile.  This keeps several
// tiles in flight per warp and removes the exposed DRAM latency of a load -> use chain.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace pfem2 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

// make the barrier initialisation visible to the async proxy (the copy engine)
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// global -> shared bulk copy; bytes is a multiple of 16, both addresses 16-byte aligned; completion is signalled
// on `bar` (complete_tx)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- tensor-map (tiled) copies: a 2-D box of the particle record array <-> shared memory, with the 64-byte swizzle.
// The record array is described as a [rows x 16 int32] tensor (one 64-byte record per row); a box of 32 rows is one
// warp tile (2 KB).  With CU_TENSOR_MAP_SWIZZLE_64B the 16-byte chunk f of row r lands at
//     r * 64 + ((f ^ ((r >> 1) & 3)) << 4)            (tile base 512-byte aligned)
// so that lane r reading "its" record with four 128-bit shared loads is bank-conflict free (a linear 64-byte row
// stride would be a 4-way conflict).  The copy engine moves the bytes: no LSU wavefronts for the global side.
__device__ __forceinline__ void tma_load_tile_2d(void *smem_dst, const void *tmap, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_tile_2d(const void *tmap, int x, int y, const void *smem_src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(x), "r"(y),
                 "r"(smem_u32(smem_src))
                 : "memory");
}
// the same with 32-bit shared-window addresses (no generic pointers in the hot loop)
__device__ __forceinline__ void tma_load_tile_2d(uint32_t smem_dst, const void *tmap, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_dst),
                 "l"(tmap), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_store_tile_2d(const void *tmap, int x, int y, uint32_t smem_src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(x), "r"(y), "r"(smem_src)
                 : "memory");
}
// four arbitrary rows of the [rows x 64 B] record tensor -> 256 contiguous bytes of shared memory (box {16, 1}: measured,
// a 4-row box raises an illegal instruction)
__device__ __forceinline__ void tma_gather4_rows(uint32_t smem_dst, const void *tmap, int4 rows, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                     smem_dst),
                 "l"(tmap), "r"(0), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra WAIT_%=;\n\t}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}
// 128-bit shared-memory accesses by 32-bit shared-window address
__device__ __forceinline__ int4 lds128(uint32_t addr)
{
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, int4 v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned lds32(uint32_t addr)
{
    unsigned v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, unsigned v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N bulk groups of this thread have finished READING their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// per-thread 16-byte asynchronous copy global -> shared (LDGSTS), L2-only caching
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

} // namespace pfem2
