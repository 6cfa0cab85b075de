// Thin inline-PTX wrappers for the sm_100a async-copy machinery used by the kernels:
// mbarrier (arrive/expect_tx/try_wait) and TMA tiled tensor copies (cp.async.bulk.tensor).
#pragma once
#include <cuda.h>
#include <cstdint>

namespace spb
{
    __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

    __device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    }
    __device__ __forceinline__ void fence_mbar_init()
    {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __device__ __forceinline__ void fence_proxy_async()
    {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    }
    __device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
    {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "SPB_WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra SPB_DONE_%=;\n\t"
            "bra SPB_WAIT_%=;\n\t"
            "SPB_DONE_%=:\n\t"
            "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
    }

    // global -> shared, 4-D tile, completion on an mbarrier (SASS: UTMALDG)
    __device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar,
                                                int c0, int c1, int c2, int c3)
    {
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            :: "r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    // shared -> global, 4-D tile (SASS: UTMASTG); out-of-range elements are clipped by the hardware
    __device__ __forceinline__ void tma_store_4d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1, int c2, int c3)
    {
        asm volatile(
            "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
            :: "l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    // shared -> global with fp64 add (rhs += tile), used by the `increment` trait
    __device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1, int c2, int c3)
    {
        asm volatile(
            "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
            :: "l"(tmap), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    // global -> L2 only (no shared memory, no barrier): warms the tile a step before the real load
    __device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* tmap, int c0, int c1, int c2, int c3)
    {
        asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
            :: "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    __device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
    template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
    template <int N> __device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
    __device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tmap)
    {
        asm volatile("prefetch.tensormap [%0];" :: "l"(tmap) : "memory");
    }
}
