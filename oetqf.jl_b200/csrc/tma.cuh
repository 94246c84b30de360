// tma.cuh -- minimal inline-PTX wrappers for the sm_100a async-copy machinery used by the matvec:
// 1-D bulk TMA (cp.async.bulk, SASS UBLKCP) completing on an mbarrier, and cache-hinted vector loads.
#pragma once
#include <cstdint>

namespace oq {

__device__ __forceinline__ uint32_t smem_addr(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}

// make the barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// order generic-proxy memory operations before subsequent async-proxy (TMA) operations
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async;" ::: "memory");
}

// global -> shared bulk copy; bytes and both addresses must be multiples of 16
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_addr(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}

// L2 eviction-priority policies for bulk copies (the encodings `createpolicy.fractional.L2::evict_*.b64 p, 1.0`
// produces; the same constants as CUTLASS' TMA::CacheHintSm100)
constexpr unsigned long long kL2EvictNormal = 0x1000000000000000ull;
constexpr unsigned long long kL2EvictFirst = 0x12F0000000000000ull;
constexpr unsigned long long kL2EvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar,
                                                 unsigned long long policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_addr(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
        : "memory");
}

// Programmatic dependent launch: a kernel launched with programmatic stream serialisation may start while its
// predecessor still runs; it must execute pdl_wait() before touching anything the predecessor produces.
__device__ __forceinline__ void pdl_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// lets the dependent kernel of this grid start launching now (its pdl_wait still waits for our completion)
__device__ __forceinline__ void pdl_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// streaming 128-bit load of matrix data: read-only path, do not allocate in L1 (each byte is used once)
__device__ __forceinline__ double2 ldg_stream(const double* p)
{
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// L2-coherent load (bypasses L1) for data written by other CTAs / peers during this kernel
__device__ __forceinline__ double ld_cg(const double* p)
{
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

}  // namespace oq
