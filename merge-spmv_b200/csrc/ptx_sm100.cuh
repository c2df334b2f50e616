// ptx_sm100.cuh -- every inline-PTX instruction the sm_100a kernels use, behind one-line wrappers:
// mbarrier, cp.async.bulk (TMA; SASS UBLKCP), the cross-proxy fence, named barriers, L2 cache policies,
// the cache-hinted x gathers, system-scope stores / loads for the NVLink carry exchange, the global timer.  Kernel code contains no asm of its own, so this file is
// the complete list of architecture-specific instructions of the library.
//
// Test seam: tests/emu/ (a host-side SIMT interpreter used ONLY by the CPU test-suite to run the
// kernel logic without a GPU) replaces this header through MSPMV_PTX_HEADER.  The product build
// never defines it; libmergespmv.so contains only the code below.
#pragma once
#ifdef MSPMV_PTX_HEADER
#include MSPMV_PTX_HEADER
#else

#include <cuda_runtime.h>
#include <stdint.h>

// the block's dynamic shared memory (one spelling for nvcc, another for the CPU interpreter)
#define MSPMV_DYNAMIC_SHARED(name) extern __shared__ __align__(128) unsigned char name[]

namespace mspmv {

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// try_wait parks the warp in hardware for a bounded time and returns false when that time is up
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// System-scope accesses for data exchanged with peer GPUs over NVLink (carry exchange): a relaxed
// 64-bit store into peer memory, and an acquire load to poll a flag a peer writes into local memory.
__device__ __forceinline__ void st_relaxed_sys_u64(uint64_t* p, uint64_t v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u64(uint64_t* p, uint64_t v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys_u64(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// nanosecond wall clock shared by all SMs (bounded spin-waits), and a loud exit
__device__ __forceinline__ uint64_t global_timer_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trap_kernel() { asm volatile("trap;"); }

// x gathers.  x is the only reused data: every gather carries an L2::evict_last policy while the value /
// index / row-offset streams are L2::evict_first, so x stays L2-resident even when it is tens of MB
// (profiles/tuning_r01.txt: 20M-column power-law 12.98 -> 5.78 ms).  Two L1 policies, chosen per warp by
// the span of its columns: ld_gather for scattered columns (no L1 allocation: random gathers have no reuse,
// and L1 capacity is what bounds the number of misses in flight -- profiles/microbench_r01.txt), ld_gather_l1
// for narrow spans (banded / FEM-like locality: lines are re-used).
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float ld_gather(const float* p, uint64_t pol)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_gather(const double* p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_gather_l1(const float* p, uint64_t pol)
{
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_gather_l1(const double* p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

}  // namespace mspmv

#endif  // MSPMV_PTX_HEADER
