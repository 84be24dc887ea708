// cds_ptx.cuh -- the few PTX wrappers the pipelined match kernels share: mbarriers, TMA bulk copies, named barriers.
#ifndef CDS_PTX_CUH
#define CDS_PTX_CUH

#include <stdint.h>

namespace cds {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// one contiguous global span -> shared memory, completing `bytes` on the mbarrier (TMA bulk copy, SASS UBLKCP)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// The same wait with the waiting warp kept out of the issue slots: mode 1 passes a suspend-time hint (ns) to try_wait (measured: the
// hardware still returns at once, the loop executes as many instructions as the plain one); mode >= 2 polls with test_wait and sleeps
// `mode` nanoseconds in between.  (The plain loop above is 24 % of all instructions the v19 candidate kernel executes,
// profiles/r01_v19_cand_source_hotspots.txt -- and none of its time: the slots it takes are idle anyway.)
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity, int mode)
{
    uint32_t ok;
    if (mode == 1) {
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        } while (!ok);
        return;
    }
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep((unsigned) mode);
    }
}
// L2 eviction policies for the bulk copies and the list loads: the streamed target planes should not push the mask group's word
// lists (re-read for every target) out of L2
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_load_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint2 ldg_hint_v2(const void *p, uint64_t policy)
{
    uint2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ uint4 ldg_hint_v4(const void *p, uint64_t policy)
{
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ uint32_t ldg_hint_u32(const void *p, uint64_t policy)
{
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ uint32_t ldg_hint_u16(const void *p, uint64_t policy)
{
    uint16_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(policy));
    return v;
}

__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// barrier 1 over the N consumer threads (the producer warp never joins it)
template <int N>
__device__ __forceinline__ void consumer_barrier()
{
    asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory");
}

}  // namespace cds
#endif
