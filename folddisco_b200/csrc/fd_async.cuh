// fd_async.cuh -- sm_100a bulk asynchronous copies (TMA engine, 1-D form) and mbarrier primitives.
//
// The posting-list scan stages the 64-byte granules of the compressed posting lists into shared memory with
// cp.async.bulk (SASS: UBLKCP) and waits for them on transaction-counting mbarriers (SASS: SYNCS), one two-stage ring
// per warp, so that the HBM latency of step s + 1 hides under the varint decode of step s.
#pragma once
#include <stdint.h>

namespace fda {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// makes the initialised barriers visible to the async proxy (the TMA engine completes transactions on them)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// one arrival + announce `bytes` of asynchronous transactions for the current phase
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// global -> shared bulk copy of `bytes` (multiple of 16; src and dst 16-byte aligned); completion is signalled as
// `bytes` of transaction on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// fire-and-forget reductions on a 32-bit shared-memory address (votes of the posting scan)
__device__ __forceinline__ void red_add_shared(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_or_shared(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

} // namespace fda
