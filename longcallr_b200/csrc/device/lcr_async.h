/* lcr_async.h — mbarrier / asynchronous-copy primitives shared by the kernels (PTX; SASS: SYNCS.*, UBLKCP, LDGSTS). */
#ifndef LCR_ASYNC_H
#define LCR_ASYNC_H

#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    /* try_wait suspends the thread in hardware up to the hint (ns) before it reports false: waiting warps issue almost nothing */
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "LCR_MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra LCR_MBAR_DONE_%=;\n\t"
        "bra LCR_MBAR_WAIT_%=;\n"
        "LCR_MBAR_DONE_%=:\n\t}" ::"r"(bar), "r"(parity), "r"(20000u) : "memory");
}
/* a wait for threads that run ahead of their partners: back off between probes instead of taking issue slots */
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(64);
    }
}
/* 16-byte asynchronous copy global -> shared (LDGSTS, L2 only); its completion is tied to an mbarrier by cp_async_arrive */
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
/* one pending arrival of the barrier is delivered when all earlier cp.async of this thread have landed */
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
/* 1-D bulk copy global -> shared (TMA engine), completion counted in bytes on an mbarrier; 16-byte aligned addresses and size */
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

#endif
