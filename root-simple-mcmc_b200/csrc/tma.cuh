// tma.cuh -- TMA bulk copy (cp.async.bulk) and mbarrier helpers shared by the
// event-likelihood kernels (fake_likelihood.cuh, unbinned_likelihood.cuh) and the
// staged proposal kernel (proposal_staged.cuh).  sm_90+ PTX; SASS: UBLKCP / SYNCS.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace smcmc {

__device__ __forceinline__ uint32_t smemAddr(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smemAddr(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tmaLoad1D(void* dstSmem, const void* srcGlobal, uint32_t bytes,
                                          uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smemAddr(dstSmem)),
        "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
        : "memory");
}

// Shared memory -> global bulk store (one bulk async-group per call site):
// the generic-proxy writes to the source must be ordered before it with
// fenceProxyAsync() by the writing threads.
__device__ __forceinline__ void fenceProxyAsync() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tmaStore1D(void* dstGlobal, const void* srcSmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstGlobal),
                 "r"(smemAddr(srcSmem)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// the source buffer may be reused / the CTA may exit
__device__ __forceinline__ void tmaStoreWaitRead() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// the stored bytes are visible to this thread's later generic loads
__device__ __forceinline__ void tmaStoreWaitAll() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace smcmc
