// pfb_bulk.cuh -- one-instruction bulk copies global -> shared memory (cp.async.bulk, the TMA engine's
// non-tensor path: UBLKCP in SASS) completing on an mbarrier.  One elected thread issues the copy, every thread of the
// CTA waits on the barrier's phase; nothing passes through registers and no thread spends issue slots on the copy.
#pragma once
#include <stdint.h>

namespace bulk {

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the async proxy must see the initialised barrier
}

// the calling thread arrives on the barrier, announces `bytes` and starts the copy (dst, src 16-byte aligned, bytes a
// multiple of 16).  Shared memory that other threads have read or written through ordinary loads / stores must be
// handed over to the async proxy first: call after a __syncthreads(); the fence below orders the CTA's earlier accesses
__device__ __forceinline__ void load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// every thread: wait for the phase with this parity; a copy that never lands traps instead of hanging the GPU
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!try_wait(bar, parity))
        if (++spins > (1u << 24)) __trap();
}

}  // namespace bulk
