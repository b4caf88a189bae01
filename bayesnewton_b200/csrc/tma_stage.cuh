// Staging of a read-only global array into shared memory by the TMA engine: one elected thread issues
// cp.async.bulk (SASS UBLKCP) for the whole array and the CTA waits on the mbarrier the copy signals -- no thread
// spends load / store instructions or registers on the copy (the probit log-density table, 128 KB per CTA).
#pragma once
#include <cstdint>
#include "real.cuh"

namespace BN_NS {

// all threads of the CTA call this; returns when `bytes` (a multiple of 16, dst / src 16-byte aligned) are in shared memory
__device__ __forceinline__ void tma_stage_to_smem(void* smem_dst, const void* gsrc, uint32_t bytes) {
    __shared__ __align__(8) uint64_t bar;
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    const uint32_t dst_a = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // the async proxy sees the initialised barrier
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        constexpr uint32_t kPiece = 32768;
        for (uint32_t off = 0; off < bytes; off += kPiece) {
            const uint32_t n = (bytes - off < kPiece) ? bytes - off : kPiece;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst_a + off), "l"(reinterpret_cast<const char*>(gsrc) + off), "r"(n), "r"(bar_a)
                         : "memory");
        }
    }
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar_a) : "memory");
        if (spins > (1u << 24)) __trap();  // a copy that never lands is an error, not a hang
    }
}

}  // namespace BN_NS
