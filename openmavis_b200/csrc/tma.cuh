// tma.cuh — bulk asynchronous copies (TMA, cp.async.bulk -> SASS UBLKCP) with mbarrier completion, used to stage image tiles
// in shared memory row by row: one elected warp issues one copy per row, nobody spends registers or issue slots on loads,
// and the CTA overlaps its own setup (zeroing score maps, loading taps) with the copies in flight.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace orbx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy; dst, src and bytes must be multiples of 16
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// global -> shared tile copy through a 3-D tensor map (cuTensorMapEncodeTiled): ONE request moves the whole box; coordinates are
// element indices (x, y, image); parts of the box outside the tensor are zero-filled.  dst must be 128-byte aligned.
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// Blocks until the given phase of the barrier completes.  The suspend-time hint lets the hardware park the thread instead of
// spinning on try_wait (the spin was 29 % of the blur kernel's executed instructions before).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(phase), "r"(0x989680u)
        : "memory");
}

}  // namespace orbx
