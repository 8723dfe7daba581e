// TMA (cp.async.bulk.tensor) plumbing for the image-tile kernels: tensor-map encoding on the host through
// the driver entry point (no -lcuda link dependency) and the mbarrier / bulk-tensor PTX on the device.
// A tile load is one instruction issued by one thread; out-of-image bytes arrive as zeros and the kernels
// patch the few border tiles afterwards (BORDER_REFLECT_101 is not something TMA can produce).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace sfe {

// ---- host -----------------------------------------------------------------------------------------
// 3-D u8 tensor {width, height, images} with byte strides {pitch, image_stride}; box {box_w, box_h, 1}.
// Returns false (and leaves *map untouched) when the layout cannot be described: TMA needs a 16-byte
// aligned base and strides that are multiples of 16.
bool tma_encode_u8_3d(CUtensorMap *map, const void *base, uint64_t width, uint64_t height, uint64_t images, uint64_t pitch,
                      uint64_t image_stride, uint32_t box_w, uint32_t box_h);

inline bool tma_layout_ok(const void *base, uint64_t pitch, uint64_t image_stride) {
    return ((uintptr_t)base & 15) == 0 && (pitch & 15) == 0 && (image_stride & 15) == 0;
}

// ---- device ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// box at element coordinates (x, y, z) -> shared memory; completion is signalled on `bar` in bytes
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

// Wait for phase `parity` of the barrier.  A load that never completes (a bad tensor map) traps after a
// bounded number of polls instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0;; spin++) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spin > (1u << 22)) __trap();
    }
}
#endif

}  // namespace sfe
