// Probe: what does tcgen05.ld.32x32b.x32.pack::16b return?  Columns 0 .. 127 of every lane are written with (column + 256 * (lane & 63))
// (tcgen05.st), then read back packed; prints how the 32 registers of lanes 0 and 1 map to columns.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) probe(uint32_t *out) {
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base + ((uint32_t)(32 * warp) << 16);
    for (int c0 = 0; c0 < 128; c0 += 8) {
        uint32_t r[8];
        for (int j = 0; j < 8; j++) r[j] = (uint32_t)(c0 + j) + 256u * (uint32_t)(tid & 63) - ((c0 + j) % 5 == 0 ? 300u : 0u);  // some negative values
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(base + c0), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                     "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                     : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t v[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(base)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; i++) out[tid * 32 + i] = v[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}
int main() {
    uint32_t *d, h[128 * 32];
    cudaMalloc(&d, sizeof(h));
    probe<<<1, 128>>>(d);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int lane : {0, 1, 33}) {
        printf("thread %d (values written: column + %d, minus 300 when column %% 5 == 0):\n ", lane, 256 * (lane & 63));
        for (int i = 0; i < 32; i++) printf(" r%d=(%d,%d)", i, (int)(short)(h[lane * 32 + i] & 0xFFFF), (int)(short)(h[lane * 32 + i] >> 16));
        printf("\n");
    }
    return 0;
}
