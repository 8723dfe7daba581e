// Probe: sustained rate of tcgen05.mma.kind::i8 issued back to back by one thread (M = 128, K = 32 per instruction), for
// N = 128 and N = 256, with the A operand in shared memory or in tensor memory.  Prints clocks per instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/tc_mma_rate tools/tc_mma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr uint32_t kLBO = 128, kSBO = 2048;
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | (uint64_t)(kLBO >> 4) << 16 | (uint64_t)(kSBO >> 4) << 32 | 1ull << 46;
}
template <int N>
__host__ __device__ constexpr uint32_t idesc() { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }

template <int N, bool kATmem>
__global__ void __launch_bounds__(128) rate(long long *out, int iters) {
    extern __shared__ __align__(1024) uint8_t sm[];  // A tile 32 KB | B tile N x 256 B
    __shared__ uint32_t tmem_base;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < (32768 + N * 256) / 16; i += 128) ((uint4 *)sm)[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t a_addr = smem_u32(sm), b_addr = smem_u32(sm) + 32768;
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            const uint32_t d = tmem + (uint32_t)((it & 1) * (N == 128 ? 128 : 0));  // N = 128: alternate two accumulators
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint64_t bd = smem_desc(b_addr + k * 2 * kLBO);
                if (kATmem) {
                    asm volatile(
                        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(d),
                        "r"(tmem + 256u + 8u * k), "l"(bd), "r"(idesc<N>()), "r"((uint32_t)(k > 0)), "r"(0u)
                        : "memory");
                } else {
                    asm volatile(
                        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(d),
                        "l"(smem_desc(a_addr + k * 2 * kLBO)), "l"(bd), "r"(idesc<N>()), "r"((uint32_t)(k > 0)), "r"(0u)
                        : "memory");
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done)
                         : "r"(smem_u32(&bar)), "r"(0u)
                         : "memory");
        out[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, bool kATmem>
void run(const char *name) {
    long long *d, h = 0;
    cudaMalloc(&d, 8);
    const int iters = 2000, smem = 32768 + N * 256 + 1024;
    cudaFuncSetAttribute(rate<N, kATmem>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; rep++) rate<N, kATmem><<<1, 128, smem>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / (iters * 8.0);
    printf("%-28s %s  %.1f clk per MMA  (%.0f int8 MAC/clk/SM)\n", name, cudaGetErrorString(e), per, 128.0 * N * 32 / per);
    cudaFree(d);
}

int main() {
    run<128, false>("N=128, A in shared memory");
    run<256, false>("N=256, A in shared memory");
    run<128, true>("N=128, A in tensor memory");
    run<256, true>("N=256, A in tensor memory");
    return 0;
}
