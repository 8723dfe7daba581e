// Probe: how fast can the warps of an SM read tensor memory?  W warps each issue tcgen05.ld (32 lanes of their quadrant) back to
// back -- .32x32b.x32 (32 columns -> 32 registers) or .32x32b.x64.pack::16b (64 columns, the low halves of two adjacent columns
// packed into one register -> 32 registers) -- optionally while one thread keeps the tensor pipe busy with int8 MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/tc_tmem_ld_rate tools/tc_tmem_ld_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr uint32_t kLBO = 128, kSBO = 2048;
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | (uint64_t)(kLBO >> 4) << 16 | (uint64_t)(kSBO >> 4) << 32 | 1ull << 46;
}
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

#define OUT32(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),   \
                 "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),   \
                 "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define REGS32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"

template <int kPack>
__global__ void __launch_bounds__(1024) rate(long long *out, int iters, int with_mma) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint32_t tmem_base;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 65536 / 16; i += blockDim.x) ((uint4 *)sm)[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const long long t0 = clock64();
    if (warp == nwarps - 1) {  // the last warp: MMA issuer (columns 256 .. 511) or idle
        if (with_mma && (tid & 31) == 0) {
            const uint32_t a_addr = smem_u32(sm), b_addr = smem_u32(sm) + 32768;
            for (int it = 0; it < iters / 2; it++) {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    asm volatile(
                        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem + 256u + 128u * (it & 1)),
                        "l"(smem_desc(a_addr + k * 2 * kLBO)), "l"(smem_desc(b_addr + k * 2 * kLBO)), "r"(kIdesc), "r"((uint32_t)(k > 0)), "r"(0u)
                        : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(done)
                             : "r"(smem_u32(&bar)), "r"(0u)
                             : "memory");
            out[1] = clock64() - t0;
        }
    } else {
        // 4 warps per 64-column block (one per lane quadrant); blocks of 64 columns inside columns 0 .. 255
        const uint32_t addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(64 * ((warp >> 2) & 3));
        uint32_t acc = 0;
        for (int it = 0; it < iters; it++) {
            uint32_t v[32];
            if (kPack)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " REGS32 ", [%32];" : OUT32(v) : "r"(addr) : "memory");
            else
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " REGS32 ", [%32];" : OUT32(v) : "r"(addr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] ^ v[31];
        }
        if (acc == 0x12345678u) out[3] = acc;
        if (tid == 0) out[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long *d, h[4];
    cudaMalloc(&d, 32);
    cudaFuncSetAttribute(rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
    cudaFuncSetAttribute(rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
    const int iters = 2000;
    for (int pack = 0; pack < 2; pack++)
        for (int with_mma = 0; with_mma < 2; with_mma++)
            for (int w : {4, 8, 16}) {
                cudaMemset(d, 0, 32);
                if (pack) rate<1><<<1, (w + 1) * 32, 66560>>>(d, iters, with_mma);
                else rate<0><<<1, (w + 1) * 32, 66560>>>(d, iters, with_mma);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
                const int cols = pack ? 64 : 32;
                printf("%-14s %2d warps, MMA %s: %s  %.0f clk per load per warp, %.0f accumulators/clk/SM (%.0f B/clk of TMEM cells)",
                       pack ? "x64.pack::16b" : "x32", w, with_mma ? "on " : "off", cudaGetErrorString(e), (double)h[0] / iters,
                       32.0 * cols * iters * w / (double)h[0], 128.0 * cols * iters * w / (double)h[0]);
                if (with_mma) printf("   | %.1f clk per MMA", (double)h[1] / (iters / 2 * 8.0));
                printf("\n");
            }
    return 0;
}
