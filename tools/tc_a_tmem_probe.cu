// Probe: tcgen05.mma kind::i8 with the A operand in tensor memory.  Which TMEM layout does the instruction expect for A?
// Hypothesis: row m on lane m, K packed four int8 to a 32-bit column (column c = k 4c .. 4c + 3).  One CTA, 128 threads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/tc_a_tmem_probe tools/tc_a_tmem_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr uint32_t kLBO = 128, kSBO = 2048;
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | (uint64_t)(kLBO >> 4) << 16 | (uint64_t)(kSBO >> 4) << 32 | 1ull << 46;
}
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__global__ void __launch_bounds__(128) probe(const int8_t *A, const int8_t *B, int *D, int ksteps) {
    // A: [128][32 * ksteps] row-major int8; B: [128 rows n][32 * ksteps] (K-major: D[m][n] = sum_k A[m][k] B[n][k])
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint32_t tmem_base;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = 32 * ksteps;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // B tile into shared memory: core matrices of 8 rows x 16 bytes, (row >> 3) * SBO + chunk * LBO + (row & 7) * 16; one tile per
    // 32-wide k step is laid out as the kernel does it: chunk index runs over the whole K (16 chunks for K = 256)
    for (int i = tid; i < 128 * (K / 16); i += 128) {
        const int row = i / (K / 16), chunk = i % (K / 16);
        *(uint4 *)(sm + (row >> 3) * kSBO + chunk * kLBO + (row & 7) * 16) = *(const uint4 *)(B + (size_t)row * K + 16 * chunk);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    // A into TMEM columns 256 ..: lane = row (warp w owns lanes 32 w .. 32 w + 31), column c = k 4c .. 4c + 3
    const uint32_t a_tmem = tmem + 256;
    for (int c0 = 0; c0 < K / 4; c0 += 8) {
        uint32_t r[8];
        for (int j = 0; j < 8; j++) r[j] = *(const uint32_t *)(A + (size_t)tid * K + 4 * (c0 + j));
        const uint32_t addr = a_tmem + ((uint32_t)(32 * warp) << 16) + c0;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                     "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                     : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        for (int ks = 0; ks < ksteps; ks++) {
            const uint64_t bd = smem_desc(smem_u32(sm) + 2 * ks * kLBO);
            const uint32_t at = a_tmem + 8 * ks;
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "setp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
                "}\n" ::"r"(tmem),
                "r"(at), "l"(bd), "r"(kIdesc), "r"((uint32_t)(ks > 0)), "r"(0u)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {   // wait for the MMAs
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n"
                : "=r"(done)
                : "r"(smem_u32(&bar)), "r"(0u)
                : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 8) {
        uint32_t v[8];
        const uint32_t addr = tmem + ((uint32_t)(32 * warp) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(addr)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; j++) D[tid * 128 + c0 + j] = (int)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    for (int ksteps : {1, 8}) {
        const int K = 32 * ksteps;
        int8_t *hA = (int8_t *)malloc(128 * K), *hB = (int8_t *)malloc(128 * K);
        srand(7);
        for (int i = 0; i < 128 * K; i++) { hA[i] = (int8_t)(rand() % 5 - 2); hB[i] = (int8_t)(rand() % 3 - 1); }
        int8_t *dA, *dB; int *dD;
        cudaMalloc(&dA, 128 * K); cudaMalloc(&dB, 128 * K); cudaMalloc(&dD, 128 * 128 * 4);
        cudaMemcpy(dA, hA, 128 * K, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 128 * K, cudaMemcpyHostToDevice);
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 256 + 1024);
        probe<<<1, 128, 128 * 256 + 1024>>>(dA, dB, dD, ksteps);
        cudaError_t e = cudaDeviceSynchronize();
        int *hD = (int *)malloc(128 * 128 * 4);
        cudaMemcpy(hD, dD, 128 * 128 * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < 128; n++) {
                int s = 0;
                for (int k = 0; k < K; k++) s += hA[m * K + k] * hB[n * K + k];
                if (s != hD[m * 128 + n] && bad++ < 5) printf("  D[%d][%d] = %d, expected %d\n", m, n, hD[m * 128 + n], s);
            }
        printf("ksteps %d: %s, %d of 16384 wrong\n", ksteps, cudaGetErrorString(e), bad);
    }
    return 0;
}
