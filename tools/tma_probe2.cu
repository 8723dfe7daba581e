// TMA bring-up variants (see tma_probe.cu).  usage: tma_probe2 rank box_w box_h promo
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap M, int rank, int x, int y, int z, int bytes, uint8_t *out, int *status) {
    extern __shared__ __align__(128) uint8_t tile[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        if (rank == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(tile)), "l"(&M), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(tile)), "l"(&M), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
    int ok = 0;
    for (int spin = 0; spin < 1000000 && !ok; spin++) {
        uint32_t done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        ok = done;
    }
    if (threadIdx.x == 0) status[0] = ok;
    if (ok) for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char **argv) {
    const int rank = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), promo = atoi(argv[4]);
    const int w = 1241, h = 376, n = 4, pitch = 1248;
    std::vector<uint8_t> img((size_t)pitch * h * n);
    for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *dout; int *dst;
    cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 256 * 256); cudaMalloc(&dst, 8); cudaMemset(dst, 0xff, 8);
    alignas(64) CUtensorMap M;
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)(rank == 2 ? h * n : h), (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUresult r = ((Fn)fp)(&M, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank %d box %dx%d promo %d: encode=%d q=%d\n", rank, bw, bh, promo, (int)r, (int)q);
    const int x = 100, y = 50, z = rank == 2 ? 0 : 2;
    probe<<<1, 128, bw * bh + 1024>>>(M, rank, x, y, z, bw * bh, dout, dst);
    cudaError_t e = cudaDeviceSynchronize();
    int st[2] = {-9, -9}; cudaMemcpy(st, dst, 8, cudaMemcpyDeviceToHost);
    std::vector<uint8_t> out(bw * bh); cudaMemcpy(out.data(), dout, out.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < bh; rr++) for (int k = 0; k < bw; k++) {
        const int X = x + k, Y = y + rr;
        const uint8_t want = (X < w) ? img[((size_t)z * h + Y) * pitch + X] : 0;
        bad += out[rr * bw + k] != want;
    }
    printf("  err=%s done=%d mismatches=%d\n", cudaGetErrorString(e), st[0], bad);
    return 0;
}
