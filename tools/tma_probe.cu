// Stand-alone check of the TMA plumbing in slam-toolkit_b200/csrc/sfe_tma.cuh: one box load of a pitched u8
// image stack, including out-of-image coordinates.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/bin/tma_probe tools/tma_probe.cu \
//        slam-toolkit_b200/csrc/sfe_context.cu -Iinclude
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../slam-toolkit_b200/csrc/sfe_common.cuh"
#include "../slam-toolkit_b200/csrc/sfe_tma.cuh"

using namespace sfe;

struct Maps { CUtensorMap m[2]; };

__global__ void probe(const __grid_constant__ Maps M, const CUtensorMap *gmaps, int mode, int which, int x, int y, int z, int box_w,
                      int box_h, uint8_t *out, int *status) {
    extern __shared__ __align__(128) uint8_t tile[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        status[1] = (int)(smem_u32(tile) & 127);
        mbar_init(&bar, 1);
        if (mode == 1) { status[0] = -2; return; }                       // only barrier init
        mbar_expect_tx(&bar, box_w * box_h);
        if (mode == 2) { status[0] = -3; return; }                       // + expect_tx
        const CUtensorMap *mp = mode == 3 ? &gmaps[which] : &M.m[which];  // mode 3: descriptor in global memory
        tma_load_3d(tile, mp, &bar, x, y, z);
    }
    __syncthreads();
    const uint32_t addr = smem_u32(&bar);
    int ok = 0;
    for (int spin = 0; spin < 1000000 && !ok; spin++) {
        uint32_t done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(addr), "r"(0) : "memory");
        ok = done;
    }
    if (threadIdx.x == 0) status[0] = ok;
    if (ok) for (int i = threadIdx.x; i < box_w * box_h; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char **argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int w = 1241, h = 376, n = 4, pitch = 1248, bw = 48, bh = 46;
    std::vector<uint8_t> img((size_t)pitch * h * n);
    for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *dout; int *dst;
    cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 256 * 256); cudaMalloc(&dst, 8); cudaMemset(dst, 0xff, 8);
    Maps M;
    bool e0 = tma_encode_u8_3d(&M.m[0], d, w, h, n, pitch, (size_t)pitch * h, bw, bh);
    bool e1 = tma_encode_u8_3d(&M.m[1], d, w, h, n, pitch, (size_t)pitch * h, 144, 38);
    printf("encode: %d %d\n", e0, e1);
    CUtensorMap *gm; cudaMalloc(&gm, sizeof(M)); cudaMemcpy(gm, &M, sizeof(M), cudaMemcpyHostToDevice);
    int fails = 0;
    const int cases[][6] = {{0, 100, 50, 2, bw, bh}, {0, 1215, 350, 3, bw, bh}, {1, -4, -3, 0, 144, 38}, {1, 1148, 349, 1, 144, 38}};
    for (auto &c : cases) {
        cudaMemset(dst, 0xff, 8);
        probe<<<1, 128, c[4] * c[5]>>>(M, gm, mode, c[0], c[1], c[2], c[3], c[4], c[5], dout, dst);
        cudaError_t e = cudaDeviceSynchronize();
        int st[2]; cudaMemcpy(st, dst, 8, cudaMemcpyDeviceToHost);
        std::vector<uint8_t> out(c[4] * c[5]); cudaMemcpy(out.data(), dout, out.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < c[5]; r++) for (int k = 0; k < c[4]; k++) {
            const int X = c[1] + k, Y = c[2] + r;
            const uint8_t want = (X >= 0 && X < w && Y >= 0 && Y < h) ? img[((size_t)c[3] * h + Y) * pitch + X] : 0;
            bad += out[r * c[4] + k] != want;
        }
        printf("case map%d (%d,%d,%d) box %dx%d: err=%s done=%d smem_align=%d mismatches=%d\n", c[0], c[1], c[2], c[3], c[4], c[5],
               cudaGetErrorString(e), st[0], st[1], bad);
        fails += (e != cudaSuccess) || st[0] != 1 || bad;
    }
    printf(fails ? "TMA PROBE FAILED\n" : "TMA PROBE OK\n");
    return fails != 0;
}
