// Integer-pipe microbenchmarks for B200 (popc / lop3 / iadd / imnmx / shared-memory byte loads):
// the peaks that bound FAST and brute-force Hamming are not in MEASURED_PEAKS.json.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
    uint32_t acc = 0;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) a[i] = __popc(a[i]) + seed;              // POPC (+IADD)
            if (OP == 1) a[i] = (a[i] ^ acc) & (a[(i + 1) & 7] | seed); // LOP3
            if (OP == 2) a[i] = a[i] + a[(i + 1) & 7] + seed;      // IADD3
            if (OP == 3) a[i] = min(a[i] ^ seed, a[(i + 1) & 7]);  // LOP + IMNMX
            if (OP == 4) a[i] = a[i] * seed + a[(i + 1) & 7];      // IMAD
            if (OP == 5) a[i] = __popc(a[i] ^ a[(i + 1) & 7]) + (a[i] << 1); // xor+popc+shift-add: hamming-like mix
            if (OP == 6) a[i] = __vimin3_s16x2(a[i], a[(i + 1) & 7], seed);   // VIMNMX3.S16x2
            if (OP == 7) a[i] = __vimax3_s32(a[i], a[(i + 1) & 7], seed);     // VIMNMX3
            if (OP == 8) a[i] = __byte_perm(a[i], a[(i + 1) & 7], seed);      // PRMT
            if (OP == 10) a[i] = __vabsdiffu4(a[i], a[(i + 1) & 7]) + 0u;       // VABSDIFF4.U8
            if (OP == 11) a[i] = __vsadu4(a[i], a[(i + 1) & 7]) + seed;        // VABSDIFF4.U8.ACC (sum of abs diffs)
            if (OP == 12) {                                                   // VABSDIFF4 + LOP3 interleaved
                if (i & 1) a[i] = __vabsdiffu4(a[i], a[(i + 1) & 7]);
                else a[i] = (a[i] ^ seed) & a[(i + 1) & 7];
            }
            if (OP == 13) {                                                   // Harley-Seal mix: 2 lop3 + 1 popc per slot
                const uint32_t x = a[i] ^ seed, y = a[(i + 1) & 7];
                a[i] = __popc(x ^ y ^ acc) + ((x & y) | (acc & (x ^ y)));
            }
            if (OP == 14) { int d; asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a[i]), "r"(a[(i + 1) & 7]), "r"(seed)); a[i] = d; }  // IDP.4A
            if (OP == 15) { int d; asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a[i]), "r"(a[(i + 1) & 7]), "r"(seed)); a[i] = d; }  // IDP.2A
            if (OP == 16) {                                                   // IDP.4A + LOP3 interleaved
                if (i & 1) { int d; asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a[i]), "r"(a[(i + 1) & 7]), "r"(seed)); a[i] = d; }
                else a[i] = (a[i] ^ seed) & a[(i + 1) & 7];
            }
            if (OP == 9) {                                                    // IMAD (fma pipe) + LOP3 (alu pipe) interleaved
                if (i & 1) a[i] = a[i] * seed + a[(i + 1) & 7];
                else a[i] = (a[i] ^ seed) & a[(i + 1) & 7];
            }
        }
        acc += a[0];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) acc += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(256) lds_bytes(uint32_t *out, int stride) {
    __shared__ uint8_t tile[72 * 72];
    for (int i = threadIdx.x; i < 72 * 72; i += 256) tile[i] = (uint8_t)(i * 7);
    __syncthreads();
    uint32_t acc = 0;
    int p = 3 * 72 + 3 + (threadIdx.x & 31) + (threadIdx.x >> 5) * 72;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc += tile[p + ((i * stride) & 63)];
        p = (p + acc) % (60 * 72) + 3 * 72;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
void run(const char *name, int ops_per_iter, uint32_t *d) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    dim3 grid(sms * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<grid, 256>>>(d, 3); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<OP><<<grid, 256>>>(d, 3 + r);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = 5.0 * grid.x * 256.0 * ITERS * 8 * ops_per_iter;
    printf("%-28s %8.2f Tlane-op/s  (%.1f lane-op/clk/SM at max clock %d MHz; %d SMs)\n", name, lane_ops / ms / 1e9,
           lane_ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000, sms);
}

int main() {
    uint32_t *d; cudaMalloc(&d, 148 * 8 * 256 * 4 * 4);
    run<0>("popc(+iadd)", 1, d);
    run<1>("lop3 x2", 2, d);
    run<2>("iadd3", 1, d);
    run<3>("lop+imnmx", 2, d);
    run<4>("imad", 1, d);
    run<5>("xor+popc+shl-add", 3, d);
    run<6>("vimnmx3.s16x2", 1, d);
    run<7>("vimnmx3.s32", 1, d);
    run<8>("prmt", 1, d);
    run<9>("imad|lop3 interleaved", 1, d);
    run<10>("vabsdiff4.u8", 1, d);
    run<11>("vabsdiff4.u8.acc", 1, d);
    run<12>("vabsdiff4|lop3 interleaved", 1, d);
    run<13>("xor,csa(2 lop3),popc,iadd", 5, d);
    run<14>("idp.4a", 1, d);
    run<15>("idp.2a", 1, d);
    run<16>("idp.4a|lop3 interleaved", 1, d);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    lds_bytes<<<sms * 8, 256>>>(d, 5); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) lds_bytes<<<sms * 8, 256>>>(d, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lds = 5.0 * sms * 8 * 256.0 * ITERS * 16;
    printf("%-28s %8.2f Tlane-LDS.U8/s (%.1f lanes/clk/SM)\n", "lds.u8 (fast-like)", lds / ms / 1e9, lds / (ms * 1e-3) / sms / (clk * 1e3));
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
