// The CUDA programming guide's TMA example (libcu++ wrappers), u8 or i32, plus a 1-D bulk copy.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cuda/ptx>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

template <typename T, int BW, int BH>
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, T *out, int mode, const T *src1d) {
    __shared__ alignas(128) T smem_buffer[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        if (mode == 0) cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        else cde::cp_async_bulk_global_to_shared(&smem_buffer, src1d, sizeof(smem_buffer), bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = smem_buffer[i / BW][i % BW];
}

template <typename T, int BW, int BH>
void run(CUtensorMapDataType dt, const char *name, int mode, int x0 = 64, int w = 1024, int pitch_elems = 0) {
    const int h = 256;
    if (!pitch_elems) pitch_elems = w;
    std::vector<T> img((size_t)pitch_elems * h);
    for (size_t i = 0; i < img.size(); i++) img[i] = (T)((i * 2654435761u) >> 13);
    T *d, *dout;
    cudaMalloc(&d, img.size() * sizeof(T)); cudaMemcpy(d, img.data(), img.size() * sizeof(T), cudaMemcpyHostToDevice);
    cudaMalloc(&dout, BW * BH * sizeof(T));
    alignas(64) CUtensorMap M;
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h}, strides[1] = {pitch_elems * sizeof(T)};
    const cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
    CUresult r = ((Fn)fp)(&M, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    kernel<T, BW, BH><<<1, 128>>>(M, x0, 32, dout, mode, d);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<T> out(BW * BH); cudaMemcpy(out.data(), dout, out.size() * sizeof(T), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < BH; rr++) for (int k = 0; k < BW; k++) bad += out[rr * BW + k] != (mode == 0 ? ((x0 + k) >= 0 && (x0 + k) < w ? img[(size_t)(32 + rr) * pitch_elems + x0 + k] : (T)0) : img[rr * BW + k]);
    printf("%s x0=%d w=%d pitch=%d mode %d: encode=%d err=%s mismatches=%d\n", name, x0, w, pitch_elems, mode, (int)r, cudaGetErrorString(e), bad);
}

int main(int argc, char **argv) {
    const int t = atoi(argv[1]);
    if (t == 0) run<int, 64, 8>(CU_TENSOR_MAP_DATA_TYPE_INT32, "i32 64x8", 0);
    if (t == 1) run<uint8_t, 64, 8>(CU_TENSOR_MAP_DATA_TYPE_UINT8, "u8 64x8", 0);
    if (t == 2) run<int, 64, 8>(CU_TENSOR_MAP_DATA_TYPE_INT32, "i32 bulk1d", 1);
    if (t == 3) run<uint8_t, 64, 8>(CU_TENSOR_MAP_DATA_TYPE_UINT8, "u8 64x8", 0, 100);
    if (t == 4) run<uint8_t, 64, 8>(CU_TENSOR_MAP_DATA_TYPE_UINT8, "u8 64x8", 0, 64, 1241, 1248);
    if (t == 5) run<uint8_t, 64, 8>(CU_TENSOR_MAP_DATA_TYPE_UINT8, "u8 64x8", 0, 101, 1241, 1248);
    if (t == 6) run<uint8_t, 48, 46>(CU_TENSOR_MAP_DATA_TYPE_UINT8, "u8 48x46", 0, 101, 1241, 1248);
    if (t == 7) run<uint8_t, 144, 38>(CU_TENSOR_MAP_DATA_TYPE_UINT8, "u8 144x38", 0, -4, 1241, 1248);
    if (t == 8) run<uint8_t, 64, 8>(CU_TENSOR_MAP_DATA_TYPE_UINT8, "u8 64x8", 0, 1200, 1241, 1248);
    return 0;
}
