// General entry points of the sfe C ABI: status strings, pinned/device memory, events.
#include <mutex>

#include "sfe_common.cuh"
#include "sfe_tma.cuh"

namespace sfe {
// cuTensorMapEncodeTiled through the runtime's driver entry point: libsfe.so does not link libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

bool primary_context_active(int device) {
    typedef CUresult (*GetDeviceFn)(CUdevice *, int);
    typedef CUresult (*CtxStateFn)(CUdevice, unsigned int *, int *);
    static GetDeviceFn get_device = nullptr;
    static CtxStateFn ctx_state = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuDeviceGet", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            get_device = (GetDeviceFn)p;
        if (cudaGetDriverEntryPoint("cuDevicePrimaryCtxGetState", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            ctx_state = (CtxStateFn)p;
    });
    if (!get_device || !ctx_state) return true;  // cannot tell: behave like a plain guard
    CUdevice d;
    unsigned int flags = 0;
    int active = 0;
    if (get_device(&d, device) != CUDA_SUCCESS || ctx_state(d, &flags, &active) != CUDA_SUCCESS) return true;
    return active != 0;
}

bool tma_encode_u8_3d(CUtensorMap *map, const void *base, uint64_t width, uint64_t height, uint64_t images, uint64_t pitch,
                      uint64_t image_stride, uint32_t box_w, uint32_t box_h) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || !tma_layout_ok(base, pitch, image_stride) || box_w % 16 || box_w > 256 || box_h > 256 || box_h == 0) return false;
    if (images > 1 && image_stride < pitch * height) return false;
    const cuuint64_t dims[3] = {width, height, images};
    const cuuint64_t strides[2] = {pitch, images > 1 ? image_stride : pitch * height};
    const cuuint32_t box[3] = {box_w, box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    alignas(64) CUtensorMap tmp;
    const CUresult r = fn(&tmp, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    *map = tmp;
    return true;
}

static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace sfe

using namespace sfe;

extern "C" {

int sfe_abi_version(void) { return SFE_ABI_VERSION; }

const char *sfe_status_string(int status) {
    switch (status) {
        case SFE_OK: return "ok";
        case SFE_ERR_BAD_ARG: return "bad argument";
        case SFE_ERR_CAPACITY: return "capacity exceeded";
        case SFE_ERR_CUDA: return "CUDA error";
        case SFE_ERR_NO_DEVICE: return "no CUDA device (no CPU fallback exists)";
        case SFE_ERR_UNSUPPORTED: return "unsupported parameters";
        default: return "unknown status";
    }
}

const char *sfe_last_error(void) { return g_err; }

int sfe_device_count(int *count) {
    SFE_REQUIRE(count, SFE_ERR_BAD_ARG, "null argument");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return SFE_ERR_NO_DEVICE;
    }
    return SFE_OK;
}

int sfe_host_alloc(void **ptr, size_t bytes) {
    SFE_REQUIRE(ptr && bytes > 0, SFE_ERR_BAD_ARG, "bad argument");
    SFE_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable));
    return SFE_OK;
}

int sfe_host_alloc_ex(void **ptr, size_t bytes, int flags) {
    SFE_REQUIRE(ptr && bytes > 0, SFE_ERR_BAD_ARG, "bad argument");
    SFE_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | ((flags & SFE_HOST_WRITE_COMBINED) ? cudaHostAllocWriteCombined : 0)));
    return SFE_OK;
}

int sfe_host_free(void *ptr) {
    if (ptr) SFE_CUDA(cudaFreeHost(ptr));
    return SFE_OK;
}

int sfe_device_alloc(int device, void **ptr, size_t bytes) {
    SFE_REQUIRE(ptr && bytes > 0, SFE_ERR_BAD_ARG, "bad argument");
    DeviceGuard g(device);
    SFE_REQUIRE(g.ok, SFE_ERR_NO_DEVICE, "cannot select device");
    SFE_CUDA(cudaMalloc(ptr, bytes));
    return SFE_OK;
}

int sfe_device_free(int device, void *ptr) {
    DeviceGuard g(device);
    SFE_REQUIRE(g.ok, SFE_ERR_NO_DEVICE, "cannot select the handle's device");
    if (ptr) SFE_CUDA(cudaFree(ptr));
    return SFE_OK;
}

int sfe_copy_to_device(int device, void *dst_dev, const void *src_host, size_t bytes) {
    SFE_REQUIRE(dst_dev && src_host, SFE_ERR_BAD_ARG, "null argument");
    DeviceGuard g(device);
    SFE_REQUIRE(g.ok, SFE_ERR_NO_DEVICE, "cannot select the handle's device");
    SFE_CUDA(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
    return SFE_OK;
}

int sfe_copy_to_host(int device, void *dst_host, const void *src_dev, size_t bytes) {
    SFE_REQUIRE(dst_host && src_dev, SFE_ERR_BAD_ARG, "null argument");
    DeviceGuard g(device);
    SFE_REQUIRE(g.ok, SFE_ERR_NO_DEVICE, "cannot select the handle's device");
    SFE_CUDA(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
    return SFE_OK;
}

int sfe_event_create(int device, sfe_event **ev) {
    SFE_REQUIRE(ev, SFE_ERR_BAD_ARG, "null argument");
    DeviceGuard g(device);
    SFE_REQUIRE(g.ok, SFE_ERR_NO_DEVICE, "cannot select device");
    sfe_event *e = new sfe_event{device, nullptr};
    cudaError_t err = cudaEventCreate(&e->ev);
    if (err != cudaSuccess) {
        set_error("cudaEventCreate: %s", cudaGetErrorString(err));
        delete e;
        return SFE_ERR_CUDA;
    }
    *ev = e;
    return SFE_OK;
}

int sfe_event_destroy(sfe_event *ev) {
    if (!ev) return SFE_OK;
    DeviceGuard g(ev->device);
    SFE_REQUIRE(g.ok, SFE_ERR_NO_DEVICE, "cannot select the handle's device");
    cudaEventDestroy(ev->ev);
    delete ev;
    return SFE_OK;
}

int sfe_event_elapsed_ms(sfe_event *start, sfe_event *stop, float *ms) {
    SFE_REQUIRE(start && stop && ms, SFE_ERR_BAD_ARG, "null argument");
    DeviceGuard g(stop->device);
    SFE_REQUIRE(g.ok, SFE_ERR_NO_DEVICE, "cannot select the handle's device");
    SFE_CUDA(cudaEventSynchronize(stop->ev));
    SFE_CUDA(cudaEventElapsedTime(ms, start->ev, stop->ev));
    return SFE_OK;
}

int sfe_hamming256(const void *a, const void *b) {  // DescriptorDistance, include/orb_extractor.h:87-103
    uint64_t x[4], y[4];
    memcpy(x, a, 32);
    memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) + __builtin_popcountll(x[2] ^ y[2]) +
           __builtin_popcountll(x[3] ^ y[3]);
}

}  // extern "C"
