// General entry points of the sfe C ABI: status strings, pinned/device memory, events.
#include <mutex>

#include <algorithm>

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

// Copy-only probe of the host <-> device path: what a pipelined host call moves per step, without kernels.
int sfe_copy_probe(int device, size_t h2d_bytes, size_t d2h_bytes, int chunks, double seconds, int flags, double *h2d_gbs,
                   double *d2h_gbs) {
    SFE_REQUIRE(h2d_gbs && d2h_gbs && chunks >= 1 && seconds > 0 && (h2d_bytes > 0 || d2h_bytes > 0), SFE_ERR_BAD_ARG, "bad argument");
    DeviceGuard g(device);
    SFE_REQUIRE(g.ok, SFE_ERR_NO_DEVICE, "cannot select device");
    void *hi = nullptr, *ho = nullptr, *di = nullptr, *dout = nullptr;
    cudaStream_t si = nullptr, so = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr, e3 = nullptr;
    int rc = SFE_OK;
    auto ok = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == SFE_OK) { set_error("sfe_copy_probe: %s: %s", what, cudaGetErrorString(e)); rc = SFE_ERR_CUDA; }
        return e == cudaSuccess;
    };
    const unsigned hflags = cudaHostAllocPortable | ((flags & SFE_HOST_WRITE_COMBINED) ? cudaHostAllocWriteCombined : 0);
    if (h2d_bytes) { ok(cudaHostAlloc(&hi, h2d_bytes, hflags), "host alloc"); ok(cudaMalloc(&di, h2d_bytes), "device alloc"); }
    if (d2h_bytes) { ok(cudaHostAlloc(&ho, d2h_bytes, cudaHostAllocPortable), "host alloc"); ok(cudaMalloc(&dout, d2h_bytes), "device alloc"); }
    ok(cudaStreamCreateWithFlags(&si, cudaStreamNonBlocking), "stream");
    ok(cudaStreamCreateWithFlags(&so, cudaStreamNonBlocking), "stream");
    ok(cudaEventCreate(&e0), "event"); ok(cudaEventCreate(&e1), "event"); ok(cudaEventCreate(&e2), "event"); ok(cudaEventCreate(&e3), "event");
    if (rc == SFE_OK) {
        if (hi) memset(hi, 1, h2d_bytes);
        if (dout) cudaMemset(dout, 2, d2h_bytes);
        auto step = [&]() {
            for (int c = 0; c < chunks; c++) {
                const size_t a = h2d_bytes * c / chunks, b = h2d_bytes * (c + 1) / chunks;
                if (b > a) cudaMemcpyAsync((char *)di + a, (char *)hi + a, b - a, cudaMemcpyHostToDevice, si);
                const size_t p = d2h_bytes * c / chunks, q = d2h_bytes * (c + 1) / chunks;
                if (q > p) cudaMemcpyAsync((char *)ho + p, (char *)dout + p, q - p, cudaMemcpyDeviceToHost, so);
            }
        };
        step();
        cudaStreamSynchronize(si); cudaStreamSynchronize(so);
        // calibrate one step, then time a fixed number of steps with events on each stream
        cudaEventRecord(e0, si); cudaEventRecord(e2, so);
        step();
        cudaEventRecord(e1, si); cudaEventRecord(e3, so);
        cudaStreamSynchronize(si); cudaStreamSynchronize(so);
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e2, e3);
        const double one = std::max((double)std::max(a, b), 0.05);
        const int steps = (int)std::min(std::max(seconds * 1e3 / one, 2.0), 100000.0);
        cudaEventRecord(e0, si); cudaEventRecord(e2, so);
        for (int i = 0; i < steps; i++) step();
        cudaEventRecord(e1, si); cudaEventRecord(e3, so);
        ok(cudaStreamSynchronize(si), "sync"); ok(cudaStreamSynchronize(so), "sync");
        cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e2, e3);
        *h2d_gbs = h2d_bytes && a > 0 ? (double)h2d_bytes * steps / (a * 1e-3) / 1e9 : 0.0;
        *d2h_gbs = d2h_bytes && b > 0 ? (double)d2h_bytes * steps / (b * 1e-3) / 1e9 : 0.0;
    }
    if (hi) cudaFreeHost(hi);
    if (ho) cudaFreeHost(ho);
    if (di) cudaFree(di);
    if (dout) cudaFree(dout);
    if (si) cudaStreamDestroy(si);
    if (so) cudaStreamDestroy(so);
    for (cudaEvent_t e : {e0, e1, e2, e3}) if (e) cudaEventDestroy(e);
    return rc;
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
