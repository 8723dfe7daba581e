// Shared host-side plumbing for the sfe C ABI (error reporting, CUDA checks, handles).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sfe.h"

namespace sfe {

void set_error(const char *fmt, ...);

#define SFE_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            ::sfe::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? SFE_ERR_NO_DEVICE \
                                                                                  : SFE_ERR_CUDA; \
        }                                                                                     \
    } while (0)

#define SFE_REQUIRE(cond, status, msg)                           \
    do {                                                         \
        if (!(cond)) {                                           \
            ::sfe::set_error("%s: %s", __func__, msg);           \
            return status;                                       \
        }                                                        \
    } while (0)

// Is the primary context of `device` already alive in this process?  (driver API through the runtime's entry points)
bool primary_context_active(int device);

// RAII device-scope guard: every ABI call runs on its handle's device and gives the caller's device back -- but only
// when that device really is in use.  A fresh host thread reports device 0 as current without having touched it, and
// since CUDA 12 cudaSetDevice() creates the primary context eagerly: restoring "device 0" blindly made every rank of a
// multi-GPU job build a context on GPU 0 the first time a worker thread called in (measured: 0.5 s stall, ~0.5 GB).
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; return; }
        ok = (prev == dev) || cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev && primary_context_active(prev)) cudaSetDevice(prev);
    }
};

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

static inline int div_up(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Programmatic dependent launch: a kernel launched with `pdl` may be scheduled while its predecessor on the stream is still
// running (its CTAs then sit in pdl_enter() until the predecessor has completed and flushed), which hides the launch latency
// between the small dependent kernels of a one-image / one-pair call.  Every kernel launched this way calls pdl_enter() first.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

}  // namespace sfe

// First statement of every kernel that may be launched with programmatic stream serialization: let the successor be scheduled
// as soon as all CTAs of this grid are running, then wait for the predecessor grid's completion (no-ops otherwise).
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

struct sfe_event {
    int device;
    cudaEvent_t ev;
};

// Device-side 256-bit Hamming distance of two descriptors held as 8 words.
__device__ __forceinline__ int hamming8(const uint32_t a[8], const uint32_t b[8]) {
    int d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d += __popc(a[i] ^ b[i]);
    return d;
}

// The same distance with half the POPCs: the popc pipe issues 16 lanes/clk/SM against 64 for LOP3 (tools/microbench.cu),
// so the eight XOR words first go through a carry-save adder tree (sum = a ^ b ^ c, carry = majority: one LOP3 each):
//   d = popc(ones) + popc(x7) + 2 popc(twos) + 4 popc(fours)
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t &sum, uint32_t &carry) {
    sum = a ^ b ^ c;
    carry = (a & b) | (c & (a ^ b));
}
__device__ __forceinline__ int hamming8_csa(const uint32_t a[8], uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, uint32_t b4,
                                            uint32_t b5, uint32_t b6, uint32_t b7) {
    uint32_t s1, c1, s2, c2, ones, c3, twos, fours;
    csa(a[0] ^ b0, a[1] ^ b1, a[2] ^ b2, s1, c1);
    csa(a[3] ^ b3, a[4] ^ b4, a[5] ^ b5, s2, c2);
    csa(s1, s2, a[6] ^ b6, ones, c3);
    csa(c1, c2, c3, twos, fours);
    return __popc(ones) + __popc(a[7] ^ b7) + 2 * __popc(twos) + 4 * __popc(fours);
}
