// common.cuh — error handling, the context, stream-ordered device buffers.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/magnetite_b200.h"

namespace mag {

struct Failure {          // thrown inside the library, caught at every extern "C" entry
    int code;
    std::string msg;
};

[[noreturn]] inline void fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    throw Failure{code, buf};
}

#define MAG_CUDA(expr)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess)                                                             \
            ::mag::fail(e_ == cudaErrorMemoryAllocation ? MAG_ERR_OOM : MAG_ERR_CUDA,      \
                        "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__,      \
                        __LINE__, cudaGetErrorString(e_));                                 \
    } while (0)

#define MAG_KERNEL_CHECK() MAG_CUDA(cudaGetLastError())

// Launch on the context's current stream and count it (mag_stats.kernel_launches).
#define MAG_LAUNCH(ctx, kern, grid, block, smem, ...)                                      \
    do {                                                                                   \
        kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                     \
        MAG_KERNEL_CHECK();                                                                \
        (ctx)->launches++;                                                                 \
    } while (0)

constexpr int kWarp = 32;

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

struct Comm;   // dist.cuh

}  // namespace mag

// The opaque context of the C ABI.
struct mag_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;      // stream of the current call (own or caller's)
    cudaMemPool_t pool = nullptr;
    uint64_t launches = 0;              // kernels launched by the current call
    mag::Comm *comm = nullptr;
    // pinned host scratch for scalar read-backs
    double *h_scal = nullptr;
};

namespace mag {

// Stream-ordered device buffer (cudaMallocAsync on the context's pool: repeated
// solves reuse the same physical memory without touching the OS allocator).
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    mag_ctx *ctx = nullptr;
    DevBuf() = default;
    DevBuf(mag_ctx *c, size_t count) { alloc(c, count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), ctx(o.ctx) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; ctx = o.ctx; o.p = nullptr; o.n = 0; }
        return *this;
    }
    void alloc(mag_ctx *c, size_t count) {
        release();
        ctx = c; n = count;
        size_t bytes = (count ? count : 1) * sizeof(T);
        MAG_CUDA(cudaMallocAsync((void **)&p, bytes, c->stream));
    }
    void zero() { if (p) MAG_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), ctx->stream)); }
    void release() {
        if (p) { cudaFreeAsync(p, ctx->stream); p = nullptr; n = 0; }
    }
    ~DevBuf() { release(); }
    size_t bytes() const { return n * sizeof(T); }
};

template <class T>
inline void copy_to_device(mag_ctx *ctx, T *dst, const T *src, size_t n, bool src_on_device) {
    if (!n) return;
    MAG_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T),
                             src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                             ctx->stream));
}
template <class T>
inline void copy_from_device(mag_ctx *ctx, T *dst, const T *src, size_t n, bool dst_on_device) {
    if (!n) return;
    MAG_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T),
                             dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                             ctx->stream));
}

struct EventTimer {       // CUDA-event phase timer on the run stream
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s;
    explicit EventTimer(cudaStream_t st) : s(st) {
        MAG_CUDA(cudaEventCreate(&a));
        MAG_CUDA(cudaEventCreate(&b));
    }
    ~EventTimer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    void start() { MAG_CUDA(cudaEventRecord(a, s)); }
    float stop() {        // synchronises on the stop event
        float ms = 0.f;
        MAG_CUDA(cudaEventRecord(b, s));
        MAG_CUDA(cudaEventSynchronize(b));
        MAG_CUDA(cudaEventElapsedTime(&ms, a, b));
        return ms;
    }
};

}  // namespace mag
