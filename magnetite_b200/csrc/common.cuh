// common.cuh — error handling, the context, stream-ordered device buffers.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
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

// The same, as a PROGRAMMATIC dependent of the launch before it on the stream (also inside a stream capture: the
// graph edge becomes a programmatic one): the grid may be scheduled while the previous kernel drains, and its
// kernels call pdl_wait() before they touch anything the previous kernel wrote.  Hides the launch gap (~1.5 us)
// between the kernels of the iteration loop; `dep` false = the plain launch.
#define MAG_LAUNCH_DEP(ctx, dep, kern, grid, block, smem, ...)                             \
    do {                                                                                   \
        cudaLaunchConfig_t cfg_ = {};                                                      \
        cfg_.gridDim = dim3(grid); cfg_.blockDim = dim3(block);                            \
        cfg_.dynamicSmemBytes = (smem); cfg_.stream = (ctx)->stream;                       \
        cudaLaunchAttribute attr_[1];                                                      \
        attr_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                  \
        attr_[0].val.programmaticStreamSerializationAllowed = 1;                           \
        cfg_.attrs = attr_; cfg_.numAttrs = (dep) ? 1 : 0;                                 \
        MAG_CUDA(cudaLaunchKernelEx(&cfg_, kern, __VA_ARGS__));                            \
        MAG_KERNEL_CHECK();                                                                \
        (ctx)->launches++;                                                                 \
    } while (0)

// First statement of a kernel launched with MAG_LAUNCH_DEP: returns once the previous kernel on the stream has
// completed and its writes are visible (a no-op for a plain launch).  An early griddepcontrol.launch_dependents in
// the kernels was measured too (2 GPUs, 4000 x 2000): no gain over the implicit trigger at exit, so it is not used.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

constexpr int kWarp = 32;

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

struct Comm;   // comm.cuh

// Device heap of a context: first-fit free list with coalescing over a few large
// cudaMalloc slabs that are kept for the life of the context.  A repeated
// solve issues the same allocation sequence and therefore lands on the same
// addresses without ever touching the driver allocator again (the CUDA
// stream-ordered pool re-created multi-GB blocks between steps, which cost
// more than the whole assembly).  Blocks are recycled at host time; that is
// safe because every call runs on one stream and every entry point is blocking.
class DeviceHeap {
public:
    void *alloc(size_t bytes);
    void release(void *p);
    void destroy();
    size_t reserved() const { return reserved_; }
private:
    struct Slab { char *base; size_t size; };
    std::vector<Slab> slabs_;
    std::map<char *, size_t> free_;     // start -> size, coalesced within a slab
    std::map<char *, size_t> used_;
    size_t reserved_ = 0;
    void add_slab(size_t at_least);
};

}  // namespace mag

// The opaque context of the C ABI.
struct mag_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;      // stream of the current call (own or caller's)
    mag::DeviceHeap heap;
    uint64_t launches = 0;              // kernels launched by the current call
    int tune = 0;                       // MAG_TUNE debug switches (see PcgScalars::tune)
    bool rs_attr_set = false;           // radix sort (64-bit keys): dynamic shared memory opt-in done on this device
    bool rs32_attr_set = false;         // same, 32-bit keys
    mag::Comm *comm = nullptr;
    // side stream for host<->device copies that run beside kernels of the run stream (aux_fork / aux_join below)
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool aux_busy = false;              // copies may still be in flight on aux_stream
    // pinned host scratch for scalar read-backs
    double *h_scal = nullptr;
};

namespace mag {

inline void DeviceHeap::add_slab(size_t at_least) {
    size_t want = std::max<size_t>(at_least, std::max<size_t>(reserved_ / 2, (size_t)256 << 20));
    want = (want + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    char *base = nullptr;
    cudaError_t e = cudaMalloc((void **)&base, want);
    if (e != cudaSuccess && want > at_least) {      // growth headroom refused: take the bare minimum
        cudaGetLastError();
        want = (at_least + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
        e = cudaMalloc((void **)&base, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(MAG_ERR_OOM, "device heap: cudaMalloc of %zu MiB failed (%zu MiB already reserved): %s",
             want >> 20, reserved_ >> 20, cudaGetErrorString(e));
    }
    slabs_.push_back({base, want});
    free_[base] = want;
    reserved_ += want;
}

inline void *DeviceHeap::alloc(size_t bytes) {
    bytes = (bytes + 511) & ~(size_t)511;
    for (int attempt = 0; attempt < 2; ++attempt) {
        for (auto it = free_.begin(); it != free_.end(); ++it) {
            if (it->second >= bytes) {
                char *p = it->first;
                const size_t rest = it->second - bytes;
                free_.erase(it);
                if (rest) free_[p + bytes] = rest;
                used_[p] = bytes;
                return p;
            }
        }
        add_slab(bytes);
    }
    fail(MAG_ERR_OOM, "device heap: allocation of %zu bytes failed", bytes);
}

inline void DeviceHeap::release(void *ptr) {
    char *p = static_cast<char *>(ptr);
    auto u = used_.find(p);
    if (u == used_.end()) return;
    size_t size = u->second;
    used_.erase(u);
    // blocks never straddle slabs, and neighbours in different slabs must not merge
    const Slab *slab = nullptr;
    for (const Slab &s : slabs_) if (p >= s.base && p < s.base + s.size) { slab = &s; break; }
    auto next = free_.lower_bound(p);
    if (next != free_.end() && next->first == p + size && slab && next->first < slab->base + slab->size) {
        size += next->second;
        next = free_.erase(next);
    }
    if (next != free_.begin()) {
        auto prev = std::prev(next);
        if (prev->first + prev->second == p && slab && prev->first >= slab->base) {
            prev->second += size;
            return;
        }
    }
    free_[p] = size;
}

inline void DeviceHeap::destroy() {
    for (const Slab &s : slabs_) cudaFree(s.base);
    slabs_.clear(); free_.clear(); used_.clear(); reserved_ = 0;
}

// Device buffer carved from the context's heap.
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
        p = static_cast<T *>(c->heap.alloc((count ? count : 1) * sizeof(T)));
    }
    void zero() { if (p) MAG_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), ctx->stream)); }
    void release() {
        if (p) { ctx->heap.release(p); p = nullptr; n = 0; }
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

// Work enqueued on ctx->aux_stream after this call starts once everything queued on the run stream so far is done.
inline void aux_fork(mag_ctx *ctx) {
    MAG_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    MAG_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    ctx->aux_busy = true;
}
// The run stream continues once everything queued on the aux stream so far is done.
inline void aux_join(mag_ctx *ctx) {
    MAG_CUDA(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
    MAG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
}
// Before memory the aux stream may still touch is handed back (error paths: the join never happened).
inline void aux_drain(mag_ctx *ctx) {
    if (ctx && ctx->aux_busy) { cudaStreamSynchronize(ctx->aux_stream); ctx->aux_busy = false; }
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
