// coarse.cuh — optional two-level preconditioner  M^-1 = D^-1 + P Ac^-1 P^T  (SURVEY §8(f) rank 4:
// "better preconditioning ... the real lever on time-to-solve once SpMV is at roofline").
//
// The reference has no preconditioner and the north star names Jacobi; this is an opt-in extra
// (mag_options.precond = 2).  Jacobi-PCG needs ~4.3*nx iterations on an nx-wide plate because the
// smooth error modes converge slowly; a coarse space that carries them makes the count depend on the
// size of an aggregate (H/h) instead of the size of the domain (L/h).
//
//   aggregates  nodes are binned geometrically into an nbx x nby grid of boxes (any 2-D mesh has
//               coordinates, so this needs no graph algorithm);
//   P           three columns per aggregate — the rigid-body modes of plane elasticity restricted to
//               the aggregate: x-translation, y-translation, rotation about the box centre (scaled by
//               1/H).  P is never stored: a row carries `mode` (= 3*aggregate + axis) and `rot`;
//   Ac = P^T K_ff P   dense nc x nc (nc = 3*nbx*nby <= ~6k), accumulated WITHOUT atomics: one CTA per
//               aggregate, one thread per (neighbour box, alpha, beta) entry, rows visited in a fixed
//               order — so the preconditioner, and with it the whole solve, stays bit-reproducible;
//   Ac^-1       explicit inverse (cuSOLVER potrf + potri, loaded lazily: a plain library
//               factorisation in the setup, not on the hot path);
//   apply       w = P^T r (segmented sums over rows sorted by aggregate), y = Ac^-1 w (dense GEMV,
//               one warp per row, also returns w.y), then z = D^-1 r + P y is formed inside the
//               p-update kernel; r.z = r.D^-1 r + w.y needs no extra pass over the fine vectors.
//   multi-GPU   Ac and w are summed over ranks with NCCL (w: nc doubles per iteration), every rank
//               applies Ac^-1 redundantly.
#pragma once
#include <dlfcn.h>

#include "comm.cuh"
#include "common.cuh"
#include "pcg.cuh"
#include "radix_sort.cuh"
#include "spmv.cuh"

namespace mag {

struct CoarseSpace {
    bool ready = false;
    uint32_t nbx = 0, nby = 0, n_agg = 0, nc = 0;
    double x0 = 0, y0 = 0, hx = 1, hy = 1;
    DevBuf<uint32_t> mode;        // per GLOBAL reduced column: 3*aggregate + axis
    DevBuf<double> rot;           // per GLOBAL reduced column: rotation-mode coefficient
    DevBuf<uint32_t> perm;        // local rows sorted by aggregate
    DevBuf<uint32_t> agg_ptr;     // n_agg+1 segment starts into perm
    DevBuf<double> Ainv;          // nc x nc
    DevBuf<double> Ac_compact;    // n_agg x 81 (setup only)
    DevBuf<double> w, y;          // nc
    DevBuf<double> partials;      // gemv dot partials
    DevBuf<unsigned> ticket;
};

// bounding box of the nodes: per-CTA min/max, finished on the host (tiny)
__global__ void bbox_kernel(const double2 *__restrict__ xy, size_t n, double *__restrict__ out /*[grid][4]*/) {
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double2 p = xy[i];
        xmin = fmin(xmin, p.x); xmax = fmax(xmax, p.x);
        ymin = fmin(ymin, p.y); ymax = fmax(ymax, p.y);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, off));
        xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, off));
        ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, off));
        ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, off));
    }
    __shared__ double s[8][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s[warp][0] = xmin; s[warp][1] = xmax; s[warp][2] = ymin; s[warp][3] = ymax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            xmin = fmin(xmin, s[w][0]); xmax = fmax(xmax, s[w][1]);
            ymin = fmin(ymin, s[w][2]); ymax = fmax(ymax, s[w][3]);
        }
        out[blockIdx.x * 4 + 0] = xmin; out[blockIdx.x * 4 + 1] = xmax;
        out[blockIdx.x * 4 + 2] = ymin; out[blockIdx.x * 4 + 3] = ymax;
    }
}

// For every DOF with an unknown displacement: mode / rot of its reduced column.
__global__ void coarse_colinfo_kernel(const double2 *__restrict__ xy, const uint8_t *__restrict__ known,
                                      const uint32_t *__restrict__ colmap, size_t n_dof, double x0, double y0,
                                      double hx, double hy, uint32_t nbx, uint32_t nby,
                                      uint32_t *__restrict__ mode, double *__restrict__ rot) {
    const size_t d = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_dof) return;
    const uint32_t node = (uint32_t)(d >> 1), ax = (uint32_t)(d & 1);
    if ((known[node] >> ax) & 1u) return;                // displacement prescribed: not a column
    const double2 p = xy[node];
    uint32_t bx = (uint32_t)fmin(fmax(floor((p.x - x0) / hx), 0.0), (double)(nbx - 1));
    uint32_t by = (uint32_t)fmin(fmax(floor((p.y - y0) / hy), 0.0), (double)(nby - 1));
    const uint32_t agg = by * nbx + bx;
    const double xc = x0 + ((double)bx + 0.5) * hx, yc = y0 + ((double)by + 0.5) * hy;
    const uint32_t c = colmap[d];
    mode[c] = 3u * agg + ax;
    rot[c] = ax ? (p.x - xc) / hx : -(p.y - yc) / hy;
}

__global__ void coarse_rowkeys_kernel(const uint32_t *__restrict__ mode, uint32_t n_rows, uint32_t row_lo,
                                      uint64_t *__restrict__ keys, uint32_t *__restrict__ payload) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    keys[i] = mode[row_lo + i] / 3u;
    payload[i] = i;
}

// agg_ptr[a] = first position in the sorted key array whose key >= a
__global__ void coarse_segments_kernel(const uint64_t *__restrict__ keys, uint32_t n_rows, uint32_t n_agg,
                                       uint32_t *__restrict__ agg_ptr) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a > n_agg) return;
    uint32_t lo = 0, hi = n_rows;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (keys[mid] < a) lo = mid + 1; else hi = mid;
    }
    agg_ptr[a] = lo;
}

// Galerkin product: CTA = aggregate I; thread t < 81 owns entry (neighbour k, alpha, beta) of the 3x27
// block row and walks the aggregate's rows and their CSR entries in a fixed order.  `far` counts
// couplings outside the 3x3 neighbourhood (boxes smaller than an element): the caller then refuses.
__global__ void __launch_bounds__(96)
coarse_galerkin_kernel(const uint32_t *__restrict__ agg_ptr, const uint32_t *__restrict__ perm,
                       const uint32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                       const double *__restrict__ val, const uint32_t *__restrict__ mode,
                       const double *__restrict__ rot, uint32_t row_lo, uint32_t nbx, uint32_t nby,
                       uint32_t nc, double *__restrict__ Ac, int *__restrict__ far) {
    const uint32_t I = blockIdx.x;
    const int t = threadIdx.x;
    const int k = t / 9, alpha = (t % 9) / 3, beta = t % 3;
    const int bx = (int)(I % nbx), by = (int)(I / nbx);
    const int jx = bx + (k % 3) - 1, jy = by + (k / 3) - 1;
    const bool live = t < 81 && jx >= 0 && jy >= 0 && jx < (int)nbx && jy < (int)nby;
    const uint32_t J = live ? (uint32_t)(jy * (int)nbx + jx) : 0xffffffffu;
    double acc = 0.0;
    int far_local = 0;
    for (uint32_t s = agg_ptr[I]; s < agg_ptr[I + 1]; ++s) {
        const uint32_t i = perm[s];
        const uint32_t mi = mode[row_lo + i];
        const int axi = (int)(mi % 3u);
        const double pia = (alpha == axi) ? 1.0 : (alpha == 2 ? rot[row_lo + i] : 0.0);
        for (uint32_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const uint32_t c = (uint32_t)col[e];
            const uint32_t mj = mode[c];
            const uint32_t aj = mj / 3u;
            if (t == 0) {
                const int dx = (int)(aj % nbx) - bx, dy = (int)(aj / nbx) - by;
                if (dx < -1 || dx > 1 || dy < -1 || dy > 1) far_local = 1;
            }
            if (aj != J || pia == 0.0) continue;
            const int axj = (int)(mj % 3u);
            const double pjb = (beta == axj) ? 1.0 : (beta == 2 ? rot[c] : 0.0);
            acc = fma(pia * val[e], pjb, acc);
        }
    }
    // compact block row: Ac[I][k][alpha][beta], 81 doubles per aggregate (summed over ranks before it
    // is expanded to the dense matrix: 1.3 MB on the wire instead of nc^2 * 8 = 302 MB)
    if (t < 81) Ac[(size_t)I * 81 + t] = live ? acc : 0.0;
    if (t == 0 && far_local) *far = 1;
    (void)nc;
}

// dense[3I+alpha][3J+beta] = compact[I][k][alpha][beta] for the (up to) nine neighbours J of I
__global__ void coarse_expand_kernel(const double *__restrict__ compact, uint32_t nbx, uint32_t nby, uint32_t nc,
                                     double *__restrict__ dense) {
    const uint32_t I = blockIdx.x;
    const int t = threadIdx.x;
    if (t >= 81) return;
    const int k = t / 9, alpha = (t % 9) / 3, beta = t % 3;
    const int jx = (int)(I % nbx) + (k % 3) - 1, jy = (int)(I / nbx) + (k / 3) - 1;
    if (jx < 0 || jy < 0 || jx >= (int)nbx || jy >= (int)nby) return;
    const uint32_t J = (uint32_t)(jy * (int)nbx + jx);
    dense[(size_t)(3u * I + alpha) * nc + 3u * J + beta] = compact[(size_t)I * 81 + t];
}

// empty aggregates (holes, boxes outside the part) and modes without support: unit diagonal
__global__ void coarse_fix_diagonal_kernel(double *__restrict__ Ac, uint32_t nc) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nc && Ac[(size_t)k * nc + k] == 0.0) Ac[(size_t)k * nc + k] = 1.0;
}
// potri leaves one triangle: mirror it
__global__ void coarse_mirror_kernel(double *__restrict__ A, uint32_t nc) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c < nc && r < c) A[(size_t)c * nc + r] = A[(size_t)r * nc + c];
}

// w = P^T r over the local rows: CTA = aggregate, fixed partition + fixed tree => deterministic.
// 512 threads per aggregate: the gathers through `perm` are latency-bound, so the sequential depth
// per thread (rows of the aggregate / 512) is what sets the time.
constexpr int kRestrictThreads = 512;
__global__ void __launch_bounds__(kRestrictThreads)
coarse_restrict_kernel(const uint32_t *__restrict__ agg_ptr, const uint32_t *__restrict__ perm,
                       const uint32_t *__restrict__ mode, const double *__restrict__ rot,
                       const double *__restrict__ r, uint32_t row_lo, double *__restrict__ w,
                       const PcgScalars *__restrict__ sc) {
    if (sc->stop) return;
    const uint32_t I = blockIdx.x;
    double a[3] = {0.0, 0.0, 0.0};
    const uint32_t s1 = agg_ptr[I + 1];
    for (uint32_t s = agg_ptr[I] + threadIdx.x; s < s1; s += 2 * kRestrictThreads) {
        const uint32_t s2 = s + kRestrictThreads;                 // two independent gathers in flight
        const uint32_t g0 = row_lo + perm[s], g1 = s2 < s1 ? row_lo + perm[s2] : 0u;
        const double r0 = r[g0], r1 = s2 < s1 ? r[g1] : 0.0;
        const uint32_t m0 = mode[g0], m1 = s2 < s1 ? mode[g1] : 0u;
        const double t0 = rot[g0], t1 = s2 < s1 ? rot[g1] : 0.0;
        a[m0 % 3u] += r0;
        a[2] = fma(t0, r0, a[2]);
        if (s2 < s1) {
            a[m1 % 3u] += r1;
            a[2] = fma(t1, r1, a[2]);
        }
    }
    __shared__ double red[3][kRestrictThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a[j] += __shfl_xor_sync(0xffffffffu, a[j], off);
        if (lane == 0) red[j][warp] = a[j];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int j = threadIdx.x;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < kRestrictThreads / 32; ++k) s += red[j][k];
        w[3u * I + j] = s;
    }
}

// y = Ainv w (one warp per row) and wy = w.y (deterministic grid sum) -> sc->wy
__global__ void __launch_bounds__(256)
coarse_gemv_kernel(const double *__restrict__ Ainv, const double *__restrict__ w, double *__restrict__ y,
                   uint32_t nc, double *__restrict__ partials, unsigned *__restrict__ ticket, PcgScalars *sc,
                   double *__restrict__ wy_out) {
    if (sc->stop) return;
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    double dot = 0.0;
    for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < nc; row += warps) {
        const double *a = Ainv + (size_t)row * nc;
        double acc = 0.0;
        for (uint32_t c = lane; c < nc; c += 32) acc = fma(__ldcs(a + c), __ldg(w + c), acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) {
            y[row] = acc;
            dot = fma(__ldg(w + row), acc, dot);
        }
    }
    double v[1] = {dot};
    double tot[1] = {0.0};
    if (grid_sum_256<1>(v, partials, ticket, tot)) *wy_out = tot[0];
}

// ---- cuSOLVER, loaded on first use ---------------------------------------------------------------
struct CusolverApi {
    void *lib = nullptr;
    int (*create)(void **) = nullptr;
    int (*destroy)(void *) = nullptr;
    int (*set_stream)(void *, cudaStream_t) = nullptr;
    int (*potrf_buf)(void *, int, int, double *, int, int *) = nullptr;
    int (*potrf)(void *, int, int, double *, int, double *, int, int *) = nullptr;
    int (*potri_buf)(void *, int, int, double *, int, int *) = nullptr;
    int (*potri)(void *, int, int, double *, int, double *, int, int *) = nullptr;
};

inline CusolverApi &cusolver_api() {
    static CusolverApi api;
    if (api.lib) return api;
    const char *names[] = {"libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11"};
    for (const char *n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) fail(MAG_ERR_CUDA, "the two-level preconditioner needs libcusolver.so.11: %s", dlerror());
    auto sym = [&](const char *s) {
        void *p = dlsym(api.lib, s);
        if (!p) fail(MAG_ERR_CUDA, "libcusolver lacks %s", s);
        return p;
    };
    api.create = reinterpret_cast<int (*)(void **)>(sym("cusolverDnCreate"));
    api.destroy = reinterpret_cast<int (*)(void *)>(sym("cusolverDnDestroy"));
    api.set_stream = reinterpret_cast<int (*)(void *, cudaStream_t)>(sym("cusolverDnSetStream"));
    api.potrf_buf = reinterpret_cast<int (*)(void *, int, int, double *, int, int *)>(sym("cusolverDnDpotrf_bufferSize"));
    api.potrf = reinterpret_cast<int (*)(void *, int, int, double *, int, double *, int, int *)>(sym("cusolverDnDpotrf"));
    api.potri_buf = reinterpret_cast<int (*)(void *, int, int, double *, int, int *)>(sym("cusolverDnDpotri_bufferSize"));
    api.potri = reinterpret_cast<int (*)(void *, int, int, double *, int, double *, int, int *)>(sym("cusolverDnDpotri"));
    return api;
}

// In-place inverse of the SPD matrix A (nc x nc; symmetric, so row/column major coincide).
inline void spd_inverse(mag_ctx *ctx, double *A, uint32_t nc) {
    CusolverApi &cs = cusolver_api();
    // one handle per context, kept: creating it (cuBLAS initialisation, library load) costs ~0.1 s on one
    // GPU and ~2 s when eight processes do it at once
    if (!ctx->cusolver && cs.create(&ctx->cusolver) != 0) fail(MAG_ERR_CUDA, "cusolverDnCreate failed");
    void *h = ctx->cusolver;
    cs.set_stream(h, ctx->stream);
    const int kLower = 0;   // CUBLAS_FILL_MODE_LOWER
    int l1 = 0, l2 = 0;
    cs.potrf_buf(h, kLower, (int)nc, A, (int)nc, &l1);
    cs.potri_buf(h, kLower, (int)nc, A, (int)nc, &l2);
    DevBuf<double> work(ctx, (size_t)std::max(l1, l2) + 1);
    DevBuf<int> info(ctx, 1);
    int h_info = 0;
    int rc = cs.potrf(h, kLower, (int)nc, A, (int)nc, work.p, l1, info.p);
    MAG_CUDA(cudaMemcpyAsync(&h_info, info.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (rc != 0 || h_info != 0) {
        fail(MAG_ERR_INDEFINITE, "coarse matrix is not positive definite (potrf info %d): the two-level "
                                 "preconditioner needs an SPD system (counter-clockwise mesh); use precond = 1", h_info);
    }
    rc = cs.potri(h, kLower, (int)nc, A, (int)nc, work.p, l2, info.p);
    MAG_CUDA(cudaMemcpyAsync(&h_info, info.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (rc != 0 || h_info != 0) fail(MAG_ERR_INDEFINITE, "potri failed (info %d)", h_info);
    // column-major LOWER == row-major UPPER: entries (r, c) with c >= r are valid; mirror to c < r
    dim3 grid(cdiv(nc, 256), nc);
    MAG_LAUNCH(ctx, coarse_mirror_kernel, grid, 256, 0, A, nc);
}

}  // namespace mag
