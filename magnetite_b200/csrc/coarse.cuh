// coarse.cuh — two-level preconditioner  M^-1 = D^-1 + P Ac^-1 P^T  ("better
// preconditioning ... the real lever on time-to-solve once SpMV is at roofline").
//
// The reference has no preconditioner and Jacobi-PCG needs ~4.3*nx iterations on an nx-wide plate because the
// smooth error modes converge slowly; a coarse space that carries them makes the count depend on the size of
// an aggregate (H/h) instead of the size of the domain (L/h): 17 203 -> 545 iterations at 16 M DOF.
//
//   aggregates  nodes are binned geometrically into an nbx x nby grid of boxes (any 2-D mesh has
//               coordinates, so this needs no graph algorithm), numbered along the SHORTER grid side first, so
//               that Ac is banded with half bandwidth hb = 3*(min(nbx,nby)+1)+2;
//   P           three columns per aggregate — the rigid-body modes of plane elasticity restricted to
//               the aggregate: x-translation, y-translation, rotation about the box centre (scaled by
//               1/H).  P is never stored: a row carries `mode` (= 3*aggregate + axis) and `rot`;
//   Ac = P^T K_ff P   accumulated WITHOUT atomics as a compact 9-point block stencil (one CTA per aggregate,
//               one thread per (neighbour box, alpha, beta) entry, rows visited in a fixed order — so the
//               preconditioner, and with it the whole solve, stays bit-reproducible), summed over the ranks once;
//   Ac = L L^T  banded Cholesky in ONE CTA whose (hb+1)^2 window lives in shared memory (~2 ms for 6144
//               unknowns; a non-positive pivot = the system is not SPD = the solve falls back to Jacobi);
//   Ac^-1       only the ROWS a rank needs — those of the aggregates its own rows (and halo) touch: 1/R of
//               them on R GPUs — by banded forward/backward substitutions, one warp per right-hand side, the
//               factor's rows shared by a CTA's 32 warps through shared memory.  A dense row is the
//               GPU-friendly form of a coarse solve: applying it is a fully parallel GEMV, where triangular
//               solves would be 2*nc sequential steps per iteration;
//   apply       w = P^T r over the local rows (segmented sums, rows sorted by aggregate), every rank's
//               partial w goes STRAIGHT INTO EVERY RANK's buffer as self-validating LL words over NVLink
//               (no NCCL call inside the iteration); the GEMV kernel adds the partials in rank order, applies
//               its rows of Ac^-1 and posts its share of w.y; z = D^-1 r + P y is formed inside the p-update
//               kernel and r.z = r.D^-1 r + w.y needs no extra pass over the fine vectors.
// No library: round 1 inverted the dense Ac with cuSOLVER potrf/potri (100 ms, on every rank, 302 MB read by
// every rank in every iteration).
#pragma once
#include "comm.cuh"
#include "common.cuh"
#include "pcg.cuh"
#include "radix_sort.cuh"
#include "spmv.cuh"

namespace mag {

constexpr uint32_t kCoarseMaxAgg = 2048;
constexpr uint32_t kCoarseMax = 3 * kCoarseMaxAgg;          // coarse unknowns at most

struct CoarseGrid {
    uint32_t nbx = 1, nby = 1;
    int x_fast = 1;                      // aggregate index = by*nbx + bx (x fast) or bx*nby + by
    double x0 = 0, y0 = 0, hx = 1, hy = 1;
    __host__ __device__ uint32_t index(uint32_t bx, uint32_t by) const { return x_fast ? by * nbx + bx : bx * nby + by; }
    __host__ __device__ void coords(uint32_t I, int &bx, int &by) const {
        if (x_fast) { bx = (int)(I % nbx); by = (int)(I / nbx); } else { by = (int)(I % nby); bx = (int)(I / nby); }
    }
    __host__ __device__ bool neighbour(uint32_t I, int k, uint32_t &J) const {      // k = (dy+1)*3 + (dx+1)
        int bx, by;
        coords(I, bx, by);
        const int jx = bx + (k % 3) - 1, jy = by + (k / 3) - 1;
        if (jx < 0 || jy < 0 || jx >= (int)nbx || jy >= (int)nby) return false;
        J = index((uint32_t)jx, (uint32_t)jy);
        return true;
    }
};

// Where the partial restrictions of all ranks meet: wbuf[(parity*kMaxRanks + src)*kCoarseMax + j], LL words.
struct CoarseLinks {
    int n = 1, me = 0;
    LLWord *wbuf[kMaxRanks];           // rank r's buffer as mapped in this process
};
constexpr size_t kCoarseWbufWords = 2ull * kMaxRanks * kCoarseMax;

struct CoarseSpace {
    bool ready = false, failed = false;   // failed: Ac is not positive definite (not an SPD system): Jacobi instead
    CoarseGrid grid;
    uint32_t n_agg = 0, nc = 0, hb = 0;
    uint32_t n_lagg = 0, m = 0;
    DevBuf<uint32_t> mode;        // per GLOBAL reduced column: 3*aggregate + axis
    DevBuf<double> rot;           // per GLOBAL reduced column: rotation-mode coefficient
    DevBuf<uint32_t> perm_ax;     // local rows sorted by aggregate: (row << 2) | axis
    DevBuf<double> rot_perm;      // rot of those rows, in that order
    DevBuf<uint32_t> agg_ptr;     // n_agg+1 segment starts into perm_ax
    DevBuf<uint32_t> lagg;        // the aggregates that have local rows
    DevBuf<uint32_t> crow;        // coarse unknowns whose row of Ac^-1 this rank applies, ascending (m)
    DevBuf<uint8_t> wy_mine;      // per crow entry: this rank adds w_j*y_j to the global w.y
    DevBuf<uint16_t> touch;       // per aggregate: ranks (bit mask) with own rows in it
    DevBuf<double> Ac_compact;    // n_agg x 81 (setup only)
    DevBuf<double> Ainv;          // m x nc
    DevBuf<double> y;             // nc (entries listed in crow are valid)
    DevBuf<LLWord> wbuf_local;    // single rank without a shared slab
    DevBuf<double> partials;
    DevBuf<unsigned> ticket;
    CoarseLinks links;
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
                                      const uint32_t *__restrict__ colmap, size_t n_dof, CoarseGrid g,
                                      uint32_t *__restrict__ mode, double *__restrict__ rot) {
    const size_t d = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_dof) return;
    const uint32_t node = (uint32_t)(d >> 1), ax = (uint32_t)(d & 1);
    if ((known[node] >> ax) & 1u) return;                // displacement prescribed: not a column
    const double2 p = xy[node];
    const uint32_t bx = (uint32_t)fmin(fmax(floor((p.x - g.x0) / g.hx), 0.0), (double)(g.nbx - 1));
    const uint32_t by = (uint32_t)fmin(fmax(floor((p.y - g.y0) / g.hy), 0.0), (double)(g.nby - 1));
    const uint32_t agg = g.index(bx, by);
    const double xc = g.x0 + ((double)bx + 0.5) * g.hx, yc = g.y0 + ((double)by + 0.5) * g.hy;
    const uint32_t c = colmap[d];
    mode[c] = 3u * agg + ax;
    rot[c] = ax ? (p.x - xc) / g.hx : -(p.y - yc) / g.hy;
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

// rows in aggregate order with everything the restriction needs next to them: (row << 2) | axis and rot
__global__ void coarse_pack_rows_kernel(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ mode,
                                        const double *__restrict__ rot, uint32_t n_rows, uint32_t row_lo,
                                        uint32_t *__restrict__ perm_ax, double *__restrict__ rot_perm) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_rows) return;
    const uint32_t row = perm[s], g = row_lo + row;
    perm_ax[s] = (row << 2) | (mode[g] % 3u);
    rot_perm[s] = rot[g];
}

// need[a] = 1 for every aggregate a column of [ext_lo, ext_hi) belongs to (owned rows and halo)
__global__ void coarse_needed_kernel(const uint32_t *__restrict__ mode, uint32_t ext_lo, uint32_t ext_hi,
                                     uint8_t *__restrict__ need) {
    const uint32_t i = ext_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ext_hi) need[mode[i] / 3u] = 1;
}

// Galerkin product: CTA = aggregate I; thread t < 81 owns entry (neighbour k, alpha, beta) of the 3x27
// block row and walks the aggregate's rows and their CSR entries in a fixed order.  `far` counts
// couplings outside the 3x3 neighbourhood (boxes smaller than an element): the caller then refuses.
__global__ void __launch_bounds__(96)
coarse_galerkin_kernel(const uint32_t *__restrict__ agg_ptr, const uint32_t *__restrict__ perm_ax,
                       const uint32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                       const double *__restrict__ val, const uint32_t *__restrict__ mode,
                       const double *__restrict__ rot, uint32_t row_lo, CoarseGrid g,
                       double *__restrict__ Ac, int *__restrict__ far) {
    const uint32_t I = blockIdx.x;
    const int t = threadIdx.x;
    const int k = t / 9, alpha = (t % 9) / 3, beta = t % 3;
    uint32_t J = 0xffffffffu;
    const bool live = t < 81 && g.neighbour(I, k, J);
    if (!live) J = 0xffffffffu;
    int bx, by;
    g.coords(I, bx, by);
    double acc = 0.0;
    int far_local = 0;
    for (uint32_t s = agg_ptr[I]; s < agg_ptr[I + 1]; ++s) {
        const uint32_t i = perm_ax[s] >> 2;
        const int axi = (int)(perm_ax[s] & 3u);
        const double pia = (alpha == axi) ? 1.0 : (alpha == 2 ? rot[row_lo + i] : 0.0);
        for (uint32_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const uint32_t c = (uint32_t)col[e];
            const uint32_t mj = mode[c];
            const uint32_t aj = mj / 3u;
            if (t == 0) {
                int cx, cy;
                g.coords(aj, cx, cy);
                if (cx - bx < -1 || cx - bx > 1 || cy - by < -1 || cy - by > 1) far_local = 1;
            }
            if (aj != J || pia == 0.0) continue;
            const int axj = (int)(mj % 3u);
            const double pjb = (beta == axj) ? 1.0 : (beta == 2 ? rot[c] : 0.0);
            acc = fma(pia * val[e], pjb, acc);
        }
    }
    // compact block row: Ac[I][k][alpha][beta], 81 doubles per aggregate (what is summed over the ranks)
    if (t < 81) Ac[(size_t)I * 81 + t] = live ? acc : 0.0;
    if (t == 0 && far_local) *far = 1;
}

// Lower band of Ac, row-wise: band[row*(hb+1) + k] = Ac[row][row-k], k = 0..hb, from the compact block rows.
__global__ void coarse_band_kernel(const double *__restrict__ compact, CoarseGrid g, uint32_t hb,
                                   double *__restrict__ band) {
    const uint32_t I = blockIdx.x;
    const int t = threadIdx.x;
    if (t >= 81) return;
    const int k = t / 9, alpha = (t % 9) / 3, beta = t % 3;
    uint32_t J;
    if (!g.neighbour(I, k, J)) return;
    const uint32_t row = 3u * I + alpha, c = 3u * J + beta;
    if (c > row) return;                                       // upper triangle: its mirror image is stored
    band[(size_t)row * (hb + 1) + (row - c)] = compact[(size_t)I * 81 + t];
}
// empty aggregates (holes, boxes outside the part) and modes without support: unit diagonal
__global__ void coarse_fix_diagonal_kernel(double *__restrict__ band, uint32_t nc, uint32_t hb) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nc && band[(size_t)k * (hb + 1)] == 0.0) band[(size_t)k * (hb + 1)] = 1.0;
}

// In-place banded Cholesky Ac = L L^T (right-looking), one CTA.  The rows j..j+hb that step j touches sit in a
// circular shared-memory window win[(row % W)*W + k] = A[row][row-k], W = hb+1; row j leaves for global memory
// when it is final and row j+W takes its slot.  *not_spd is set at the first non-positive pivot.
__global__ void __launch_bounds__(1024)
band_cholesky_kernel(double *__restrict__ band, uint32_t n, uint32_t hb, double *__restrict__ invd,
                     int *__restrict__ not_spd) {
    extern __shared__ double chol_smem[];
    const uint32_t W = hb + 1;
    double *win = chol_smem;                 // W*W
    double *colv = chol_smem + (size_t)W * W;   // W: column j below the diagonal, scaled
    __shared__ double s_d;
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    for (uint32_t e = tid; e < W * W; e += nt) {
        const uint32_t row = e / W, k = e % W;
        win[e] = row < n ? band[(size_t)row * W + k] : 0.0;
    }
    __syncthreads();
    for (uint32_t j = 0; j < n; ++j) {
        double *rowj = win + (size_t)(j % W) * W;
        if (tid == 0) {
            double d = rowj[0];
            if (!(d > 0.0)) { *not_spd = 1; d = 1.0; }
            d = sqrt(d);
            rowj[0] = d;
            s_d = d;
            invd[j] = 1.0 / d;
        }
        __syncthreads();
        const double d = s_d;
        const uint32_t pmax = min(hb, n - 1 - j);            // rows j+1 .. j+pmax hold column j
        for (uint32_t p = 1 + tid; p <= pmax; p += nt) {
            double *ri = win + (size_t)((j + p) % W) * W;
            const double l = ri[p] / d;
            ri[p] = l;
            colv[p] = l;
        }
        __syncthreads();
        // trailing update: A[j+p][j+q] -= l_p * l_q for 1 <= q <= p <= pmax
        for (uint32_t e = tid; e < pmax * pmax; e += nt) {
            const uint32_t p = e / pmax + 1, q = e % pmax + 1;
            if (q <= p) win[(size_t)((j + p) % W) * W + (p - q)] -= colv[p] * colv[q];
        }
        __syncthreads();
        // row j is final: store it, and bring row j+W into its slot
        for (uint32_t k = tid; k < W; k += nt) {
            band[(size_t)j * W + k] = rowj[k];
            rowj[k] = (j + W < n) ? band[(size_t)(j + W) * W + k] : 0.0;
        }
        __syncthreads();
    }
}

// upper[i*(hb+1) + k] = L[i+k][i] = lower[(i+k)*(hb+1) + k]: the rows of L^T, for coalesced backward substitutions
__global__ void band_transpose_kernel(const double *__restrict__ lower, uint32_t n, uint32_t hb,
                                      double *__restrict__ upper) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t W = hb + 1;
    if (e >= (size_t)n * W) return;
    const uint32_t i = (uint32_t)(e / W), k = (uint32_t)(e % W);
    upper[e] = (i + k < n) ? lower[(size_t)(i + k) * W + k] : 0.0;
}

// Rows of Ac^-1: out[r][0..n) = Ac^-1 e_c for c = crow[r] (Ac is symmetric).  One warp per right-hand side; the
// CTA's 32 warps step through the factor together so that its rows are read once per CTA (shared memory, chunks of
// kInvChunk rows) instead of once per warp.  Each warp keeps the last hb+1 entries of its vector in shared memory.
constexpr int kInvWarps = 32, kInvChunk = 8;
__global__ void __launch_bounds__(kInvWarps * 32)
band_inverse_rows_kernel(const double *__restrict__ lower, const double *__restrict__ upper,
                         const double *__restrict__ invd, uint32_t n, uint32_t hb,
                         const uint32_t *__restrict__ crow, uint32_t m, double *__restrict__ out) {
    extern __shared__ double inv_smem[];
    const uint32_t W = hb + 1;
    double *rows = inv_smem;                                       // kInvChunk * W
    double *zwin_all = inv_smem + (size_t)kInvChunk * W;           // kInvWarps * W
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * kInvWarps + warp;
    const bool live = r < m;
    const uint32_t c = live ? crow[r] : 0xffffffffu;
    double *zwin = zwin_all + (size_t)warp * W;
    double *orow = out + (size_t)(live ? r : 0) * n;
    for (uint32_t k = lane; k < W; k += 32) zwin[k] = 0.0;
    const uint32_t c_first = crow[blockIdx.x * kInvWarps];         // crow ascends: nothing happens before this row
    // ---- forward: L z = e_c ----
    if (live)
        for (uint32_t i = lane; i < c_first; i += 32) orow[i] = 0.0;
    for (uint32_t ib = c_first; ib < n; ib += kInvChunk) {
        __syncthreads();
        const uint32_t nrows = min((uint32_t)kInvChunk, n - ib);
        for (uint32_t e = threadIdx.x; e < nrows * W; e += blockDim.x) rows[e] = lower[(size_t)ib * W + e];
        __syncthreads();
        if (live) {
            uint32_t pos = ib % W;                                   // slot of row i in the circular window
            for (uint32_t q = 0; q < nrows; ++q, pos = (pos + 1 == W) ? 0u : pos + 1) {
                const uint32_t i = ib + q;
                const double *Li = rows + (size_t)q * W;
                double s = 0.0;
                // slots of rows before the first one hold zeros (they belong to rows still to come)
                for (uint32_t k = 1 + lane; k <= hb; k += 32) s = fma(Li[k], zwin[pos >= k ? pos - k : pos + W - k], s);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                const double z = ((i == c ? 1.0 : 0.0) - s) * invd[i];
                __syncwarp();
                if (lane == 0) zwin[pos] = z;
                __syncwarp();
            }
            if ((uint32_t)lane < nrows) orow[ib + lane] = zwin[(ib + lane) % W];
        }
    }
    // ---- backward: L^T x = z, in place ----
    __syncthreads();
    for (uint32_t k = lane; k < W; k += 32) zwin[k] = 0.0;        // x beyond the end is zero
    for (uint32_t top = n; top > 0; top -= min((uint32_t)kInvChunk, top)) {
        const uint32_t nrows = min((uint32_t)kInvChunk, top), ib = top - nrows;   // rows ib .. top-1, descending
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < nrows * W; e += blockDim.x) rows[e] = upper[(size_t)ib * W + e];
        __syncthreads();
        if (live) {
            double zreg = 0.0;
            if ((uint32_t)lane < nrows) zreg = orow[ib + lane];
            uint32_t pos = (top - 1) % W;
            for (uint32_t qq = nrows; qq > 0; --qq, pos = (pos == 0) ? W - 1 : pos - 1) {
                const uint32_t q = qq - 1, i = ib + q;
                const double *Ui = rows + (size_t)q * W;
                double s = 0.0;
                for (uint32_t k = 1 + lane; k <= hb; k += 32) s = fma(Ui[k], zwin[pos + k >= W ? pos + k - W : pos + k], s);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                const double zi = __shfl_sync(0xffffffffu, zreg, (int)q);
                const double x = (zi - s) * invd[i];
                __syncwarp();
                if (lane == 0) zwin[pos] = x;
                __syncwarp();
            }
            if ((uint32_t)lane < nrows) orow[ib + lane] = zwin[(ib + lane) % W];
        }
    }
}

// w = P^T r over the local rows: CTA = one aggregate that has local rows, fixed partition + fixed tree =>
// deterministic.  The three sums go to EVERY rank's buffer (LL words; the own rank included).
constexpr int kRestrictThreads = 512;
__global__ void __launch_bounds__(kRestrictThreads)
coarse_restrict_kernel(const uint32_t *__restrict__ lagg, const uint32_t *__restrict__ agg_ptr,
                       const uint32_t *__restrict__ perm_ax, const double *__restrict__ rot_perm,
                       const double *__restrict__ r, uint32_t row_lo, int step, CoarseLinks links,
                       const PcgScalars *__restrict__ sc) {
    if (sc->stop) return;
    const uint32_t I = lagg[blockIdx.x];
    double a[3] = {0.0, 0.0, 0.0};
    const uint32_t s1 = agg_ptr[I + 1];
    for (uint32_t s = agg_ptr[I] + threadIdx.x; s < s1; s += 2 * kRestrictThreads) {
        const uint32_t s2 = s + kRestrictThreads;                 // two independent gathers in flight
        const bool two = s2 < s1;
        const uint32_t pa0 = perm_ax[s], pa1 = two ? perm_ax[s2] : 0u;
        const double t0 = rot_perm[s], t1 = two ? rot_perm[s2] : 0.0;
        const double r0 = r[row_lo + (pa0 >> 2)], r1 = two ? r[row_lo + (pa1 >> 2)] : 0.0;
        a[pa0 & 3u] += r0;
        a[2] = fma(t0, r0, a[2]);
        if (two) {
            a[pa1 & 3u] += r1;
            a[2] = fma(t1, r1, a[2]);
        }
    }
    __shared__ double red[3][kRestrictThreads / 32];
    __shared__ double tot[3];
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
        tot[j] = s;
    }
    __syncthreads();
    if ((int)threadIdx.x < 3 * links.n) {
        const int dst = threadIdx.x / 3, j = threadIdx.x % 3;
        const uint32_t seq = ll_seq(sc, step < 0 ? 0ull : sc->chunk_base + (unsigned long long)step + 1ull);
        const int parity = step < 0 ? 0 : (step & 1);
        LLWord *slot = links.wbuf[dst] + ((size_t)parity * kMaxRanks + links.me) * kCoarseMax + 3u * I + j;
        ll_store(slot, tot[j], seq);
    }
}

// y = (rows of Ac^-1) w and this rank's share of w.y.  Every CTA first builds the complete w in shared memory
// (the partials of the ranks that touch an aggregate, added in rank order: the same bits on every rank), then
// its warps take rows.  The share of w.y goes to *wy_out, or into the mailboxes when links.n > 0.
__global__ void __launch_bounds__(256)
coarse_apply_kernel(const double *__restrict__ Ainv, const uint32_t *__restrict__ crow,
                    const uint8_t *__restrict__ wy_mine, const uint16_t *__restrict__ touch, uint32_t m, uint32_t nc,
                    int step, CoarseLinks clinks, PeerLinks links, double *__restrict__ y,
                    double *__restrict__ partials, unsigned *__restrict__ ticket, PcgScalars *sc,
                    double *__restrict__ wy_out) {
    if (sc->stop) return;
    extern __shared__ double w_s[];
    const uint32_t seq = ll_seq(sc, step < 0 ? 0ull : sc->chunk_base + (unsigned long long)step + 1ull);
    const int parity = step < 0 ? 0 : (step & 1);
    const LLWord *mine = clinks.wbuf[clinks.me] + (size_t)parity * kMaxRanks * kCoarseMax;
    for (uint32_t j = threadIdx.x; j < nc; j += blockDim.x) {
        uint32_t tm = touch[j / 3u];
        double s = 0.0;
        while (tm) {
            const int src = __ffs(tm) - 1;
            tm &= tm - 1u;
            s += ll_wait(mine + (size_t)src * kCoarseMax + j, seq, sc);
        }
        w_s[j] = s;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    double dot = 0.0;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < m; i += warps) {
        const double *a = Ainv + (size_t)i * nc;
        double acc = 0.0;
        for (uint32_t c = lane; c < nc; c += 32) acc = fma(__ldcs(a + c), w_s[c], acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) {
            const uint32_t row = crow[i];
            y[row] = acc;
            if (wy_mine[i]) dot = fma(w_s[row], acc, dot);
        }
    }
    double v[1] = {dot};
    double tot[1] = {0.0};
    const bool last = grid_sum_256<1>(v, partials, ticket, tot);
    if (links.n) {
        if (grid_is_last_cta()) mailbox_post(links, kMailWy, parity, tot[0], 0.0, seq);
    } else if (last) {
        *wy_out = tot[0];
    }
}

}  // namespace mag
