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
    DevBuf<float> rot;            // per GLOBAL reduced column: rotation-mode coefficient.  fp32 on purpose: P is DEFINED by
                                  // these numbers (restriction, prolongation and Galerkin product read the same ones), so
                                  // their precision costs nothing but 4 bytes per row in two kernels of every iteration
    DevBuf<uint32_t> perm_ax;     // local rows sorted by aggregate: (row << 2) | axis
    DevBuf<float> rot_perm;       // rot of those rows, in that order
    DevBuf<uint32_t> agg_ptr;     // n_agg+1 segment starts into perm_ax
    DevBuf<uint32_t> lagg;        // the aggregates that have local rows
    DevBuf<uint32_t> crow;        // coarse unknowns whose row of Ac^-1 this rank applies, ascending (m)
    DevBuf<uint8_t> wy_mine;      // per crow entry: this rank adds w_j*y_j to the global w.y
    DevBuf<uint16_t> touch;       // per aggregate: ranks (bit mask) with own rows in it
    DevBuf<double> Ac_compact;    // n_agg x 81 (setup only)
    DevBuf<double> Ainv;          // m x nc
    DevBuf<double> y;             // nc (entries listed in crow are valid)
    DevBuf<double> w;             // nc: the complete restriction P^T r
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
                                      uint32_t *__restrict__ mode, float *__restrict__ rot) {
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
    rot[c] = (float)(ax ? (p.x - xc) / g.hx : -(p.y - yc) / g.hy);
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
                                        const float *__restrict__ rot, uint32_t n_rows, uint32_t row_lo,
                                        uint32_t *__restrict__ perm_ax, float *__restrict__ rot_perm) {
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

// Galerkin product Ac = P^T (K_ff P) for the local rows: CTA = aggregate I, warp w takes the aggregate's rows
// w, w+8, ... one at a time.  Lane (k, beta) < 27 forms t = (K P)[row][neighbour box k, mode beta] from the row's
// entries (each lane loads one entry, the entries are broadcast with shuffles in CSR order) and adds
// P[row][alpha] * t to its three sums (alpha = the row's axis, and the rotation mode).  A fixed assignment of
// rows to warps, a fixed order inside a warp and a fixed order over the warps: the same bits in every run.
// `far` reports couplings outside the 3x3 neighbourhood (boxes smaller than an element): the caller refuses.
constexpr int kGalerkinWarps = 8;
__global__ void __launch_bounds__(kGalerkinWarps * 32)
coarse_galerkin_kernel(const uint32_t *__restrict__ agg_ptr, const uint32_t *__restrict__ perm_ax,
                       const uint32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                       const double *__restrict__ val, const uint32_t *__restrict__ mode,
                       const float *__restrict__ rot, uint32_t row_lo, CoarseGrid g,
                       double *__restrict__ Ac, int *__restrict__ far) {
    __shared__ double part[kGalerkinWarps][81];
    const uint32_t I = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = lane / 3, beta = lane % 3;                     // lanes 27..31 only load entries
    int bx, by;
    g.coords(I, bx, by);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;                         // alpha = x-translation, y-translation, rotation
    int far_local = 0;
    for (uint32_t s = agg_ptr[I] + warp; s < agg_ptr[I + 1]; s += kGalerkinWarps) {
        const uint32_t pa = perm_ax[s], i = pa >> 2;
        const int axi = (int)(pa & 3u);
        const double roti = (double)rot[row_lo + i];
        const uint32_t e0 = rowptr[i], e1 = rowptr[i + 1];
        double t = 0.0;
        for (uint32_t base = e0; base < e1; base += 32) {
            const uint32_t e = base + lane;
            int kk = -1, axj = 0;
            double v = 0.0, rc = 0.0;
            if (e < e1) {
                const uint32_t c = (uint32_t)col[e];
                const uint32_t mj = mode[c];
                v = val[e]; rc = (double)rot[c];
                axj = (int)(mj % 3u);
                int cx, cy;
                g.coords(mj / 3u, cx, cy);
                const int dx = cx - bx, dy = cy - by;
                if (dx < -1 || dx > 1 || dy < -1 || dy > 1) far_local = 1;
                else kk = (dy + 1) * 3 + (dx + 1);
            }
            const int cnt = (int)min(32u, e1 - base);
            for (int q = 0; q < cnt; ++q) {
                const int kq = __shfl_sync(0xffffffffu, kk, q), aq = __shfl_sync(0xffffffffu, axj, q);
                const double vq = __shfl_sync(0xffffffffu, v, q), rq = __shfl_sync(0xffffffffu, rc, q);
                if (kq == k) {
                    if (beta == aq) t += vq;
                    else if (beta == 2) t = fma(vq, rq, t);
                }
            }
        }
        if (axi == 0) a0 += t; else a1 += t;
        a2 = fma(roti, t, a2);
    }
    if (lane < 27) {
        part[warp][k * 9 + 0 * 3 + beta] = a0;
        part[warp][k * 9 + 1 * 3 + beta] = a1;
        part[warp][k * 9 + 2 * 3 + beta] = a2;
    }
    if (far_local) *far = 1;
    __syncthreads();
    // compact block row: Ac[I][k][alpha][beta], 81 doubles per aggregate (what is summed over the ranks);
    // neighbours outside the grid never receive anything and stay zero
    if (threadIdx.x < 81) {
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < kGalerkinWarps; ++q) sum += part[q][threadIdx.x];
        Ac[(size_t)I * 81 + threadIdx.x] = sum;
    }
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

// In-place banded Cholesky Ac = L L^T (right-looking), one CTA, THREE columns (one aggregate) per step.  The rows
// j..j+hb+2 a step touches sit in a circular shared-memory window win[(row % R)*Wk + k] = A[row][row-k], R = hb+3 row
// slots of Wk = hb+1 entries; the three rows that enter the window travel through registers during the step.
// Two barriers per step.  The trailing update is bound by shared-memory bandwidth: with three columns at once every
// window entry is read and written once per aggregate instead of once per column (8.8 -> see profiles).
// n is a multiple of 3 (three modes per aggregate).  *not_spd is set at the first non-positive pivot.
__global__ void __launch_bounds__(1024)
band_cholesky_kernel(double *__restrict__ band, uint32_t n, uint32_t hb, double *__restrict__ invd,
                     int *__restrict__ not_spd) {
    extern __shared__ double chol_smem[];
    const uint32_t Wk = hb + 1, R = hb + 3;
    double *win = chol_smem;                        // R * Wk
    double *colv = chol_smem + (size_t)R * Wk;      // 3 * R: the three scaled columns, colv[c*R + p]
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t tx = tid & 31u, ty = tid >> 5;                    // 32 x 32 threads for the trailing update
    for (uint32_t e = tid; e < R * Wk; e += nt) {
        const uint32_t row = e / Wk, k = e % Wk;
        win[e] = row < n ? band[(size_t)row * Wk + k] : 0.0;
    }
    __syncthreads();
    uint32_t slot_j = 0;                                             // j % R
    for (uint32_t j = 0; j < n; j += 3) {
        // the rows that enter the window after this step: on their way while the step runs (3*Wk <= 1024 threads)
        double incoming = 0.0;
        {
            const uint32_t rr = tid / Wk, k = tid % Wk;
            if (rr < 3 && j + R + rr < n) incoming = band[(size_t)(j + R + rr) * Wk + k];
        }
        const uint32_t s0 = slot_j, s1 = (slot_j + 1 >= R) ? slot_j + 1 - R : slot_j + 1,
                       s2 = (slot_j + 2 >= R) ? slot_j + 2 - R : slot_j + 2;
        // 3 x 3 diagonal block, by every thread (the same bits everywhere)
        const double a00 = win[(size_t)s0 * Wk], a10 = win[(size_t)s1 * Wk + 1], a11 = win[(size_t)s1 * Wk];
        const double a20 = win[(size_t)s2 * Wk + 2], a21 = win[(size_t)s2 * Wk + 1], a22 = win[(size_t)s2 * Wk];
        bool bad = !(a00 > 0.0);
        const double r0 = rsqrt(bad ? 1.0 : a00), l00 = (bad ? 1.0 : a00) * r0;
        const double l10 = a10 * r0, l20 = a20 * r0;
        double d1 = a11 - l10 * l10;
        if (!(d1 > 0.0)) { bad = true; d1 = 1.0; }
        const double r1 = rsqrt(d1), l11 = d1 * r1;
        const double l21 = (a21 - l20 * l10) * r1;
        double d2 = a22 - l20 * l20 - l21 * l21;
        if (!(d2 > 0.0)) { bad = true; d2 = 1.0; }
        const double r2 = rsqrt(d2), l22 = d2 * r2;
        if (bad && tid == 0) *not_spd = 1;
        // panel: rows j+p, p = 3 .. pmax, hold (up to) three entries of the block column: k = p - c
        const uint32_t pmax = min(hb + 2, n - 1 - j);
        for (uint32_t p = 3 + tid; p <= pmax; p += nt) {
            uint32_t slot = slot_j + p;
            if (slot >= R) slot -= R;
            double *ri = win + (size_t)slot * Wk;
            const double x0 = p <= hb ? ri[p] : 0.0, x1 = p - 1 <= hb ? ri[p - 1] : 0.0, x2 = ri[p - 2];
            const double m0 = x0 * r0;
            const double m1 = (x1 - m0 * l10) * r1;
            const double m2 = (x2 - m0 * l20 - m1 * l21) * r2;
            if (p <= hb) ri[p] = m0;
            if (p - 1 <= hb) ri[p - 1] = m1;
            ri[p - 2] = m2;
            colv[p] = m0; colv[R + p] = m1; colv[2 * R + p] = m2;
        }
        __syncthreads();
        // the three rows of the block are final: store them, and put rows j+R .. j+R+2 into their slots — nothing
        // below touches those slots before the barrier at the end of the step
        if (tid == 0) { invd[j] = r0; invd[j + 1] = r1; invd[j + 2] = r2; }
        {
            const uint32_t rr = tid / Wk, k = tid % Wk;
            if (rr < 3) {
                const uint32_t slot = rr == 0 ? s0 : (rr == 1 ? s1 : s2);
                double v = win[(size_t)slot * Wk + k];
                if (rr == 0 && k == 0) v = l00;
                if (rr == 1) { if (k == 0) v = l11; else if (k == 1) v = l10; }
                if (rr == 2) { if (k == 0) v = l22; else if (k == 1) v = l21; else if (k == 2) v = l20; }
                band[(size_t)(j + rr) * Wk + k] = v;
                win[(size_t)slot * Wk + k] = incoming;
            }
        }
        // trailing update: A[j+p][j+q] -= sum_c l_pc * l_qc for 3 <= q <= p <= pmax, in 32 x 32 tiles of (p, q)
        double lq0[5], lq1[5], lq2[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const uint32_t q = tx + 3 + 32u * u;
            const bool ok = q <= pmax;
            lq0[u] = ok ? colv[q] : 0.0; lq1[u] = ok ? colv[R + q] : 0.0; lq2[u] = ok ? colv[2 * R + q] : 0.0;
        }
        for (uint32_t p0 = 3; p0 <= pmax; p0 += 32) {
            const uint32_t p = p0 + ty;
            if (p > pmax) break;
            uint32_t slot = slot_j + p;
            if (slot >= R) slot -= R;
            double *rp = win + (size_t)slot * Wk;
            const double lp0 = colv[p], lp1 = colv[R + p], lp2 = colv[2 * R + p];
#pragma unroll
            for (int u = 0; u < 5; ++u) {
                const uint32_t q = tx + 3 + 32u * u;
                if (q <= p) rp[p - q] -= lp0 * lq0[u] + lp1 * lq1[u] + lp2 * lq2[u];
            }
        }
        __syncthreads();
        slot_j += 3;
        if (slot_j >= R) slot_j -= R;
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
// CTA's warps step through the factor together so that its rows are read once per CTA (shared memory, chunks of
// kInvChunk rows, the next chunk prefetched through registers while the current one is used).  The last hb
// entries of a warp's vector live in REGISTERS, systolically: lane l, group g holds z_{i-1-l-32g}; after a step
// every value moves one lane up (shuffle) and the new entry enters at lane 0.
constexpr int kInvChunk = 32, kInvGroups = 5;                       // 5 x 32 = 160 >= hb (<= 3*46+2 = 140)
constexpr uint32_t kInvMaxBand = 32 * kInvGroups;
template <int kInvWarps>              // 16: many right-hand sides (fewer CTAs read the factor); 8: few (shorter steps)
__global__ void __launch_bounds__(kInvWarps * 32)
band_inverse_rows_kernel(const double *__restrict__ lower, const double *__restrict__ upper,
                         const double *__restrict__ invd, uint32_t n, uint32_t hb,
                         const uint32_t *__restrict__ crow, uint32_t m, double *__restrict__ out) {
    extern __shared__ double inv_smem[];
    const uint32_t W = hb + 1;
    double *rows[2] = {inv_smem, inv_smem + (size_t)kInvChunk * W};       // two chunks of factor rows
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * kInvWarps + warp;
    const bool live = r < m;
    const uint32_t c = live ? crow[r] : 0xffffffffu;
    double *orow = out + (size_t)(live ? r : 0) * n;
    const uint32_t c_first = crow[blockIdx.x * kInvWarps];           // crow ascends: nothing happens before this row
    const uint32_t per_thread = (kInvChunk * W + blockDim.x - 1) / blockDim.x;   // <= 32*141/(32*kInvWarps)
    double stash[18];                                               // kInvChunk*W / (32*kInvWarps) <= 32*141/256
    auto fetch = [&](const double *src, uint32_t row0, uint32_t nrows) {          // global -> registers
#pragma unroll
        for (uint32_t u = 0; u < 144 / kInvWarps; ++u) {
            const uint32_t e = threadIdx.x + u * blockDim.x;
            stash[u] = (u < per_thread && e < nrows * W) ? src[(size_t)row0 * W + e] : 0.0;
        }
    };
    auto commit = [&](double *dst, uint32_t nrows) {                              // registers -> shared
#pragma unroll
        for (uint32_t u = 0; u < 144 / kInvWarps; ++u) {
            const uint32_t e = threadIdx.x + u * blockDim.x;
            if (u < per_thread && e < nrows * W) dst[e] = stash[u];
        }
    };
    double hist[kInvGroups];                                         // hist[g] of lane l = vector entry at distance 1+l+32g
    // ---- forward: L z = e_c, rows c_first .. n-1 ascending ----
    if (live)
        for (uint32_t i = lane; i < c_first; i += 32) orow[i] = 0.0;
#pragma unroll
    for (int g = 0; g < kInvGroups; ++g) hist[g] = 0.0;
    {
        const uint32_t first_rows = min((uint32_t)kInvChunk, n - c_first);
        fetch(lower, c_first, first_rows);
        commit(rows[0], first_rows);
    }
    __syncthreads();
    int buf = 0;
    for (uint32_t ib = c_first; ib < n; ib += kInvChunk, buf ^= 1) {
        const uint32_t nrows = min((uint32_t)kInvChunk, n - ib);
        const uint32_t nb = ib + kInvChunk;
        const uint32_t next_rows = nb < n ? min((uint32_t)kInvChunk, n - nb) : 0u;
        if (next_rows) fetch(lower, nb, next_rows);
        if (live) {
            double keep = 0.0;                                       // entry ib + lane, for the coalesced store
            for (uint32_t q = 0; q < nrows; ++q) {
                const uint32_t i = ib + q;
                const double *Li = rows[buf] + (size_t)q * W;
                double s = 0.0;
#pragma unroll
                for (int g = 0; g < kInvGroups; ++g) {
                    const uint32_t k = 1u + lane + 32u * g;
                    if (k <= hb) s = fma(Li[k], hist[g], s);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                const double z = ((i == c ? 1.0 : 0.0) - s) * invd[i];
                if ((uint32_t)lane == q) keep = z;
                // shift the history one lane up; the new entry enters at lane 0 of group 0
#pragma unroll
                for (int g = kInvGroups - 1; g >= 0; --g) {
                    const double up = __shfl_up_sync(0xffffffffu, hist[g], 1);
                    const double carry = g ? __shfl_sync(0xffffffffu, hist[g - 1], 31) : z;
                    hist[g] = lane ? up : carry;
                }
            }
            if ((uint32_t)lane < nrows) orow[ib + lane] = keep;
        }
        if (next_rows) commit(rows[buf ^ 1], next_rows);
        __syncthreads();
    }
    // ---- backward: L^T x = z, in place, rows n-1 .. 0 descending ----
#pragma unroll
    for (int g = 0; g < kInvGroups; ++g) hist[g] = 0.0;             // x beyond the end is zero
    {
        const uint32_t nrows = min((uint32_t)kInvChunk, n), ib = n - nrows;
        fetch(upper, ib, nrows);
        commit(rows[0], nrows);
    }
    __syncthreads();
    buf = 0;
    for (uint32_t top = n; top > 0; buf ^= 1) {
        const uint32_t nrows = min((uint32_t)kInvChunk, top), ib = top - nrows;       // rows ib .. top-1
        const uint32_t next_rows = min((uint32_t)kInvChunk, ib), nb = ib - next_rows;
        if (next_rows) fetch(upper, nb, next_rows);
        if (live) {
            const double zreg = (uint32_t)lane < nrows ? orow[ib + lane] : 0.0;
            double keep = 0.0;
            for (uint32_t qq = nrows; qq > 0; --qq) {
                const uint32_t q = qq - 1, i = ib + q;
                const double *Ui = rows[buf] + (size_t)q * W;
                double s = 0.0;
#pragma unroll
                for (int g = 0; g < kInvGroups; ++g) {
                    const uint32_t k = 1u + lane + 32u * g;
                    if (k <= hb) s = fma(Ui[k], hist[g], s);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                const double x = (__shfl_sync(0xffffffffu, zreg, (int)q) - s) * invd[i];
                if ((uint32_t)lane == q) keep = x;
#pragma unroll
                for (int g = kInvGroups - 1; g >= 0; --g) {
                    const double up = __shfl_up_sync(0xffffffffu, hist[g], 1);
                    const double carry = g ? __shfl_sync(0xffffffffu, hist[g - 1], 31) : x;
                    hist[g] = lane ? up : carry;
                }
            }
            if ((uint32_t)lane < nrows) orow[ib + lane] = keep;
        }
        if (next_rows) commit(rows[buf ^ 1], next_rows);
        __syncthreads();
        top = ib;
    }
}

// w = P^T r over the local rows: CTA = one aggregate that has local rows, fixed partition + fixed tree =>
// deterministic.  The three sums go to EVERY rank's buffer (LL words; the own rank included).
constexpr int kRestrictThreads = 512;
__global__ void __launch_bounds__(kRestrictThreads)
coarse_restrict_kernel(const uint32_t *__restrict__ lagg, const uint32_t *__restrict__ agg_ptr,
                       const uint32_t *__restrict__ perm_ax, const float *__restrict__ rot_perm,
                       const double *__restrict__ r, uint32_t row_lo, int step, CoarseLinks links,
                       const PcgScalars *__restrict__ sc) {
    pdl_wait();
    if (sc->stop) return;
    const uint32_t I = lagg[blockIdx.x];
    double a[3] = {0.0, 0.0, 0.0};
    const uint32_t s1 = agg_ptr[I + 1];
    for (uint32_t s = agg_ptr[I] + threadIdx.x; s < s1; s += 4 * kRestrictThreads) {
        uint32_t pa[4];                                           // four independent gathers in flight
        double tq[4], rq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t su = s + u * kRestrictThreads;
            pa[u] = su < s1 ? perm_ax[su] : 0xffffffffu;
            tq[u] = su < s1 ? (double)rot_perm[su] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) rq[u] = pa[u] != 0xffffffffu ? r[row_lo + (pa[u] >> 2)] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {                             // fixed order: the same sums in every run
            if (pa[u] == 0xffffffffu) continue;
            if (pa[u] & 1u) a[1] += rq[u]; else a[0] += rq[u];
            a[2] = fma(tq[u], rq[u], a[2]);
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

// w = sum over the ranks of their partial restrictions, added in rank order (the same bits on every rank).
__global__ void __launch_bounds__(256)
coarse_gather_w_kernel(const uint16_t *__restrict__ touch, uint32_t nc, int step, CoarseLinks clinks,
                       double *__restrict__ w, PcgScalars *sc) {
    pdl_wait();
    if (sc->stop) return;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nc) return;
    const uint32_t seq = ll_seq(sc, step < 0 ? 0ull : sc->chunk_base + (unsigned long long)step + 1ull);
    const int parity = step < 0 ? 0 : (step & 1);
    const LLWord *mine = clinks.wbuf[clinks.me] + (size_t)parity * kMaxRanks * kCoarseMax;
    uint32_t tm = touch[j / 3u];
    double s = 0.0;
    while (tm) {
        const int src = __ffs(tm) - 1;
        tm &= tm - 1u;
        s += ll_wait(mine + (size_t)src * kCoarseMax + j, seq, sc);
    }
    w[j] = s;
}

// y = (this rank's rows of Ac^-1) w, one CTA per row (a row is 48 KB: a single warp streaming it is bound by its
// own load latency, which is all that is left when 8 GPUs share the rows), and this rank's share of w.y — to
// *wy_out, or into the mailboxes when links.n > 0.  Fixed partition of a row over the threads, fixed reduction
// tree, partial sums of the CTAs added in CTA order: the same bits in every run.
__global__ void __launch_bounds__(256)
coarse_apply_kernel(const double *__restrict__ Ainv, const uint32_t *__restrict__ crow,
                    const uint8_t *__restrict__ wy_mine, const double *__restrict__ w, uint32_t m, uint32_t nc,
                    int step, PeerLinks links, double *__restrict__ y, double *__restrict__ partials,
                    unsigned *__restrict__ ticket, PcgScalars *sc, double *__restrict__ wy_out) {
    pdl_wait();
    if (sc->stop) return;
    __shared__ double red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double dot = 0.0;
    for (uint32_t i = blockIdx.x; i < m; i += gridDim.x) {
        const double *a = Ainv + (size_t)i * nc;
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        uint32_t c = threadIdx.x;
        for (; c + 1792 < nc; c += 2048) {                           // eight loads in flight per thread
            double av[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) av[u] = __ldcs(a + c + 256 * u);
            acc0 = fma(av[0], __ldg(w + c), acc0);           acc1 = fma(av[1], __ldg(w + c + 256), acc1);
            acc2 = fma(av[2], __ldg(w + c + 512), acc2);     acc3 = fma(av[3], __ldg(w + c + 768), acc3);
            acc0 = fma(av[4], __ldg(w + c + 1024), acc0);    acc1 = fma(av[5], __ldg(w + c + 1280), acc1);
            acc2 = fma(av[6], __ldg(w + c + 1536), acc2);    acc3 = fma(av[7], __ldg(w + c + 1792), acc3);
        }
        for (; c < nc; c += 256) acc0 = fma(__ldcs(a + c), __ldg(w + c), acc0);
        double acc = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double sum = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) sum += red[q];
            const uint32_t row = crow[i];
            y[row] = sum;
            if (wy_mine[i]) dot = fma(__ldg(w + row), sum, dot);
        }
        __syncthreads();
    }
    double v[1] = {dot};                                         // thread 0 carries the CTA's share, the rest zero
    double tot[1] = {0.0};
    const bool last = grid_sum_256<1>(v, partials, ticket, tot);
    if (links.n) {
        const uint32_t seq = ll_seq(sc, step < 0 ? 0ull : sc->chunk_base + (unsigned long long)step + 1ull);
        if (grid_is_last_cta()) mailbox_post(links, kMailWy, step < 0 ? 0 : (step & 1), tot[0], 0.0, seq);
    } else if (last) {
        *wy_out = tot[0];
    }
}

// The same with one WARP per row: the better shape when there are enough rows to fill the machine with warps
// (one GPU holds all 6144 rows: 58 us against 74 us for the CTA-per-row kernel, 302 MB at 5.2 TB/s).
__global__ void __launch_bounds__(256)
coarse_apply_warp_kernel(const double *__restrict__ Ainv, const uint32_t *__restrict__ crow,
                         const uint8_t *__restrict__ wy_mine, const double *__restrict__ w, uint32_t m, uint32_t nc,
                         int step, PeerLinks links, double *__restrict__ y, double *__restrict__ partials,
                         unsigned *__restrict__ ticket, PcgScalars *sc, double *__restrict__ wy_out) {
    pdl_wait();
    if (sc->stop) return;
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    double dot = 0.0;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < m; i += warps) {
        const double *a = Ainv + (size_t)i * nc;
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        uint32_t c = lane;
        for (; c + 96 < nc; c += 128) {
            const double a0 = __ldcs(a + c), a1 = __ldcs(a + c + 32), a2 = __ldcs(a + c + 64), a3 = __ldcs(a + c + 96);
            acc0 = fma(a0, __ldg(w + c), acc0);
            acc1 = fma(a1, __ldg(w + c + 32), acc1);
            acc2 = fma(a2, __ldg(w + c + 64), acc2);
            acc3 = fma(a3, __ldg(w + c + 96), acc3);
        }
        for (; c < nc; c += 32) acc0 = fma(__ldcs(a + c), __ldg(w + c), acc0);
        double acc = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) {
            const uint32_t row = crow[i];
            y[row] = acc;
            if (wy_mine[i]) dot = fma(__ldg(w + row), acc, dot);
        }
    }
    double v[1] = {dot};
    double tot[1] = {0.0};
    const bool last = grid_sum_256<1>(v, partials, ticket, tot);
    if (links.n) {
        const uint32_t seq = ll_seq(sc, step < 0 ? 0ull : sc->chunk_base + (unsigned long long)step + 1ull);
        if (grid_is_last_cta()) mailbox_post(links, kMailWy, step < 0 ? 0 : (step & 1), tot[0], 0.0, seq);
    } else if (last) {
        *wy_out = tot[0];
    }
}

}  // namespace mag
