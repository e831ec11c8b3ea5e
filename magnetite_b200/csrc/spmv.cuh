// spmv.cuh — y = K_ff x in fp64 (replaces `&CsrMatrix * DVector`, reference
// src/solver.rs:31-36, nalgebra-sparse spmm_csr_dense).
//
// Two device formats of the same matrix:
//   CSR      rowptr u32 / col i32 / val f64 — the assembly product and the
//            parity artefact (bit-exact pattern vs the reference's CSR);
//   SELL-32  the solver format: rows in slices of 32 (one warp), stored
//            column-major inside the slice and padded to the slice's longest
//            row, so lane l of a warp streams val/col at base + k*32 + l —
//            every load instruction of the val and col streams is one fully
//            coalesced 256 B / 128 B request.  The x gathers go through L1/L2
//            (ld.global.nc): neighbouring rows of a mesh matrix touch
//            neighbouring columns, so they coalesce too.  When the matrix band
//            is < 32768 the columns are stored as 16-bit offsets from the row,
//            two per 32-bit word (10 instead of 12 bytes per entry).
// The SpMV that runs inside CG also produces the partial dot product p.q, so q
// is not re-read for it.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "scan.cuh"

namespace mag {

struct CsrMatrix {                 // local rows [row_lo, row_lo+n_rows) x global cols
    uint32_t n_rows = 0, row_lo = 0;
    uint64_t n_cols = 0, nnz = 0;
    DevBuf<uint32_t> rowptr;       // n_rows+1
    DevBuf<int32_t> col;
    DevBuf<double> val;
};

struct SellMatrix {
    uint32_t n_rows = 0, n_slices = 0, row_lo = 0;
    uint64_t entries = 0;          // padded entries (multiple of 32)
    bool narrow = false;           // columns stored as int16 offsets from the row (band < 32768)
    DevBuf<uint32_t> slice_off;    // n_slices+1, in units of 32 entries
    DevBuf<int32_t> col;           // wide index stream   (narrow == false)
    DevBuf<uint32_t> pcol;         // narrow index stream (narrow == true): two 16-bit offsets from the global row per word
    DevBuf<double> val;
    uint64_t index_bytes() const { return entries * (narrow ? 2ull : 4ull); }
};

// ---- CSR -> SELL-32 --------------------------------------------------------
__global__ void sell_width_kernel(const uint32_t *__restrict__ rowptr, uint32_t n_rows,
                                  uint32_t n_slices, int round_even, uint32_t *__restrict__ width) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t len = (row < n_rows) ? rowptr[row + 1] - rowptr[row] : 0u;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, off));
    const uint32_t s = row >> 5;
    if (round_even) len = (len + 1u) & ~1u;      // narrow index stream packs two offsets per word
    if ((threadIdx.x & 31) == 0 && s < n_slices) width[s] = len;
}

// largest |col - global row| of a CSR block: decides whether 16-bit column offsets suffice
__global__ void band_width_kernel(const uint32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                  uint32_t n_rows, uint32_t row_lo, int *__restrict__ band) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    int b = 0;
    if (row < n_rows)
        for (uint32_t p = rowptr[row]; p < rowptr[row + 1]; ++p) b = max(b, abs(col[p] - (int)(row_lo + row)));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, off));
    if ((threadIdx.x & 31) == 0 && b) atomicMax(band, b);
}

// Wide: sval/scol[base + k*32 + lane] = value / absolute column of entry k.  Narrow: the slice width is
// even; pcol[base/2 + (k/2)*32 + lane] packs the 16-bit offsets (column - global row) of entries k and
// k+1, and their two values sit next to each other at sval[2*(base/2 + (k/2)*32 + lane) + {0,1}].
template <bool NARROW>
__global__ void sell_fill_kernel(const uint32_t *__restrict__ rowptr, const int32_t *__restrict__ ccol,
                                 const double *__restrict__ cval, uint32_t n_rows, uint32_t row_lo,
                                 const uint32_t *__restrict__ slice_off, int32_t *__restrict__ scol,
                                 uint32_t *__restrict__ pcol, double *__restrict__ sval) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t s = row >> 5, lane = threadIdx.x & 31;
    // whole warps stay together: a slice is written by the warp that owns it
    const uint32_t n_slices = (n_rows + 31) >> 5;
    if (s >= n_slices) return;
    const uint32_t w = slice_off[s + 1] - slice_off[s];
    const size_t base = (size_t)slice_off[s] * 32 + lane;
    const size_t pbase = (size_t)slice_off[s] * 16 + lane;
    const uint32_t p0 = (row < n_rows) ? rowptr[row] : 0u;
    const uint32_t len = (row < n_rows) ? rowptr[row + 1] - p0 : 0u;
    // padding multiplies 0.0 by x at the row's own column (x is padded by 32 finite entries)
    const int32_t grow = (int32_t)(row_lo + row);
    uint32_t pack = 0;
    for (uint32_t k = 0; k < w; ++k) {
        const bool real = k < len;
        const int32_t c = real ? ccol[p0 + k] : grow;
        // narrow: values travel in pairs too (entries k, k+1 of a lane are adjacent: one 16-byte load)
        const size_t vpos = NARROW ? (pbase + (size_t)(k >> 1) * 32) * 2 + (k & 1u) : base + (size_t)k * 32;
        sval[vpos] = real ? cval[p0 + k] : 0.0;
        if (NARROW) {
            const uint32_t d = (uint32_t)(uint16_t)(int16_t)(c - grow);
            if (k & 1u) pcol[pbase + (size_t)(k >> 1) * 32] = pack | (d << 16);
            else pack = d;
        } else {
            scol[base + (size_t)k * 32] = c;
        }
    }
}

// allow_narrow: store columns as packed 16-bit offsets from the row when the band is < 32768 (two
// offsets per 32-bit word, slice widths rounded up to even).  16 M-DOF plate: 2.18 GB instead of 2.56 GB
// per SpMV and 359 us instead of 407 us — but only since the index loads are software-pipelined: before
// that the narrow kernel was latency-bound and 4 % SLOWER than the wide one despite moving 15 % less.
inline void build_sell(mag_ctx *ctx, const CsrMatrix &A, SellMatrix &S, bool allow_narrow = true) {
    S.n_rows = A.n_rows; S.row_lo = A.row_lo;
    S.n_slices = (A.n_rows + 31) / 32;
    S.slice_off.alloc(ctx, (size_t)S.n_slices + 1);
    if (S.n_slices == 0) {
        S.entries = 0; S.narrow = false; S.slice_off.zero(); S.col.alloc(ctx, 0); S.val.alloc(ctx, 0);
        return;
    }
    const unsigned blocks = cdiv((size_t)S.n_slices * 32, 256);
    int h_band = 1 << 30;
    if (allow_narrow) {
        DevBuf<int> band(ctx, 1);
        band.zero();
        MAG_LAUNCH(ctx, band_width_kernel, blocks, 256, 0, (const uint32_t *)A.rowptr.p, (const int32_t *)A.col.p,
                   A.n_rows, A.row_lo, band.p);
        MAG_CUDA(cudaMemcpyAsync(&h_band, band.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    S.narrow = allow_narrow && h_band < 32768;
    MAG_LAUNCH(ctx, sell_width_kernel, blocks, 256, 0, (const uint32_t *)A.rowptr.p, A.n_rows,
               S.n_slices, S.narrow ? 1 : 0, S.slice_off.p);
    exclusive_scan_u32(ctx, S.slice_off.p, S.n_slices, S.slice_off.p, (size_t)S.n_slices + 1);
    uint32_t groups = 0;
    MAG_CUDA(cudaMemcpyAsync(&groups, S.slice_off.p + S.n_slices, sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    S.entries = (uint64_t)groups * 32;
    S.val.alloc(ctx, S.entries);
    if (S.narrow) {
        S.pcol.alloc(ctx, S.entries / 2);
        MAG_LAUNCH(ctx, sell_fill_kernel<true>, blocks, 256, 0, (const uint32_t *)A.rowptr.p,
                   (const int32_t *)A.col.p, (const double *)A.val.p, A.n_rows, A.row_lo,
                   (const uint32_t *)S.slice_off.p, (int32_t *)nullptr, S.pcol.p, S.val.p);
    } else {
        S.col.alloc(ctx, S.entries);
        MAG_LAUNCH(ctx, sell_fill_kernel<false>, blocks, 256, 0, (const uint32_t *)A.rowptr.p,
                   (const int32_t *)A.col.p, (const double *)A.val.p, A.n_rows, A.row_lo,
                   (const uint32_t *)S.slice_off.p, S.col.p, (uint32_t *)nullptr, S.val.p);
    }
}

// ---- deterministic grid-wide sums ------------------------------------------
// Every CTA writes its partial sum(s); the last CTA to arrive (ticket counter)
// adds all partials in a fixed order, so the result does not depend on which
// CTA happens to be last.  NV = number of simultaneous sums (1 or 2).
// CTA-wide flag set by grid_sum_256: this CTA arrived last (one static __shared__ per kernel)
__device__ __forceinline__ bool &cta_is_last_flag() {
    __shared__ bool flag;
    return flag;
}
__device__ __forceinline__ bool grid_is_last_cta() { return cta_is_last_flag(); }

template <int NV>
__device__ __forceinline__ bool grid_sum_256(double (&v)[NV], double *__restrict__ partials,
                                             unsigned *__restrict__ ticket, double (&total)[NV],
                                             bool system_scope = false) {
    __shared__ double red[NV][8];
    bool &is_last = cta_is_last_flag();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
        if (lane == 0) red[i][warp] = v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[i][w];
            partials[(size_t)i * gridDim.x + blockIdx.x] = s;
        }
        if (system_scope) __threadfence_system();   // also orders this CTA's peer stores before the ticket
        else __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    if (system_scope) __threadfence_system(); else __threadfence();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += 256)
            s += __ldcg(&partials[(size_t)i * gridDim.x + b]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) red[i][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[i][w];
            total[i] = s;
        }
        *ticket = 0;      // ready for the next launch
    }
    __syncthreads();
    return threadIdx.x == 0;
}

// ---- SELL-32 SpMV ----------------------------------------------------------
// Index batch of 4 entries of one lane: four absolute columns (wide) or two words holding four
// 16-bit offsets from the row's own column (narrow).
template <class IDX> struct IdxBatch;
template <> struct IdxBatch<int32_t> {
    int c0, c1, c2, c3;
    __device__ __forceinline__ void load(const int32_t *c, uint32_t k) {
        c0 = __ldcs(c + (size_t)(k + 0) * 32); c1 = __ldcs(c + (size_t)(k + 1) * 32);
        c2 = __ldcs(c + (size_t)(k + 2) * 32); c3 = __ldcs(c + (size_t)(k + 3) * 32);
    }
    __device__ __forceinline__ int o0() const { return c0; }
    __device__ __forceinline__ int o1() const { return c1; }
    __device__ __forceinline__ int o2() const { return c2; }
    __device__ __forceinline__ int o3() const { return c3; }
};
template <> struct IdxBatch<int16_t> {
    uint32_t p0, p1;
    __device__ __forceinline__ void load(const uint32_t *pc, uint32_t k) {
        p0 = __ldcs(pc + (size_t)(k >> 1) * 32); p1 = __ldcs(pc + (size_t)((k >> 1) + 1) * 32);
    }
    __device__ __forceinline__ int o0() const { return (int)(short)(p0 & 0xffffu); }
    __device__ __forceinline__ int o1() const { return (int)p0 >> 16; }
    __device__ __forceinline__ int o2() const { return (int)(short)(p1 & 0xffffu); }
    __device__ __forceinline__ int o3() const { return (int)p1 >> 16; }
};

// One warp per slice, grid-stride over slices.  x is indexed by global column; y by local row.
// If DOT, also accumulates sum_i x[row_lo+i]*y[i].
// The kernel is latency-bound before it is bandwidth-bound (ncu: long-scoreboard stalls dominate), and
// the x gathers depend on the index loads, so the index batch of the NEXT four entries is loaded while
// the current batch's values and gathers are in flight: one memory round trip per batch instead of two.
template <bool DOT, class IDX>
__device__ __forceinline__ double sell_rows(const uint32_t *__restrict__ slice_off,
                                            const IDX *__restrict__ scol,
                                            const double *__restrict__ sval,
                                            const double *__restrict__ x, double *__restrict__ y,
                                            uint32_t n_rows, uint32_t n_slices, uint32_t row_lo) {
    constexpr bool kNarrow = sizeof(IDX) == 2;     // IDX = int16_t selects the packed-offset stream
    using Word = typename std::conditional<kNarrow, uint32_t, int32_t>::type;
    const int lane = threadIdx.x & 31;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    double dot = 0.0;
    for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
        const uint32_t o0 = __ldg(&slice_off[s]), o1 = __ldg(&slice_off[s + 1]);
        const double *v = sval + (size_t)o0 * 32 + lane;
        const Word *c = reinterpret_cast<const Word *>(scol) + (size_t)o0 * (kNarrow ? 16 : 32) + lane;
        const uint32_t w = o1 - o0;
        const uint32_t row = s * 32 + lane;
        const double *xb = kNarrow ? x + (row_lo + row) : x;    // narrow offsets are relative to the row
        double acc0 = 0.0, acc1 = 0.0;
        uint32_t k = 0;
        IdxBatch<IDX> cur, nxt;
        if (w >= 4) nxt.load(c, 0);
        if constexpr (kNarrow) {
            // vectorised value stream: one 16-byte load per lane covers two entries (512 B per warp request)
            const double2 *v2 = reinterpret_cast<const double2 *>(sval) + (size_t)o0 * 16 + lane;
            for (; k + 4 <= w; k += 4) {
                cur = nxt;
                if (k + 8 <= w) nxt.load(c, k + 4);             // next batch's indices, in flight early
                const double2 a01 = __ldcs(v2 + (size_t)(k >> 1) * 32), a23 = __ldcs(v2 + (size_t)((k >> 1) + 1) * 32);
                const double x0 = __ldg(xb + cur.o0()), x1 = __ldg(xb + cur.o1());
                const double x2 = __ldg(xb + cur.o2()), x3 = __ldg(xb + cur.o3());
                acc0 = fma(a01.x, x0, acc0); acc1 = fma(a01.y, x1, acc1);
                acc0 = fma(a23.x, x2, acc0); acc1 = fma(a23.y, x3, acc1);
            }
            for (; k < w; k += 2) {        // w is even in narrow mode
                const uint32_t p0 = __ldcs(reinterpret_cast<const uint32_t *>(c) + (size_t)(k >> 1) * 32);
                const double2 a01 = __ldcs(v2 + (size_t)(k >> 1) * 32);
                acc0 = fma(a01.x, __ldg(xb + (int)(short)(p0 & 0xffffu)), acc0);
                acc1 = fma(a01.y, __ldg(xb + ((int)p0 >> 16)), acc1);
            }
        } else {
            for (; k + 4 <= w; k += 4) {
                cur = nxt;
                if (k + 8 <= w) nxt.load(c, k + 4);             // next batch's indices, in flight early
                const double a0 = __ldcs(v + (size_t)(k + 0) * 32), a1 = __ldcs(v + (size_t)(k + 1) * 32);
                const double a2 = __ldcs(v + (size_t)(k + 2) * 32), a3 = __ldcs(v + (size_t)(k + 3) * 32);
                const double x0 = __ldg(xb + cur.o0()), x1 = __ldg(xb + cur.o1());
                const double x2 = __ldg(xb + cur.o2()), x3 = __ldg(xb + cur.o3());
                acc0 = fma(a0, x0, acc0); acc1 = fma(a1, x1, acc1);
                acc0 = fma(a2, x2, acc0); acc1 = fma(a3, x3, acc1);
            }
            for (; k < w; ++k)
                acc0 = fma(__ldcs(v + (size_t)k * 32), __ldg(xb + (int)__ldcs(reinterpret_cast<const int32_t *>(c) + (size_t)k * 32)), acc0);
        }
        if (row < n_rows) {
            const double yi = acc0 + acc1;
            y[row] = yi;
            if (DOT) dot = fma(__ldg(x + row_lo + row), yi, dot);
        }
    }
    return dot;
}

template <class IDX>
__global__ void __launch_bounds__(256, 6)
spmv_sell_kernel(const uint32_t *__restrict__ slice_off, const IDX *__restrict__ scol,
                 const double *__restrict__ sval, const double *__restrict__ x,
                 double *__restrict__ y, uint32_t n_rows, uint32_t n_slices, uint32_t row_lo) {
    (void)sell_rows<false, IDX>(slice_off, scol, sval, x, y, n_rows, n_slices, row_lo);
}

// ---- scalar CSR SpMV (thread per row; parity path and format comparison) ---
__global__ void __launch_bounds__(256)
spmv_csr_kernel(const uint32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
                uint32_t n_rows) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    double acc = 0.0;
    for (uint32_t p = rowptr[row]; p < rowptr[row + 1]; ++p)
        acc = __dadd_rn(acc, __dmul_rn(val[p], __ldg(x + col[p])));   // reference order, no FMA
    y[row] = acc;
}

// True residual of the owned rows: out[0] = sum_i (b_i - (K_ff x)_i)^2, out[1] = sum_i b_i^2, with the row
// products in the reference's order (ascending column, product then sum, no FMA) and a deterministic grid sum.
__global__ void __launch_bounds__(256)
true_residual_kernel(const uint32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                     const double *__restrict__ val, const double *__restrict__ x, const double *__restrict__ b,
                     uint32_t n_rows, double *__restrict__ partials, unsigned *__restrict__ ticket,
                     double *__restrict__ out) {
    double v[2] = {0.0, 0.0};
    for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (uint32_t p = rowptr[row]; p < rowptr[row + 1]; ++p)
            acc = __dadd_rn(acc, __dmul_rn(val[p], __ldg(x + col[p])));
        const double bi = b[row], d = bi - acc;
        v[0] = fma(d, d, v[0]);
        v[1] = fma(bi, bi, v[1]);
    }
    double tot[2] = {0.0, 0.0};
    if (grid_sum_256<2>(v, partials, ticket, tot)) { out[0] = tot[0]; out[1] = tot[1]; }
}

inline unsigned sell_grid(const mag_ctx *ctx, uint32_t n_slices) {
    const unsigned need = cdiv(n_slices, 8);
    const unsigned cap = (unsigned)ctx->sm_count * 6u;
    return need < cap ? (need ? need : 1u) : cap;
}

}  // namespace mag
