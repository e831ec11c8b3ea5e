// pcg.cuh — conjugate gradients on the device (replaces argmin's
// ConjugateGradient + Executor at reference src/solver.rs:141-157).
//
// Three kernels per iteration, all scalars (alpha, beta, residual, iteration
// count, stop flag) live in device memory, so an iteration needs no host round
// trip; `check_every` iterations are captured in one CUDA graph and the host
// only polls the stop flag between graph launches.  Once the flag is set every
// kernel returns immediately, so the result is the iterate at the exact
// stopping iteration regardless of the chunk size.
//
//   A: q = K p            and  pq  = p.q            (SELL SpMV + fused dot)
//   B: alpha = rz/pq;  x += alpha p;  r -= alpha q;  rz' = r.(Dinv r);  rr = r.r
//      last CTA: iteration count, stop test
//   C: beta = rz'/rz;  p = Dinv r + beta p
//
// compat mode (reference semantics): no preconditioner, x0 = 0, stop when the
// cost (||r||_2, or r.r) is <= 1e-4 absolute or after 1e7 iterations
// (src/solver.rs:18-19, 143, 153-154).  argmin carries r with the opposite
// sign (r = A x - b, p = -r + beta p); the iterates are identical.
#pragma once
#include "common.cuh"
#include "spmv.cuh"

namespace mag {

struct PcgScalars {
    double rz[2];        // r.z of the current / next iteration (index = iteration parity)
    double pq;
    double rr;           // r.r after the last completed iteration
    double thr2;         // stop when rr <= thr2
    double first_pq;     // sign tells negative-definite systems
    unsigned long long iter, max_iter;
    int stop;            // 1: converged, 2: max_iter, 3: breakdown
    unsigned ticket_a, ticket_b;
    int pad;
};

struct PcgWork {
    uint32_t n = 0;              // local rows
    uint32_t row_lo = 0;         // global index of local row 0
    DevBuf<double> x, r, q, dinv;
    DevBuf<double> p_store;      // global-indexed direction vector (owned part + halo)
    double *p = nullptr;         // = p_store.p (index by global reduced row)
    DevBuf<double> partials;     // 2 * grid
    DevBuf<PcgScalars> scal;
    unsigned grid_vec = 1, grid_spmv = 1;
};

__global__ void __launch_bounds__(256, 6)
pcg_spmv_kernel(const uint32_t *__restrict__ slice_off, const int32_t *__restrict__ scol,
                const double *__restrict__ sval, const double *__restrict__ p,
                double *__restrict__ q, uint32_t n_rows, uint32_t n_slices, uint32_t row_lo,
                double *__restrict__ partials, PcgScalars *__restrict__ sc) {
    if (sc->stop) return;
    double v[1] = {sell_rows<true>(slice_off, scol, sval, p, q, n_rows, n_slices, row_lo)};
    double tot[1];
    if (grid_sum_256<1>(v, partials, &sc->ticket_a, tot)) {
        sc->pq = tot[0];
        if (sc->iter == 0) sc->first_pq = tot[0];
    }
}

// same, scalar CSR (format comparison)
__global__ void __launch_bounds__(256)
pcg_spmv_csr_kernel(const uint32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                    const double *__restrict__ val, const double *__restrict__ p,
                    double *__restrict__ q, uint32_t n_rows, uint32_t row_lo,
                    double *__restrict__ partials, PcgScalars *__restrict__ sc) {
    if (sc->stop) return;
    double dot = 0.0;
    for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows;
         row += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (uint32_t k = rowptr[row]; k < rowptr[row + 1]; ++k) acc = fma(val[k], __ldg(p + col[k]), acc);
        q[row] = acc;
        dot = fma(__ldg(p + row_lo + row), acc, dot);
    }
    double v[1] = {dot};
    double tot[1];
    if (grid_sum_256<1>(v, partials, &sc->ticket_a, tot)) {
        sc->pq = tot[0];
        if (sc->iter == 0) sc->first_pq = tot[0];
    }
}

__global__ void __launch_bounds__(256)
pcg_update_xr_kernel(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
                     const double *__restrict__ q, const double *__restrict__ dinv, uint32_t n,
                     uint32_t row_lo, int parity, double *__restrict__ partials,
                     PcgScalars *__restrict__ sc) {
    if (sc->stop) return;
    const double pq = sc->pq;
    const double alpha = sc->rz[parity] / pq;
    double v[2] = {0.0, 0.0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double pi = p[row_lo + i], qi = q[i];
        const double xi = fma(alpha, pi, x[i]);
        const double ri = fma(-alpha, qi, r[i]);
        x[i] = xi;
        r[i] = ri;
        v[0] = fma(ri * dinv[i], ri, v[0]);   // r.z with z = Dinv r
        v[1] = fma(ri, ri, v[1]);
    }
    double tot[2];
    if (grid_sum_256<2>(v, partials, &sc->ticket_b, tot)) {
        sc->rz[parity ^ 1] = tot[0];
        sc->rr = tot[1];
        const unsigned long long it = sc->iter + 1;
        sc->iter = it;
        if (!(pq != 0.0) || !(tot[1] == tot[1])) sc->stop = 3;      // breakdown / NaN
        else if (tot[1] <= sc->thr2) sc->stop = 1;
        else if (it >= sc->max_iter) sc->stop = 2;
    }
}

__global__ void __launch_bounds__(256)
pcg_update_p_kernel(double *__restrict__ p, const double *__restrict__ r,
                    const double *__restrict__ dinv, uint32_t n, uint32_t row_lo, int parity,
                    const PcgScalars *__restrict__ sc) {
    if (sc->stop) return;
    const double beta = sc->rz[parity ^ 1] / sc->rz[parity];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        p[row_lo + i] = fma(beta, p[row_lo + i], r[i] * dinv[i]);
}

// x = 0, r = b, dinv from the diagonal, p = Dinv r, rz = r.z, rr = r.r
__global__ void __launch_bounds__(256)
pcg_init_kernel(double *__restrict__ x, double *__restrict__ r, double *__restrict__ p,
                double *__restrict__ dinv, const double *__restrict__ b,
                const double *__restrict__ diag, int jacobi, uint32_t n, uint32_t row_lo,
                double *__restrict__ partials, PcgScalars *__restrict__ sc) {
    double v[2] = {0.0, 0.0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double d = diag[i];
        const double di = (jacobi && d != 0.0) ? 1.0 / d : 1.0;
        const double bi = b[i];
        dinv[i] = di;
        x[i] = 0.0;
        r[i] = bi;
        const double z = bi * di;
        p[row_lo + i] = z;
        v[0] = fma(bi, z, v[0]);
        v[1] = fma(bi, bi, v[1]);
    }
    double tot[2];
    if (grid_sum_256<2>(v, partials, &sc->ticket_b, tot)) {
        sc->rz[0] = tot[0];
        sc->rr = tot[1];
    }
}

}  // namespace mag
