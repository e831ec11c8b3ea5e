// bc.cuh — Dirichlet elimination on the device (replaces the dense known/unknown
// partition of reference src/solver.rs:340-432 and the dense->COO scan of
// src/solver.rs:126-137), plus the post-solve scatter and reaction forces
// (src/solver.rs:444-473).
//
// Reference semantics kept exactly:
//   rows of K_ff  = DOFs whose force is known   (solver.rs:380-383), ascending
//   cols of K_ff  = DOFs whose displacement is unknown (solver.rs:389-396), ascending
//   rhs_i = sum over known-displacement cols, ascending, of -(K[i,c]*u_c), + f_i
//           (solver.rs:390-391, 402, 427-432)
//   K_ff keeps only entries with k != 0.0 (solver.rs:132)
//   reaction f_i = sum over all cols ascending of K[i,c]*u_c (solver.rs:462-466)
#pragma once
#include "assembly.cuh"
#include "common.cuh"
#include "scan.cuh"

namespace mag {

__device__ __forceinline__ bool dof_u_known(const uint8_t *known, uint32_t dof) {
    return (known[dof >> 1] >> (dof & 1)) & 1u;            // MAG_KNOWN_UX=1, UY=2
}
__device__ __forceinline__ bool dof_f_known(const uint8_t *known, uint32_t dof) {
    return (known[dof >> 1] >> (2 + (dof & 1))) & 1u;      // MAG_KNOWN_FX=4, FY=8
}

// rowflag[d] = force known, colflag[d] = displacement unknown.  *unpaired is set when some DOF has both
// or neither (the reference's mesher never produces that, mesher.rs:881-900; its solver only compares
// the two counts, solver.rs:380-396): such a system can be assembled and exported, not solved.
__global__ void dof_flags_kernel(const uint8_t *__restrict__ known, size_t n_dof,
                                 uint32_t *__restrict__ rowflag, uint32_t *__restrict__ colflag,
                                 int *__restrict__ unpaired) {
    const size_t d = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_dof) return;
    const uint32_t rf = dof_f_known(known, (uint32_t)d) ? 1u : 0u;
    const uint32_t cf = dof_u_known(known, (uint32_t)d) ? 0u : 1u;
    rowflag[d] = rf;
    colflag[d] = cf;
    if (rf != cf) *unpaired = 1;
}

// colid[d] = reduced column of DOF d, or 0xffffffff where the displacement is prescribed (one 8-byte load per
// column node tells the fused assembly both the column ids and the prescribed flags).
__global__ void col_ids_kernel(const uint8_t *__restrict__ known, const uint32_t *__restrict__ colmap, size_t n_dof,
                               uint32_t *__restrict__ colid) {
    const size_t d = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_dof) return;
    colid[d] = dof_u_known(known, (uint32_t)d) ? 0xffffffffu : colmap[d];
}

// Pass 1 (fill == 0): count kept entries of every owned reduced row.
// Pass 2 (fill == 1): write col/val, the rhs and the diagonal.
// One thread per owned DOF; its BSR row is walked in ascending column order.
template <int FILL>
__global__ void __launch_bounds__(256)
eliminate_kernel(const uint32_t *__restrict__ browptr, const uint32_t *__restrict__ bcol,
                 const double *__restrict__ bval, uint32_t node_lo, uint32_t n_owned_dof,
                 const uint8_t *__restrict__ known, const uint32_t *__restrict__ rowmap,
                 const uint32_t *__restrict__ colmap, const double *__restrict__ ux,
                 const double *__restrict__ uy, const double *__restrict__ fx,
                 const double *__restrict__ fy, int drop_zeros, uint32_t row_lo,
                 uint32_t *__restrict__ row_nnz, const uint32_t *__restrict__ rowptr,
                 int32_t *__restrict__ col, double *__restrict__ val, double *__restrict__ rhs,
                 double *__restrict__ diag) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_owned_dof) return;
    const uint32_t ln = t >> 1, ax = t & 1u;
    const uint32_t dof = 2u * node_lo + t;
    if (!dof_f_known(known, dof)) return;
    const uint32_t r = rowmap[dof] - row_lo;       // local reduced row
    const uint32_t gr = rowmap[dof];               // global reduced row (diagonal test)
    uint32_t cnt = 0;
    uint32_t w = FILL ? rowptr[r] : 0u;
    double s = 0.0, dg = 0.0;
    for (uint32_t b = browptr[ln]; b < browptr[ln + 1]; ++b) {
        const uint32_t cn = bcol[b];
        const double2 kv = *reinterpret_cast<const double2 *>(bval + (size_t)b * 4 + ax * 2);
        const uint8_t kn = known[cn];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const double k = a ? kv.y : kv.x;
            const bool uknown = (kn >> a) & 1u;
            if (uknown) {
                if (FILL) {
                    const double u = a ? uy[cn] : ux[cn];
                    s = __dadd_rn(s, __dmul_rn(__dmul_rn(k, u), -1.0));
                }
            } else if (!drop_zeros || k != 0.0) {
                if (FILL) {
                    const uint32_t c = colmap[2u * cn + a];
                    col[w] = (int32_t)c;
                    val[w] = k;
                    if (c == gr) dg = k;
                    ++w;
                }
                ++cnt;
            }
        }
    }
    if (FILL) {
        rhs[r] = __dadd_rn(s, ax ? fy[ln + node_lo] : fx[ln + node_lo]);
        diag[r] = dg;
    } else {
        row_nnz[r] = cnt;
    }
}

// solver.rs:444-454 — U[d] = prescribed value, or the solution at colmap[d].
__global__ void scatter_solution_kernel(const uint8_t *__restrict__ known,
                                        const uint32_t *__restrict__ colmap,
                                        const double *__restrict__ bc_ux, const double *__restrict__ bc_uy,
                                        const double *__restrict__ xsol, size_t n_nodes,
                                        double *__restrict__ ux, double *__restrict__ uy) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const uint8_t k = known[i];
    ux[i] = (k & MAG_KNOWN_UX) ? bc_ux[i] : xsol[colmap[2 * i]];
    uy[i] = (k & MAG_KNOWN_UY) ? bc_uy[i] : xsol[colmap[2 * i + 1]];
}

// Inverse of the scatter: xsol[colmap[d]] = U[d] for every DOF whose displacement is unknown.
__global__ void gather_solution_kernel(const uint8_t *__restrict__ known, const uint32_t *__restrict__ colmap,
                                       const double *__restrict__ ux, const double *__restrict__ uy,
                                       size_t n_nodes, double *__restrict__ xsol) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const uint8_t k = known[i];
    if (!(k & MAG_KNOWN_UX)) xsol[colmap[2 * i]] = ux[i];
    if (!(k & MAG_KNOWN_UY)) xsol[colmap[2 * i + 1]] = uy[i];
}

// solver.rs:457-473 — forces: prescribed where known, else the full-row product.
__global__ void __launch_bounds__(256)
reactions_kernel(const uint32_t *__restrict__ browptr, const uint32_t *__restrict__ bcol,
                 const double *__restrict__ bval, uint32_t node_lo, uint32_t n_owned_dof,
                 const uint8_t *__restrict__ known, const double *__restrict__ bc_fx,
                 const double *__restrict__ bc_fy, const double *__restrict__ ux,
                 const double *__restrict__ uy, double *__restrict__ fx, double *__restrict__ fy) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_owned_dof) return;
    const uint32_t ln = t >> 1, ax = t & 1u;
    const uint32_t node = ln + node_lo;
    double f;
    if (dof_f_known(known, 2u * node + ax)) {
        f = ax ? bc_fy[node] : bc_fx[node];
    } else {
        f = 0.0;
        for (uint32_t b = browptr[ln]; b < browptr[ln + 1]; ++b) {
            const uint32_t cn = bcol[b];
            const double2 kv = *reinterpret_cast<const double2 *>(bval + (size_t)b * 4 + ax * 2);
            f = __dadd_rn(f, __dmul_rn(kv.x, ux[cn]));
            f = __dadd_rn(f, __dmul_rn(kv.y, uy[cn]));
        }
    }
    if (ax) fy[node] = f; else fx[node] = f;
}

// Expand the BSR full matrix to scalar CSR (parity export of K_total).
__global__ void bsr_to_csr_kernel(const uint32_t *__restrict__ browptr, const uint32_t *__restrict__ bcol,
                                  const double *__restrict__ bval, uint32_t n_rows_nodes,
                                  int64_t *__restrict__ rowptr, int32_t *__restrict__ col,
                                  double *__restrict__ val) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n_rows_nodes) return;
    const uint32_t ln = t >> 1, ax = t & 1u;
    const uint32_t b0 = browptr[ln], b1 = browptr[ln + 1];
    const size_t base = (size_t)4 * b0 + (size_t)ax * 2 * (b1 - b0);
    rowptr[t] = (int64_t)base;
    if (t == 2 * n_rows_nodes - 1) rowptr[t + 1] = (int64_t)(base + 2 * (size_t)(b1 - b0));
    for (uint32_t b = b0; b < b1; ++b) {
        const size_t o = base + 2 * (size_t)(b - b0);
        col[o] = (int32_t)(2 * bcol[b]);
        col[o + 1] = (int32_t)(2 * bcol[b] + 1);
        val[o] = bval[(size_t)b * 4 + ax * 2];
        val[o + 1] = bval[(size_t)b * 4 + ax * 2 + 1];
    }
}

}  // namespace mag
