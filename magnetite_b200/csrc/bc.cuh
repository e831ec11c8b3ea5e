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

// Pass 1, one thread per owned DOF: the number of entries its K_ff row keeps, its right-hand side and its
// diagonal.  The BSR row is walked in ascending column order (reads only: this pass streams K at HBM speed).
__global__ void __launch_bounds__(256)
eliminate_rows_kernel(const uint32_t *__restrict__ browptr, const uint32_t *__restrict__ bcol,
                      const double *__restrict__ bval, uint32_t node_lo, uint32_t n_owned_dof,
                      const uint8_t *__restrict__ known, const uint32_t *__restrict__ rowmap,
                      const uint2 *__restrict__ colid2, const double *__restrict__ ux,
                      const double *__restrict__ uy, const double *__restrict__ fx,
                      const double *__restrict__ fy, int drop_zeros, uint32_t row_lo,
                      uint32_t *__restrict__ row_nnz, double *__restrict__ rhs, double *__restrict__ diag) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_owned_dof) return;
    const uint32_t ln = t >> 1, ax = t & 1u;
    const uint32_t dof = 2u * node_lo + t;
    if (!dof_f_known(known, dof)) return;
    const uint32_t gr = rowmap[dof];               // global reduced row
    const uint32_t r = gr - row_lo;                // local reduced row
    uint32_t cnt = 0;
    double s = 0.0, dg = 0.0;
    for (uint32_t b = browptr[ln]; b < browptr[ln + 1]; ++b) {
        const uint32_t cn = bcol[b];
        const double2 kv = *reinterpret_cast<const double2 *>(bval + (size_t)b * 4 + ax * 2);
        const uint2 cid = __ldg(colid2 + cn);      // reduced columns of (cn, x), (cn, y); ~0: displacement prescribed
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const double k = a ? kv.y : kv.x;
            const uint32_t c = a ? cid.y : cid.x;
            if (c == 0xffffffffu) {
                const double u = a ? uy[cn] : ux[cn];
                s = __dadd_rn(s, __dmul_rn(__dmul_rn(k, u), -1.0));      // solver.rs:390-391
            } else if (!drop_zeros || k != 0.0) {                        // solver.rs:132
                if (c == gr) dg = k;
                ++cnt;
            }
        }
    }
    row_nnz[r] = cnt;
    rhs[r] = __dadd_rn(s, ax ? fy[ln + node_lo] : fx[ln + node_lo]);     // solver.rs:427-432
    diag[r] = dg;
}

// Pass 2, one thread per 2x2 BLOCK of K: its (up to) two entries of the node's x row and two of its y row go to
// their places in K_ff.  Blocks are sorted by (row node, column node), so the lanes of a warp that share a row
// node are neighbours and write neighbouring entries — the stores of a warp fall into a few runs instead of one
// sector per lane, which is what a thread per ROW produced (2.55 ms of the round-1 elimination; this: see
// profiles/).  An entry's position = its row's start (rowptr) + the kept entries of the blocks before it in the
// same row: counted with ballots inside the warp, and by the row's first lane for blocks left of the warp.
__global__ void __launch_bounds__(256)
eliminate_fill_blocks_kernel(const uint32_t *__restrict__ browptr, const uint32_t *__restrict__ brow,
                             const uint32_t *__restrict__ bcol, const double *__restrict__ bval, uint32_t n_blocks,
                             uint32_t node_lo, const uint8_t *__restrict__ known, const uint32_t *__restrict__ rowmap,
                             const uint2 *__restrict__ colid2, int drop_zeros, uint32_t row_lo,
                             const uint32_t *__restrict__ rowptr, int32_t *__restrict__ col, double *__restrict__ val) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = b < n_blocks;
    const uint32_t ln = valid ? brow[b] : 0xffffffffu;
    uint32_t kn = 0;
    uint2 gr = make_uint2(0, 0), cid = make_uint2(0xffffffffu, 0xffffffffu);
    double2 v0 = make_double2(0.0, 0.0), v1 = v0;
    if (valid) {
        const uint32_t node = node_lo + ln;
        kn = known[node];
        gr = *reinterpret_cast<const uint2 *>(rowmap + 2u * (size_t)node);
        cid = __ldg(colid2 + bcol[b]);
        const double2 *src = reinterpret_cast<const double2 *>(bval + (size_t)b * 4);
        v0 = __ldcs(src); v1 = __ldcs(src + 1);
    }
    const bool ex0 = (kn & MAG_KNOWN_FX) != 0, ex1 = (kn & MAG_KNOWN_FY) != 0;      // rows of K_ff: force known
    const bool kx0 = ex0 && cid.x != 0xffffffffu && (!drop_zeros || v0.x != 0.0);
    const bool kx1 = ex0 && cid.y != 0xffffffffu && (!drop_zeros || v0.y != 0.0);
    const bool ky0 = ex1 && cid.x != 0xffffffffu && (!drop_zeros || v1.x != 0.0);
    const bool ky1 = ex1 && cid.y != 0xffffffffu && (!drop_zeros || v1.y != 0.0);
    // kept entries of the lanes left of me that belong to the same row node
    const uint32_t grp = __match_any_sync(0xffffffffu, ln);
    const uint32_t left = grp & ((1u << lane) - 1u);
    uint32_t px = __popc(__ballot_sync(0xffffffffu, kx0) & left) + __popc(__ballot_sync(0xffffffffu, kx1) & left);
    uint32_t py = __popc(__ballot_sync(0xffffffffu, ky0) & left) + __popc(__ballot_sync(0xffffffffu, ky1) & left);
    // ... and of the row's blocks that sit left of this warp: counted once, by the row's first lane here
    const int first_lane = __ffs(grp) - 1;
    uint32_t lx = 0, ly = 0;
    if (valid && lane == first_lane) {
        const uint32_t b0 = browptr[ln];
        for (uint32_t bb = b0; bb < b; ++bb) {
            const uint2 c2 = __ldg(colid2 + bcol[bb]);
            const double2 *src = reinterpret_cast<const double2 *>(bval + (size_t)bb * 4);
            const double2 w0 = src[0], w1 = src[1];
            lx += (ex0 && c2.x != 0xffffffffu && (!drop_zeros || w0.x != 0.0)) + (ex0 && c2.y != 0xffffffffu && (!drop_zeros || w0.y != 0.0));
            ly += (ex1 && c2.x != 0xffffffffu && (!drop_zeros || w1.x != 0.0)) + (ex1 && c2.y != 0xffffffffu && (!drop_zeros || w1.y != 0.0));
        }
    }
    px += __shfl_sync(0xffffffffu, lx, first_lane);
    py += __shfl_sync(0xffffffffu, ly, first_lane);
    if (kx0 || kx1) {
        const uint32_t pos = rowptr[gr.x - row_lo] + px;
        if (kx0) { col[pos] = (int32_t)cid.x; val[pos] = v0.x; }
        if (kx1) { col[pos + (kx0 ? 1u : 0u)] = (int32_t)cid.y; val[pos + (kx0 ? 1u : 0u)] = v0.y; }
    }
    if (ky0 || ky1) {
        const uint32_t pos = rowptr[gr.y - row_lo] + py;
        if (ky0) { col[pos] = (int32_t)cid.x; val[pos] = v1.x; }
        if (ky1) { col[pos + (ky0 ? 1u : 0u)] = (int32_t)cid.y; val[pos + (ky0 ? 1u : 0u)] = v1.y; }
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
