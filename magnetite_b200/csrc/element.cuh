// element.cuh — per-triangle kernels: signed area, CST stiffness K_e = B^T D B t A,
// and stress recovery.  One thread per triangle, fp64, plane-stress D and the
// thickness in constant memory, node coordinates gathered as one 16-byte
// (x,y) load per node, connectivity read as three coalesced uint32 streams.
//
// Arithmetic follows the reference operation by operation (explicit
// __dmul_rn/__dadd_rn/__ddiv_rn so nvcc never contracts to FMA), which makes
// K_e, the exact-zero pattern and the stresses bit-identical to the CPU
// restatement, not merely within 1e-12:
//   area    src/solver.rs:187-193     B  src/solver.rs:204-230
//   D       src/solver.rs:240-250     K_e src/solver.rs:263-278
//   stress  src/solver.rs:496-535
#pragma once
#include "common.cuh"

namespace mag {

struct MaterialConst {
    double D[9];      // row-major 3x3, already scaled by E/(1-nu^2)
    double t;         // part_thickness
};
__constant__ MaterialConst c_mat;

// Host side of compute_stress_strain_matrix (solver.rs:240-250); this file is
// compiled with -ffp-contract=off for the host pass.
inline MaterialConst make_material(const mag_material &m) {
    MaterialConst c;
    const double nu = m.poisson_ratio;
    const double base[9] = {1.0, nu, 0.0, nu, 1.0, 0.0, 0.0, 0.0, (1.0 - nu) / 2.0};
    volatile double one_minus = 1.0 - nu * nu;
    const double scale = m.youngs_modulus / one_minus;
    for (int i = 0; i < 9; ++i) {
        volatile double v = base[i] * scale;
        c.D[i] = v;
    }
    c.t = m.part_thickness;
    return c;
}

inline void upload_material(mag_ctx *ctx, const mag_material &m) {
    MaterialConst c = make_material(m);
    MAG_CUDA(cudaMemcpyToSymbolAsync(c_mat, &c, sizeof c, 0, cudaMemcpyHostToDevice, ctx->stream));
}

__device__ __forceinline__ double fmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double fadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double fsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double fdiv(double a, double b) { return __ddiv_rn(a, b); }

struct Tri {
    double x0, y0, x1, y1, x2, y2;
};

__device__ __forceinline__ Tri load_tri(const double2 *__restrict__ xy, uint32_t a, uint32_t b,
                                        uint32_t c) {
    const double2 p0 = __ldg(&xy[a]), p1 = __ldg(&xy[b]), p2 = __ldg(&xy[c]);
    return Tri{p0.x, p0.y, p1.x, p1.y, p2.x, p2.y};
}

// solver.rs:192 — 0.5 * (x0*(y1-y2) + x1*(y2-y0) + x2*(y0-y1)), left to right.
__device__ __forceinline__ double tri_area(const Tri &t) {
    const double s = fadd(fadd(fmul(t.x0, fsub(t.y1, t.y2)), fmul(t.x1, fsub(t.y2, t.y0))),
                          fmul(t.x2, fsub(t.y0, t.y1)));
    return fmul(0.5, s);
}

// solver.rs:213-227 — B as 3 rows of 6, every entry (zeros included) / (2A).
__device__ __forceinline__ void tri_B(const Tri &t, double area, double B[3][6]) {
    const double b1 = fsub(t.y1, t.y2), b2 = fsub(t.y2, t.y0), b3 = fsub(t.y0, t.y1);
    const double g1 = fsub(t.x2, t.x1), g2 = fsub(t.x0, t.x2), g3 = fsub(t.x1, t.x0);
    const double den = fmul(2.0, area);
    const double z = fdiv(0.0, den);
    const double qb1 = fdiv(b1, den), qb2 = fdiv(b2, den), qb3 = fdiv(b3, den);
    const double qg1 = fdiv(g1, den), qg2 = fdiv(g2, den), qg3 = fdiv(g3, den);
    B[0][0] = qb1; B[0][1] = z;   B[0][2] = qb2; B[0][3] = z;   B[0][4] = qb3; B[0][5] = z;
    B[1][0] = z;   B[1][1] = qg1; B[1][2] = z;   B[1][3] = qg2; B[1][4] = z;   B[1][5] = qg3;
    B[2][0] = qg1; B[2][1] = qb1; B[2][2] = qg2; B[2][3] = qb2; B[2][4] = qg3; B[2][5] = qb3;
}

constexpr int kElemThreads = 128;
constexpr int kKeStride = 37;   // odd stride (in doubles): conflict-free 64-bit smem stores

__global__ void validate_conn_kernel(const uint32_t *__restrict__ n0, const uint32_t *__restrict__ n1,
                                     const uint32_t *__restrict__ n2, size_t n_elems,
                                     uint32_t n_nodes, int *__restrict__ bad) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    if (n0[e] >= n_nodes || n1[e] >= n_nodes || n2[e] >= n_nodes) *bad = 1;
}

__global__ void pack_xy_kernel(const double *__restrict__ x, const double *__restrict__ y,
                               double2 *__restrict__ xy, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) xy[i] = make_double2(x[i], y[i]);
}

__global__ void element_area_kernel(const double2 *__restrict__ xy, const uint32_t *__restrict__ n0,
                                    const uint32_t *__restrict__ n1, const uint32_t *__restrict__ n2,
                                    size_t n_elems, double *__restrict__ area) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    area[e] = tri_area(load_tri(xy, n0[e], n1[e], n2[e]));
}

// K_e for local element i (global id elist ? elist[i] : i).
// block_major = 1: out[(i*9 + lr*3 + lc)*4 + a*2 + b] = K_e[2lr+a][2lc+b]  (assembly layout:
//                  every 2x2 node-pair block is one aligned 32-byte sector);
// block_major = 0: out[i*36 + r*6 + c] (row-major, the parity export).
// Results are staged through shared memory so the 288 B per element leave the SM
// as fully coalesced 8-byte stores.
__global__ void __launch_bounds__(kElemThreads)
element_stiffness_kernel(const double2 *__restrict__ xy, const uint32_t *__restrict__ n0,
                         const uint32_t *__restrict__ n1, const uint32_t *__restrict__ n2,
                         const uint32_t *__restrict__ elist, size_t n_local, int block_major,
                         double *__restrict__ out) {
    __shared__ double sk[kElemThreads * kKeStride];
    const size_t first = (size_t)blockIdx.x * kElemThreads;
    const size_t i = first + threadIdx.x;
    if (i < n_local) {
        const size_t e = elist ? elist[i] : i;
        const Tri t = load_tri(xy, n0[e], n1[e], n2[e]);
        const double area = tri_area(t);
        double B[3][6];
        tri_B(t, area, B);
        // BtD = B^T * D  (6x3): k ascending, product then sum (nalgebra gemv/axcpy order)
        double BtD[6][3];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double s = fmul(B[0][r], c_mat.D[0 * 3 + c]);
                s = fadd(fmul(B[1][r], c_mat.D[1 * 3 + c]), s);
                s = fadd(fmul(B[2][r], c_mat.D[2 * 3 + c]), s);
                BtD[r][c] = s;
            }
        double *row = sk + threadIdx.x * kKeStride;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                double s = fmul(BtD[r][0], B[0][c]);
                s = fadd(fmul(BtD[r][1], B[1][c]), s);
                s = fadd(fmul(BtD[r][2], B[2][c]), s);
                s = fmul(s, area);          // solver.rs:276
                s = fmul(s, c_mat.t);       // solver.rs:277
                const int slot = block_major ? (((r >> 1) * 3 + (c >> 1)) * 4 + (r & 1) * 2 + (c & 1))
                                             : (r * 6 + c);
                row[slot] = s;
            }
    }
    __syncthreads();
    const size_t remaining = n_local - first;
    const int n_here = remaining < (size_t)kElemThreads ? (int)remaining : kElemThreads;
    double *dst = out + first * 36;
    for (int j = threadIdx.x; j < n_here * 36; j += kElemThreads) {
        const int el = j / 36, k = j - el * 36;
        dst[j] = sk[el * kKeStride + k];
    }
}

// B of every element, row-major 3x6 (solver.rs:204-230) — parity export of the pub function.
__global__ void strain_displacement_kernel(const double2 *__restrict__ xy, const uint32_t *__restrict__ n0,
                                           const uint32_t *__restrict__ n1, const uint32_t *__restrict__ n2,
                                           size_t n_elems, double *__restrict__ out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    const Tri t = load_tri(xy, n0[e], n1[e], n2[e]);
    double B[3][6];
    tri_B(t, tri_area(t), B);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) out[e * 18 + r * 6 + c] = B[r][c];
}

// solver.rs:496-535.  sigma = (D*B)*u_e; sign = -1 iff sx+sy < 1.0;
// stress = sqrt(sx^2 + sy^2) * sign.  sigma3 (optional) receives sx,sy,txy.
__global__ void __launch_bounds__(256)
stress_kernel(const double2 *__restrict__ xy, const uint32_t *__restrict__ n0,
              const uint32_t *__restrict__ n1, const uint32_t *__restrict__ n2, size_t n_elems,
              const double *__restrict__ ux, const double *__restrict__ uy,
              double *__restrict__ stress, double *__restrict__ sigma3) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    const uint32_t a = n0[e], b = n1[e], c = n2[e];
    const Tri t = load_tri(xy, a, b, c);
    double B[3][6];
    tri_B(t, tri_area(t), B);
    const double ue[6] = {__ldg(&ux[a]), __ldg(&uy[a]), __ldg(&ux[b]),
                          __ldg(&uy[b]), __ldg(&ux[c]), __ldg(&uy[c])};
    double s[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double db = fmul(c_mat.D[r * 3 + 0], B[0][j]);
            db = fadd(fmul(c_mat.D[r * 3 + 1], B[1][j]), db);
            db = fadd(fmul(c_mat.D[r * 3 + 2], B[2][j]), db);
            acc = (j == 0) ? fmul(db, ue[0]) : fadd(fmul(db, ue[j]), acc);
        }
        s[r] = acc;
    }
    const double sign = (fadd(s[0], s[1]) < 1.0) ? -1.0 : 1.0;
    stress[e] = fmul(__dsqrt_rn(fadd(fmul(s[0], s[0]), fmul(s[1], s[1]))), sign);
    if (sigma3) {
        sigma3[3 * e] = s[0];
        sigma3[3 * e + 1] = s[1];
        sigma3[3 * e + 2] = s[2];
    }
}

}  // namespace mag
