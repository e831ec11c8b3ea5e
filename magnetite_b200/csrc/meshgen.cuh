// meshgen.cuh — synthetic meshes generated directly in HBM, so the
// benchmark sizes (16 M DOF and up) never exist on the host.  Emulates the
// mesher semantics the solver depends on: 0-based node ids, CCW triangles with
// area >= 1 (untouched by check_ccw, reference src/mesher.rs:522-526), node
// defaults ux=uy=None, fx=fy=Some(0.0) (src/mesher.rs:615-624) and the
// tensile-example boundary rules (examples/tensile-example/input.json:10-33).
#pragma once
#include "common.cuh"

namespace mag {

// Plate(nx,ny,h): node (i,j) -> id j*(nx+1)+i at (i*h, j*h).
// i == 0  : ux = uy = 0 known, forces unknown
// i == nx : ux = ux_right known, fy = 0 known (uy free, fx unknown)
// else    : fx = fy = 0 known
__global__ void plate_nodes_kernel(uint32_t nx, uint32_t ny, double h, double ux_right,
                                   double *__restrict__ x, double *__restrict__ y,
                                   double *__restrict__ ux, double *__restrict__ uy,
                                   double *__restrict__ fx, double *__restrict__ fy,
                                   uint8_t *__restrict__ known) {
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)(nx + 1) * (ny + 1);
    if (id >= n) return;
    const uint32_t i = (uint32_t)(id % (nx + 1)), j = (uint32_t)(id / (nx + 1));
    x[id] = (double)i * h;
    y[id] = (double)j * h;
    ux[id] = 0.0; uy[id] = 0.0; fx[id] = 0.0; fy[id] = 0.0;
    uint8_t k;
    if (i == 0) k = MAG_KNOWN_UX | MAG_KNOWN_UY;
    else if (i == nx) { k = MAG_KNOWN_UX | MAG_KNOWN_FY; ux[id] = ux_right; }
    else k = MAG_KNOWN_FX | MAG_KNOWN_FY;
    known[id] = k;
}

// cell (i,j): a = id(i,j), b = a+1, c = a+nx+1, d = c+1 -> [a,b,d], [a,d,c]; cell-major.
__global__ void plate_elems_kernel(uint32_t nx, uint32_t ny, uint32_t *__restrict__ n0,
                                   uint32_t *__restrict__ n1, uint32_t *__restrict__ n2) {
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (size_t)nx * ny) return;
    const uint32_t i = (uint32_t)(cell % nx), j = (uint32_t)(cell / nx);
    const uint32_t a = j * (nx + 1) + i, b = a + 1, c = a + nx + 1, d = c + 1;
    n0[2 * cell] = a; n1[2 * cell] = b; n2[2 * cell] = d;
    n0[2 * cell + 1] = a; n1[2 * cell + 1] = d; n2[2 * cell + 1] = c;
}

// ---- perforated plate: Plate(nx,ny,h) minus the cells whose centre lies within radius*h of the
// lattice points ((k+0.5)*pitch*h, (l+0.5)*pitch*h); unreferenced nodes dropped, survivors
// renumbered in row-major order.  Same formulas as meshgen.py.
__device__ __forceinline__ bool perforated_keep_cell(uint32_t i, uint32_t j, double h, double P, double R2) {
    const double cx = ((double)i + 0.5) * h, cy = ((double)j + 0.5) * h;
    const double gx = (floor(cx / P) + 0.5) * P, gy = (floor(cy / P) + 0.5) * P;
    const double dx = cx - gx, dy = cy - gy;
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) > R2;    // no FMA: identical to the numpy generator
}

__global__ void perforated_flags_kernel(uint32_t nx, uint32_t ny, double h, double P, double R2,
                                        uint32_t *__restrict__ cell_keep, uint32_t *__restrict__ node_used) {
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_nodes = (size_t)(nx + 1) * (ny + 1);
    if (id >= n_nodes) return;
    const uint32_t i = (uint32_t)(id % (nx + 1)), j = (uint32_t)(id / (nx + 1));
    if (i < nx && j < ny) cell_keep[(size_t)j * nx + i] = perforated_keep_cell(i, j, h, P, R2) ? 1u : 0u;
    bool used = false;                       // a node survives if any of its (up to) four cells does
    for (int dj = -1; dj <= 0; ++dj)
        for (int di = -1; di <= 0; ++di) {
            const long ci = (long)i + di, cj = (long)j + dj;
            if (ci >= 0 && cj >= 0 && ci < (long)nx && cj < (long)ny)
                used = used || perforated_keep_cell((uint32_t)ci, (uint32_t)cj, h, P, R2);
        }
    node_used[id] = used ? 1u : 0u;
}

__global__ void perforated_nodes_kernel(uint32_t nx, uint32_t ny, double h, double ux_right,
                                        const uint32_t *__restrict__ node_pos, double *__restrict__ x,
                                        double *__restrict__ y, double *__restrict__ ux, double *__restrict__ uy,
                                        double *__restrict__ fx, double *__restrict__ fy, uint8_t *__restrict__ known) {
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (size_t)(nx + 1) * (ny + 1)) return;
    if (node_pos[id + 1] == node_pos[id]) return;
    const uint32_t i = (uint32_t)(id % (nx + 1)), j = (uint32_t)(id / (nx + 1));
    const uint32_t o = node_pos[id];
    x[o] = (double)i * h; y[o] = (double)j * h;
    ux[o] = (i == nx) ? ux_right : 0.0; uy[o] = 0.0; fx[o] = 0.0; fy[o] = 0.0;
    known[o] = (i == 0) ? (MAG_KNOWN_UX | MAG_KNOWN_UY) : (i == nx) ? (MAG_KNOWN_UX | MAG_KNOWN_FY)
                                                                   : (MAG_KNOWN_FX | MAG_KNOWN_FY);
}

__global__ void perforated_elems_kernel(uint32_t nx, uint32_t ny, const uint32_t *__restrict__ cell_pos,
                                        const uint32_t *__restrict__ node_pos, uint32_t *__restrict__ n0,
                                        uint32_t *__restrict__ n1, uint32_t *__restrict__ n2) {
    const size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (size_t)nx * ny) return;
    if (cell_pos[cell + 1] == cell_pos[cell]) return;
    const uint32_t i = (uint32_t)(cell % nx), j = (uint32_t)(cell / nx);
    const size_t a = (size_t)j * (nx + 1) + i;
    const uint32_t na = node_pos[a], nb = node_pos[a + 1], nc = node_pos[a + nx + 1], nd = node_pos[a + nx + 2];
    const size_t e = 2 * (size_t)cell_pos[cell];
    n0[e] = na; n1[e] = nb; n2[e] = nd;
    n0[e + 1] = na; n1[e + 1] = nd; n2[e + 1] = nc;
}

}  // namespace mag

struct mag_devmesh {
    mag_ctx *ctx = nullptr;
    uint64_t n_nodes = 0, n_elems = 0;
    mag::DevBuf<double> x, y, ux, uy, fx, fy;
    mag::DevBuf<uint32_t> n0, n1, n2;
    mag::DevBuf<uint8_t> known;
};
