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

}  // namespace mag

struct mag_devmesh {
    mag_ctx *ctx = nullptr;
    uint64_t n_nodes = 0, n_elems = 0;
    mag::DevBuf<double> x, y, ux, uy, fx, fy;
    mag::DevBuf<uint32_t> n0, n1, n2;
    mag::DevBuf<uint8_t> known;
};
