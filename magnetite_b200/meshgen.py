"""Synthetic meshes for parity tests and benchmarks; gmsh is not needed.

They emulate the mesher semantics the solver depends on: 0-based node ids, node defaults
ux=uy=None / fx=fy=Some(0.0) (reference src/mesher.rs:615-624), CCW triangles whose area is
>= 1 so `check_ccw` (src/mesher.rs:522-526) leaves them alone, and the tensile-example
boundary rules (examples/tensile-example/input.json:10-33).
"""
from __future__ import annotations

import numpy as np

from .datatypes import KNOWN_FX, KNOWN_FY, KNOWN_UX, KNOWN_UY, MeshSoA, ModelMetadata

EXAMPLE_MATERIAL = ModelMetadata(youngs_modulus=69e9, poisson_ratio=0.33, part_thickness=0.5)


def _tensile_bcs(x, i_left, i_right, ux_right):
    n = x.shape[0]
    ux = np.zeros(n); uy = np.zeros(n); fx = np.zeros(n); fy = np.zeros(n)
    known = np.full(n, KNOWN_FX | KNOWN_FY, np.uint8)          # free nodes: fx = fy = Some(0)
    known[i_left] = KNOWN_UX | KNOWN_UY                         # clamped
    known[i_right] = KNOWN_UX | KNOWN_FY                        # ux prescribed, fy = 0
    ux[i_right] = ux_right
    return ux, uy, fx, fy, known


def plate(nx: int, ny: int, h: float = 2.0, ux_right: float = 3.0) -> MeshSoA:
    """Plate(nx,ny,h): node (i,j) -> id j*(nx+1)+i at (i*h, j*h); cell (a,b,c,d) ->
    triangles [a,b,d], [a,d,c], cell-major.  Left edge clamped, right edge ux=ux_right, fy=0."""
    ii, jj = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    x = (ii * float(h)).ravel().astype(np.float64)
    y = (jj * float(h)).ravel().astype(np.float64)
    ci, cj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    a = (cj * (nx + 1) + ci).ravel().astype(np.uint32)
    b, c = a + 1, a + np.uint32(nx + 1)
    d = c + 1
    n0 = np.empty(2 * a.size, np.uint32); n1 = np.empty_like(n0); n2 = np.empty_like(n0)
    n0[0::2], n1[0::2], n2[0::2] = a, b, d
    n0[1::2], n1[1::2], n2[1::2] = a, d, c
    ids = np.arange((nx + 1) * (ny + 1))
    left, right = ids[ids % (nx + 1) == 0], ids[ids % (nx + 1) == nx]
    ux, uy, fx, fy, known = _tensile_bcs(x, left, right, ux_right)
    return MeshSoA(x, y, n0, n1, n2, ux, uy, fx, fy, known,
                   {"kind": "plate", "nx": nx, "ny": ny, "h": h})


def jitter(mesh: MeshSoA, frac: float = 0.2, seed: int = 12345) -> MeshSoA:
    """Displace interior nodes by U(-frac*h, frac*h)^2 (NumPy PCG64(seed)); boundary nodes stay so
    the box rules still select the same nodes."""
    m = mesh.copy()
    h = float(m.meta.get("h", 1.0))
    rng = np.random.Generator(np.random.PCG64(seed))
    interior = (m.x > m.x.min()) & (m.x < m.x.max()) & (m.y > m.y.min()) & (m.y < m.y.max())
    dx = rng.uniform(-frac * h, frac * h, m.n_nodes)
    dy = rng.uniform(-frac * h, frac * h, m.n_nodes)
    m.x[interior] += dx[interior]
    m.y[interior] += dy[interior]
    m.meta["jitter"] = frac
    return m


def perforated_plate(nx: int, ny: int, h: float = 2.0, pitch: int = 64, radius: int = 16,
                     ux_right: float = 3.0) -> MeshSoA:
    """Plate with circular holes: cells whose centre lies within `radius*h` of the lattice points
    at pitch `pitch*h` (offset pitch/2) are removed, unreferenced nodes dropped and the survivors
    renumbered in row-major order.  Hole boundaries are traction-free."""
    full = plate(nx, ny, h, ux_right)
    ci, cj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    cx = (ci.ravel() + 0.5) * h
    cy = (cj.ravel() + 0.5) * h
    P, R = pitch * h, radius * h
    gx = (np.floor(cx / P) + 0.5) * P
    gy = (np.floor(cy / P) + 0.5) * P
    keep_cell = (cx - gx) ** 2 + (cy - gy) ** 2 > R * R
    keep_el = np.repeat(keep_cell, 2)
    n0, n1, n2 = full.n0[keep_el], full.n1[keep_el], full.n2[keep_el]
    used = np.zeros(full.n_nodes, bool)
    used[n0] = True; used[n1] = True; used[n2] = True
    newid = np.cumsum(used, dtype=np.int64) - 1
    sel = np.flatnonzero(used)
    rn = lambda a: newid[a].astype(np.uint32)
    return MeshSoA(full.x[sel], full.y[sel], rn(n0), rn(n1), rn(n2), full.ux[sel], full.uy[sel],
                   full.fx[sel], full.fy[sel], full.known[sel],
                   {"kind": "perforated", "nx": nx, "ny": ny, "h": h, "pitch": pitch, "radius": radius})


def patch_square(side: float = 2.0, eps_x: float = 0.005) -> MeshSoA:
    """KAT-3: one square, two triangles, uniaxial strain eps_x: node0 clamped,
    node3 ux=0 / fy=0, nodes 1,2 ux = eps_x*side / fy=0.  Exact solution uy = -nu*eps_x*y."""
    s = float(side)
    x = np.array([0.0, s, s, 0.0]); y = np.array([0.0, 0.0, s, s])
    n0 = np.array([0, 0], np.uint32); n1 = np.array([1, 2], np.uint32); n2 = np.array([2, 3], np.uint32)
    ux = np.array([0.0, eps_x * s, eps_x * s, 0.0]); uy = np.zeros(4); fx = np.zeros(4); fy = np.zeros(4)
    known = np.array([KNOWN_UX | KNOWN_UY, KNOWN_UX | KNOWN_FY, KNOWN_UX | KNOWN_FY, KNOWN_UX | KNOWN_FY], np.uint8)
    return MeshSoA(x, y, n0, n1, n2, ux, uy, fx, fy, known, {"kind": "patch", "h": s})
