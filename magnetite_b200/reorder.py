"""Node renumbering around the solve for meshes whose ids carry no locality.

The reference keeps gmsh's node tags as ids (src/mesher.rs:663-671), so K has no band structure on
the example geometries.  The GPU path is correct for any numbering, but the row-block partition, the
halo extents (mag_halo_plan) and the 16-bit SELL column offsets want neighbouring nodes to have
neighbouring ids.  `rcm` asks the library (mag_reorder_rcm, csrc/reorder.cpp, host only) for a
reverse Cuthill-McKee permutation; `permute_mesh` applies it; `unpermute_nodal` maps per-node results
back, so callers keep the reference's numbering: DOF = 2*node + axis of the ORIGINAL ids.

Element order and the local node order inside every element are untouched: areas, orientation,
per-element stress and the order in which element contributions are summed into an entry of K
(ascending element index, solver.rs:299-323) stay what they were.  What does change is the order in
which prescribed columns are summed into the rhs (ascending NEW column, solver.rs:427) and the order
of the dot products, so results agree with the un-permuted solve to rounding, not bit for bit.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _lib
from .datatypes import MeshSoA
from .error import MagnetiteError


def _host_check(rc: int, what: str) -> None:
    if rc != 0:
        msg = (_lib.load().mag_host_last_error() or b"").decode("utf-8", "replace")
        raise MagnetiteError.Solver(f"{what}: {msg}", code=rc)


def mesh_band(mesh: MeshSoA) -> int:
    """max |a - b| over the node pairs of every element (half-bandwidth of K in 2x2 node blocks)."""
    m = mesh.normalised()
    band = C.c_uint64()
    _host_check(_lib.load().mag_mesh_band(m.n_nodes, m.n_elems, _lib.ptr(m.n0), _lib.ptr(m.n1), _lib.ptr(m.n2),
                                          C.byref(band)), "mag_mesh_band")
    return int(band.value)


def rcm(mesh: MeshSoA) -> Tuple[np.ndarray, int, int]:
    """(new_of_old uint32[n_nodes], band_before, band_after) — reverse Cuthill-McKee on the node graph."""
    m = mesh.normalised()
    new_of_old = np.empty(m.n_nodes, np.uint32)
    before, after = C.c_uint64(), C.c_uint64()
    _host_check(_lib.load().mag_reorder_rcm(m.n_nodes, m.n_elems, _lib.ptr(m.n0), _lib.ptr(m.n1), _lib.ptr(m.n2),
                                            _lib.ptr(new_of_old), C.byref(before), C.byref(after)), "mag_reorder_rcm")
    return new_of_old, int(before.value), int(after.value)


def permute_mesh(mesh: MeshSoA, new_of_old: np.ndarray) -> MeshSoA:
    """The same mesh with node i renamed new_of_old[i]; elements keep their order and orientation."""
    m = mesh.normalised()
    p = np.asarray(new_of_old, np.int64)
    if p.shape != (m.n_nodes,) or (m.n_nodes and not np.array_equal(np.sort(p), np.arange(m.n_nodes))):
        raise ValueError("new_of_old is not a permutation of the node ids")

    def move(a):
        out = np.empty_like(a)
        out[p] = a
        return out

    pu = p.astype(np.uint32)
    return MeshSoA(move(m.x), move(m.y), pu[m.n0], pu[m.n1], pu[m.n2], move(m.ux), move(m.uy), move(m.fx),
                   move(m.fy), move(m.known), dict(m.meta))


def unpermute_nodal(values: np.ndarray, new_of_old: np.ndarray) -> np.ndarray:
    """Per-node results of the permuted mesh, back in the original node order."""
    return np.ascontiguousarray(np.asarray(values)[np.asarray(new_of_old, np.int64)])
