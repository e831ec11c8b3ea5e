"""Host-side mirror of the reference's `solver` module (src/solver.rs) over the C ABI.

    run(nodes, elements, model_metadata)          <- solver::run            (solver.rs:543-586)
    compute_element_area(element, nodes)          <- solver::compute_element_area (solver.rs:187-193)
    DOF, MAX_CG_ITER, TARGET_CG_COST              <- solver.rs:17-19

`run` has the reference's signature and side effects: it fills ux/uy/fx/fy of
every node and `stress` of every element in place, prints the same "info:" lines,
and raises MagnetiteError (kind "Solver") where the reference returns Err or
panics.  All arithmetic happens on the GPU in libmagnetite_b200.so; this module
only flattens AoS -> SoA and back (what the Rust shim in rust/ does around the
same FFI call).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (MagMaterial, MagMesh, MagOptions, MagResult, MagStats, check, default_options,
                   load, ptr)
from .datatypes import Element, MeshSoA, ModelMetadata, Node
from .error import MagnetiteError

DOF = 2                       # solver.rs:17
MAX_CG_ITER = int(1e7)        # solver.rs:18
TARGET_CG_COST = 1e-4         # solver.rs:19

_default_ctx: Optional[_lib.Context] = None


def default_context() -> _lib.Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = _lib.Context(0)
    return _default_ctx


def _mesh_struct(m: MeshSoA, on_device: bool = False) -> MagMesh:
    s = MagMesh()
    s.n_nodes, s.n_elems = m.n_nodes, m.n_elems
    s.x, s.y = ptr(m.x), ptr(m.y)
    s.n0, s.n1, s.n2 = ptr(m.n0), ptr(m.n1), ptr(m.n2)
    s.ux, s.uy, s.fx, s.fy = ptr(m.ux), ptr(m.uy), ptr(m.fx), ptr(m.fy)
    s.known = ptr(m.known)
    s.on_device = 1 if on_device else 0
    return s


def _material(meta: ModelMetadata) -> MagMaterial:
    return MagMaterial(float(meta.youngs_modulus), float(meta.poisson_ratio), float(meta.part_thickness))


@dataclass
class Solution:
    ux: np.ndarray
    uy: np.ndarray
    fx: np.ndarray
    fy: np.ndarray
    stress: np.ndarray
    sigma: Optional[np.ndarray]
    stats: dict


class System:
    """An assembled system resident on the GPU (mag_system): full K (2x2 BSR), K_ff (CSR +
    SELL-32), rhs and the DOF maps."""

    def __init__(self, mesh: MeshSoA, meta: ModelMetadata, ctx: Optional[_lib.Context] = None,
                 options: Optional[MagOptions] = None, on_device: bool = False):
        self.ctx = ctx or default_context()
        self.mesh = mesh if on_device else mesh.normalised()
        self.meta = meta
        self._h = C.c_void_p()
        st = MagStats()
        ms = _mesh_struct(self.mesh, on_device)
        mat = _material(meta)
        opt = options or default_options()
        check(load().mag_assemble(self.ctx.handle, C.byref(ms), C.byref(mat), C.byref(opt),
                                  C.byref(self._h), C.byref(st)), "mag_assemble")
        self.assemble_stats = st.as_dict()
        self.n_nodes, self.n_elems = int(st.n_nodes), int(st.n_elems)
        self.n_dof, self.n_free = int(st.n_dof), int(st.n_free)
        self.nnz, self.nnz_structural = int(st.nnz), int(st.nnz_structural)

    def close(self):
        if self._h is not None and self._h.value:
            load().mag_system_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- parity exports ------------------------------------------------------
    def export_kff(self):
        """(rowptr int64, col int32, val f64, rhs f64, free_map int64) of K_ff — the matrix the
        reference hands to CG (solver.rs:126-137) and the rhs of solver.rs:427-432."""
        rowptr = np.empty(self.n_free + 1, np.int64)
        col = np.empty(self.nnz, np.int32)
        val = np.empty(self.nnz, np.float64)
        rhs = np.empty(self.n_free, np.float64)
        fmap = np.empty(self.n_dof, np.int64)
        check(load().mag_system_export_kff(self._h, ptr(rowptr), ptr(col), ptr(val), ptr(rhs), ptr(fmap)),
              "mag_system_export_kff")
        return rowptr, col, val, rhs, fmap

    def export_full(self):
        """Structural CSR of the full K (solver.rs:290-331), DOF = 2*node + axis."""
        rowptr = np.empty(self.n_dof + 1, np.int64)
        col = np.empty(self.nnz_structural, np.int32)
        val = np.empty(self.nnz_structural, np.float64)
        check(load().mag_system_export_full(self._h, ptr(rowptr), ptr(col), ptr(val)), "mag_system_export_full")
        return rowptr, col, val

    def spmv(self, x: np.ndarray, fmt: int = 2) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float64)
        if x.shape[0] != self.n_free:
            raise ValueError("x has the wrong length")
        y = np.empty(self.n_free, np.float64)
        check(load().mag_system_spmv(self._h, fmt, ptr(x), ptr(y)), "mag_system_spmv")
        return y

    def true_residual(self, ux: np.ndarray, uy: np.ndarray):
        """(sum (b - K_ff x)^2, sum b^2) over the rows this rank owns for a returned displacement field
        (mag_system_residual: x rebuilt through the free-DOF map, row products in the reference's order)."""
        ux, uy = np.ascontiguousarray(ux, np.float64), np.ascontiguousarray(uy, np.float64)
        rr, bb = C.c_double(), C.c_double()
        check(load().mag_system_residual(self._h, ptr(ux), ptr(uy), 0, C.byref(rr), C.byref(bb)), "mag_system_residual")
        return float(rr.value), float(bb.value)

    def spmv_bench(self, reps: int = 50, fmt: int = 2):
        ms = C.c_float()
        nbytes = C.c_uint64()
        check(load().mag_system_spmv_bench(self._h, fmt, reps, C.byref(ms), C.byref(nbytes)), "mag_system_spmv_bench")
        return float(ms.value), int(nbytes.value)

    # ---- solve -----------------------------------------------------------------
    def solve(self, options: Optional[MagOptions] = None, want_sigma: bool = False,
              out: Optional[dict] = None, allow_not_converged: bool = False) -> Solution:
        opt = options or default_options()
        n, e = self.n_nodes, self.n_elems
        res = MagResult()
        if out is not None:      # caller-provided (possibly device) buffers
            bufs = out
            res.on_device = 1 if out.get("on_device") else 0
        else:
            bufs = {k: np.empty(n, np.float64) for k in ("ux", "uy", "fx", "fy")}
            bufs["stress"] = np.empty(e, np.float64)
            bufs["sigma"] = np.empty((e, 3), np.float64) if want_sigma else None
        res.ux, res.uy, res.fx, res.fy = (ptr(bufs[k]) for k in ("ux", "uy", "fx", "fy"))
        res.stress = ptr(bufs["stress"])
        res.sigma = ptr(bufs.get("sigma"))
        st = MagStats()
        allow = (_lib.MAG_ERR_NOT_CONVERGED,) if allow_not_converged else ()
        check(load().mag_system_solve(self._h, C.byref(opt), C.byref(res), C.byref(st)), "mag_system_solve", allow)
        return Solution(bufs["ux"], bufs["uy"], bufs["fx"], bufs["fy"], bufs["stress"], bufs.get("sigma"),
                        st.as_dict())


def solve_soa(mesh: MeshSoA, meta: ModelMetadata, ctx: Optional[_lib.Context] = None,
              options: Optional[MagOptions] = None, want_sigma: bool = False, reorder: bool = False) -> Solution:
    """One call through mag_solve with host buffers (the end-to-end path the Rust shim takes).

    `reorder=True` renumbers the nodes with reverse Cuthill-McKee first when that narrows the band
    (meshes in gmsh order) and returns the results in the ORIGINAL numbering; see
    reorder.py for what that does and does not change."""
    if reorder:
        from . import reorder as R
        new_of_old, before, after = R.rcm(mesh)
        info = {"applied": after < before, "band_before": before, "band_after": after}
        if not info["applied"]:
            sol = solve_soa(mesh, meta, ctx, options, want_sigma)
        else:
            sol = solve_soa(R.permute_mesh(mesh, new_of_old), meta, ctx, options, want_sigma)
            for k in ("ux", "uy", "fx", "fy"):
                setattr(sol, k, R.unpermute_nodal(getattr(sol, k), new_of_old))
        sol.stats["reorder"] = info
        return sol
    ctx = ctx or default_context()
    m = mesh.normalised()
    n, e = m.n_nodes, m.n_elems
    ms, mat = _mesh_struct(m), _material(meta)
    opt = options or default_options()
    bufs = {k: np.empty(n, np.float64) for k in ("ux", "uy", "fx", "fy")}
    stress = np.empty(e, np.float64)
    sigma = np.empty((e, 3), np.float64) if want_sigma else None
    res = MagResult(ptr(bufs["ux"]), ptr(bufs["uy"]), ptr(bufs["fx"]), ptr(bufs["fy"]), ptr(stress),
                    ptr(sigma), 0)
    st = MagStats()
    check(load().mag_solve(ctx.handle, C.byref(ms), C.byref(mat), C.byref(opt), C.byref(res), C.byref(st)),
          "mag_solve")
    return Solution(bufs["ux"], bufs["uy"], bufs["fx"], bufs["fy"], stress, sigma, st.as_dict())


def virtual_rank_solve(mesh: MeshSoA, meta: ModelMetadata, nranks: int, ctx: Optional[_lib.Context] = None,
                       options: Optional[MagOptions] = None) -> Solution:
    """An `nranks`-way row-block partitioned solve emulated on ONE GPU (mag_debug_virtual_solve):
    the partition, kernels and peer halo stores of the multi-GPU path, with the allreduce done by
    a kernel.  Test hook."""
    ctx = ctx or default_context()
    m = mesh.normalised()
    n, e = m.n_nodes, m.n_elems
    ms, mat = _mesh_struct(m), _material(meta)
    opt = options or default_options()
    bufs = {k: np.empty(n, np.float64) for k in ("ux", "uy", "fx", "fy")}
    stress = np.empty(e, np.float64)
    res = MagResult(ptr(bufs["ux"]), ptr(bufs["uy"]), ptr(bufs["fx"]), ptr(bufs["fy"]), ptr(stress), None, 0)
    st = MagStats()
    check(load().mag_debug_virtual_solve(ctx.handle, C.byref(ms), C.byref(mat), C.byref(opt), nranks,
                                         C.byref(res), C.byref(st)), "mag_debug_virtual_solve")
    return Solution(bufs["ux"], bufs["uy"], bufs["fx"], bufs["fy"], stress, None, st.as_dict())


def element_stiffness(mesh: MeshSoA, meta: ModelMetadata, ctx: Optional[_lib.Context] = None) -> np.ndarray:
    """K_e of every element, (E, 6, 6) row-major (solver.rs:263-278)."""
    ctx = ctx or default_context()
    m = mesh.normalised()
    ms, mat = _mesh_struct(m), _material(meta)
    ke = np.empty((m.n_elems, 6, 6), np.float64)
    check(load().mag_element_stiffness(ctx.handle, C.byref(ms), C.byref(mat), ptr(ke)), "mag_element_stiffness")
    return ke


def element_areas(mesh: MeshSoA, ctx: Optional[_lib.Context] = None) -> np.ndarray:
    """Signed areas of all elements in one launch (solver.rs:187-193)."""
    ctx = ctx or default_context()
    m = mesh.normalised()
    ms = _mesh_struct(m)
    area = np.empty(m.n_elems, np.float64)
    check(load().mag_element_area(ctx.handle, C.byref(ms), ptr(area)), "mag_element_area")
    return area


def stress_soa(mesh: MeshSoA, meta: ModelMetadata, ux, uy, want_sigma=False, ctx=None):
    """Stress recovery only (solver.rs:496-535)."""
    ctx = ctx or default_context()
    m = mesh.normalised()
    ms, mat = _mesh_struct(m), _material(meta)
    ux = np.ascontiguousarray(ux, np.float64); uy = np.ascontiguousarray(uy, np.float64)
    s = np.empty(m.n_elems, np.float64)
    sig = np.empty((m.n_elems, 3), np.float64) if want_sigma else None
    check(load().mag_stress(ctx.handle, C.byref(ms), C.byref(mat), ptr(ux), ptr(uy), ptr(s), ptr(sig)), "mag_stress")
    return (s, sig) if want_sigma else s


def strain_displacement_matrices(mesh: MeshSoA, ctx: Optional[_lib.Context] = None) -> np.ndarray:
    """B of every element, (E, 3, 6) (solver.rs:204-230), in one launch."""
    ctx = ctx or default_context()
    m = mesh.normalised()
    ms = _mesh_struct(m)
    b = np.empty((m.n_elems, 3, 6), np.float64)
    check(load().mag_strain_displacement(ctx.handle, C.byref(ms), ptr(b)), "mag_strain_displacement")
    return b


# ---------------------------------------------------------------------------
# the reference's public functions
# ---------------------------------------------------------------------------
def compute_stress_strain_matrix(poisson_ratio: float, youngs_modulus: float) -> np.ndarray:
    """solver::compute_stress_strain_matrix (solver.rs:240-250): plane-stress D, 3x3."""
    d = np.empty(9, np.float64)
    check(load().mag_stress_strain(float(poisson_ratio), float(youngs_modulus), ptr(d)), "mag_stress_strain")
    return d.reshape(3, 3)


def compute_strain_displacement_matrix(element: Element, nodes: Sequence[Node], element_area: float = None) -> np.ndarray:
    """solver::compute_strain_displacement_matrix (solver.rs:204-230): B of one element, 3x6.  The
    reference takes the (signed) area as an argument; here it is recomputed on the device."""
    tri = [nodes[i] for i in element.nodes]
    return strain_displacement_matrices(MeshSoA.from_aos(tri, [Element([0, 1, 2])]))[0]


def compute_element_area(element: Element, nodes: Sequence[Node]) -> float:
    """solver::compute_element_area (solver.rs:187-193) — signed area of one element.
    (mesher.check_ccw uses the batched `element_areas` instead of calling this per element.)"""
    tri = [nodes[i] for i in element.nodes]
    m = MeshSoA.from_aos(tri, [Element([0, 1, 2])])
    return float(element_areas(m)[0])


def run(nodes: List[Node], elements: List[Element], model_metadata: ModelMetadata,
        options: Optional[MagOptions] = None, quiet: bool = False, reorder: bool = False) -> None:
    """solver::run (solver.rs:543-586): fills node.ux/uy/fx/fy and element.stress in place.

    Default options reproduce the reference's solver semantics (plain CG from x0 = 0, absolute
    cost 1e-4, 1e7 iterations: `compat=1`).  Pass options for the north-star Jacobi-PCG, and
    `reorder=True` for meshes in gmsh order (results stay in the caller's numbering).
    """
    say = (lambda *_: None) if quiet else print
    mesh = MeshSoA.from_aos(nodes, elements)
    opt = options or default_options(compat=1)
    say("info: building element stiffness matrices...")           # solver.rs:551
    say("info: building total stiffness matrix...")               # solver.rs:570
    say("info: setting up system...")                             # solver.rs:416
    say("info: solving...")                                       # solver.rs:437
    sol = solve_soa(mesh, model_metadata, options=opt, reorder=reorder)
    st = sol.stats
    say(f"info: finished conjugate gradient approximation in {st['iters']} iterations")  # :101-104
    say("info: solved system in {:.3f} seconds".format(st["ms_solve"] / 1e3))            # :441
    say("info: solve complete")                                   # solver.rs:484
    for i, nd in enumerate(nodes):                                # solver.rs:476-482
        nd.ux, nd.uy = float(sol.ux[i]), float(sol.uy[i])
        nd.fx, nd.fy = float(sol.fx[i]), float(sol.fy[i])
    for i, el in enumerate(elements):                             # solver.rs:532-533
        el.stress = float(sol.stress[i])
