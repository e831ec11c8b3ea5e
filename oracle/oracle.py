"""ctypes wrapper of the C restatement (oracle/magnetite_oracle.c).

TEST INFRASTRUCTURE ONLY — PARITY UNPINNED (see magnetite_oracle.h).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "_build" / "libmagnetite_oracle.so"

KNOWN_UX, KNOWN_UY, KNOWN_FX, KNOWN_FY = 1, 2, 4, 8
COST_NORM, COST_SQ = 0, 1


class OrcMesh(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("n_elems", C.c_uint64), ("x", C.c_void_p), ("y", C.c_void_p),
                ("n0", C.c_void_p), ("n1", C.c_void_p), ("n2", C.c_void_p), ("ux", C.c_void_p),
                ("uy", C.c_void_p), ("fx", C.c_void_p), ("fy", C.c_void_p), ("known", C.c_void_p)]


class OrcMaterial(C.Structure):
    _fields_ = [("youngs_modulus", C.c_double), ("poisson_ratio", C.c_double), ("part_thickness", C.c_double)]


class OrcCsr(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("n_cols", C.c_uint64), ("nnz", C.c_uint64),
                ("rowptr", C.POINTER(C.c_int64)), ("col", C.POINTER(C.c_int32)), ("val", C.POINTER(C.c_double))]


class OrcCgOptions(C.Structure):
    _fields_ = [("max_iter", C.c_uint64), ("target_cost", C.c_double), ("cost_kind", C.c_int),
                ("jacobi", C.c_int), ("rel_tol", C.c_double)]


class OrcStats(C.Structure):
    _fields_ = [("iters", C.c_uint64), ("final_cost", C.c_double), ("b_norm", C.c_double),
                ("n_free", C.c_uint64), ("n_constrained", C.c_uint64), ("nnz_ff", C.c_uint64),
                ("nnz_structural", C.c_uint64), ("t_elem", C.c_double), ("t_asm", C.c_double),
                ("t_part", C.c_double), ("t_solve", C.c_double), ("t_react", C.c_double),
                ("t_stress", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class OrcResult(C.Structure):
    _fields_ = [("ux", C.c_void_p), ("uy", C.c_void_p), ("fx", C.c_void_p), ("fy", C.c_void_p),
                ("stress", C.c_void_p)]


_lib = None


def build(force: bool = False):
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < (_HERE / "magnetite_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(LIB_PATH))
        lib.orc_element_area.restype = C.c_double
        lib.orc_element_area.argtypes = [C.POINTER(OrcMesh), C.c_uint64]
        lib.orc_last_error.restype = C.c_char_p
        _lib = lib
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def _check(rc):
    if rc != 0:
        raise OracleError(rc, (load().orc_last_error() or b"").decode())


def _p(a):
    return None if a is None else a.ctypes.data


def cg_options(max_iter=10_000_000, target_cost=1e-4, cost_kind=COST_NORM, jacobi=0, rel_tol=0.0):
    return OrcCgOptions(int(max_iter), float(target_cost), int(cost_kind), int(jacobi), float(rel_tol))


class Mesh:
    """Keeps the numpy arrays alive next to the C struct.  `m` is any object with x,y,n0,n1,n2,
    ux,uy,fx,fy,known arrays (magnetite_b200.datatypes.MeshSoA works)."""

    def __init__(self, m):
        c = np.ascontiguousarray
        self.x, self.y = c(m.x, np.float64), c(m.y, np.float64)
        self.n0, self.n1, self.n2 = c(m.n0, np.uint32), c(m.n1, np.uint32), c(m.n2, np.uint32)
        self.ux, self.uy = c(m.ux, np.float64), c(m.uy, np.float64)
        self.fx, self.fy = c(m.fx, np.float64), c(m.fy, np.float64)
        self.known = c(m.known, np.uint8)
        self.n_nodes, self.n_elems = self.x.shape[0], self.n0.shape[0]
        self.c = OrcMesh(self.n_nodes, self.n_elems, _p(self.x), _p(self.y), _p(self.n0), _p(self.n1),
                         _p(self.n2), _p(self.ux), _p(self.uy), _p(self.fx), _p(self.fy), _p(self.known))


def material(meta):
    return OrcMaterial(float(meta.youngs_modulus), float(meta.poisson_ratio), float(meta.part_thickness))


def _csr_to_numpy(A: OrcCsr):
    n, nnz = int(A.n_rows), int(A.nnz)
    rowptr = np.ctypeslib.as_array(A.rowptr, shape=(n + 1,)).copy()
    col = np.ctypeslib.as_array(A.col, shape=(max(nnz, 1),))[:nnz].copy()
    val = np.ctypeslib.as_array(A.val, shape=(max(nnz, 1),))[:nnz].copy()
    return rowptr, col, val


def element_area(m: Mesh):
    lib = load()
    return np.array([lib.orc_element_area(C.byref(m.c), e) for e in range(m.n_elems)])


def element_stiffness(m: Mesh, meta) -> np.ndarray:
    ke = np.empty((m.n_elems, 6, 6), np.float64)
    mat = material(meta)
    load().orc_element_stiffness(C.byref(m.c), C.byref(mat), C.c_void_p(_p(ke)))
    return ke


def stress_strain(nu, E):
    d = np.empty(9)
    load().orc_stress_strain(C.c_double(nu), C.c_double(E), C.c_void_p(_p(d)))
    return d.reshape(3, 3)


def assemble_dense(m: Mesh, ke: np.ndarray) -> np.ndarray:
    n = 2 * m.n_nodes
    K = np.empty((n, n), np.float64, order="F")      # column-major like nalgebra's DMatrix
    load().orc_assemble_dense(C.byref(m.c), C.c_void_p(_p(np.ascontiguousarray(ke))), C.c_void_p(K.ctypes.data))
    return K


def assemble_sparse(m: Mesh, ke: np.ndarray):
    """Structural CSR of the full K: (rowptr, col, val)."""
    A = OrcCsr()
    _check(load().orc_assemble_sparse(C.byref(m.c), C.c_void_p(_p(np.ascontiguousarray(ke))), C.byref(A)))
    out = _csr_to_numpy(A)
    load().orc_csr_free(C.byref(A))
    return out


def _numpy_to_csr(rowptr, col, val, n_cols):
    rowptr = np.ascontiguousarray(rowptr, np.int64); col = np.ascontiguousarray(col, np.int32)
    val = np.ascontiguousarray(val, np.float64)
    A = OrcCsr(rowptr.shape[0] - 1, n_cols, col.shape[0],
               rowptr.ctypes.data_as(C.POINTER(C.c_int64)), col.ctypes.data_as(C.POINTER(C.c_int32)),
               val.ctypes.data_as(C.POINTER(C.c_double)))
    return A, (rowptr, col, val)


def partition(m: Mesh, K, dense: bool):
    """K_ff CSR, rhs, free_map.  K is the dense matrix (dense=True) or the (rowptr,col,val) triple."""
    n = 2 * m.n_nodes
    rhs = np.empty(n, np.float64)
    fmap = np.empty(n, np.int64)
    nf = C.c_uint64()
    Kff = OrcCsr()
    if dense:
        _check(load().orc_partition_dense(C.byref(m.c), C.c_void_p(K.ctypes.data), C.byref(Kff),
                                          C.c_void_p(_p(rhs)), C.c_void_p(_p(fmap)), C.byref(nf)))
    else:
        A, keep = _numpy_to_csr(*K, n)
        _check(load().orc_partition_sparse(C.byref(m.c), C.byref(A), C.byref(Kff), C.c_void_p(_p(rhs)),
                                           C.c_void_p(_p(fmap)), C.byref(nf)))
    out = _csr_to_numpy(Kff)
    load().orc_csr_free(C.byref(Kff))
    return out, rhs[: nf.value].copy(), fmap


def cg(csr, b, opt: OrcCgOptions):
    rowptr, col, val = csr
    n = rowptr.shape[0] - 1
    A, keep = _numpy_to_csr(rowptr, col, val, n)
    b = np.ascontiguousarray(b, np.float64)
    x = np.zeros(n)
    it = C.c_uint64(); cost = C.c_double()
    _check(load().orc_cg(C.byref(A), C.c_void_p(_p(b)), C.c_void_p(_p(x)), C.byref(opt), C.byref(it), C.byref(cost)))
    return x, int(it.value), float(cost.value)


def spmv(csr, x):
    rowptr, col, val = csr
    n = rowptr.shape[0] - 1
    A, keep = _numpy_to_csr(rowptr, col, val, n)
    x = np.ascontiguousarray(x, np.float64)
    y = np.empty(n)
    load().orc_spmv(C.byref(A), C.c_void_p(_p(x)), C.c_void_p(_p(y)))
    return y


def stress(m: Mesh, meta, ux, ux_y, want_sigma=False):
    mat = material(meta)
    ux = np.ascontiguousarray(ux, np.float64); uy = np.ascontiguousarray(ux_y, np.float64)
    s = np.empty(m.n_elems); sig = np.empty((m.n_elems, 3)) if want_sigma else None
    load().orc_stress(C.byref(m.c), C.byref(mat), C.c_void_p(_p(ux)), C.c_void_p(_p(uy)), C.c_void_p(_p(s)),
                      C.c_void_p(_p(sig)))
    return (s, sig) if want_sigma else s


def run(m: Mesh, meta, opt: OrcCgOptions = None, dense: bool = False):
    """solver::run restated: returns dict(ux,uy,fx,fy,stress,stats)."""
    opt = opt or cg_options()
    mat = material(meta)
    n, e = m.n_nodes, m.n_elems
    out = {k: np.empty(n) for k in ("ux", "uy", "fx", "fy")}
    out["stress"] = np.empty(e)
    res = OrcResult(_p(out["ux"]), _p(out["uy"]), _p(out["fx"]), _p(out["fy"]), _p(out["stress"]))
    st = OrcStats()
    _check(load().orc_run(C.byref(m.c), C.byref(mat), C.byref(opt), 1 if dense else 0, C.byref(res), C.byref(st)))
    out["stats"] = st.as_dict()
    return out
