"""ctypes binding of libmagnetite_b200.so (include/magnetite_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device
is visible, every compute entry point raises.  Nothing here imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from .error import MagnetiteError

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libmagnetite_b200.so"

ABI_VERSION = 3
MAG_OK = 0
MAG_ERR_CUDA, MAG_ERR_OOM, MAG_ERR_BAD_BC, MAG_ERR_BAD_INDEX = -1, -2, -3, -4
MAG_ERR_INDEFINITE, MAG_ERR_NOT_CONVERGED, MAG_ERR_NCCL, MAG_ERR_BAD_ARG = -5, -6, -7, -8

c_double_p = C.POINTER(C.c_double)
c_u32_p = C.POINTER(C.c_uint32)
c_u8_p = C.POINTER(C.c_uint8)


class MagMesh(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("n_elems", C.c_uint64),
                ("x", C.c_void_p), ("y", C.c_void_p),
                ("n0", C.c_void_p), ("n1", C.c_void_p), ("n2", C.c_void_p),
                ("ux", C.c_void_p), ("uy", C.c_void_p), ("fx", C.c_void_p), ("fy", C.c_void_p),
                ("known", C.c_void_p), ("on_device", C.c_int32)]


class MagMaterial(C.Structure):
    _fields_ = [("youngs_modulus", C.c_double), ("poisson_ratio", C.c_double),
                ("part_thickness", C.c_double)]


class MagOptions(C.Structure):
    _fields_ = [("rel_tol", C.c_double), ("abs_tol", C.c_double), ("max_iter", C.c_uint64),
                ("precond", C.c_int32), ("compat", C.c_int32), ("cost_kind", C.c_int32),
                ("drop_exact_zeros", C.c_int32), ("check_every", C.c_int32),
                ("spmv_format", C.c_int32), ("want_sigma", C.c_int32), ("allreduce", C.c_int32),
                ("coarse_aggregates", C.c_int32), ("assembly", C.c_int32),
                ("result_scope", C.c_int32), ("reserved0", C.c_int32),
                ("stream", C.c_void_p)]


class MagResult(C.Structure):
    _fields_ = [("ux", C.c_void_p), ("uy", C.c_void_p), ("fx", C.c_void_p), ("fy", C.c_void_p),
                ("stress", C.c_void_p), ("sigma", C.c_void_p), ("on_device", C.c_int32)]


class MagStats(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("n_elems", C.c_uint64), ("n_dof", C.c_uint64),
                ("n_free", C.c_uint64), ("n_constrained", C.c_uint64),
                ("nnz_structural", C.c_uint64), ("nnz", C.c_uint64), ("sell_entries", C.c_uint64),
                ("iters", C.c_uint64), ("final_residual", C.c_double), ("b_norm", C.c_double),
                ("converged", C.c_int32), ("negative_definite", C.c_int32),
                ("ms_upload", C.c_float), ("ms_elem", C.c_float), ("ms_sort", C.c_float),
                ("ms_reduce", C.c_float), ("ms_bc", C.c_float), ("ms_format", C.c_float),
                ("ms_solve", C.c_float), ("ms_post", C.c_float), ("ms_download", C.c_float),
                ("ms_total", C.c_float), ("kernel_launches", C.c_uint64),
                ("spmv_bytes", C.c_uint64), ("prof", C.c_double * 8),
                ("ms_coarse_setup", C.c_float), ("n_coarse", C.c_uint32),
                ("sell_index_bits", C.c_uint32), ("precond_used", C.c_uint32)]

    def as_dict(self) -> dict:
        d = {name: getattr(self, name) for name, _ in self._fields_}
        d["prof"] = list(self.prof)
        return d


# every symbol include/magnetite_b200.h declares: (restype, argtypes)
_vp = C.c_void_p
_SIGNATURES = {
    "mag_abi_version": (C.c_int, []),
    "mag_last_error": (C.c_char_p, []),
    "mag_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "mag_ctx_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "mag_ctx_destroy": (None, [_vp]),
    "mag_options_default": (None, [C.POINTER(MagOptions)]),
    "mag_solve": (C.c_int, [_vp, C.POINTER(MagMesh), C.POINTER(MagMaterial), C.POINTER(MagOptions),
                            C.POINTER(MagResult), C.POINTER(MagStats)]),
    "mag_assemble": (C.c_int, [_vp, C.POINTER(MagMesh), C.POINTER(MagMaterial), C.POINTER(MagOptions),
                               C.POINTER(_vp), C.POINTER(MagStats)]),
    "mag_system_solve": (C.c_int, [_vp, C.POINTER(MagOptions), C.POINTER(MagResult), C.POINTER(MagStats)]),
    "mag_system_free": (None, [_vp]),
    "mag_system_info": (C.c_int, [_vp, C.POINTER(MagStats)]),
    "mag_system_export_kff": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mag_system_export_full": (C.c_int, [_vp, _vp, _vp, _vp]),
    "mag_element_stiffness": (C.c_int, [_vp, C.POINTER(MagMesh), C.POINTER(MagMaterial), _vp]),
    "mag_element_area": (C.c_int, [_vp, C.POINTER(MagMesh), _vp]),
    "mag_strain_displacement": (C.c_int, [_vp, C.POINTER(MagMesh), _vp]),
    "mag_stress_strain": (C.c_int, [C.c_double, C.c_double, _vp]),
    "mag_stress": (C.c_int, [_vp, C.POINTER(MagMesh), C.POINTER(MagMaterial), _vp, _vp, _vp, _vp]),
    "mag_system_spmv": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "mag_system_spmv_bench": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
    "mag_system_residual": (C.c_int, [_vp, _vp, _vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mag_csv_output": (C.c_int, [C.c_char_p, C.c_char_p, C.c_uint64, _vp, _vp, _vp, _vp, C.c_uint64, _vp, _vp, _vp, _vp]),
    "mag_format_f64": (C.c_size_t, [C.c_double, C.c_char_p]),
    "mag_host_last_error": (C.c_char_p, []),
    "mag_reorder_rcm": (C.c_int, [C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "mag_mesh_band": (C.c_int, [C.c_uint64, C.c_uint64, _vp, _vp, _vp, C.POINTER(C.c_uint64)]),
    "mag_devmesh_plate": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.POINTER(_vp)]),
    "mag_devmesh_perforated": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_double, C.c_uint32, C.c_uint32, C.c_double,
                                         C.POINTER(_vp)]),
    "mag_devmesh_download": (C.c_int, [_vp] + [_vp] * 10),
    "mag_devmesh_view": (C.c_int, [_vp, C.POINTER(MagMesh)]),
    "mag_devmesh_free": (None, [_vp]),
    "mag_comm_unique_id": (C.c_int, [_vp]),
    "mag_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "mag_comm_rank": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mag_partition_nodes": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "mag_halo_plan": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int32, C.POINTER(C.c_int32)]),
    "mag_debug_sort_pairs": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_int]),
    "mag_debug_exclusive_scan": (C.c_int, [_vp, _vp, _vp, C.c_uint64]),
    "mag_debug_virtual_solve": (C.c_int, [_vp, C.POINTER(MagMesh), C.POINTER(MagMaterial), C.POINTER(MagOptions),
                                          C.c_int, C.POINTER(MagResult), C.POINTER(MagStats)]),
}

_lib = None


def _preload_nccl():
    """libmagnetite_b200.so needs libnccl.so.2; prefer the copy torch bundles (already mapped if
    torch was imported), else whatever the loader finds."""
    try:
        import nvidia.nccl  # type: ignore
        for base in nvidia.nccl.__path__:
            cand = Path(base) / "lib" / "libnccl.so.2"
            if cand.exists():
                C.CDLL(str(cand), mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("MAGNETITE_B200_LIB", LIB_PATH))
    if not path.exists():
        raise MagnetiteError.Solver(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C magnetite_b200/csrc` (there is no CPU fallback)")
    _preload_nccl()
    lib = C.CDLL(str(path))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library drift
        fn.restype = res
        fn.argtypes = args
    if lib.mag_abi_version() != ABI_VERSION:
        raise MagnetiteError.Solver(f"{path} has ABI version {lib.mag_abi_version()}, this binding expects {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def declared_symbols():
    return list(_SIGNATURES)


def last_error() -> str:
    return (load().mag_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str = "", allow=()):
    if rc == MAG_OK or rc in allow:
        return rc
    msg = last_error()
    raise MagnetiteError.Solver(f"{what}: {msg}" if what else msg, code=rc)


def ptr(a):
    """Address of a numpy array, a torch tensor, a raw int address, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(f"cannot take the address of {type(a)!r}")


class Context:
    """One mag_ctx (device + stream + memory pool).  Fails loudly without a GPU."""

    def __init__(self, device: int = 0):
        lib = load()
        h = _vp()
        check(lib.mag_ctx_create(C.byref(h), device), "mag_ctx_create")
        self._h = h
        self.device = device

    @property
    def handle(self):
        if self._h is None:
            raise MagnetiteError.Solver("context already destroyed")
        return self._h

    def close(self):
        if self._h is not None:
            load().mag_ctx_destroy(self._h)
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


def default_options(**overrides) -> MagOptions:
    o = MagOptions()
    load().mag_options_default(C.byref(o))
    for k, v in overrides.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown solver option {k!r}")
        setattr(o, k, v)
    return o
