"""All-host-cores Jacobi-PCG on the oracle's K_ff (oracle_mt.c) — bench/test infrastructure, information only.

Not a parity artefact and not the reference's algorithm (the reference is single-threaded): bench.py reports it
beside the 1-thread port so the CPU side of the comparison is not understated.  Built on demand."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "libmagnetite_oracle_mt.so"
_lib = None


def load():
    global _lib
    if _lib is None:
        src = HERE / "oracle_mt.c"
        if not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
            subprocess.run(["make", "-C", str(HERE), "mt"], check=True, capture_output=True)
        lib = C.CDLL(str(LIB))
        lib.orc_mt_host_cores.restype = C.c_int
        lib.orc_mt_pcg.restype = C.c_int
        lib.orc_mt_pcg.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                   C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        _lib = lib
    return _lib


def host_cores() -> int:
    """Cores this process may use (the affinity mask when the platform has one)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return int(load().orc_mt_host_cores())


def pcg(csr, b, rel_tol=1e-9, max_iter=10_000_000, threads=0):
    """(x, iterations, final ||r||_2) for the CSR triple (rowptr int64, col int32, val f64)."""
    rowptr, col, val = (np.ascontiguousarray(a, t) for a, t in zip(csr, (np.int64, np.int32, np.float64)))
    b = np.ascontiguousarray(b, np.float64)
    n = b.shape[0]
    x = np.empty(n, np.float64)
    it, res = C.c_uint64(), C.c_double()
    rc = load().orc_mt_pcg(n, rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, b.ctypes.data, x.ctypes.data,
                           float(rel_tol), int(max_iter), int(threads if threads > 0 else host_cores()), C.byref(it), C.byref(res))
    if rc != 0:
        raise RuntimeError(f"orc_mt_pcg failed with code {rc}")
    return x, int(it.value), float(res.value)
