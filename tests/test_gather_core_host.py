"""CPU check of the gather assembly's per-node core (magnetite_b200/csrc/gather_core.h — the functions the
CUDA kernels of gather.cuh call, compiled here by g++ into a test harness, tests/native/gather_host_test.cpp):
the block rows it produces equal the oracle's full K (solver.rs:290-331) bit for bit, on every mesh family
the GPU parity tests use, for whole meshes and for the node ranges of a 3-way partition.  The kernels'
launch glue itself is covered by tests/test_gpu_parity.py (default mode, `assembly=0`)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from magnetite_b200 import dist as mdist, meshgen, solver
from magnetite_b200.datatypes import MeshSoA
from oracle import oracle as O

HERE = Path(__file__).resolve().parent
GOLDEN = HERE / "golden"
META = meshgen.EXAMPLE_MATERIAL


@pytest.fixture(scope="module")
def harness(built):
    src = HERE / "native" / "gather_host_test.cpp"
    out = HERE / "native" / "_build" / "libgather_host.so"
    core = HERE.parent / "magnetite_b200" / "csrc" / "gather_core.h"
    out.parent.mkdir(exist_ok=True)
    if not out.exists() or out.stat().st_mtime < max(src.stat().st_mtime, core.stat().st_mtime):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-Wextra", "-shared", "-fPIC",
                        "-o", str(out), str(src)], check=True)
    lib = C.CDLL(str(out))
    vp = C.c_void_p
    lib.gather_host_assemble.restype = C.c_uint64
    lib.gather_host_assemble.argtypes = [C.c_uint64, C.c_uint64, vp, vp, vp, vp, vp, vp, C.c_double, C.c_uint32,
                                         C.c_uint32, C.c_int, vp, vp, vp, C.c_uint64]
    return lib


def _gather_rows(lib, mesh, meta, lo, hi, use_elist):
    """CSR (rowptr, col, val) over the DOF rows [2*lo, 2*hi) from the harness' block rows."""
    m = mesh.normalised()
    D = np.ascontiguousarray(solver.compute_stress_strain_matrix(meta.poisson_ratio, meta.youngs_modulus))
    p = lambda a: a.ctypes.data                                               # noqa: E731
    browptr = np.zeros(hi - lo + 1, np.uint32)
    args = (m.n_nodes, m.n_elems, p(m.x), p(m.y), p(m.n0), p(m.n1), p(m.n2), p(D), float(meta.part_thickness), lo, hi,
            1 if use_elist else 0, p(browptr))
    nb = lib.gather_host_assemble(*args, None, None, 0)
    bcol = np.zeros(max(nb, 1), np.uint32)
    bval = np.full(max(nb, 1) * 4, np.nan)
    assert lib.gather_host_assemble(*args, p(bcol), p(bval), nb) == nb == browptr[-1]
    bcol, bval = bcol[:nb], bval[: 4 * nb].reshape(nb, 2, 2)
    per_node = np.diff(browptr.astype(np.int64))
    rowptr = np.concatenate([[0], np.cumsum(np.repeat(2 * per_node, 2))])
    col = np.empty(4 * nb, np.int32)
    val = np.empty(4 * nb)
    for i in range(hi - lo):                                                  # node row -> two DOF rows
        b0, b1 = int(browptr[i]), int(browptr[i + 1])
        cols = np.stack([2 * bcol[b0:b1], 2 * bcol[b0:b1] + 1], 1).ravel()
        for a in range(2):
            s = int(rowptr[2 * i + a])
            col[s:s + cols.size] = cols
            val[s:s + cols.size] = bval[b0:b1, a, :].ravel()
    return rowptr, col, val


def _oracle_rows(mesh, meta, lo, hi):
    om = O.Mesh(mesh)
    rp, col, val = O.assemble_sparse(om, O.element_stiffness(om, meta))
    s, e = int(rp[2 * lo]), int(rp[2 * hi])
    return rp[2 * lo: 2 * hi + 1] - rp[2 * lo], col[s:e], val[s:e]


def _same(a, b):
    return all(np.array_equal(x, y, equal_nan=(x.dtype.kind == "f")) for x, y in zip(a, b))


def _fan(spokes=40):
    """A hub with `spokes` neighbours (more than kMaxCols = 16: the one-column-at-a-time path) next to a strip
    of ordinary triangles; the hub has the HIGHEST id so its row is built from scattered low ids."""
    ang = np.linspace(0, 2 * np.pi, spokes, endpoint=False)
    x = np.concatenate([3 * np.cos(ang) + 0.1 * np.sin(5 * ang), [0.05]]); y = np.concatenate([3 * np.sin(ang), [-0.02]])
    hub = spokes
    n0 = np.full(spokes, hub); n1 = np.arange(spokes); n2 = (np.arange(spokes) + 1) % spokes
    rng = np.random.default_rng(4)
    order = rng.permutation(spokes)                                           # element order unrelated to the geometry
    n = spokes + 1
    known = np.full(n, 12, np.uint8); known[:2] = 3
    z = np.zeros(n)
    return MeshSoA(x, y, n0[order].astype(np.uint32), n1[order].astype(np.uint32), n2[order].astype(np.uint32),
                   z, z.copy(), z.copy(), z.copy(), known)


def _degenerate():
    """Repeated nodes inside an element (area 0: inf/NaN entries, the same ones in the same places) and an
    element listed twice."""
    m = meshgen.jitter(meshgen.plate(5, 4)).copy()
    m.n1[3] = m.n0[3]                      # corner 0 == corner 1
    m.n2[7] = m.n1[7] = m.n0[7]            # all three corners equal
    m.n0[11], m.n1[11], m.n2[11] = m.n0[10], m.n1[10], m.n2[10]      # duplicate element
    return m


def _example(name):
    g = np.load(GOLDEN / f"{name}.npz")
    return MeshSoA(g["x"], g["y"], g["n0"], g["n1"], g["n2"], g["bc_ux"], g["bc_uy"], g["bc_fx"], g["bc_fy"], g["known"])


def _clockwise():
    m = meshgen.jitter(meshgen.plate(9, 6)).copy()
    m.n1, m.n2 = m.n2.copy(), m.n1.copy()   # every element reversed: negative areas, -0.0 in B
    return m


MESHES = {
    "plate_20x10": lambda: meshgen.plate(20, 10),
    "plate_33x17_h0.3": lambda: meshgen.plate(33, 17, h=0.3),
    "jitter_31x19": lambda: meshgen.jitter(meshgen.plate(31, 19)),
    "perforated_96x48": lambda: meshgen.perforated_plate(96, 48, pitch=16, radius=4),
    "clockwise_9x6": _clockwise,
    "fan_40": _fan,
    "fan_17": lambda: _fan(17),             # 18 columns in the hub row: just past kMaxCols
    "fan_15": lambda: _fan(15),             # 16 columns: the last size the thread-local path takes
    "degenerate": _degenerate,
    "example_linkedin": lambda: _example("example_linkedin"),      # Delaunay mesh in gmsh-like order
    "example_tensile": lambda: _example("example_tensile"),        # all clockwise after check_ccw
}


@pytest.mark.parametrize("name", list(MESHES))
def test_gather_core_matches_the_oracle_bit_for_bit(harness, name):
    mesh = MESHES[name]()
    n = mesh.n_nodes
    want = _oracle_rows(mesh, META, 0, n)
    assert _same(_gather_rows(harness, mesh, META, 0, n, use_elist=False), want)
    assert _same(_gather_rows(harness, mesh, META, 0, n, use_elist=True), want)
    for r in range(3):                                             # the row blocks of a 3-rank run, with their element lists
        lo, hi = mdist.partition_nodes(n, 3, r)
        assert _same(_gather_rows(harness, mesh, META, lo, hi, use_elist=True), _oracle_rows(mesh, META, lo, hi)), (r, lo, hi)


def test_gather_core_edge_cases(harness):
    empty = MeshSoA(*(np.zeros(0, t) for t in (np.float64, np.float64, np.uint32, np.uint32, np.uint32, np.float64,
                                                 np.float64, np.float64, np.float64, np.uint8)))
    rp, col, val = _gather_rows(harness, empty, META, 0, 0, False)
    assert list(rp) == [0] and col.size == 0 and val.size == 0
    # nodes that no element references have empty rows; a rank that owns only such nodes assembles nothing
    m = meshgen.plate(3, 2).copy()
    extra = 5
    pad = lambda a, v: np.concatenate([a, np.full(extra, v, a.dtype)])      # noqa: E731
    m2 = MeshSoA(pad(m.x, 99.0), pad(m.y, 99.0), m.n0, m.n1, m.n2, pad(m.ux, 0), pad(m.uy, 0), pad(m.fx, 0), pad(m.fy, 0),
                 pad(m.known, 12))
    n = m.n_nodes
    assert _same(_gather_rows(harness, m2, META, 0, n + extra, True), _oracle_rows(m2, META, 0, n + extra))
    rp, col, val = _gather_rows(harness, m2, META, n, n + extra, True)
    assert not rp.any() and col.size == 0
    # another material: D and t reach the core as arguments
    meta = META.__class__(210e9, 0.25, 2.0)
    j = meshgen.jitter(meshgen.plate(6, 5))
    assert _same(_gather_rows(harness, j, meta, 0, j.n_nodes, False), _oracle_rows(j, meta, 0, j.n_nodes))
