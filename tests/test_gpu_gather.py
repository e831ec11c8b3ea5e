"""GPU: the two assembly modes of mag_options.assembly leave exactly the same system — the full K, K_ff, the rhs
and the DOF maps bit for bit — hence bit-identical solves, for whole meshes and for the row blocks of the partitioned
path:  0 (default) gather: incidence lists sorted by node, one thread per node row adds recomputed K_e rows;
1 sorted COO keys: K_e per triangle, 9 keys per triangle, stable sort, warp-shuffle segmented reduction.

This file compares GPU paths with each other.  Their link to the ORACLE is transitive and stated here on purpose:
the per-node core the gather kernels call (gather_core.h: count_cols, fill_row, ke_rows) is compiled by g++ into
tests/test_gather_core_host.py and compared there with the oracle bit for bit, and the default mode (0) is what every
test of tests/test_gpu_parity.py runs against the oracle and the golden fixtures (MAGNETITE_B200_TEST_ASSEMBLY=1 runs
that whole suite against the sorted-key mode)."""
from pathlib import Path

import numpy as np
import pytest

from magnetite_b200 import _lib, meshgen, solver
from magnetite_b200.datatypes import MeshSoA

pytestmark = pytest.mark.gpu
META = meshgen.EXAMPLE_MATERIAL
GOLDEN = Path(__file__).resolve().parent / "golden"


def _example(name):
    g = np.load(GOLDEN / f"{name}.npz")
    return MeshSoA(g["x"], g["y"], g["n0"], g["n1"], g["n2"], g["bc_ux"], g["bc_uy"], g["bc_fx"], g["bc_fy"], g["known"])


def _fan(spokes=40):
    """A hub with more neighbours than a thread keeps locally (kMaxCols = 16), clamped on two rim nodes and
    pulled on a third."""
    ang = np.linspace(0, 2 * np.pi, spokes, endpoint=False)
    x = np.concatenate([3 * np.cos(ang) + 0.1 * np.sin(5 * ang), [0.05]]); y = np.concatenate([3 * np.sin(ang), [-0.02]])
    order = np.random.default_rng(4).permutation(spokes)
    n0 = np.full(spokes, spokes)[order]; n1 = np.arange(spokes)[order]; n2 = ((np.arange(spokes) + 1) % spokes)[order]
    n = spokes + 1
    known = np.full(n, 12, np.uint8); known[:2] = 3
    z = np.zeros(n)
    fx = z.copy(); fx[spokes // 2] = 1e6
    return MeshSoA(x, y, n0.astype(np.uint32), n1.astype(np.uint32), n2.astype(np.uint32), z, z.copy(), fx, z.copy(), known)


MESHES = {
    "plate_20x10": lambda: meshgen.plate(20, 10),
    "plate_33x17_h0.3": lambda: meshgen.plate(33, 17, h=0.3),
    "jitter_31x19": lambda: meshgen.jitter(meshgen.plate(31, 19)),
    "perforated_96x48": lambda: meshgen.perforated_plate(96, 48, pitch=16, radius=4),
    "plate_257x65": lambda: meshgen.plate(257, 65),
    "fan_40": _fan,                                      # 41 columns in the hub row: past the thread-local table (16)
    "fan_15": lambda: _fan(15),                          # 16 columns: the table's last size
    "example_linkedin": lambda: _example("example_linkedin"),
    "example_tensile": lambda: _example("example_tensile"),
}


def _bits(a):
    return a.view(np.uint64) if a.dtype == np.float64 else a


@pytest.mark.parametrize("mode", [0])
@pytest.mark.parametrize("name", list(MESHES))
def test_gather_assembly_is_bit_identical_to_the_sorted_key_assembly(ctx, name, mode):
    mesh = MESHES[name]()
    jac = dict(precond=1)                                           # the same solver on both sides
    with solver.System(mesh, META, ctx, options=_lib.default_options(assembly=1)) as A, \
            solver.System(mesh, META, ctx, options=_lib.default_options(assembly=mode)) as G:
        assert (G.n_free, G.nnz, G.nnz_structural) == (A.n_free, A.nnz, A.nnz_structural)
        for a, g in zip(A.export_full(), G.export_full()):
            assert np.array_equal(_bits(a), _bits(g))
        for a, g in zip(A.export_kff(), G.export_kff()):
            assert np.array_equal(_bits(a), _bits(g))
        sa, sg = A.solve(_lib.default_options(assembly=1, **jac)), G.solve(_lib.default_options(assembly=mode, **jac))
    for k in ("ux", "uy", "fx", "fy", "stress"):
        assert np.array_equal(getattr(sa, k), getattr(sg, k)), k
    assert sa.stats["iters"] == sg.stats["iters"]


@pytest.mark.parametrize("mode", [0])
@pytest.mark.parametrize("R", [2, 5])
def test_gather_assembly_on_row_blocks(ctx, R, mode):
    """The partitioned path (element lists, owned node ranges) through R virtual ranks on one GPU."""
    mesh = meshgen.jitter(meshgen.plate(48, 21))
    a = solver.virtual_rank_solve(mesh, META, R, ctx, _lib.default_options(assembly=1, precond=1))
    g = solver.virtual_rank_solve(mesh, META, R, ctx, _lib.default_options(assembly=mode, precond=1))
    for k in ("ux", "uy", "fx", "fy", "stress"):
        assert np.array_equal(getattr(a, k), getattr(g, k)), k
    one = solver.solve_soa(mesh, META, ctx, _lib.default_options(assembly=mode, rel_tol=1e-12))
    tight = solver.virtual_rank_solve(mesh, META, R, ctx, _lib.default_options(assembly=mode, rel_tol=1e-12))
    u1, ur = np.concatenate([one.ux, one.uy]), np.concatenate([tight.ux, tight.uy])
    assert np.linalg.norm(ur - u1) / np.linalg.norm(u1) < 1e-9


@pytest.mark.parametrize("mode", [0])
def test_gather_assembly_through_mag_solve_and_empty_mesh(ctx, mode):
    mesh = meshgen.plate(24, 12)
    a = solver.solve_soa(mesh, META, ctx, _lib.default_options(compat=1, assembly=1))
    g = solver.solve_soa(mesh, META, ctx, _lib.default_options(compat=1, assembly=mode))
    for k in ("ux", "uy", "fx", "fy", "stress"):
        assert np.array_equal(getattr(a, k), getattr(g, k)), k
    empty = MeshSoA(*(np.zeros(0, t) for t in (np.float64, np.float64, np.uint32, np.uint32, np.uint32, np.float64,
                                                 np.float64, np.float64, np.float64, np.uint8)))
    sol = solver.solve_soa(empty, META, ctx, _lib.default_options(assembly=mode))
    assert sol.ux.size == 0 and sol.stress.size == 0 and sol.stats["iters"] == 0
    # an isolated node that no element references has an empty incidence list and keeps an empty matrix row
    iso = meshgen.plate(4, 3).copy()
    iso = MeshSoA(np.append(iso.x, 99.0), np.append(iso.y, 99.0), iso.n0, iso.n1, iso.n2, np.append(iso.ux, 0.0),
                  np.append(iso.uy, 0.0), np.append(iso.fx, 0.0), np.append(iso.fy, 0.0), np.append(iso.known, 3).astype(np.uint8))
    a = solver.solve_soa(iso, META, ctx, _lib.default_options(compat=1, assembly=1))
    g = solver.solve_soa(iso, META, ctx, _lib.default_options(compat=1, assembly=mode))
    for k in ("ux", "uy", "fx", "fy", "stress"):
        assert np.array_equal(getattr(a, k), getattr(g, k)), k
