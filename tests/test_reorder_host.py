"""Host-side node renumbering (mag_reorder_rcm, csrc/reorder.cpp; magnetite_b200/reorder.py) — SURVEY §8(e):
meshes in gmsh order get a bandwidth-reducing permutation before the row blocks are cut, results come
back in the caller's numbering.  CPU only: the oracle stands in for the GPU solver here, so what is
checked is the permutation logic itself; tests/test_gpu_parity.py runs the same through mag_solve."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from magnetite_b200 import _lib, dist as mdist, meshgen, reorder
from magnetite_b200.datatypes import MeshSoA
from magnetite_b200.error import MagnetiteError
from oracle import oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden"


def _example(name):
    g = np.load(GOLDEN / f"{name}.npz")
    mesh = MeshSoA(g["x"], g["y"], g["n0"], g["n1"], g["n2"], g["bc_ux"], g["bc_uy"], g["bc_fx"], g["bc_fy"], g["known"])
    return g, mesh, meshgen.EXAMPLE_MATERIAL.__class__(*g["material"])


def _shuffled(mesh, seed=3):
    p = np.random.default_rng(seed).permutation(mesh.n_nodes).astype(np.uint32)
    return reorder.permute_mesh(mesh, p), p


def test_rcm_is_a_permutation_and_restores_the_plate_band(built):
    nx, ny = 60, 30
    mesh, _ = _shuffled(meshgen.plate(nx, ny))
    p, before, after = reorder.rcm(mesh)
    assert p.dtype == np.uint32 and np.array_equal(np.sort(p), np.arange(mesh.n_nodes))
    assert before > mesh.n_nodes // 2                      # a random numbering has no band
    assert after <= min(nx, ny) + 2                        # the plate numbered along its short side
    assert reorder.mesh_band(reorder.permute_mesh(mesh, p)) == after == 31
    p2, b2, a2 = reorder.rcm(mesh)                         # deterministic
    assert np.array_equal(p, p2) and (b2, a2) == (before, after)


def test_rcm_is_at_least_as_tight_as_scipy_on_the_example_meshes(built):
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    for name, limit in (("example_tensile", 60), ("example_linkedin", 110), ("example_cover", 60)):
        _, m, _ = _example(name)
        p, before, after = reorder.rcm(m)
        r = np.concatenate([m.n0, m.n1, m.n2]).astype(np.int64); c = np.concatenate([m.n1, m.n2, m.n0]).astype(np.int64)
        A = sp.csr_matrix((np.ones(2 * len(r)), (np.concatenate([r, c]), np.concatenate([c, r]))), shape=(m.n_nodes,) * 2)
        perm = reverse_cuthill_mckee(A, symmetric_mode=True)
        inv = np.empty(m.n_nodes, np.int64); inv[perm] = np.arange(m.n_nodes)
        scipy_band = reorder.mesh_band(reorder.permute_mesh(m, inv))
        assert before > m.n_nodes * 0.9 and after <= limit and after <= scipy_band * 1.25, (name, before, after, scipy_band)


def test_permute_mesh_keeps_elements_and_moves_node_payloads(built):
    mesh = meshgen.jitter(meshgen.plate(7, 5))
    pm, p = _shuffled(mesh, seed=9)
    assert np.array_equal(O.element_area(O.Mesh(pm)), O.element_area(O.Mesh(mesh)))      # same triangles, same orientation
    for k in ("x", "y", "ux", "uy", "fx", "fy", "known"):
        assert np.array_equal(reorder.unpermute_nodal(getattr(pm, k), p), getattr(mesh, k)), k
    for k in ("n0", "n1", "n2"):
        assert np.array_equal(getattr(pm, k), p[getattr(mesh, k)])
    with pytest.raises(ValueError):
        reorder.permute_mesh(mesh, np.zeros(mesh.n_nodes, np.uint32))
    with pytest.raises(ValueError):
        reorder.permute_mesh(mesh, p[:-1])


@pytest.mark.parametrize("name", ["example_tensile", "example_linkedin", "example_cover"])
def test_reordered_solve_matches_the_golden_solve_in_the_original_numbering(built, name):
    """solve(permuted mesh), un-permuted == the committed golden solve of the mesh as the mesher numbered it,
    within the north-star tolerances (u 1e-9 relative L2, stress 1e-8)."""
    g, mesh, meta = _example(name)
    p, before, after = reorder.rcm(mesh)
    res = O.run(O.Mesh(reorder.permute_mesh(mesh, p)), meta, O.cg_options(), dense=False)
    assert res["stats"]["nnz_ff"] == int(g["nnz_ff"][0])            # same matrix up to a symmetric permutation
    u = np.concatenate([reorder.unpermute_nodal(res["ux"], p), reorder.unpermute_nodal(res["uy"], p)])
    u_ref = np.concatenate([g["ux"], g["uy"]])
    assert np.linalg.norm(u - u_ref) / np.linalg.norm(u_ref) < 1e-9
    assert np.abs(res["stress"] - g["stress"]).max() / np.abs(g["stress"]).max() < 1e-8      # element order untouched
    f = np.concatenate([reorder.unpermute_nodal(res["fx"], p), reorder.unpermute_nodal(res["fy"], p)])
    f_ref = np.concatenate([g["fx"], g["fy"]])
    assert np.abs(f - f_ref).max() / np.abs(f_ref).max() < 1e-7


def _halo_nodes(m, nranks):
    """Nodes every rank reads from other ranks when rows are cut into contiguous node blocks (the
    column extent of a block, as mag_halo_plan receives it, minus the block itself)."""
    n = m.n_nodes
    cuts = [mdist.partition_nodes(n, nranks, r)[0] for r in range(nranks)] + [n]
    conn = np.stack([m.n0, m.n1, m.n2], 1).astype(np.int64)
    lo, hi = np.full(n, n), np.zeros(n, np.int64)
    for k in range(3):
        np.minimum.at(lo, conn[:, k], conn.min(1))
        np.maximum.at(hi, conn[:, k], conn.max(1))
    total = 0
    for r in range(nranks):
        a, b = cuts[r], cuts[r + 1]
        if b > a:
            total += (a - min(lo[a:b].min(), a)) + (max(hi[a:b].max() + 1, b) - b)
    return total


def test_rcm_shrinks_the_halo_of_an_8_way_row_block_partition(built):
    _, mesh, _ = _example("example_linkedin")
    p, _, _ = reorder.rcm(mesh)
    before, after = _halo_nodes(mesh, 8), _halo_nodes(reorder.permute_mesh(mesh, p), 8)
    assert before > 4 * mesh.n_nodes          # gmsh-like order: every block reaches across most of the vector
    assert after < mesh.n_nodes // 4 and after * 20 < before


def test_rcm_edge_cases(built):
    lib = _lib.load()
    # empty mesh
    empty = MeshSoA(*(np.zeros(0, t) for t in (np.float64, np.float64, np.uint32, np.uint32, np.uint32, np.float64,
                                                 np.float64, np.float64, np.float64, np.uint8)))
    p, before, after = reorder.rcm(empty)
    assert p.shape == (0,) and before == after == 0
    # two components + nodes no element references: a valid permutation, unreferenced nodes last in their order
    n0 = np.array([0, 1, 7, 8], np.uint32); n1 = np.array([1, 2, 8, 9], np.uint32); n2 = np.array([5, 5, 11, 11], np.uint32)
    new_of_old = np.empty(12, np.uint32)
    b0, b1 = C.c_uint64(), C.c_uint64()
    assert lib.mag_reorder_rcm(12, 4, _lib.ptr(n0), _lib.ptr(n1), _lib.ptr(n2), _lib.ptr(new_of_old), C.byref(b0), C.byref(b1)) == 0
    assert np.array_equal(np.sort(new_of_old), np.arange(12))
    assert list(new_of_old[[3, 4, 6, 10]]) == [8, 9, 10, 11]
    comp_a, comp_b = new_of_old[[0, 1, 2, 5]], new_of_old[[7, 8, 9, 11]]
    assert comp_a.max() - comp_a.min() == 3 and comp_b.max() - comp_b.min() == 3      # components stay contiguous
    assert b0.value == 5 and b1.value <= 3
    # a degenerate element (repeated node) is tolerated; band outputs are optional
    d0 = np.array([0, 2], np.uint32); d1 = np.array([0, 1], np.uint32); d2 = np.array([1, 0], np.uint32)
    out = np.empty(3, np.uint32)
    assert lib.mag_reorder_rcm(3, 2, _lib.ptr(d0), _lib.ptr(d1), _lib.ptr(d2), _lib.ptr(out), None, None) == 0
    assert np.array_equal(np.sort(out), np.arange(3))
    # an element that references a node past the end: MAG_ERR_BAD_INDEX with a message, nothing written
    bad = np.array([0, 12], np.uint32)
    rc = lib.mag_reorder_rcm(12, 2, _lib.ptr(bad), _lib.ptr(n1), _lib.ptr(n2), _lib.ptr(new_of_old), None, None)
    assert rc == _lib.MAG_ERR_BAD_INDEX and b"element 1" in lib.mag_host_last_error()
    mesh = meshgen.plate(3, 2)
    mesh.n2 = mesh.n2.copy(); mesh.n2[4] = 1000
    with pytest.raises(MagnetiteError, match="references a node"):
        reorder.rcm(mesh)
    with pytest.raises(MagnetiteError, match="references a node"):
        reorder.mesh_band(mesh)


def test_cpp_cli_band_mode_agrees_with_the_python_binding(built, tmp_path):
    """`magnetite_b200 --band geom.msh` (C++ host layer -> mag_reorder_rcm, no GPU) reports the same band as the
    Python binding for the same mesh file."""
    import subprocess
    from magnetite_b200 import geometry
    root = Path(__file__).resolve().parent.parent
    subprocess.run(["make", "-C", str(root / "host")], check=True, capture_output=True)
    _, mesh, _ = _example("example_cover")
    conn = np.stack([mesh.n0, mesh.n1, mesh.n2], 1).astype(int)
    geometry.write_msh(str(tmp_path / "geom.msh"), mesh.x, mesh.y, conn)
    r = subprocess.run([str(root / "host" / "magnetite_b200"), "--band", str(tmp_path / "geom.msh")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    nodes, elements = geometry.parse_mesh(str(tmp_path / "geom.msh"))            # the numbering the CLI saw
    _, before, after = reorder.rcm(MeshSoA.from_aos(nodes, elements))
    assert r.stdout.split() == ["nodes", str(len(nodes)), "elements", str(len(elements)), "band", str(before), "rcm", str(after)]
    assert after * 20 < before
