"""CPU: pins the oracle (oracle/magnetite_oracle.c) against analytic known-answer tests derived
from the reference's formulas (SURVEY §8c — the reference has no tests or golden vectors of
its own: PARITY UNPINNED), an independent numpy restatement, a direct solve, and the committed
golden fixtures."""
from pathlib import Path

import numpy as np
import pytest

from magnetite_b200 import meshgen
from magnetite_b200.datatypes import MeshSoA
from oracle import oracle as O
from oracle import reference_semantics as R

META = meshgen.EXAMPLE_MATERIAL
NU, E, T = META.poisson_ratio, META.youngs_modulus, META.part_thickness
GOLDEN = Path(__file__).resolve().parent / "golden"


def tri_mesh(pts, conn=((0, 1, 2),)):
    pts = np.asarray(pts, float)
    conn = np.asarray(conn, np.uint32)
    n = len(pts)
    z = np.zeros(n)
    return MeshSoA(pts[:, 0].copy(), pts[:, 1].copy(), conn[:, 0].copy(), conn[:, 1].copy(), conn[:, 2].copy(),
                   z.copy(), z.copy(), z.copy(), z.copy(), np.zeros(n, np.uint8))


def conn_of(m):
    return np.stack([m.n0, m.n1, m.n2], 1).astype(int)


def test_kat1_element_stiffness_values():
    m = tri_mesh([(0, 0), (2, 0), (0, 2)])
    om = O.Mesh(m)
    assert O.element_area(om)[0] == 2.0
    D = O.stress_strain(NU, E)
    assert D[0, 0] == pytest.approx(7.7432386937492981e10, rel=1e-15)
    assert D[0, 1] == pytest.approx(2.5552687689372684e10, rel=1e-15)
    assert D[2, 2] == pytest.approx(2.5939849624060146e10, rel=1e-15)
    ke = O.element_stiffness(om, META)[0]
    row0 = [2.5843059140388283e10, 1.2873134328358208e10, -1.9358096734373245e10,
            -6.4849624060150366e9, -6.4849624060150366e9, -6.3881719223431711e9]
    np.testing.assert_allclose(ke[0], row0, rtol=1e-15)
    assert ke[2, 2] == pytest.approx(1.9358096734373245e10, rel=1e-15)
    assert ke[3, 3] == ke[3, 4] == ke[4, 4] == pytest.approx(6.4849624060150366e9, rel=1e-15)
    for r, c in ((2, 3), (2, 4), (3, 5), (4, 5)):
        assert ke[r, c] == 0.0 and ke[c, r] == 0.0          # exact zeros


def test_kat2_orientation_flips_sign():
    ke = O.element_stiffness(O.Mesh(tri_mesh([(0, 0), (2, 0), (0, 2)])), META)[0]
    m_rev = tri_mesh([(0, 2), (2, 0), (0, 0)])
    om = O.Mesh(m_rev)
    assert O.element_area(om)[0] == -2.0
    ke_rev = O.element_stiffness(om, META)[0]
    P = np.zeros((6, 6))
    for a, b in ((0, 2), (1, 1), (2, 0)):
        P[2 * a, 2 * b] = P[2 * a + 1, 2 * b + 1] = 1
    np.testing.assert_allclose(ke_rev, -P @ ke @ P.T, rtol=1e-14, atol=0)


def test_element_properties_rigid_body_and_symmetry():
    rng = np.random.default_rng(7)
    pts = rng.uniform(-3, 3, (30, 2))
    conn = rng.integers(0, 30, (40, 3))
    conn = conn[(conn[:, 0] != conn[:, 1]) & (conn[:, 1] != conn[:, 2]) & (conn[:, 0] != conn[:, 2])]
    m = tri_mesh(pts, conn)
    om = O.Mesh(m)
    area = O.element_area(om)
    keep = np.abs(area) > 0.3
    ke = O.element_stiffness(om, META)[keep]
    scale = np.abs(ke).max(axis=(1, 2))
    tx = np.array([1, 0, 1, 0, 1, 0.0]); ty = np.array([0, 1, 0, 1, 0, 1.0])
    assert (np.abs(ke @ tx).max(axis=1) / scale).max() < 1e-13
    assert (np.abs(ke @ ty).max(axis=1) / scale).max() < 1e-13
    assert (np.abs(ke - ke.transpose(0, 2, 1)).max(axis=(1, 2)) / scale).max() < 1e-13


def test_c_oracle_matches_numpy_twin_bitwise():
    m = meshgen.jitter(meshgen.plate(5, 4))
    om = O.Mesh(m)
    ke = O.element_stiffness(om, META)
    for e, el in enumerate(conn_of(m)):
        assert np.array_equal(ke[e], R.element_stiffness(m.x, m.y, el, NU, E, T))
    K = O.assemble_dense(om, ke)
    assert np.array_equal(K, R.assemble(m.x, m.y, conn_of(m), NU, E, T))


def test_dense_and_sparse_modes_are_bit_identical():
    for m in (meshgen.plate(8, 6), meshgen.jitter(meshgen.plate(7, 5)),
              meshgen.perforated_plate(16, 16, pitch=8, radius=2)):
        om = O.Mesh(m)
        ke = O.element_stiffness(om, META)
        Kd = O.assemble_dense(om, ke)
        rp, col, val = O.assemble_sparse(om, ke)
        dense_from_sparse = np.zeros_like(Kd)
        for r in range(len(rp) - 1):
            dense_from_sparse[r, col[rp[r]:rp[r + 1]]] = val[rp[r]:rp[r + 1]]
        assert np.array_equal(Kd, dense_from_sparse)
        (a, b, c), rhs_d, fm_d = O.partition(om, Kd, dense=True)
        (a2, b2, c2), rhs_s, fm_s = O.partition(om, (rp, col, val), dense=False)
        assert np.array_equal(a, a2) and np.array_equal(b, b2) and np.array_equal(c, c2)
        assert np.array_equal(rhs_d, rhs_s) and np.array_equal(fm_d, fm_s)
        rd, rs = O.run(om, META, dense=True), O.run(om, META, dense=False)
        for k in ("ux", "uy", "fx", "fy", "stress"):
            assert np.array_equal(rd[k], rs[k]), k
        assert rd["stats"]["iters"] == rs["stats"]["iters"]


def test_kat5_pattern_counts():
    m = meshgen.plate(8, 6)
    om = O.Mesh(m)
    rp, col, val = O.assemble_sparse(om, O.element_stiffness(om, META))
    assert len(val) == 1516 and np.count_nonzero(val) == 1320
    assert rp[-1] == 1516 and len(rp) == 2 * m.n_nodes + 1


def test_kat3_patch_test_analytic():
    m = meshgen.patch_square(2.0, 0.005)
    r = O.run(O.Mesh(m), META, dense=True)
    np.testing.assert_allclose(r["ux"], [0, 0.01, 0.01, 0], atol=1e-15)
    np.testing.assert_allclose(r["uy"], [0, 0, -0.0033, -0.0033], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(r["fx"], [-1.725e8, 1.725e8, 1.725e8, -1.725e8], rtol=1e-9)
    np.testing.assert_allclose(r["stress"], [3.45e8, 3.45e8], rtol=1e-9)


def test_kat4_clockwise_flip_quirk():
    g = np.load(GOLDEN / "clockwise_unit.npz")
    ccw = O.run(O.Mesh(meshgen.patch_square(1.0, 0.005)), META, dense=True)
    np.testing.assert_allclose(g["ux"], ccw["ux"], rtol=1e-9, atol=1e-15)    # displacement-driven: same u
    np.testing.assert_allclose(g["uy"], ccw["uy"], rtol=1e-9, atol=1e-15)
    assert (g["kff_val"][g["kff_col"] == np.repeat(np.arange(len(g["kff_rowptr"]) - 1), np.diff(g["kff_rowptr"]))] < 0).all()
    f = np.load(GOLDEN / "clockwise_force.npz")                              # force-driven: sign flips
    assert f["ux"][1] == pytest.approx(-5.7971014492753614e-5, rel=1e-9)
    np.testing.assert_allclose(f["stress"], [-4.0e6, -4.0e6], rtol=1e-9)


def test_cg_matches_direct_solve_and_port_mode():
    m = meshgen.jitter(meshgen.plate(12, 6))
    om = O.Mesh(m)
    r = O.run(om, META, dense=False)
    U, F, Kff, rhs = R.solve_direct(m.x, m.y, conn_of(m), m.known, m.ux, m.uy, m.fx, m.fy, NU, E, T)
    u = np.stack([r["ux"], r["uy"]], 1).ravel()
    f = np.stack([r["fx"], r["fy"]], 1).ravel()
    assert np.linalg.norm(u - U) / np.linalg.norm(U) < 1e-11
    assert np.linalg.norm(f - F) / np.linalg.norm(F) < 1e-10
    s_ref, _ = R.stress(m.x, m.y, conn_of(m), U, NU, E)
    assert np.abs(r["stress"] - s_ref).max() / np.abs(s_ref).max() < 1e-9
    assert abs(f.reshape(-1, 2).sum(0)).max() / np.abs(f).max() < 1e-10       # equilibrium
    p = O.run(om, META, O.cg_options(jacobi=1, rel_tol=1e-12), dense=False)    # port-mode PCG
    up = np.stack([p["ux"], p["uy"]], 1).ravel()
    assert np.linalg.norm(up - U) / np.linalg.norm(U) < 1e-9


def test_cg_executor_semantics():
    rp = np.array([0, 1, 2], np.int64); col = np.array([0, 1], np.int32); val = np.array([2.0, 4.0])
    x, it, cost = O.cg((rp, col, val), np.array([0.0, 0.0]), O.cg_options())
    assert it == 0 and cost == 0.0 and (x == 0).all()              # init cost already <= target
    x, it, cost = O.cg((rp, col, val), np.array([2.0, 4.0]), O.cg_options())
    np.testing.assert_allclose(x, [1.0, 1.0], rtol=1e-14)
    assert 1 <= it <= 2 and cost <= 1e-4
    x, it, cost = O.cg((rp, col, val), np.array([2.0, 4.0]), O.cg_options(max_iter=1))
    assert it == 1                                                 # max_iters honoured
    x0, it0, _ = O.cg((rp, col, val), np.array([2.0, 4.0]), O.cg_options(max_iter=0))
    assert it0 == 0 and (x0 == 0).all()                            # best_param = x0


def test_bad_inputs():
    m = meshgen.plate(3, 2)
    bad = m.copy(); bad.n1[0] = 999
    with pytest.raises(O.OracleError) as ei:
        O.run(O.Mesh(bad), META)
    assert ei.value.code == -3
    bc = m.copy(); bc.known[5] = 1 | 2 | 4 | 8
    with pytest.raises(O.OracleError) as ei:
        O.run(O.Mesh(bc), META)
    assert ei.value.code == -2


@pytest.mark.parametrize("name", ["patch_2x2", "plate_8x6", "plate_jitter_10x7", "perforated_24x16",
                                  "clockwise_unit", "clockwise_force"])
def test_oracle_reproduces_golden(name):
    g = np.load(GOLDEN / f"{name}.npz")
    m = MeshSoA(g["x"], g["y"], g["n0"], g["n1"], g["n2"], g["bc_ux"], g["bc_uy"], g["bc_fx"], g["bc_fy"], g["known"])
    om = O.Mesh(m)
    ke = O.element_stiffness(om, META)
    assert np.array_equal(ke, g["ke"])
    full = O.assemble_sparse(om, ke)
    assert all(np.array_equal(a, g[k]) for a, k in zip(full, ("full_rowptr", "full_col", "full_val")))
    (rp, col, val), rhs, fmap = O.partition(om, full, dense=False)
    assert np.array_equal(rp, g["kff_rowptr"]) and np.array_equal(col, g["kff_col"])
    assert np.array_equal(val, g["kff_val"]) and np.array_equal(rhs, g["rhs"]) and np.array_equal(fmap, g["free_map"])
    r = O.run(om, META, dense=False)
    for k in ("ux", "uy", "fx", "fy", "stress"):
        assert np.array_equal(r[k], g[k]), k
    assert r["stats"]["iters"] == int(g["iters"][0])


def test_all_cores_pcg_agrees_with_the_sequential_port():
    """oracle_mt (bench infrastructure: the port's Jacobi-PCG on pthreads) reaches the port's solution in the
    port's iteration count, for any thread count, and is deterministic for a given one."""
    from oracle import oracle_mt as MT
    mesh = meshgen.jitter(meshgen.plate(60, 30))
    om = O.Mesh(mesh)
    csr, rhs, fmap = O.partition(om, O.assemble_sparse(om, O.element_stiffness(om, META)), dense=False)
    x_ref, it_ref, _ = O.cg(csr, rhs, O.cg_options(jacobi=1, rel_tol=1e-10))
    for threads in (1, 3, 8, 10_000):                      # absurd thread counts are clamped (>= 64 rows each)
        x, it, res = MT.pcg(csr, rhs, rel_tol=1e-10, threads=threads)
        assert abs(it - it_ref) <= 2
        assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-9
        assert res <= 1e-10 * np.linalg.norm(rhs)
    a, _, _ = MT.pcg(csr, rhs, threads=4)
    b, _, _ = MT.pcg(csr, rhs, threads=4)
    assert np.array_equal(a, b)
    # a zero right-hand side needs no iteration; an empty system is legal
    x0, it0, _ = MT.pcg(csr, np.zeros_like(rhs), threads=4)
    assert it0 == 0 and not x0.any()
    xe, ite, _ = MT.pcg((np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0)), np.zeros(0), threads=4)
    assert xe.size == 0 and ite == 0


def test_oracle_matches_the_formulas_of_the_reference_documentation():
    """The reference ships no numeric vectors, but its own documentation states the element formulas independently
    of the code (under-the-hood.md:294-299 and :577-582 D; :520-524 B with the 1/(2A) factor, x_ij = x_i - x_j and
    y_ij = y_i - y_j over nodes 1..3, :592; :551 and :572 k_e = B^T D B t A; :598-604 the area as half a determinant —
    printed there with a third column of zeros, which would vanish: the ones column of solver.rs:187-193 is meant;
    :621-640 the scatter of k_e into K by node pairs, DOF = 2*node + axis).  The oracle follows the CODE's operation
    order, so agreement with these closed forms is to rounding, not bitwise: 1e-13 relative to the largest entry."""
    rng = np.random.default_rng(7)
    pts = rng.uniform(-5.0, 5.0, size=(40, 2))
    conn = np.array([rng.choice(40, 3, replace=False) for _ in range(60)], np.uint32)
    m = tri_mesh(pts, conn)
    om = O.Mesh(m)
    area = O.element_area(om)
    ke = O.element_stiffness(om, META)
    D_doc = E / (1.0 - NU ** 2) * np.array([[1.0, NU, 0.0], [NU, 1.0, 0.0], [0.0, 0.0, (1.0 - NU) / 2.0]])
    np.testing.assert_allclose(O.stress_strain(NU, E), D_doc, rtol=1e-15)
    K_doc = np.zeros((80, 80))
    for e, (a, b, c) in enumerate(conn.astype(int)):
        (x1, y1), (x2, y2), (x3, y3) = pts[a], pts[b], pts[c]
        A_doc = 0.5 * np.linalg.det(np.array([[x1, y1, 1.0], [x2, y2, 1.0], [x3, y3, 1.0]]))
        assert area[e] == pytest.approx(A_doc, rel=1e-12)
        y23, y31, y12, x32, x13, x21 = y2 - y3, y3 - y1, y1 - y2, x3 - x2, x1 - x3, x2 - x1
        B_doc = np.array([[y23, 0, y31, 0, y12, 0], [0, x32, 0, x13, 0, x21], [x32, y23, x13, y31, x21, y12]]) / (2.0 * A_doc)
        ke_doc = B_doc.T @ D_doc @ B_doc * T * A_doc
        assert np.abs(ke[e] - ke_doc).max() <= 1e-13 * np.abs(ke_doc).max()
        for lr, nr in enumerate((a, b, c)):
            for lc, nc in enumerate((a, b, c)):
                K_doc[2 * nr:2 * nr + 2, 2 * nc:2 * nc + 2] += ke_doc[2 * lr:2 * lr + 2, 2 * lc:2 * lc + 2]
    K = O.assemble_dense(om, ke)
    assert np.abs(K - K_doc).max() <= 1e-12 * np.abs(K_doc).max()
