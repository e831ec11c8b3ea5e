"""GPU: parity of the CUDA path (through the C ABI) with the oracle and the golden fixtures.

Bars (BASELINE.json north_star): CSR sparsity pattern and DOF numbering bit-exact; K_e within
1e-12 relative (measured per element as max|dK|/max|K|); displacements within 1e-9 relative L2;
stresses within 1e-8 relative (|ds|_inf/|s|_inf).  Where the device follows the reference's
operation order exactly (K_e, full K, K_ff, rhs, reactions and stresses given u) the tests also
demand bit equality.
"""
from pathlib import Path

import numpy as np
import pytest

from magnetite_b200 import _lib, meshgen, solver
from magnetite_b200.datatypes import Element, MeshSoA, Node, Vertex
from magnetite_b200.error import MagnetiteError
from oracle import oracle as O

pytestmark = pytest.mark.gpu
META = meshgen.EXAMPLE_MATERIAL
GOLDEN = Path(__file__).resolve().parent / "golden"
GOLDEN_CASES = ["patch_2x2", "plate_8x6", "plate_jitter_10x7", "perforated_24x16", "clockwise_unit",
                "clockwise_force"]


def golden_mesh(g):
    return MeshSoA(g["x"], g["y"], g["n0"], g["n1"], g["n2"], g["bc_ux"], g["bc_uy"], g["bc_fx"], g["bc_fy"], g["known"])


def rel_l2(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def ke_rel(ke, ref):
    scale = np.abs(ref).max(axis=(1, 2))
    return (np.abs(ke - ref).max(axis=(1, 2)) / scale).max()


def compat():
    return _lib.default_options(compat=1)


MESHES = {
    "plate_20x10": lambda: meshgen.plate(20, 10),
    "plate_33x17_h0.3": lambda: meshgen.plate(33, 17, h=0.3),       # non-representable coordinates
    "jitter_31x19": lambda: meshgen.jitter(meshgen.plate(31, 19)),
    "perforated_96x48": lambda: meshgen.perforated_plate(96, 48, pitch=16, radius=4),
    "plate_257x65": lambda: meshgen.plate(257, 65),                 # > one sort tile per pass, ragged slices
}


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_against_golden_fixtures(ctx, name):
    g = np.load(GOLDEN / f"{name}.npz")
    mesh = golden_mesh(g)
    ke = solver.element_stiffness(mesh, META, ctx)
    assert ke_rel(ke, g["ke"]) < 1e-12
    assert np.array_equal(ke, g["ke"]), "K_e is expected to be bit-identical"
    with solver.System(mesh, META, ctx) as S:
        rp, col, val = S.export_full()
        assert np.array_equal(rp, g["full_rowptr"]) and np.array_equal(col, g["full_col"])
        assert np.array_equal(val, g["full_val"])
        rp, col, val, rhs, fmap = S.export_kff()
        assert np.array_equal(rp, g["kff_rowptr"]) and np.array_equal(col, g["kff_col"])   # pattern: bit-exact
        assert np.array_equal(fmap, g["free_map"])                                          # DOF numbering
        assert np.array_equal(val, g["kff_val"]) and np.array_equal(rhs, g["rhs"])
        sol = S.solve(compat())
    u, ur = np.concatenate([sol.ux, sol.uy]), np.concatenate([g["ux"], g["uy"]])
    assert rel_l2(u, ur) < 1e-9
    assert np.abs(sol.stress - g["stress"]).max() / np.abs(g["stress"]).max() < 1e-8
    assert np.array_equal(np.sign(sol.stress), np.sign(g["stress"]))
    f, fr = np.concatenate([sol.fx, sol.fy]), np.concatenate([g["fx"], g["fy"]])
    assert np.abs(f - fr).max() / np.abs(fr).max() < 1e-8


@pytest.mark.parametrize("name", list(MESHES))
def test_pipeline_against_oracle(ctx, name):
    mesh = MESHES[name]()
    om = O.Mesh(mesh)
    ke_ref = O.element_stiffness(om, META)
    ke = solver.element_stiffness(mesh, META, ctx)
    assert ke_rel(ke, ke_ref) < 1e-12 and np.array_equal(ke, ke_ref)
    assert np.array_equal(solver.element_areas(mesh, ctx), O.element_area(om))
    full_ref = O.assemble_sparse(om, ke_ref)
    (rp_r, col_r, val_r), rhs_r, fmap_r = O.partition(om, full_ref, dense=False)
    with solver.System(mesh, META, ctx) as S:
        full = S.export_full()
        for a, b in zip(full, full_ref):
            assert np.array_equal(a, b)
        rp, col, val, rhs, fmap = S.export_kff()
        assert np.array_equal(rp, rp_r) and np.array_equal(col, col_r) and np.array_equal(fmap, fmap_r)
        assert np.array_equal(val, val_r) and np.array_equal(rhs, rhs_r)
        assert S.nnz == len(val_r) and S.nnz_structural == len(full_ref[2])
        # SpMV: both device formats against the oracle's sequential row sums
        x = np.random.default_rng(3).normal(size=S.n_free)
        y_ref = O.spmv((rp_r, col_r, val_r), x)
        assert np.array_equal(S.spmv(x, fmt=1), y_ref)                     # CSR kernel keeps the order
        assert rel_l2(S.spmv(x, fmt=2), y_ref) < 1e-14                     # SELL uses FMA
        ref = O.run(om, META, O.cg_options(), dense=False)
        sol = S.solve(compat(), want_sigma=True)
        sol_csr = S.solve(_lib.default_options(compat=1, spmv_format=1))
    u, ur = np.concatenate([sol.ux, sol.uy]), np.concatenate([ref["ux"], ref["uy"]])
    assert rel_l2(u, ur) < 1e-9
    assert rel_l2(np.concatenate([sol_csr.ux, sol_csr.uy]), ur) < 1e-9
    assert np.abs(sol.stress - ref["stress"]).max() / np.abs(ref["stress"]).max() < 1e-8
    f, fr = np.concatenate([sol.fx, sol.fy]), np.concatenate([ref["fx"], ref["fy"]])
    assert np.abs(f - fr).max() / np.abs(fr).max() < 1e-8
    assert sol.stats["converged"] == 1 and sol.stats["final_residual"] <= 1e-4
    # given the SAME displacements the device post-processing is bit-identical to the oracle
    s_dev, sig_dev = solver.stress_soa(mesh, META, ref["ux"], ref["uy"], want_sigma=True, ctx=ctx)
    s_ref, sig_ref = O.stress(om, META, ref["ux"], ref["uy"], want_sigma=True)
    assert np.array_equal(s_dev, s_ref) and np.array_equal(sig_dev, sig_ref)
    # equilibrium: reactions + applied forces sum to ~0
    assert abs(sol.fx.sum()) / np.abs(sol.fx).max() < 1e-8 and abs(sol.fy.sum()) / np.abs(sol.fx).max() < 1e-8


def test_north_star_pcg_mode(ctx):
    """Jacobi-PCG to 1e-9 relative residual (the benchmark configuration) still lands within the
    displacement tolerance of a tight solve on a small mesh when asked for a tight residual."""
    mesh = meshgen.jitter(meshgen.plate(40, 20))
    ref = O.run(O.Mesh(mesh), META, O.cg_options(), dense=False)
    ur = np.concatenate([ref["ux"], ref["uy"]])
    with solver.System(mesh, META, ctx) as S:
        tight = S.solve(_lib.default_options(rel_tol=1e-13))
        loose = S.solve(_lib.default_options())                       # rel_tol 1e-9, Jacobi
        nopre = S.solve(_lib.default_options(precond=0))
    assert rel_l2(np.concatenate([tight.ux, tight.uy]), ur) < 1e-9
    assert loose.stats["final_residual"] <= 1e-9 * loose.stats["b_norm"]
    assert loose.stats["iters"] < tight.stats["iters"]
    assert rel_l2(np.concatenate([loose.ux, loose.uy]), ur) < 1e-6
    assert nopre.stats["converged"] == 1
    port = O.run(O.Mesh(mesh), META, O.cg_options(jacobi=1, rel_tol=1e-9), dense=False)
    assert abs(int(loose.stats["iters"]) - int(port["stats"]["iters"])) <= max(3, port["stats"]["iters"] // 50)


def test_bit_identical_across_runs(ctx):
    mesh = meshgen.jitter(meshgen.plate(64, 32))
    outs = []
    for _ in range(2):
        with solver.System(mesh, META, ctx) as S:
            kff = S.export_kff()
            sol = S.solve(_lib.default_options())
        outs.append((kff, sol))
    for a, b in zip(outs[0][0], outs[1][0]):
        assert a.tobytes() == b.tobytes()
    for k in ("ux", "uy", "fx", "fy", "stress"):
        assert getattr(outs[0][1], k).tobytes() == getattr(outs[1][1], k).tobytes()
    assert outs[0][1].stats["iters"] == outs[1][1].stats["iters"]


def test_clockwise_mesh_is_negative_definite_but_solves(ctx):
    g = np.load(GOLDEN / "clockwise_unit.npz")
    sol = solver.solve_soa(golden_mesh(g), META, ctx, compat())
    assert sol.stats["negative_definite"] == 1
    assert rel_l2(np.concatenate([sol.ux, sol.uy]), np.concatenate([g["ux"], g["uy"]])) < 1e-9
    sol_j = solver.solve_soa(golden_mesh(g), META, ctx, _lib.default_options(rel_tol=1e-13))
    assert rel_l2(np.concatenate([sol_j.ux, sol_j.uy]), np.concatenate([g["ux"], g["uy"]])) < 1e-9


def test_drop_in_run_signature_and_csv(ctx, tmp_path, capsys):
    """solver::run's signature on AoS nodes/elements, then csv_output, like main.rs:64-69."""
    from magnetite_b200 import post_processor
    mesh = meshgen.plate(12, 6)
    nodes, elements = mesh.to_aos()
    assert nodes[0].fx is None and nodes[1].ux is None and elements[0].stress is None
    solver.run(nodes, elements, META)
    out = capsys.readouterr().out
    assert "info: building element stiffness matrices..." in out and "info: solve complete" in out
    assert "info: finished conjugate gradient approximation in" in out
    ref = O.run(O.Mesh(mesh), META, O.cg_options(), dense=True)
    u = np.array([[n.ux, n.uy] for n in nodes]); s = np.array([e.stress for e in elements])
    assert rel_l2(u.ravel(), np.stack([ref["ux"], ref["uy"]], 1).ravel()) < 1e-9
    assert np.abs(s - ref["stress"]).max() / np.abs(ref["stress"]).max() < 1e-8
    assert all(n.fx is not None and n.fy is not None for n in nodes)
    post_processor.csv_output(elements, nodes, str(tmp_path / "nodes.csv"), str(tmp_path / "elements.csv"), quiet=True)
    rows = (tmp_path / "nodes.csv").read_text().splitlines()
    assert rows[0] == "x,y,ux,uy" and len(rows) == len(nodes) + 1
    assert rows[1] == "0,0,0,0" and rows[13].startswith("24,0,3,")
    erows = (tmp_path / "elements.csv").read_text().splitlines()
    assert erows[0] == "n0,n1,n2,stress" and erows[1].startswith("0,1,14,")
    assert solver.compute_element_area(elements[0], nodes) == 2.0


def test_error_paths(ctx):
    mesh = meshgen.plate(6, 3)
    bad = mesh.copy(); bad.n2[4] = 10_000
    with pytest.raises(MagnetiteError) as ei:
        solver.solve_soa(bad, META, ctx)
    assert ei.value.code == _lib.MAG_ERR_BAD_INDEX and ei.value.kind == "Solver"
    bc = mesh.copy(); bc.known[3] = 15
    with pytest.raises(MagnetiteError) as ei:
        solver.solve_soa(bc, META, ctx)
    assert ei.value.code == _lib.MAG_ERR_BAD_BC
    with solver.System(meshgen.plate(30, 15), META, ctx) as S:
        with pytest.raises(MagnetiteError) as ei:
            S.solve(_lib.default_options(max_iter=4))
        assert ei.value.code == _lib.MAG_ERR_NOT_CONVERGED
        sol = S.solve(_lib.default_options(max_iter=4), allow_not_converged=True)
        assert sol.stats["iters"] == 4 and sol.stats["converged"] == 0
        ok = S.solve(_lib.default_options())                  # the context survives an error
        assert ok.stats["converged"] == 1


def test_empty_and_degenerate_inputs(ctx):
    z = np.zeros(0)
    empty = MeshSoA(z, z, np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint32), z, z, z, z,
                    np.zeros(0, np.uint8))
    sol = solver.solve_soa(empty, META, ctx)
    assert sol.ux.size == 0 and sol.stress.size == 0 and sol.stats["iters"] == 0
    # nodes but no elements and nothing to solve: everything prescribed
    m = meshgen.patch_square()
    allfixed = m.copy(); allfixed.known[:] = 3
    sol = solver.solve_soa(allfixed, META, ctx)
    ref = O.run(O.Mesh(allfixed), META, dense=True)
    assert sol.stats["n_free"] == 0 and np.array_equal(sol.ux, ref["ux"])
    assert np.array_equal(sol.fx, ref["fx"]) and np.array_equal(sol.stress, ref["stress"])
    # zero load: b = 0 -> 0 iterations, u = 0 (argmin stops on the initial cost)
    zero = meshgen.plate(5, 4, ux_right=0.0)
    sol = solver.solve_soa(zero, META, ctx, compat())
    assert sol.stats["iters"] == 0 and not sol.ux.any() and not sol.uy.any()
    # an isolated node that no element references keeps an empty matrix row
    iso = meshgen.plate(4, 3).copy()
    iso = MeshSoA(np.append(iso.x, 99.0), np.append(iso.y, 99.0), iso.n0, iso.n1, iso.n2, np.append(iso.ux, 0.0),
                  np.append(iso.uy, 0.0), np.append(iso.fx, 0.0), np.append(iso.fy, 0.0), np.append(iso.known, 3).astype(np.uint8))
    sol = solver.solve_soa(iso, META, ctx, compat())
    ref = O.run(O.Mesh(iso), META, dense=True)
    assert rel_l2(np.concatenate([sol.ux, sol.uy]), np.concatenate([ref["ux"], ref["uy"]])) < 1e-9


def test_device_generated_plate_matches_host_generator(ctx):
    import ctypes as C
    lib = _lib.load()
    dm = C.c_void_p()
    _lib.check(lib.mag_devmesh_plate(ctx.handle, 37, 11, 2.0, 3.0, C.byref(dm)), "devmesh")
    view = _lib.MagMesh()
    _lib.check(lib.mag_devmesh_view(dm, C.byref(view)), "view")
    assert view.on_device == 1 and view.n_elems == 2 * 37 * 11
    mat = _lib.MagMaterial(META.youngs_modulus, META.poisson_ratio, META.part_thickness)
    opt = _lib.default_options(rel_tol=1e-12)
    sysh = C.c_void_p(); st = _lib.MagStats()
    _lib.check(lib.mag_assemble(ctx.handle, C.byref(view), C.byref(mat), C.byref(opt), C.byref(sysh), C.byref(st)), "assemble")
    host = meshgen.plate(37, 11)
    with solver.System(host, META, ctx) as S:
        ref_kff = S.export_kff()
        n_free, nnz = S.n_free, S.nnz
    assert (st.n_free, st.nnz) == (n_free, nnz)
    rowptr = np.empty(n_free + 1, np.int64); col = np.empty(nnz, np.int32); val = np.empty(nnz); rhs = np.empty(n_free)
    _lib.check(lib.mag_system_export_kff(sysh, _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(val), _lib.ptr(rhs), None), "export")
    assert np.array_equal(rowptr, ref_kff[0]) and np.array_equal(col, ref_kff[1])
    assert np.array_equal(val, ref_kff[2]) and np.array_equal(rhs, ref_kff[3])
    lib.mag_system_free(sysh)
    lib.mag_devmesh_free(dm)


def test_full_size_properties_1m_triangles(ctx):
    """BASELINE config 3 (1000x500 cells, 1 M triangles): size-independent properties instead of a
    full oracle run — equilibrium, uniform-strain sanity, pattern counts, determinism of K_ff."""
    mesh = meshgen.plate(1000, 500)
    with solver.System(mesh, META, ctx) as S:
        assert S.n_free == 2 * 501_501 - 2 * 501 - 501
        rp, col, val, rhs, fmap = S.export_kff()
        assert (np.diff(rp) > 0).all() and (val != 0.0).all()
        rows = np.repeat(np.arange(S.n_free), np.diff(rp))
        # symmetric pattern, ascending columns
        assert ((col[1:] > col[:-1]) | (rows[1:] != rows[:-1])).all()
        import scipy.sparse as sp
        A = sp.csr_matrix((val, col, rp), shape=(S.n_free, S.n_free))
        assert abs(A - A.T).max() / abs(A).max() < 1e-12
        sol = S.solve(_lib.default_options())
        assert sol.stats["converged"] == 1
        # true residual of the returned solution
        xfree = np.empty(S.n_free)
        u = np.stack([sol.ux, sol.uy], 1).ravel()
        xfree[fmap[fmap >= 0]] = u[fmap >= 0]
        assert np.linalg.norm(A @ xfree - rhs) / np.linalg.norm(rhs) < 5e-9
    assert abs(sol.fx.sum()) / np.abs(sol.fx).max() < 1e-3   # residual 1e-9 relative, ~1e3 boundary nodes
    assert sol.ux.min() >= -1e-9 and sol.ux.max() <= 3.0 + 1e-9
    mid = np.abs(sol.stress[len(sol.stress) // 2])
    assert 0.5 < mid / (69e9 * 3.0 / 2000.0) < 1.5            # ~ E * strain in the middle of the plate


def test_two_level_solve_is_bit_identical_run_to_run_on_a_large_grid(ctx):
    """2.1 M unknowns: the p-update runs on more CTAs than are resident at once, so a CTA scheduled late would see
    any scalar the same launch rewrites (the round-1 race on r.z).  Two solves of one system and a
    solve of a second system built from the same mesh must agree in every bit and in the iteration count."""
    mesh = meshgen.plate(1500, 700)
    opt = _lib.default_options(precond=2)
    outs = []
    with solver.System(mesh, META, ctx) as S:
        assert S.n_free > 2_000_000
        outs.append(S.solve(opt)); outs.append(S.solve(opt))
        rr, bb = S.true_residual(outs[0].ux, outs[0].uy)
        assert rr <= (2e-9) ** 2 * bb
    with solver.System(mesh, META, ctx) as S:
        outs.append(S.solve(opt))
    assert outs[0].stats["precond_used"] == 2 and outs[0].stats["converged"] == 1
    for o in outs[1:]:
        assert o.stats["iters"] == outs[0].stats["iters"]
        for k in ("ux", "uy", "fx", "fy", "stress"):
            assert getattr(o, k).tobytes() == getattr(outs[0], k).tobytes(), k


def test_config3_kff_bit_exact_against_oracle(ctx):
    """BASELINE configs[2] (1 M triangles, 1000 x 500 cells) at full size: K_ff (pattern and values), rhs and the
    free-DOF numbering bit for bit against the oracle's assembly + partition (solver.rs:290-404, 126-137) — a few
    seconds of CPU, no CPU solve — and the TRUE residual of the device solve through mag_system_residual."""
    mesh = meshgen.plate(1000, 500)
    om = O.Mesh(mesh)
    (rp_r, col_r, val_r), rhs_r, fmap_r = O.partition(om, O.assemble_sparse(om, O.element_stiffness(om, META)), dense=False)
    with solver.System(mesh, META, ctx) as S:
        rp, col, val, rhs, fmap = S.export_kff()
        assert np.array_equal(rp, rp_r) and np.array_equal(col, col_r) and np.array_equal(fmap, fmap_r)
        assert np.array_equal(val.view(np.uint64), val_r.view(np.uint64))            # bit patterns, signed zeros included
        assert np.array_equal(rhs.view(np.uint64), rhs_r.view(np.uint64))
        sol = S.solve(_lib.default_options())
        rr, bb = S.true_residual(sol.ux, sol.uy)
        assert np.sqrt(rr / bb) <= 2e-9
        # the oracle's own row sums on the device result agree with the device's residual
        x = np.where(fmap >= 0, np.stack([sol.ux, sol.uy], 1).ravel(), 0.0)[fmap >= 0]
        r = rhs_r - O.spmv((rp_r, col_r, val_r), x)
        assert abs(np.dot(r, r) - rr) <= 1e-9 * rr and abs(np.dot(rhs_r, rhs_r) - bb) <= 1e-12 * bb


def test_signed_zero_entries_follow_the_reference(ctx):
    """The reference adds every contribution to a zeroed dense entry (solver.rs:295-296, 304-323): a lone -0.0
    contribution is stored as +0.0.  Clockwise triangles produce -0.0 in B (0/den with den < 0): the full K must
    match the oracle's bit patterns, not just compare equal."""
    def flipped(m):                       # what check_ccw (mesher.rs:522-526) does when every area is < 1
        m = m.copy()
        m.n0, m.n2 = m.n2.copy(), m.n0.copy()
        return m

    for mesh in (flipped(meshgen.plate(12, 7, h=0.5)), flipped(meshgen.jitter(meshgen.plate(9, 6))), meshgen.plate(9, 6)):
        om = O.Mesh(mesh)
        full_ref = O.assemble_sparse(om, O.element_stiffness(om, META))
        for assembly in (0, 1):
            with solver.System(mesh, META, ctx, _lib.default_options(assembly=assembly)) as S:
                rp, col, val = S.export_full()
            assert np.array_equal(rp, full_ref[0]) and np.array_equal(col, full_ref[1])
            assert np.array_equal(val.view(np.uint64), full_ref[2].view(np.uint64)), f"assembly {assembly}"


@pytest.mark.parametrize("R", [2, 3, 5, 8])
def test_virtual_rank_partition_matches_single_gpu(ctx, R):
    """The multi-GPU path (row blocks, redundant boundary elements, peer halo stores, allreduced
    dots) emulated with R virtual ranks on one GPU against the single-rank solve and the oracle."""
    mesh = meshgen.jitter(meshgen.plate(48, 30))
    opt = _lib.default_options(rel_tol=1e-13)
    one = solver.solve_soa(mesh, META, ctx, opt)
    many = solver.virtual_rank_solve(mesh, META, R, ctx, opt)
    assert many.stats["converged"] == 1
    assert many.stats["nnz"] == one.stats["nnz"] and many.stats["nnz_structural"] == one.stats["nnz_structural"]
    assert rel_l2(np.concatenate([many.ux, many.uy]), np.concatenate([one.ux, one.uy])) < 1e-10
    ref = O.run(O.Mesh(mesh), META, O.cg_options(), dense=False)
    assert rel_l2(np.concatenate([many.ux, many.uy]), np.concatenate([ref["ux"], ref["uy"]])) < 1e-9
    assert np.abs(many.stress - ref["stress"]).max() / np.abs(ref["stress"]).max() < 1e-8
    f, fr = np.concatenate([many.fx, many.fy]), np.concatenate([ref["fx"], ref["fy"]])
    assert np.abs(f - fr).max() / np.abs(fr).max() < 1e-8
    assert abs(int(many.stats["iters"]) - int(one.stats["iters"])) <= 3
    again = solver.virtual_rank_solve(mesh, META, R, ctx, opt)           # deterministic for a given R
    assert again.ux.tobytes() == many.ux.tobytes() and again.stress.tobytes() == many.stress.tobytes()


def test_virtual_ranks_on_perforated_plate_and_compat_mode(ctx):
    mesh = meshgen.perforated_plate(64, 40, pitch=16, radius=4)
    ref = O.run(O.Mesh(mesh), META, O.cg_options(), dense=False)
    sol = solver.virtual_rank_solve(mesh, META, 4, ctx, compat())
    assert rel_l2(np.concatenate([sol.ux, sol.uy]), np.concatenate([ref["ux"], ref["uy"]])) < 1e-9
    assert np.abs(sol.stress - ref["stress"]).max() / np.abs(ref["stress"]).max() < 1e-8


def test_two_process_nccl_solve_matches_single_gpu(ctx):
    """Real multi-process path (NCCL + CUDA IPC) when the box has >= 2 GPUs."""
    import subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(root / "tests" / "dist_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_OK" in r.stdout


@pytest.mark.parametrize("name", ["example_tensile", "example_linkedin", "example_cover"])
def test_reference_example_configs(ctx, name):
    """BASELINE configs 1-2: the reference's example geometries (stand-in mesher, reference input
    semantics incl. check_ccw's all-clockwise tensile mesh) against the oracle in faithful-dense mode."""
    g = np.load(GOLDEN / f"{name}.npz")
    mesh = golden_mesh(g)
    meta = META.__class__(*g["material"])
    sol = solver.solve_soa(mesh, meta, ctx, compat())
    u, ur = np.concatenate([sol.ux, sol.uy]), np.concatenate([g["ux"], g["uy"]])
    assert rel_l2(u, ur) < 1e-9
    assert np.abs(sol.stress - g["stress"]).max() / np.abs(g["stress"]).max() < 1e-8
    f, fr = np.concatenate([sol.fx, sol.fy]), np.concatenate([g["fx"], g["fy"]])
    assert np.abs(f - fr).max() / np.abs(fr).max() < 1e-8
    assert sol.stats["nnz"] == int(g["nnz_ff"][0])
    assert sol.stats["negative_definite"] == (1 if name == "example_tensile" else 0)
    assert abs(int(sol.stats["iters"]) - int(g["iters"][0])) <= max(5, int(g["iters"][0]) // 20)
    # the same mesh through the north-star solver (Jacobi-PCG) and through 4 virtual ranks
    pcg = solver.solve_soa(mesh, meta, ctx, _lib.default_options(rel_tol=1e-13))
    assert rel_l2(np.concatenate([pcg.ux, pcg.uy]), ur) < 1e-9
    vr = solver.virtual_rank_solve(mesh, meta, 4, ctx, _lib.default_options(rel_tol=1e-13))
    assert rel_l2(np.concatenate([vr.ux, vr.uy]), ur) < 1e-9


def test_cpp_host_layer_end_to_end(ctx, tmp_path):
    """The C++ mirror of main.rs:54-76 (solver::run + csv_output over the C ABI) writes byte-identical
    CSVs to the Python mirror: same library, deterministic kernels, same Rust-style float formatting."""
    import subprocess
    from magnetite_b200 import post_processor
    root = Path(__file__).resolve().parent.parent
    exe = root / "host" / "plate_demo"
    subprocess.run(["make", "-C", str(root / "host")], check=True, capture_output=True)   # header may have changed
    r = subprocess.run([str(exe), "12", "6", str(tmp_path / "n_cpp.csv"), str(tmp_path / "e_cpp.csv")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "info: solve complete" in r.stdout and "area of element 0: 2" in r.stdout
    nodes, elements = meshgen.plate(12, 6).to_aos()
    solver.run(nodes, elements, META, quiet=True)
    post_processor.csv_output(elements, nodes, str(tmp_path / "n_py.csv"), str(tmp_path / "e_py.csv"), quiet=True)
    assert (tmp_path / "n_cpp.csv").read_bytes() == (tmp_path / "n_py.csv").read_bytes()
    assert (tmp_path / "e_cpp.csv").read_bytes() == (tmp_path / "e_py.csv").read_bytes()


def test_device_perforated_generator_matches_host(ctx):
    import ctypes as C
    lib = _lib.load()
    dm = C.c_void_p()
    _lib.check(lib.mag_devmesh_perforated(ctx.handle, 150, 70, 2.0, 32, 8, 3.0, C.byref(dm)), "perforated")
    view = _lib.MagMesh()
    _lib.check(lib.mag_devmesh_view(dm, C.byref(view)), "view")
    host = meshgen.perforated_plate(150, 70, 2.0, pitch=32, radius=8).normalised()
    assert (view.n_nodes, view.n_elems) == (host.n_nodes, host.n_elems) and host.n_elems < 2 * 150 * 70
    N, E = host.n_nodes, host.n_elems
    d = MeshSoA(np.empty(N), np.empty(N), np.empty(E, np.uint32), np.empty(E, np.uint32), np.empty(E, np.uint32),
                np.empty(N), np.empty(N), np.empty(N), np.empty(N), np.empty(N, np.uint8))
    _lib.check(lib.mag_devmesh_download(dm, *(_lib.ptr(a) for a in (d.x, d.y, d.n0, d.n1, d.n2, d.ux, d.uy, d.fx, d.fy, d.known))), "download")
    for k in ("x", "y", "n0", "n1", "n2", "ux", "uy", "fx", "fy", "known"):
        assert np.array_equal(getattr(d, k), getattr(host, k)), k
    lib.mag_devmesh_free(dm)


TWO_LEVEL_MESHES = {
    "plate_120x60": lambda: meshgen.plate(120, 60),
    "jitter_90x50": lambda: meshgen.jitter(meshgen.plate(90, 50)),
    "perforated_128x64": lambda: meshgen.perforated_plate(128, 64, pitch=32, radius=8),
}


@pytest.mark.parametrize("name", list(TWO_LEVEL_MESHES))
def test_two_level_preconditioner_matches_oracle(ctx, name):
    """precond=2 (Jacobi + aggregation coarse space): same answer as the oracle,
    far fewer iterations than Jacobi, bit-identical run to run, also over virtual ranks."""
    mesh = TWO_LEVEL_MESHES[name]()
    ref = O.run(O.Mesh(mesh), META, O.cg_options(), dense=False)
    ur = np.concatenate([ref["ux"], ref["uy"]])
    with solver.System(mesh, META, ctx) as S:
        jac = S.solve(_lib.default_options(rel_tol=1e-13))
        two = S.solve(_lib.default_options(rel_tol=1e-13, precond=2, coarse_aggregates=32))
        again = S.solve(_lib.default_options(rel_tol=1e-13, precond=2, coarse_aggregates=32))
    assert two.stats["converged"] == 1
    assert rel_l2(np.concatenate([two.ux, two.uy]), ur) < 1e-9
    assert np.abs(two.stress - ref["stress"]).max() / np.abs(ref["stress"]).max() < 1e-8
    assert two.stats["iters"] < 0.6 * jac.stats["iters"], (two.stats["iters"], jac.stats["iters"])
    assert again.ux.tobytes() == two.ux.tobytes() and again.stats["iters"] == two.stats["iters"]
    vr = solver.virtual_rank_solve(mesh, META, 3, ctx, _lib.default_options(rel_tol=1e-13, precond=2, coarse_aggregates=32))
    assert rel_l2(np.concatenate([vr.ux, vr.uy]), ur) < 1e-9
    assert abs(int(vr.stats["iters"]) - int(two.stats["iters"])) <= max(3, two.stats["iters"] // 20)


def test_two_level_on_examples_and_indefinite_meshes(ctx):
    g = np.load(GOLDEN / "example_linkedin.npz")
    mesh = golden_mesh(g)
    meta = META.__class__(*g["material"])
    two = solver.solve_soa(mesh, meta, ctx, _lib.default_options(rel_tol=1e-13, precond=2))
    assert rel_l2(np.concatenate([two.ux, two.uy]), np.concatenate([g["ux"], g["uy"]])) < 1e-9
    assert two.stats["precond_used"] == 2
    # all-clockwise mesh: K is negative definite, the coarse matrix has no Cholesky factor: Jacobi takes over
    t = np.load(GOLDEN / "example_tensile.npz")
    fb = solver.solve_soa(golden_mesh(t), META.__class__(*t["material"]), ctx, _lib.default_options(rel_tol=1e-13, precond=2))
    assert fb.stats["precond_used"] == 1 and fb.stats["converged"] == 1 and fb.stats["negative_definite"] == 1
    assert rel_l2(np.concatenate([fb.ux, fb.uy]), np.concatenate([t["ux"], t["uy"]])) < 1e-9
    # boxes smaller than the elements (2048 aggregates on an 800-cell plate): no 9-point coarse stencil, Jacobi again
    fine = solver.solve_soa(meshgen.plate(40, 20), META, ctx, _lib.default_options(precond=2, coarse_aggregates=2048))
    assert fine.stats["precond_used"] == 1 and fine.stats["converged"] == 1
    # the default (precond 3 = auto): Jacobi below 20 000 unknowns, two-level above
    small = solver.solve_soa(meshgen.plate(40, 20), META, ctx, _lib.default_options())
    large = solver.solve_soa(meshgen.plate(160, 80), META, ctx, _lib.default_options())
    assert small.stats["precond_used"] == 1 and large.stats["precond_used"] == 2


def test_single_cluster_solve_matches_the_general_path(ctx):
    """Systems that fit one thread-block cluster's shared memory are solved by ONE kernel (small.cuh); MAG_TUNE=128
    switches it off.  Both paths run the same recurrence: same iteration counts (to rounding), same answers, for
    the reference's semantics (compat, with and without a max_iters end), Jacobi-PCG, and a negative-definite mesh."""
    import os
    os.environ["MAG_TUNE"] = "128"
    try:
        general = _lib.Context(0)
    finally:
        del os.environ["MAG_TUNE"]
    try:
        g = np.load(GOLDEN / "example_linkedin.npz")
        t = np.load(GOLDEN / "example_tensile.npz")
        cases = [(golden_mesh(g), META.__class__(*g["material"])), (golden_mesh(t), META.__class__(*t["material"])),
                 (meshgen.jitter(meshgen.plate(40, 20)), META), (meshgen.plate(3, 2), META)]
        for mesh, meta in cases:
            for opt in (dict(compat=1), dict(precond=1, rel_tol=1e-12), dict(compat=1, max_iter=37), dict(precond=0)):
                a = solver.solve_soa(mesh, meta, ctx, _lib.default_options(**opt)) if "max_iter" not in opt else None
                if a is None:                      # max_iter ends the run: both paths return argmin's best_param
                    with solver.System(mesh, meta, ctx) as S:
                        a = S.solve(_lib.default_options(**opt), allow_not_converged=True)
                    with solver.System(mesh, meta, general) as S:
                        b = S.solve(_lib.default_options(**opt), allow_not_converged=True)
                else:
                    b = solver.solve_soa(mesh, meta, general, _lib.default_options(**opt))
                assert a.stats["kernel_launches"] < b.stats["kernel_launches"]            # one CG kernel against a graph of them
                assert abs(int(a.stats["iters"]) - int(b.stats["iters"])) <= max(2, int(b.stats["iters"]) // 50), opt
                assert a.stats["converged"] == b.stats["converged"] and a.stats["negative_definite"] == b.stats["negative_definite"]
                ua, ub = np.concatenate([a.ux, a.uy]), np.concatenate([b.ux, b.uy])
                tol = 1e-9 if a.stats["converged"] else 1e-6
                assert rel_l2(ua, ub) < tol, (opt, rel_l2(ua, ub))
        again = [solver.solve_soa(cases[0][0], cases[0][1], ctx, _lib.default_options(compat=1)) for _ in range(2)]
        assert again[0].ux.tobytes() == again[1].ux.tobytes()                               # bit-identical run to run
    finally:
        general.close()


def test_programmatic_and_plain_launches_of_the_loop_agree(ctx):
    """The kernels of the CG loop are launched as programmatic dependents inside the graph (common.cuh:
    MAG_LAUNCH_DEP, every kernel starts with griddepcontrol.wait); MAG_TUNE=512 launches them plainly.  Host arrays
    of 65 536 nodes and more move their prescribed values in and their nodal results out on a side stream beside
    the kernels (system.cuh, solve.cuh: aux_fork / aux_join); MAG_TUNE=2048 keeps everything on the run stream.
    Ordering is the only thing either switch changes, so every bit of the result must be the same — Jacobi and
    two-level, one GPU and three virtual ranks."""
    import os
    os.environ["MAG_TUNE"] = str(512 + 2048)
    try:
        plain = _lib.Context(0)
    finally:
        del os.environ["MAG_TUNE"]
    try:
        mesh = meshgen.jitter(meshgen.plate(400, 180))
        assert mesh.n_nodes >= 65536
        for opt in (dict(precond=1), dict(precond=2, coarse_aggregates=64)):
            a = solver.solve_soa(mesh, META, ctx, _lib.default_options(**opt))
            b = solver.solve_soa(mesh, META, plain, _lib.default_options(**opt))
            va = solver.virtual_rank_solve(mesh, META, 3, ctx, _lib.default_options(**opt))
            vb = solver.virtual_rank_solve(mesh, META, 3, plain, _lib.default_options(**opt))
            for x, y in ((a, b), (va, vb)):
                assert x.stats["iters"] == y.stats["iters"] and x.stats["converged"] == 1
                for k in ("ux", "uy", "fx", "fy", "stress"):
                    assert getattr(x, k).tobytes() == getattr(y, k).tobytes(), (opt, k)
    finally:
        plain.close()


def test_narrow_and_wide_sell_index_streams_agree(ctx):
    """16-bit column offsets (banded numbering) and 32-bit absolute columns carry the same matrix:
    identical SpMV results and identical CG iterates; unstructured numbering falls back to 32 bits."""
    mesh = meshgen.jitter(meshgen.plate(150, 40))
    x = np.random.default_rng(5).normal(size=2 * mesh.n_nodes)
    with solver.System(mesh, META, ctx) as N, solver.System(mesh, META, ctx, options=_lib.default_options(spmv_format=3)) as W:
        assert N.assemble_stats["sell_index_bits"] == 16 and W.assemble_stats["sell_index_bits"] == 32
        xs = x[: N.n_free]
        assert rel_l2(N.spmv(xs, fmt=2), W.spmv(xs, fmt=2)) < 1e-15      # same products, remainder summed in another order
        a, b = N.solve(_lib.default_options()), W.solve(_lib.default_options(spmv_format=3))
        assert rel_l2(np.concatenate([a.ux, a.uy]), np.concatenate([b.ux, b.uy])) < 1e-7
        assert abs(int(a.stats["iters"]) - int(b.stats["iters"])) <= 3
    long_plate = meshgen.plate(16500, 1)                           # band 2*(nx+1)+2 > 32767: falls back
    with solver.System(long_plate, META, ctx) as S:
        assert S.assemble_stats["sell_index_bits"] == 32
        xs = np.random.default_rng(6).normal(size=S.n_free)
        rp, col, val, rhs, fmap = S.export_kff()
        assert rel_l2(S.spmv(xs, fmt=2), O.spmv((rp, col, val), xs)) < 1e-14


def _random_bc_mesh(seed, nx=26, ny=17):
    """Jittered plate with a random, per-DOF consistent boundary pattern: random nodes get prescribed
    displacements (possibly one axis only) and random free DOFs get non-zero applied forces."""
    rng = np.random.default_rng(seed)
    m = meshgen.jitter(meshgen.plate(nx, ny), seed=seed).copy()
    n = m.n_nodes
    m.known[:] = 4 | 8                                   # everything free, forces known (= 0)
    m.ux[:] = 0; m.uy[:] = 0; m.fx[:] = 0; m.fy[:] = 0
    fixed = rng.choice(n, size=max(4, n // 12), replace=False)
    for i in fixed:
        kind = rng.integers(0, 3)
        if kind in (0, 2):
            m.known[i] = (int(m.known[i]) & 0xfb) | 1; m.ux[i] = rng.normal() * 1e-3
        if kind in (1, 2):
            m.known[i] = (int(m.known[i]) & 0xf7) | 2; m.uy[i] = rng.normal() * 1e-3
    loaded = rng.choice(np.setdiff1d(np.arange(n), fixed), size=n // 10, replace=False)
    m.fx[loaded] = rng.normal(size=loaded.size) * 1e6
    m.fy[loaded] = rng.normal(size=loaded.size) * 1e6
    m.known[0] = 3; m.ux[0] = m.uy[0] = 0.0; m.known[1] = 3   # no rigid-body mode left
    m.fx[[0, 1]] = 0; m.fy[[0, 1]] = 0
    return m


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_boundary_patterns_and_force_loads(ctx, seed):
    mesh = _random_bc_mesh(seed)
    om = O.Mesh(mesh)
    full_ref = O.assemble_sparse(om, O.element_stiffness(om, META))
    (rp_r, col_r, val_r), rhs_r, fmap_r = O.partition(om, full_ref, dense=False)
    with solver.System(mesh, META, ctx) as S:
        rp, col, val, rhs, fmap = S.export_kff()
        assert np.array_equal(rp, rp_r) and np.array_equal(col, col_r) and np.array_equal(fmap, fmap_r)
        assert np.array_equal(val, val_r) and np.array_equal(rhs, rhs_r)          # applied forces enter the rhs
        sol = S.solve(compat())
    ref = O.run(om, META, O.cg_options(), dense=True)
    assert rel_l2(np.concatenate([sol.ux, sol.uy]), np.concatenate([ref["ux"], ref["uy"]])) < 1e-9
    assert np.abs(sol.stress - ref["stress"]).max() / np.abs(ref["stress"]).max() < 1e-8
    f, fr = np.concatenate([sol.fx, sol.fy]), np.concatenate([ref["fx"], ref["fy"]])
    assert np.abs(f - fr).max() / np.abs(fr).max() < 1e-8
    vr = solver.virtual_rank_solve(mesh, META, 3, ctx, compat())
    assert rel_l2(np.concatenate([vr.ux, vr.uy]), np.concatenate([ref["ux"], ref["uy"]])) < 1e-9


def test_rows_by_force_columns_by_displacement(ctx):
    """The reference picks ROWS by `force known` and COLUMNS by `displacement unknown`
    (solver.rs:380-396) and only needs the two COUNTS to agree.  A DOF with both known plus a DOF with
    neither keeps the counts equal while the two sets differ: the elimination must still reproduce the
    oracle's (non-symmetric) K_ff, rhs and numbering bit for bit."""
    mesh = meshgen.jitter(meshgen.plate(9, 6)).copy()
    a, b = 25, 46                                       # interior nodes with fx = fy = Some(0)
    mesh.known[a] = 1 | 4 | 8; mesh.ux[a] = 2e-3        # x: displacement AND force known
    mesh.known[b] = 8                                   # x: neither known
    om = O.Mesh(mesh)
    full_ref = O.assemble_sparse(om, O.element_stiffness(om, META))
    (rp_r, col_r, val_r), rhs_r, fmap_r = O.partition(om, full_ref, dense=False)
    with solver.System(mesh, META, ctx) as S:
        rp, col, val, rhs, fmap = S.export_kff()
    assert np.array_equal(rp, rp_r) and np.array_equal(col, col_r) and np.array_equal(val, val_r)
    assert np.array_equal(rhs, rhs_r) and np.array_equal(fmap, fmap_r)
    import scipy.sparse as sp
    A = sp.csr_matrix((val, col, rp), shape=(len(rhs), len(rhs)))
    assert abs(A - A.T).max() > 0                        # really the unsymmetric case


def test_public_element_functions(ctx):
    """The pub functions of solver.rs besides run(): area, B and D, against the numpy twin."""
    from oracle import reference_semantics as R
    mesh = meshgen.jitter(meshgen.plate(6, 4))
    Bs = solver.strain_displacement_matrices(mesh, ctx)
    conn = np.stack([mesh.n0, mesh.n1, mesh.n2], 1).astype(int)
    for e in (0, 7, len(conn) - 1):
        assert np.array_equal(Bs[e], R.B_matrix(mesh.x, mesh.y, conn[e], R.area(mesh.x, mesh.y, conn[e])))
    nodes, elements = mesh.to_aos()
    assert np.array_equal(solver.compute_strain_displacement_matrix(elements[3], nodes), Bs[3])
    assert np.array_equal(solver.compute_stress_strain_matrix(0.33, 69e9), R.D_matrix(0.33, 69e9))


@pytest.mark.parametrize("max_iter", [0, 1, 20, 21, 33, 58])
def test_compat_returns_best_param_when_max_iters_ends_the_run(ctx, max_iter):
    """argmin's Executor hands back best_param (solver.rs:167-176): when max_iters ends the run that is
    the iterate with the lowest residual norm so far, not necessarily the last one (CG's 2-norm is not
    monotone: at 21, 33 and 58 iterations on this mesh the last iterate is worse than an earlier one).
    Like the reference, compat mode does not treat that as an error."""
    mesh = meshgen.jitter(meshgen.plate(30, 15))
    ref = O.run(O.Mesh(mesh), META, O.cg_options(max_iter=max_iter), dense=False)
    sol = solver.solve_soa(mesh, META, ctx, _lib.default_options(compat=1, max_iter=max_iter))
    assert sol.stats["iters"] == max_iter and sol.stats["converged"] == 0
    ur = np.concatenate([ref["ux"], ref["uy"]])
    assert rel_l2(np.concatenate([sol.ux, sol.uy]), ur) < 1e-9
    assert abs(sol.stats["final_residual"] - ref["stats"]["final_cost"]) <= 1e-9 * ref["stats"]["final_cost"]


def test_cpp_cli_full_flow_matches_python_flow(ctx, tmp_path):
    """main.rs:54-76 in C++ (parse_mesh -> check_ccw -> apply_boundary_conditions -> solver::run ->
    csv_output) against the same flow through the Python mirror: byte-identical CSVs.  The mesh has
    triangles of area 0.5, so check_ccw flips every element (the tensile-example quirk)."""
    import subprocess
    from magnetite_b200 import geometry, mesher, post_processor
    root = Path(__file__).resolve().parent.parent
    subprocess.run(["make", "-C", str(root / "host")], check=True, capture_output=True)
    m = meshgen.jitter(meshgen.plate(22, 9, h=1.0), frac=0.1)
    xs = m.x - 11.0
    conn = np.stack([m.n0, m.n1, m.n2], 1).astype(int)
    geometry.write_msh(str(tmp_path / "geom.msh"), xs, m.y, conn)
    inp = str(GOLDEN / "tensile_input.json")
    r = subprocess.run([str(root / "host" / "magnetite_b200"), inp, "geom.msh", "--skip"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "info: loaded 230 nodes and 396 elements" in r.stdout and "info: loaded 2 boundary rules" in r.stdout
    # the same flow through the Python mirror
    nodes, elements = geometry.parse_mesh(str(tmp_path / "geom.msh"))
    mesher.check_ccw(elements, nodes)
    assert all(e.nodes != list(c) for e, c in zip(elements, conn))          # every element was flipped
    data = mesher.load_input_file(inp)
    mesher.apply_boundary_conditions(data, nodes)
    solver.run(nodes, elements, mesher.parse_input_metadata(data), quiet=True)
    post_processor.csv_output(elements, nodes, str(tmp_path / "n_py.csv"), str(tmp_path / "e_py.csv"), quiet=True)
    assert (tmp_path / "nodes.csv").read_bytes() == (tmp_path / "n_py.csv").read_bytes()
    assert (tmp_path / "elements.csv").read_bytes() == (tmp_path / "e_py.csv").read_bytes()
    assert max(n.ux for n in nodes) == 3.0 and min(n.ux for n in nodes) == 0.0
    # --reorder: renumbered around the solve, CSVs in the mesh file's numbering, same values to rounding
    (tmp_path / "ro").mkdir()
    r = subprocess.run([str(root / "host" / "magnetite_b200"), inp, str(tmp_path / "geom.msh"), "--skip", "--reorder"],
                       cwd=tmp_path / "ro", capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "info: node band " in r.stdout, r.stdout + r.stderr
    a = np.loadtxt(tmp_path / "nodes.csv", delimiter=",", skiprows=1)
    b = np.loadtxt(tmp_path / "ro" / "nodes.csv", delimiter=",", skiprows=1)
    assert np.array_equal(a[:, :2], b[:, :2]) and rel_l2(b[:, 2:], a[:, 2:]) < 1e-9
    ea = np.loadtxt(tmp_path / "elements.csv", delimiter=",", skiprows=1)
    eb = np.loadtxt(tmp_path / "ro" / "elements.csv", delimiter=",", skiprows=1)
    assert np.array_equal(ea[:, :3], eb[:, :3]) and np.abs(eb[:, 3] - ea[:, 3]).max() < 1e-8 * np.abs(ea[:, 3]).max()


def test_rcm_reordered_solve_keeps_the_callers_numbering(ctx):
    """a mesh whose node ids carry no locality (gmsh order) is renumbered on the host
    (mag_reorder_rcm) before the solve and the results come back in the caller's numbering.  The
    shuffled plate has a band of ~36 k reduced DOFs (32-bit SELL columns); after RCM it is banded
    again (16-bit offsets).  Same answers within the north-star tolerances, and equal to the solve of
    the plate in its natural numbering."""
    from magnetite_b200 import reorder
    base = meshgen.jitter(meshgen.plate(180, 100))
    perm = np.random.default_rng(21).permutation(base.n_nodes).astype(np.uint32)
    mesh = reorder.permute_mesh(base, perm)
    opt = _lib.default_options(rel_tol=1e-12)
    plain = solver.solve_soa(mesh, META, ctx, opt)
    ro = solver.solve_soa(mesh, META, ctx, opt, reorder=True)
    info = ro.stats["reorder"]
    assert info["applied"] and info["band_before"] > base.n_nodes // 2 and info["band_after"] <= 110
    assert plain.stats["sell_index_bits"] == 32 and ro.stats["sell_index_bits"] == 16
    assert ro.stats["nnz"] == plain.stats["nnz"] and ro.stats["n_free"] == plain.stats["n_free"]
    u_plain, u_ro = np.concatenate([plain.ux, plain.uy]), np.concatenate([ro.ux, ro.uy])
    assert rel_l2(u_ro, u_plain) < 1e-9
    assert np.abs(ro.stress - plain.stress).max() / np.abs(plain.stress).max() < 1e-8
    f_plain, f_ro = np.concatenate([plain.fx, plain.fy]), np.concatenate([ro.fx, ro.fy])
    assert np.abs(f_ro - f_plain).max() / np.abs(f_plain).max() < 1e-7
    nat = solver.solve_soa(base, META, ctx, opt)                       # the plate as generated
    # base node i is node perm[i] of the shuffled mesh
    assert rel_l2(np.concatenate([ro.ux[perm], ro.uy[perm]]), np.concatenate([nat.ux, nat.uy])) < 1e-9
    # a structured plate already numbered along its short side is left alone
    tall = solver.solve_soa(meshgen.plate(12, 40), META, ctx, opt, reorder=True)
    assert not tall.stats["reorder"]["applied"]
    # the drop-in entry point takes the same switch and fills the caller's lists in their order
    nodes, elements = reorder.permute_mesh(meshgen.plate(14, 9), np.random.default_rng(2).permutation(150).astype(np.uint32)).to_aos()
    nodes2, elements2 = [Node(Vertex(n.vertex.x, n.vertex.y), n.ux, n.uy, n.fx, n.fy) for n in nodes], [Element(list(e.nodes)) for e in elements]
    solver.run(nodes, elements, META, options=opt, quiet=True)
    solver.run(nodes2, elements2, META, options=opt, quiet=True, reorder=True)
    assert rel_l2(np.array([n.ux for n in nodes2] + [n.uy for n in nodes2]), np.array([n.ux for n in nodes] + [n.uy for n in nodes])) < 1e-9
    assert np.abs(np.array([e.stress for e in elements2]) - np.array([e.stress for e in elements])).max() < 1e-8 * max(abs(e.stress) for e in elements)
