"""torchrun worker: every rank solves the same plate through the comm-aware C ABI (row-block
partition, NCCL allreduce, IPC halo stores); rank 0 compares with a single-GPU virtual-rank solve
and the oracle.  Launched by tests/test_gpu_parity.py and by hand:

    torchrun --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 tests/dist_gpu_check.py
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from magnetite_b200 import _lib, dist as mdist, meshgen, solver  # noqa: E402


def main():
    rank, world, local = mdist.init_process_group("nccl")
    ctx = _lib.Context(local)
    mdist.init_comm(ctx)
    meta = meshgen.EXAMPLE_MATERIAL
    cases = [(meshgen.jitter(meshgen.plate(96, 64)), 0, 1), (meshgen.plate(300, 200), 0, 1), (meshgen.plate(300, 200), 1, 1),
             (meshgen.plate(300, 200), 0, 2), (meshgen.perforated_plate(256, 128, pitch=32, radius=8), 1, 2)]
    for mesh, allreduce, precond in cases:     # allreduce 0: peer-memory mailbox, 1: NCCL; precond 2: two-level
        opt = _lib.default_options(rel_tol=1e-12, allreduce=allreduce, precond=precond, coarse_aggregates=64 if precond == 2 else 0)
        sol = solver.solve_soa(mesh, meta, ctx, opt)            # comm-aware: this rank's row block
        sol2 = solver.solve_soa(mesh, meta, ctx, opt)
        assert sol.ux.tobytes() == sol2.ux.tobytes(), "not deterministic run to run"
        if rank == 0:
            solo = _lib.Context(local)                          # no communicator: plain single-GPU solve
            one = solver.solve_soa(mesh, meta, solo, opt)
            u, u1 = np.concatenate([sol.ux, sol.uy]), np.concatenate([one.ux, one.uy])
            err = np.linalg.norm(u - u1) / np.linalg.norm(u1)
            ferr = np.abs(np.concatenate([sol.fx - one.fx, sol.fy - one.fy])).max() / np.abs(one.fx).max()
            serr = np.abs(sol.stress - one.stress).max() / np.abs(one.stress).max()
            # both runs stopped on the same criterion (the residual at the stopping iteration is rounding noise at
            # 1e-12 relative, so its VALUE differs between summation orders; the bound is what must hold)
            for s in (sol, one):
                assert s.stats["converged"] == 1 and s.stats["final_residual"] <= 1e-12 * s.stats["b_norm"]
            # the one-process emulation of the same partition (virtual ranks): same kernels, same summation order
            emu = solver.virtual_rank_solve(mesh, meta, world, solo, opt)
            ue = np.concatenate([emu.ux, emu.uy])
            err_emu = np.linalg.norm(u - ue) / np.linalg.norm(ue)
            print(f"rank0: {mesh.n_elems} elements, world {world}, allreduce {allreduce}, precond {precond}: iters {sol.stats['iters']} "
                  f"(1 GPU: {one.stats['iters']}, emulation: {emu.stats['iters']}), |du| {err:.2e}, |df| {ferr:.2e}, |ds| {serr:.2e}, "
                  f"vs emulation |du| {err_emu:.2e} (bit-identical: {u.tobytes() == ue.tobytes()})", flush=True)
            assert err < 1e-9 and ferr < 1e-7 and serr < 1e-8
            assert abs(int(sol.stats["iters"]) - int(one.stats["iters"])) <= max(5, int(one.stats["iters"]) // 100)
            assert err_emu < 1e-9 and abs(int(sol.stats["iters"]) - int(emu.stats["iters"])) <= max(5, int(emu.stats["iters"]) // 100)
            solo.close()
    # mag_options.result_scope = 1: a rank writes only its slice of the result arrays (nodes by mag_partition_nodes,
    # elements split the same way); the slices of all ranks together are the complete result
    import ctypes as C
    mesh = meshgen.plate(300, 200).normalised()
    full = solver.solve_soa(mesh, meta, ctx, _lib.default_options(rel_tol=1e-12))
    n, e = mesh.n_nodes, mesh.n_elems
    part = {k: np.full(n, np.nan) for k in ("ux", "uy", "fx", "fy")}
    part["stress"] = np.full(e, np.nan)
    res = _lib.MagResult(*[_lib.ptr(part[k]) for k in ("ux", "uy", "fx", "fy", "stress")], None, 0)
    ms, mat = solver._mesh_struct(mesh), solver._material(meta)
    opt = _lib.default_options(rel_tol=1e-12, result_scope=1)
    _lib.check(_lib.load().mag_solve(ctx.handle, C.byref(ms), C.byref(mat), C.byref(opt), C.byref(res), None), "mag_solve(scope 1)")
    nlo, nhi = mdist.partition_nodes(n, world, rank)
    elo, ehi = mdist.partition_nodes(e, world, rank)
    for k in ("ux", "uy", "fx", "fy"):
        assert np.array_equal(part[k][nlo:nhi], getattr(full, k)[nlo:nhi]), k
        assert np.isnan(part[k][:nlo]).all() and np.isnan(part[k][nhi:]).all(), k
    assert np.array_equal(part["stress"][elo:ehi], full.stress[elo:ehi])
    assert np.isnan(part["stress"][:elo]).all() and np.isnan(part["stress"][ehi:]).all()
    import torch.distributed as dist
    dist.barrier()
    if rank == 0:
        print("result_scope=1: every rank wrote exactly its slice, bit-identical to the complete result", flush=True)
        print("DIST_OK", flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
