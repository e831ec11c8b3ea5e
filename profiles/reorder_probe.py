"""What the host-side RCM renumbering (mag_reorder_rcm) buys on the GPU: a plate whose node
ids were shuffled (no locality, like gmsh order) solved as numbered and after renumbering.

    python profiles/reorder_probe.py [nx ny]        (default 1000 500 = BASELINE config 3's size)
    python profiles/reorder_probe.py nx ny --spmv-only     shuffled vs renumbered, no solves (for 4000 2000)

Prints one JSON line per variant: band, SELL index width, isolated SpMV time (mag_system_spmv_bench),
Jacobi-PCG iterations and solve time to 1e-9.  Timing is the library's own (CUDA events on its stream)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from magnetite_b200 import _lib, meshgen, reorder, solver  # noqa: E402


def measure(name, mesh, ctx, extra, solve=True):
    with solver.System(mesh, meshgen.EXAMPLE_MATERIAL, ctx) as S:
        ms_spmv, nbytes = S.spmv_bench(reps=200 if solve else 50)
        row = {"variant": name, "n_free": S.n_free, "nnz": S.nnz, "node_band": reorder.mesh_band(mesh),
               "sell_index_bits": int(S.assemble_stats["sell_index_bits"]),
               "assembly_ms": round(float(S.assemble_stats["ms_total"]), 3),
               "spmv_ms": round(ms_spmv, 5), "spmv_algorithmic_bytes": nbytes,
               "spmv_gbs": round(nbytes / (ms_spmv * 1e-3) / 1e9, 1)}
        row.update(extra)
        if not solve:
            print(json.dumps(row), flush=True)
            return None
        sol = S.solve(_lib.default_options())
        st = sol.stats
        row.update({"pcg_iters": int(st["iters"]), "pcg_solve_ms": round(float(st["ms_solve"]), 2),
                    "rel_residual": float(st["final_residual"]) / float(st["b_norm"])})
        print(json.dumps(row), flush=True)
        return np.concatenate([sol.ux, sol.uy])


def main():
    nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 500)
    ctx = _lib.Context(0)
    base = meshgen.plate(nx, ny)
    perm = np.random.default_rng(7).permutation(base.n_nodes).astype(np.uint32)
    shuffled = reorder.permute_mesh(base, perm)
    if "--spmv-only" in sys.argv:
        del base
        measure("node ids shuffled", shuffled, ctx, {}, solve=False)
        t = time.perf_counter()
        new_of_old, before, after = reorder.rcm(shuffled)
        t_rcm = time.perf_counter() - t
        measure("shuffled ids renumbered by mag_reorder_rcm", reorder.permute_mesh(shuffled, new_of_old), ctx,
                {"rcm_host_s": round(t_rcm, 3), "band_before": before, "band_after": after}, solve=False)
        ctx.close()
        return
    u_nat = measure("plate as generated (row-major ids)", base, ctx, {})
    u_shuf = measure("node ids shuffled", shuffled, ctx, {})
    t = time.perf_counter()
    new_of_old, before, after = reorder.rcm(shuffled)
    t_rcm = time.perf_counter() - t
    t = time.perf_counter()
    renumbered = reorder.permute_mesh(shuffled, new_of_old)
    t_perm = time.perf_counter() - t
    u_rcm = measure("shuffled ids renumbered by mag_reorder_rcm", renumbered, ctx,
                    {"rcm_host_s": round(t_rcm, 3), "permute_host_s": round(t_perm, 3), "band_before": before,
                     "band_after": after})
    n = base.n_nodes
    back = lambda u, p: np.concatenate([u[:n][p], u[n:][p]])          # noqa: E731  results in the caller's numbering
    ref = np.linalg.norm(u_nat)
    print(json.dumps({"rel_l2_shuffled_vs_natural": float(np.linalg.norm(back(u_shuf, perm) - u_nat) / ref),
                      "rel_l2_rcm_vs_natural": float(np.linalg.norm(back(back(u_rcm, new_of_old), perm) - u_nat) / ref)}))
    ctx.close()


if __name__ == "__main__":
    main()
