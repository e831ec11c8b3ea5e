"""BASELINE configs 1-2 on the GPU: the reference's example geometries (stand-in mesher fixtures under
tests/golden/) through mag_solve with the reference's solver semantics (compat: plain CG, x0 = 0, absolute
cost <= 1e-4) and through the north-star Jacobi-PCG (1e-9).  The only number the reference publishes is the
linkedin "solve" time, 0.286 s on the author's laptop (readme.md:28; its timer brackets dense->CSR + CG).

    python profiles/example_probe.py

One JSON line per (example, mode): the library's own timers (CUDA events), third of three calls."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from magnetite_b200 import _lib, meshgen, solver  # noqa: E402
from magnetite_b200.datatypes import MeshSoA  # noqa: E402


def main():
    ctx = _lib.Context(0)
    for name in ("example_linkedin", "example_tensile", "example_cover"):
        g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
        mesh = MeshSoA(g["x"], g["y"], g["n0"], g["n1"], g["n2"], g["bc_ux"], g["bc_uy"], g["bc_fx"], g["bc_fy"], g["known"])
        meta = meshgen.EXAMPLE_MATERIAL.__class__(*g["material"])
        for mode, opt in (("reference semantics (plain CG, cost <= 1e-4)", dict(compat=1)), ("Jacobi-PCG 1e-9", dict())):
            for _ in range(3):
                sol = solver.solve_soa(mesh, meta, ctx, _lib.default_options(**opt))
            st = sol.stats
            u, ur = np.concatenate([sol.ux, sol.uy]), np.concatenate([g["ux"], g["uy"]])
            print(json.dumps({"example": name, "mode": mode, "nodes": mesh.n_nodes, "elements": mesh.n_elems,
                              "n_free": int(st["n_free"]), "iters": int(st["iters"]),
                              "assembly_ms": round(float(st["ms_elem"] + st["ms_sort"] + st["ms_reduce"] + st["ms_bc"] + st["ms_format"]), 3),
                              "solve_ms": round(float(st["ms_solve"]), 3), "whole_call_ms": round(float(st["ms_total"]), 3),
                              "rel_l2_vs_golden": float(np.linalg.norm(u - ur) / np.linalg.norm(ur))}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
