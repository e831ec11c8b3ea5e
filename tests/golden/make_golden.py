"""Generates tests/golden/*.npz from the oracle (oracle/magnetite_oracle.c).

The reference ships no golden vectors and cannot be run here (PARITY UNPINNED, see
oracle/magnetite_oracle.h), so these fixtures pin the ORACLE, not the Rust binary: they freeze
what the restatement produced when the suite was written, so later edits to the oracle or to
the CUDA path cannot drift together unnoticed.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from magnetite_b200 import meshgen  # noqa: E402
from magnetite_b200.datatypes import MeshSoA  # noqa: E402
from oracle import oracle as O  # noqa: E402

META = meshgen.EXAMPLE_MATERIAL


def flipped(mesh: MeshSoA) -> MeshSoA:
    """What check_ccw (reference src/mesher.rs:522-526) does to a mesh whose triangles have area
    < 1: every element's node order is reversed (clockwise; K becomes negative definite)."""
    m = mesh.copy()
    m.n0, m.n2 = m.n2.copy(), m.n0.copy()
    return m


def cases():
    yield "patch_2x2", meshgen.patch_square(2.0, 0.005)
    yield "plate_8x6", meshgen.plate(8, 6)
    yield "plate_jitter_10x7", meshgen.jitter(meshgen.plate(10, 7))
    yield "perforated_24x16", meshgen.perforated_plate(24, 16, pitch=8, radius=2)
    yield "clockwise_unit", flipped(meshgen.patch_square(1.0, 0.005))
    # force-driven load on a clockwise mesh (SURVEY KAT-4): sign of u and stress flips
    m = flipped(meshgen.patch_square(1.0, 0.0))
    m.known[1] = m.known[2] = 4 | 8          # fx, fy known
    m.fx[1] = m.fx[2] = 1e6
    m.known[3] = 1 | 8
    yield "clockwise_force", m


def main():
    out = Path(__file__).resolve().parent
    for name, mesh in cases():
        om = O.Mesh(mesh)
        ke = O.element_stiffness(om, META)
        full = O.assemble_sparse(om, ke)
        (rp, col, val), rhs, fmap = O.partition(om, full, dense=False)
        res = O.run(om, META, O.cg_options(), dense=False)
        np.savez_compressed(
            out / f"{name}.npz",
            x=mesh.x, y=mesh.y, n0=mesh.n0, n1=mesh.n1, n2=mesh.n2, bc_ux=mesh.ux, bc_uy=mesh.uy,
            bc_fx=mesh.fx, bc_fy=mesh.fy, known=mesh.known,
            material=np.array([META.youngs_modulus, META.poisson_ratio, META.part_thickness]),
            ke=ke, full_rowptr=full[0], full_col=full[1], full_val=full[2],
            kff_rowptr=rp, kff_col=col, kff_val=val, rhs=rhs, free_map=fmap,
            ux=res["ux"], uy=res["uy"], fx=res["fx"], fy=res["fy"], stress=res["stress"],
            iters=np.array([res["stats"]["iters"]]))
        print(f"{name}: {mesh.n_nodes} nodes, {mesh.n_elems} elems, nnz_ff {len(val)}, "
              f"{res['stats']['iters']} CG iterations")


if __name__ == "__main__":
    main()
