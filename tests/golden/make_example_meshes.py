"""Builds the BASELINE config-1/2 fixtures from the reference's own example inputs
(/root/reference/examples — readable only in the build container, so the result is committed):
the outline goes through the gmsh-free stand-in mesher (magnetite_b200/geometry.py), then through
the reference's input semantics (node defaults, check_ccw with its `< 1.0` quirk, the box rules of
input.json), and the ORACLE in faithful-dense mode (the reference's data structures and CG) gives
the expected outputs.

    python tests/golden/make_example_meshes.py
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/examples")

from magnetite_b200 import geometry, mesher  # noqa: E402
from magnetite_b200.datatypes import Element, MeshSoA  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {
    "example_tensile": ("tensile-example", ["vertices.csv"]),
    "example_linkedin": ("linkedin-logo", ["linkedin.svg"]),
    "example_cover": ("cover-eample", ["geom.svg"]),
}


def build(name, folder, geoms):
    data = mesher.load_input_file(str(REF / folder / "input.json"))
    meta = mesher.parse_input_metadata(data)
    containers = []
    for g in geoms:                                       # mesher.rs:947-959
        if g.endswith(".svg"):
            containers = geometry.parse_svg(str(REF / folder / g), meta.characteristic_length_min)
            break
        containers.append(geometry.parse_csv(str(REF / folder / g)))
    xs, ys, conn = geometry.standin_mesh(containers, meta.characteristic_length_min, meta.characteristic_length_max)
    nodes = mesher.default_nodes(xs, ys)
    elements = [Element([int(a), int(b), int(c)]) for a, b, c in conn]
    # check_ccw (mesher.rs:522-526) with the oracle's area (no GPU in the build container)
    soa = MeshSoA.from_aos(nodes, elements)
    area = O.element_area(O.Mesh(soa))
    for el, a in zip(elements, area):
        if a < 1.0:
            el.nodes = list(reversed(el.nodes))
    mesher.apply_boundary_conditions(data, nodes)
    mesh = MeshSoA.from_aos(nodes, elements)
    t0 = time.perf_counter()
    res = O.run(O.Mesh(mesh), meta, O.cg_options(), dense=True)       # the reference's algorithm
    dt = time.perf_counter() - t0
    st = res["stats"]
    np.savez_compressed(
        Path(__file__).resolve().parent / f"{name}.npz",
        x=mesh.x, y=mesh.y, n0=mesh.n0, n1=mesh.n1, n2=mesh.n2, bc_ux=mesh.ux, bc_uy=mesh.uy, bc_fx=mesh.fx,
        bc_fy=mesh.fy, known=mesh.known,
        material=np.array([meta.youngs_modulus, meta.poisson_ratio, meta.part_thickness]),
        ux=res["ux"], uy=res["uy"], fx=res["fx"], fy=res["fy"], stress=res["stress"],
        iters=np.array([st["iters"]]), nnz_ff=np.array([st["nnz_ff"]]),
        flipped=np.array([int((area < 1.0).sum())]),
        oracle_seconds=np.array([st["t_elem"], st["t_asm"], st["t_part"], st["t_solve"], st["t_react"], st["t_stress"]]))
    print(f"{name}: {mesh.n_nodes} nodes, {mesh.n_elems} elements, {int((area < 1.0).sum())} flipped by check_ccw, "
          f"{st['n_free']} free DOF, nnz {st['nnz_ff']}, {st['iters']} CG iterations, oracle dense {dt:.2f} s "
          f"(partition {st['t_part']:.2f} s, CG {st['t_solve']:.2f} s)")


if __name__ == "__main__":
    for name, (folder, geoms) in CASES.items():
        build(name, folder, geoms)
