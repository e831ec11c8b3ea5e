"""`python -m magnetite_b200 input.json geometry...` — the reference's main.rs:21-76 with the numerical core on
the GPU:  mesher::run (outlines -> gmsh -> mesh -> boundary rules), solver::run, post_processor::csv_output.

    python -m magnetite_b200 input.json outline.svg                 SVG with OUTER / INNER* ids (mesher.rs:26-244)
    python -m magnetite_b200 input.json outer.csv [hole.csv ...]    CSV outlines (mesher.rs:253-299)
    python -m magnetite_b200 input.json geom.msh                    a mesh gmsh already wrote (kept on disk)

Options: --skip (accepted for compatibility: the plot step, post_processor.rs:90-123, is out of scope and never
runs), --cmap (ignored likewise), --reorder (renumber the nodes around the solve), --standin
(mesh the outlines with the built-in gmsh-free mesher instead of calling gmsh).  Writes nodes.csv and
elements.csv into the working directory; on error prints `Received error: <err>` and exits 1 (main.rs:43-51)."""
from __future__ import annotations

import argparse
import sys

from . import geometry, mesher, post_processor, solver
from .error import MagnetiteError


def entry(argv) -> None:
    ap = argparse.ArgumentParser(prog="magnetite_b200", description="2D linear-elastic FEA on a B200 (Magnetite drop-in)")
    ap.add_argument("input_file", help="Input Json with boundary conditions")
    ap.add_argument("geometry_files", nargs="*", help="Geometry SVG or CSVs (or one .msh)")
    ap.add_argument("-c", "--cmap", default="coolwarm", help="cmap for python plot (ignored: no plot step)")
    ap.add_argument("-s", "--skip", action="store_true", help="skip python plot (always skipped)")
    ap.add_argument("--reorder", action="store_true", help="renumber the nodes (reverse Cuthill-McKee) around the solve")
    ap.add_argument("--standin", action="store_true", help="mesh with the built-in stand-in mesher instead of gmsh")
    args = ap.parse_intermixed_args(argv)
    geoms = args.geometry_files
    if len(geoms) == 1 and geoms[0].endswith(".msh") or args.standin:
        input_json = mesher.load_input_file(args.input_file)
        meta = mesher.parse_input_metadata(input_json)
        if args.standin:
            containers = []
            for geom in geoms:                                       # mesher.rs:946-959
                if geom.endswith(".svg"):
                    containers = geometry.parse_svg(geom, meta.characteristic_length_min)
                    break
                elif geom.endswith(".csv"):
                    containers.append(geometry.parse_csv(geom))
                else:
                    raise MagnetiteError.Input(f"Unrecognized geometry filetype {geom}")
            xs, ys, conn = geometry.standin_mesh(containers, meta.characteristic_length_min, meta.characteristic_length_max)
            nodes = mesher.default_nodes(xs, ys)
            from .datatypes import Element
            elements = [Element([int(a), int(b), int(c)]) for a, b, c in conn]
        else:
            nodes, elements = geometry.parse_mesh(geoms[0])
        mesher.check_ccw(elements, nodes)
        print(f"info: loaded {len(nodes)} nodes and {len(elements)} elements")
        mesher.apply_boundary_conditions(input_json, nodes, quiet=False)
    else:
        nodes, elements, meta = mesher.run(geoms, args.input_file)                  # main.rs:58-61
    solver.run(nodes, elements, meta, reorder=args.reorder)                         # main.rs:64
    post_processor.csv_output(elements, nodes, "nodes.csv", "elements.csv")         # main.rs:67-69


def main(argv=None) -> int:
    try:
        entry(sys.argv[1:] if argv is None else argv)
    except MagnetiteError as err:
        print(f"Received error: {err}", file=sys.stderr)                             # main.rs:46
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
