"""Input semantics of the reference's mesher that the solver path depends on
(SURVEY §8(f) rank 1): the pieces that turn a mesh + input.json into the solver's inputs, and the
entry point that strings the whole input side together.

    check_ccw                   src/mesher.rs:522-526  (flips when signed area < 1.0 — sic)
    load_input_file             src/mesher.rs:713-760
    parse_input_metadata        src/mesher.rs:769-808
    apply_boundary_conditions   src/mesher.rs:815-930  (strict > / <, later rules overwrite)
    run                         src/mesher.rs:939-974  geometry files + input.json -> nodes, elements, metadata
                                (outline parsing, .geo and the gmsh call live in geometry.py)
"""
from __future__ import annotations

import json
import sys
from typing import List, Sequence

import numpy as np

from .datatypes import (BoundaryRegion, BoundaryRule, BoundaryTarget, Element, MeshSoA,
                        ModelMetadata, Node)
from .error import MagnetiteError

F64_MIN, F64_MAX = -sys.float_info.max, sys.float_info.max     # f64::MIN / f64::MAX


def load_input_file(input_file: str) -> dict:
    try:
        with open(input_file, "r") as fh:
            text = fh.read()
    except OSError:
        raise MagnetiteError.Input(f"Unable to open input file {input_file}")
    try:
        data = json.loads(text)
    except json.JSONDecodeError as err:
        raise MagnetiteError.Input(f"Error in input file json: {err}")
    if "metadata" not in data:
        raise MagnetiteError.Input("Input json missing metadata field")
    if "boundary_conditions" not in data:
        raise MagnetiteError.Input("Input json missing boundary_conditions field in metadata section")
    for key in ("part_thickness", "material_elasticity", "poisson_ratio"):
        if key not in data["metadata"]:
            raise MagnetiteError.Input(f"Input json missing {key} field in metadata section")
    return data


def _num(v):
    return float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else None


def parse_input_metadata(input_json: dict) -> ModelMetadata:
    md = input_json["metadata"]
    e, t, nu = _num(md.get("material_elasticity")), _num(md.get("part_thickness")), _num(md.get("poisson_ratio"))
    cmin, cmax = _num(md.get("characteristic_length_min")), _num(md.get("characteristic_length_max"))
    if e is None:
        raise MagnetiteError.Input("Input json missing material elasticity")
    if nu is None:
        raise MagnetiteError.Input("Input json missing poisson ratio")
    if cmin is None:
        raise MagnetiteError.Input("Input json missing minimum characteristic length")
    if cmax is None:
        raise MagnetiteError.Input("Input json missing maximum characteristic length")
    if t is None:
        raise MagnetiteError.Input("Input json missing part thickness")   # the reference unwrap()s
    return ModelMetadata(e, nu, t, float(np.float32(cmin)), float(np.float32(cmax)))


def parse_boundary_rules(input_json: dict) -> List[BoundaryRule]:
    rules: List[BoundaryRule] = []
    for name, rule in input_json["boundary_conditions"].items():
        if "region" not in rule:
            raise MagnetiteError.Input(f"Boundary rule {name} is missing region field")
        if "targets" not in rule:
            raise MagnetiteError.Input(f"Boundary rule {name} is missing target field")
        reg = BoundaryRegion(F64_MIN, F64_MAX, F64_MIN, F64_MAX)      # mesher.rs:835-840
        for key, attr in (("x_target_min", "x_min"), ("x_target_max", "x_max"),
                          ("y_target_min", "y_min"), ("y_target_max", "y_max")):
            if key in rule["region"]:
                v = _num(rule["region"][key])
                if v is None:
                    raise MagnetiteError.Input(f"Bad value for {key} in {name}")
                setattr(reg, attr, v)
        tg = rule["targets"]
        tgt = BoundaryTarget(_num(tg.get("ux")), _num(tg.get("uy")), _num(tg.get("fx")), _num(tg.get("fy")))
        if reg.x_min > reg.x_max:
            raise MagnetiteError.Input(f"Boundary '{name}' has x_target_min greater than x_target_max")
        if reg.y_min > reg.y_max:
            raise MagnetiteError.Input(f"Boundary '{name}' has y_target_min greater than y_target_max")
        if tgt.fx is None and tgt.ux is None:
            raise MagnetiteError.Input(f"Boundary '{name}' is under-constrained in x-axis")
        if tgt.fy is None and tgt.uy is None:
            raise MagnetiteError.Input(f"Boundary '{name}' is under-constrained in y-axis")
        if tgt.fx is not None and tgt.ux is not None:
            raise MagnetiteError.Input(f"Boundary '{name}' is over-constrained in x-axis")
        if tgt.fy is not None and tgt.uy is not None:
            raise MagnetiteError.Input(f"Boundary '{name}' is over-constrained in y-axis")
        rules.append(BoundaryRule(name, reg, tgt))
    return rules


def apply_boundary_conditions(input_json: dict, nodes: Sequence[Node], quiet: bool = True) -> None:
    """mesher.rs:815-930: a node strictly inside a rule's box takes all four targets of the rule;
    rules are applied in file order, so later rules win."""
    rules = parse_boundary_rules(input_json)
    if not quiet:
        print(f"info: loaded {len(rules)} boundary rules from input file")
    for node in nodes:
        for rule in rules:
            r = rule.region
            if r.x_min < node.vertex.x < r.x_max and r.y_min < node.vertex.y < r.y_max:
                node.ux, node.uy = rule.target.ux, rule.target.uy
                node.fx, node.fy = rule.target.fx, rule.target.fy


def default_nodes(xs, ys) -> List[Node]:
    """Node defaults of parse_mesh (mesher.rs:615-624): ux=uy=None, fx=fy=Some(0.0)."""
    from .datatypes import Vertex
    return [Node(Vertex(float(a), float(b)), None, None, 0.0, 0.0) for a, b in zip(xs, ys)]


def check_ccw(elements: Sequence[Element], nodes: Sequence[Node]) -> None:
    """mesher.rs:522-526 applied to every element (mesher.rs:691-693): reverse the node order of
    any element whose signed area is < 1.0.  Areas come from the GPU in one launch."""
    from . import solver
    if not elements:
        return
    areas = solver.element_areas(MeshSoA.from_aos(nodes, elements))
    for el, a in zip(elements, areas):
        if a < 1.0:
            el.nodes = list(reversed(el.nodes))


def run(geometry_files: Sequence[str], input_file: str, quiet: bool = False):
    """mesher::run (mesher.rs:939-974): (nodes, elements, model_metadata) ready for solver.run.

    An .svg ends the list (its containers replace whatever was read before, mesher.rs:948-950), every .csv
    adds one container, anything else is an Input error.  The mesh comes from gmsh through geom.geo /
    geom.msh in the working directory, both removed afterwards like the reference does; check_ccw runs on
    the GPU (one launch for all elements) before the boundary rules are applied."""
    import os
    from . import geometry
    input_json = load_input_file(input_file)
    meta = parse_input_metadata(input_json)
    vertices: list = []
    for geom in geometry_files:
        if geom.endswith(".svg"):
            vertices = geometry.parse_svg(geom, meta.characteristic_length_min)
            break
        elif geom.endswith(".csv"):
            vertices.append(geometry.parse_csv(geom))
        else:
            raise MagnetiteError.Input(f"Unrecognized geometry filetype {geom}")
    mesh_filepath = "geom.msh"
    geometry.compute_mesh(vertices, mesh_filepath, meta.characteristic_length_min, meta.characteristic_length_max, quiet)
    nodes, elements = geometry.parse_mesh(mesh_filepath)
    check_ccw(elements, nodes)                                       # mesher.rs:691-693
    if not quiet:
        print(f"info: loaded {len(nodes)} nodes and {len(elements)} elements")   # mesher.rs:695-699
    os.remove(mesh_filepath)                                         # mesher.rs:701
    apply_boundary_conditions(input_json, nodes, quiet)
    return nodes, elements, meta
