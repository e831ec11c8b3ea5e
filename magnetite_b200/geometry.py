"""Geometry input of the reference's mesher, without gmsh.

    parse_csv        src/mesher.rs:253-299   outline vertices from a CSV with x,y columns
    parse_svg        src/mesher.rs:26-244    polygon / polyline / rect elements, ids OUTER / INNER*
    parse_mesh       src/mesher.rs:536-704   MSH 4.x ASCII ($Nodes / $Elements) -> nodes, elements
    write_msh        (inverse of parse_mesh, for round trips and for handing meshes to other tools)
    standin_mesh     a gmsh-free triangulation of the same outlines: boundary resampled at the
                     characteristic length, hexagonal interior lattice, scipy Delaunay, triangles
                     whose centroid falls outside OUTER or inside an INNER removed.

    build_geo        src/mesher.rs:305-472   outlines -> the .geo script gmsh meshes (same text as the reference writes)
    compute_mesh     src/mesher.rs:481-519   build_geo + `gmsh geom.geo -2 -o <msh>` — used when a gmsh binary exists

The stand-in does not reproduce gmsh's node placement (element counts differ from the author's
runs); it reproduces what the solver sees: 0-based node ids in file order, node defaults
ux=uy=None / fx=fy=Some(0.0), `check_ccw`, then the boundary rules.  A real `geom.msh` written by
gmsh is consumed as-is by `parse_mesh`.
"""
from __future__ import annotations

import math
import os
import subprocess
import xml.etree.ElementTree as ET
from typing import List, Sequence, Tuple

import numpy as np

from .datatypes import Element, Node, Vertex
from .error import MagnetiteError


# ---------------------------------------------------------------------------
# outlines -> gmsh (only when a gmsh binary is installed; the stand-in mesher below needs none)
# ---------------------------------------------------------------------------
def _rust_display(v, dtype) -> str:
    """Rust's `{}` for f64 / f32: shortest digits that round-trip IN THAT TYPE, positional, no ".0"."""
    v = dtype(v)
    if np.isnan(v):
        return "NaN"
    if np.isinf(v):
        return "inf" if v > 0 else "-inf"
    return np.format_float_positional(v, unique=True, trim="-")


def geo_text(vertices_containers: Sequence[Sequence[Vertex]], characteristic_length_min: float,
             characteristic_length_max: float) -> str:
    """The .geo script of build_geo (mesher.rs:305-472), character for character: container 0 is the outer
    loop, the others are holes; points, lines and loops are numbered per container with running offsets;
    with exactly two containers the surface lists the loops in reverse (mesher.rs:424-430).  Coordinates
    print as f64, the characteristic lengths as f32 (ModelMetadata, datatypes.rs:27-28)."""
    if not vertices_containers:
        raise MagnetiteError.Input("no geometry was given")          # the reference indexes [0] and panics
    f = lambda v: _rust_display(v, np.float64)                       # noqa: E731
    out = ["// Define outer points\n"]
    for i, v in enumerate(vertices_containers[0]):
        out.append(f"Point({i}) = {{ {f(v.x)}, {f(v.y)}, 0, 1.0 }};\n")
    out.append("\n// Define inner points\n")
    offset = len(vertices_containers[0])
    inner_offsets = [0]
    for verts in vertices_containers[1:]:
        inner_offsets.append(offset)
        for i, v in enumerate(verts):
            out.append(f"Point({i + offset}) = {{ {f(v.x)}, {f(v.y)}, 0, 1.0 }};\n")
        offset += len(verts)
    out.append("\n// Connect points\n")
    for i, verts in enumerate(vertices_containers):
        out.append(f"\n// Point connections for surface {i}\n")
        off = inner_offsets[i]
        for k in range(1, len(verts)):
            out.append(f"Line({k + off - 1}) = {{ {k + off - 1}, {k + off} }};\n")
        out.append(f"Line({len(verts) + off - 1}) = {{ {len(verts) + off - 1}, {off} }};\n")
    out.append("\n//Register loops\n")
    for i, verts in enumerate(vertices_containers):
        off = inner_offsets[i]
        out.append(f"Line Loop({i + 1}) = {{")
        out.extend(f"{',' if k else ''} {k + off}" for k in range(len(verts)))
        out.append(" };\n")
    out.append("\n//Define surface\n")
    out.append("Plane Surface(1) = {")
    n = len(vertices_containers)
    order = list(range(n)) if n > 2 else list(reversed(range(n)))
    out.extend(f"{',' if k else ''} {loop + 1}" for k, loop in enumerate(order))
    out.append(" };\n")
    g = lambda v: _rust_display(v, np.float32)                       # noqa: E731
    out.append("\n// Define Mesh Settings\nMesh.ElementOrder = 1;\nMesh.Algorithm  = 1;\n"
               f"Mesh.CharacteristicLengthMin = {g(characteristic_length_min)};\n"
               f"Mesh.CharacteristicLengthMax = {g(characteristic_length_max)};\nMesh 2;\n")
    return "".join(out)


def build_geo(vertices_containers: Sequence[Sequence[Vertex]], output_file: str, characteristic_length_min: float,
              characteristic_length_max: float) -> None:
    """mesher.rs:305-472."""
    text = geo_text(vertices_containers, characteristic_length_min, characteristic_length_max)
    with open(output_file, "w", newline="") as fh:
        fh.write(text)


def compute_mesh(vertices: Sequence[Sequence[Vertex]], output: str, characteristic_length_min: float,
                 characteristic_length_max: float, quiet: bool = False) -> None:
    """mesher.rs:481-519: writes geom.geo, runs `gmsh geom.geo -2 -o <output>`, deletes geom.geo.  Like the
    reference it only fails when gmsh cannot be started ("Gmsh failed: ..."), not on gmsh's exit status — a
    failed meshing run surfaces in parse_mesh as "Unable to open auto-generated mesh file"."""
    geo_filepath = "geom.geo"
    if not quiet:
        print("info: building .geo for Gmsh with {:.3f}< CL < {:.3f}".format(
            float(np.float32(characteristic_length_min)), float(np.float32(characteristic_length_max))))
    build_geo(vertices, geo_filepath, characteristic_length_min, characteristic_length_max)
    if not quiet:
        print("info: running gmsh...")
    try:
        subprocess.run(["gmsh", geo_filepath, "-2", "-o", output], capture_output=True)
    except OSError as err:
        raise MagnetiteError.Mesher(f"Gmsh failed: {err.strerror} (os error {err.errno})")
    os.remove(geo_filepath)


# ---------------------------------------------------------------------------
# outlines
# ---------------------------------------------------------------------------
def parse_csv(csv_file: str) -> List[Vertex]:
    try:
        with open(csv_file, "r") as fh:
            contents = fh.read()
    except OSError:
        raise MagnetiteError.Input(f"Unable to open csv file {csv_file}")
    headers = None
    out: List[Vertex] = []
    for line in contents.split("\n"):
        if not line:
            continue
        if headers is None:
            headers = [h.strip() for h in line.split(",")]
            if "x" not in headers or "y" not in headers:
                raise MagnetiteError.Input("Error in csv file: Missing x and/or y field")
            xi, yi = headers.index("x"), headers.index("y")
        else:
            try:
                vals = [float(v.strip()) for v in line.split(",")]
            except ValueError:
                raise MagnetiteError.Input("Non-float value in csv points")
            out.append(Vertex(vals[xi], vals[yi]))
    return out


def _local(tag: str) -> str:
    return tag.rsplit("}", 1)[-1]


def parse_svg(svg_file: str, min_element_length: float) -> List[List[Vertex]]:
    """containers[0] = OUTER outline, containers[1:] = INNER outlines (holes).  y is inverted
    (mesher.rs:73), repeated vertices and vertices closer than `min_element_length` to the previous
    one are skipped (mesher.rs:78-91); an element's id, or its parent's, must start with OUTER or
    INNER (mesher.rs:97-128)."""
    try:
        tree = ET.parse(svg_file)
    except OSError:
        raise MagnetiteError.Input(f"Unable to open svg file {svg_file}")
    root = tree.getroot()
    parent = {child: par for par in root.iter() for child in par}
    containers: List[List[Vertex]] = [[]]

    def file_under(id_: str, pts: List[Vertex]):
        if id_.strip().startswith("INNER"):
            containers.append(pts)
        elif id_.strip().startswith("OUTER"):
            if containers[0]:
                raise MagnetiteError.Input("Multiple OUTER geometries in SVG")
            containers[0] = pts
        # other ids: the reference prints a warning and skips the shape

    def resolve_id(el):
        if el.get("id") is not None:
            return el.get("id")
        par = parent.get(el)
        if par is not None and par.get("id") is not None:
            return par.get("id")
        raise MagnetiteError.Input("Error in svg file. Missing id field on polyline")

    for el in root.iter():
        if _local(el.tag) not in ("polyline", "polygon"):
            continue
        raw = el.get("points")
        if raw is None:
            raise MagnetiteError.Input(f"Error in svg file. No points in polyline element {el.get('id')!r}")
        try:
            flat = [float(t) for t in raw.split(" ") if t != ""]
        except ValueError:
            raise MagnetiteError.Input("Non-float value in svg points")
        pts: List[Vertex] = []
        for i in range(0, len(flat) - 1, 2):
            v = Vertex(flat[i], -flat[i + 1])
            if any(p.x == v.x and p.y == v.y for p in pts):
                continue
            if pts and math.hypot(pts[-1].x - v.x, pts[-1].y - v.y) < float(min_element_length):
                continue
            pts.append(v)
        file_under(resolve_id(el), pts)

    for el in root.iter():
        if _local(el.tag) != "rect":
            continue
        x, y = float(el.get("x", 0.0)), float(el.get("y", 0.0))
        if el.get("width") is None or el.get("height") is None:
            raise MagnetiteError.Input("Error in svg file. No width/height definition in rectangle.")
        w, h = float(el.get("width")), float(el.get("height"))
        file_under(resolve_id(el), [Vertex(x, -y), Vertex(x + w, -y), Vertex(x + w, -y - h), Vertex(x, -y - h)])

    if not containers[0]:
        raise MagnetiteError.Input("No OUTER geometry")
    return containers


# ---------------------------------------------------------------------------
# MSH 4.x ASCII
# ---------------------------------------------------------------------------
def parse_mesh(mesh_file: str) -> Tuple[List[Node], List[Element]]:
    """mesher.rs:536-704 without the check_ccw pass and without deleting the file: nodes are placed
    at index tag-1 with ux=uy=None, fx=fy=Some(0.0); only elements of 2-D entities are kept, with
    0-based node ids."""
    try:
        with open(mesh_file, "r") as fh:
            lines = iter(fh.read().split("\n"))
    except OSError as err:
        raise MagnetiteError.Mesher(f"Unable to open auto-generated mesh file: {err}")
    state, seen_meta = "limbo", False
    unordered: List[Node] = []
    indexes: List[int] = []
    elements: List[Element] = []
    for line in lines:
        if not line:
            continue
        if line.startswith("$End"):
            state = "limbo"
        if state == "limbo":
            seen_meta = False
            if line.startswith("$Entities"):
                state = "entities"
            elif line.startswith("$Node"):
                state = "nodes"
            elif line.startswith("$Elements"):
                state = "elements"
            continue
        if state == "nodes":
            if not seen_meta:
                seen_meta = True
                continue
            n_local = int(line.split(" ")[3])
            tags = [int(next(lines)) for _ in range(n_local)]
            for i in range(n_local):
                c = [float(v) for v in next(lines).split(" ")]
                unordered.append(Node(Vertex(c[0], c[1]), None, None, 0.0, 0.0))
                indexes.append(tags[i] - 1)
        elif state == "elements":
            if not seen_meta:
                seen_meta = True
                continue
            head = [int(v) for v in line.split(" ")]
            entity_dim, n_el = head[0], head[3]
            for _ in range(n_el):
                meta = [int(v) for v in next(lines).strip().split(" ")]
                if entity_dim == 2:
                    elements.append(Element([meta[1] - 1, meta[2] - 1, meta[3] - 1]))
    nodes: List[Node] = [None] * len(unordered)   # type: ignore[list-item]
    for idx, nd in zip(indexes, unordered):
        nodes[idx] = nd
    if any(n is None for n in nodes):
        raise MagnetiteError.Mesher("node tags are not dense 1..N")
    return nodes, elements


def write_msh(path: str, xs: Sequence[float], ys: Sequence[float], conn: Sequence[Sequence[int]]) -> None:
    """Minimal MSH 4.1 ASCII file with one 2-D entity."""
    n, e = len(xs), len(conn)
    with open(path, "w") as fh:
        fh.write("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
        fh.write(f"$Nodes\n1 {n} 1 {n}\n2 1 0 {n}\n")
        fh.write("".join(f"{i + 1}\n" for i in range(n)))
        fh.write("".join(f"{float(x)!r} {float(y)!r} 0\n" for x, y in zip(xs, ys)))
        fh.write("$EndNodes\n")
        fh.write(f"$Elements\n1 {e} 1 {e}\n2 1 2 {e}\n")
        fh.write("".join(f"{i + 1} {c[0] + 1} {c[1] + 1} {c[2] + 1}\n" for i, c in enumerate(conn)))
        fh.write("$EndElements\n")


# ---------------------------------------------------------------------------
# gmsh-free stand-in mesher
# ---------------------------------------------------------------------------
def _poly(container: Sequence[Vertex]) -> np.ndarray:
    return np.array([[v.x, v.y] for v in container], float)


def _inside(poly: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """Even-odd point-in-polygon for an (m,2) array of points."""
    x, y = pts[:, 0], pts[:, 1]
    inside = np.zeros(len(pts), bool)
    xj, yj = poly[-1]
    for xi, yi in poly:
        cond = (yi > y) != (yj > y)
        with np.errstate(divide="ignore", invalid="ignore"):
            xint = (xj - xi) * (y - yi) / (yj - yi) + xi
        inside ^= cond & (x < xint)
        xj, yj = xi, yi
    return inside


def _resample(poly: np.ndarray, h: float) -> np.ndarray:
    out = []
    for a, b in zip(poly, np.roll(poly, -1, axis=0)):
        k = max(1, int(math.ceil(np.linalg.norm(b - a) / h)))
        out.extend(a + (b - a) * (t / k) for t in range(k))
    return np.array(out)


def _dist_to_segments(pts: np.ndarray, poly: np.ndarray) -> np.ndarray:
    d = np.full(len(pts), np.inf)
    for a, b in zip(poly, np.roll(poly, -1, axis=0)):
        ab = b - a
        t = np.clip(((pts - a) @ ab) / max(ab @ ab, 1e-300), 0.0, 1.0)
        d = np.minimum(d, np.linalg.norm(pts - (a + t[:, None] * ab), axis=1))
    return d


def standin_mesh(containers: Sequence[Sequence[Vertex]], cl_min: float, cl_max: float):
    """(xs, ys, conn) for the region inside containers[0] and outside containers[1:]."""
    from scipy.spatial import Delaunay
    h = float(cl_max) if cl_min <= 0 else float(cl_min)      # the fine end of the gmsh size band
    polys = [_poly(c) for c in containers]
    boundary = np.vstack([_resample(p, h) for p in polys])
    lo, hi = polys[0].min(0), polys[0].max(0)
    ys = np.arange(lo[1] + 0.5 * h, hi[1], h * math.sqrt(3) / 2)
    rows = []
    for k, yy in enumerate(ys):
        xs = np.arange(lo[0] + (0.5 if k % 2 else 0.0) * h + 0.25 * h, hi[0], h)
        rows.append(np.stack([xs, np.full_like(xs, yy)], 1))
    lattice = np.vstack(rows)
    keep = _inside(polys[0], lattice)
    for p in polys[1:]:
        keep &= ~_inside(p, lattice)
    for p in polys:
        keep &= _dist_to_segments(lattice, p) > 0.55 * h
    pts = np.vstack([boundary, lattice[keep]])
    tri = Delaunay(pts).simplices
    cen = pts[tri].mean(1)
    ok = _inside(polys[0], cen)
    for p in polys[1:]:
        ok &= ~_inside(p, cen)
    a, b, c = pts[tri[:, 0]], pts[tri[:, 1]], pts[tri[:, 2]]
    area2 = np.abs((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1]))
    ok &= area2 > 1e-6 * h * h
    tri = tri[ok]
    used = np.zeros(len(pts), bool)
    used[tri.ravel()] = True
    newid = np.cumsum(used) - 1
    return pts[used, 0].copy(), pts[used, 1].copy(), newid[tri].astype(np.int64)
