"""CPU: geometry input semantics of the reference's mesher (parse_csv, parse_svg, MSH-4 reader) and the
gmsh-free stand-in mesher; the BASELINE config-1/2 fixtures stay reproducible by the oracle."""
from pathlib import Path

import numpy as np
import pytest

from magnetite_b200 import geometry, mesher, meshgen
from magnetite_b200.datatypes import MeshSoA
from magnetite_b200.error import MagnetiteError
from oracle import oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden"

SVG = """<?xml version="1.0"?>
<svg xmlns="http://www.w3.org/2000/svg" viewBox="0 0 100 60">
  <g id="OUTER"><rect width="100" height="60"/></g>
  <polygon id="INNER-1" points="20 20 20.01 20 40 20 40 40 20 40 20 20"/>
  <polyline id="decoration" points="1 1 2 2"/>
  <g id="INNER-2"><rect x="60" y="10" width="10" height="10"/></g>
</svg>"""


def test_parse_csv(tmp_path):
    p = tmp_path / "v.csv"
    p.write_text("y, x\n4.5,-11\n4.5,-10\n\n-4.5, -10\n")
    v = geometry.parse_csv(str(p))
    assert [(a.x, a.y) for a in v] == [(-11.0, 4.5), (-10.0, 4.5), (-10.0, -4.5)]
    p.write_text("a,b\n1,2\n")
    with pytest.raises(MagnetiteError, match="Missing x and/or y"):
        geometry.parse_csv(str(p))
    with pytest.raises(MagnetiteError, match="Unable to open"):
        geometry.parse_csv(str(tmp_path / "nope.csv"))


def test_parse_svg_rules(tmp_path):
    p = tmp_path / "g.svg"
    p.write_text(SVG)
    c = geometry.parse_svg(str(p), 0.5)
    assert len(c) == 3                                            # OUTER + two INNER, "decoration" skipped
    assert [(v.x, v.y) for v in c[0]] == [(0, -0.0), (100, -0.0), (100, -60), (0, -60)]   # rect via parent id, y inverted
    inner = [(v.x, v.y) for v in c[1]]
    assert inner == [(20, -20), (40, -20), (40, -40), (20, -40)]  # near-duplicate (< min length) and repeat dropped
    assert [(v.x, v.y) for v in c[2]] == [(60, -10), (70, -10), (70, -20), (60, -20)]
    p.write_text(SVG.replace('id="OUTER"', 'id="SHELL"'))
    with pytest.raises(MagnetiteError, match="No OUTER geometry"):
        geometry.parse_svg(str(p), 0.5)
    p.write_text(SVG.replace('id="INNER-1"', 'id="OUTER-2"'))
    with pytest.raises(MagnetiteError, match="Multiple OUTER"):
        geometry.parse_svg(str(p), 0.5)


def test_msh_round_trip(tmp_path):
    m = meshgen.jitter(meshgen.plate(5, 3))
    conn = np.stack([m.n0, m.n1, m.n2], 1).astype(int)
    path = tmp_path / "geom.msh"
    geometry.write_msh(str(path), m.x, m.y, conn)
    nodes, elements = geometry.parse_mesh(str(path))
    assert len(nodes) == m.n_nodes and len(elements) == m.n_elems
    assert all(n.vertex.x == x and n.vertex.y == y for n, x, y in zip(nodes, m.x, m.y))
    assert all(n.ux is None and n.uy is None and n.fx == 0.0 and n.fy == 0.0 for n in nodes)   # mesher.rs:615-624
    assert [e.nodes for e in elements] == conn.tolist()
    with pytest.raises(MagnetiteError, match="Unable to open"):
        geometry.parse_mesh(str(tmp_path / "missing.msh"))


def test_standin_mesher_respects_holes(tmp_path):
    p = tmp_path / "g.svg"
    p.write_text(SVG)
    c = geometry.parse_svg(str(p), 0.5)
    xs, ys, conn = geometry.standin_mesh(c, 0.0, 4.0)
    a = 0.5 * np.abs((xs[conn[:, 1]] - xs[conn[:, 0]]) * (ys[conn[:, 2]] - ys[conn[:, 0]])
                     - (xs[conn[:, 2]] - xs[conn[:, 0]]) * (ys[conn[:, 1]] - ys[conn[:, 0]]))
    assert abs(a.sum() - (100 * 60 - 20 * 20 - 10 * 10)) < 1e-6 * 6000      # area of the region, holes excluded
    assert conn.min() == 0 and conn.max() == len(xs) - 1 and len(np.unique(conn)) == len(xs)
    cx, cy = xs[conn].mean(1), ys[conn].mean(1)
    assert not ((cx > 20) & (cx < 40) & (cy < -20) & (cy > -40)).any()


@pytest.mark.parametrize("name,flipped", [("example_tensile", True), ("example_linkedin", False), ("example_cover", False)])
def test_example_fixtures_reproducible(name, flipped):
    g = np.load(GOLDEN / f"{name}.npz")
    mesh = MeshSoA(g["x"], g["y"], g["n0"], g["n1"], g["n2"], g["bc_ux"], g["bc_uy"], g["bc_fx"], g["bc_fy"], g["known"])
    meta = meshgen.EXAMPLE_MATERIAL.__class__(*g["material"])
    om = O.Mesh(mesh)
    area = O.element_area(om)
    assert (area < 0).all() if flipped else (area >= 1.0).all()          # check_ccw's `< 1.0`
    assert int(g["flipped"][0]) == (mesh.n_elems if flipped else 0)
    res = O.run(om, meta, O.cg_options(), dense=False)                   # sparse mode == dense mode bit for bit
    for k in ("ux", "uy", "fx", "fy", "stress"):
        assert np.array_equal(res[k], g[k]), k
    assert res["stats"]["iters"] == int(g["iters"][0]) and res["stats"]["nnz_ff"] == int(g["nnz_ff"][0])
    # the boundary rules selected something on both ends
    assert ((mesh.known & 3) == 3).sum() > 3 and ((mesh.known & 1) == 1).sum() > ((mesh.known & 3) == 3).sum() - 1
