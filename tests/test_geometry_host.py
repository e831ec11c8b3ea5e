"""CPU: geometry input semantics of the reference's mesher (parse_csv, parse_svg, MSH-4 reader) and the
gmsh-free stand-in mesher; the BASELINE config-1/2 fixtures stay reproducible by the oracle."""
import re
from pathlib import Path

import numpy as np
import pytest

from magnetite_b200 import geometry, mesher, meshgen
from magnetite_b200.datatypes import MeshSoA, Vertex
from magnetite_b200.error import MagnetiteError
from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent
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


FAKE_GMSH = r'''#!{python}
"""Stand-in for the gmsh binary in tests: `gmsh geom.geo -2 -o out.msh`.  Parses the .geo script the way gmsh
reads it (points, lines, line loops, the plane surface, the characteristic lengths), checks that every loop
is closed, and meshes the outlines with the repo's gmsh-free stand-in mesher."""
import re, sys
sys.path.insert(0, {root!r})
from magnetite_b200 import geometry
from magnetite_b200.datatypes import Vertex
geo, out = sys.argv[1], sys.argv[sys.argv.index("-o") + 1]
assert sys.argv[2] == "-2"
text = open(geo).read()
pts = {{int(m[1]): (float(m[2]), float(m[3])) for m in re.finditer(r"Point\((\d+)\) = \{{ (\S+), (\S+), 0, 1\.0 \}};", text)}}
lines = {{int(m[1]): (int(m[2]), int(m[3])) for m in re.finditer(r"Line\((\d+)\) = \{{ (\d+), (\d+) \}};", text)}}
loops = {{int(m[1]): [int(v) for v in m[2].split(",")] for m in re.finditer(r"Line Loop\((\d+)\) = \{{([^}}]*)\}};", text)}}
surface = [int(v) for v in re.search(r"Plane Surface\(1\) = \{{([^}}]*)\}};", text)[1].split(",")]
cl_min = float(re.search(r"Mesh\.CharacteristicLengthMin = (\S+);", text)[1])
cl_max = float(re.search(r"Mesh\.CharacteristicLengthMax = (\S+);", text)[1])
assert "Mesh.ElementOrder = 1;" in text and "Mesh 2;" in text and sorted(surface) == sorted(loops)
containers = {{}}
for lid, ids in loops.items():
    for a, b in zip(ids, ids[1:] + ids[:1]):            # consecutive lines share a point and the loop closes
        assert lines[a][1] == lines[b][0], (lid, a, b)
    containers[lid] = [Vertex(*pts[lines[i][0]]) for i in ids]
xs, ys, conn = geometry.standin_mesh([containers[k] for k in sorted(containers)], cl_min, cl_max)
geometry.write_msh(out, xs, ys, conn)
'''


def _install_fake_gmsh(tmp_path, monkeypatch):
    import os
    import stat
    import sys
    bindir = tmp_path / "bin"
    bindir.mkdir()
    exe = bindir / "gmsh"
    exe.write_text(FAKE_GMSH.format(python=sys.executable, root=str(ROOT)))
    exe.chmod(exe.stat().st_mode | stat.S_IXUSR)
    monkeypatch.setenv("PATH", str(bindir) + os.pathsep + os.environ.get("PATH", ""))


def test_geo_script_matches_the_reference_format():
    """build_geo (mesher.rs:305-472) character for character on one outer loop with one hole (loops listed in
    reverse, mesher.rs:424-430) and on three containers (ascending); f64 coordinates, f32 lengths."""
    sq = [Vertex(0, 0), Vertex(10, 0), Vertex(10, 5.5), Vertex(0.1 + 0.2, 5.5)]
    hole = [Vertex(2, 2), Vertex(3, 2), Vertex(3, 3)]
    text = geometry.geo_text([sq, hole], 0.0, float(np.float32(0.3)))
    assert text == (
        "// Define outer points\n"
        "Point(0) = { 0, 0, 0, 1.0 };\nPoint(1) = { 10, 0, 0, 1.0 };\nPoint(2) = { 10, 5.5, 0, 1.0 };\n"
        "Point(3) = { 0.30000000000000004, 5.5, 0, 1.0 };\n"
        "\n// Define inner points\n"
        "Point(4) = { 2, 2, 0, 1.0 };\nPoint(5) = { 3, 2, 0, 1.0 };\nPoint(6) = { 3, 3, 0, 1.0 };\n"
        "\n// Connect points\n"
        "\n// Point connections for surface 0\n"
        "Line(0) = { 0, 1 };\nLine(1) = { 1, 2 };\nLine(2) = { 2, 3 };\nLine(3) = { 3, 0 };\n"
        "\n// Point connections for surface 1\n"
        "Line(4) = { 4, 5 };\nLine(5) = { 5, 6 };\nLine(6) = { 6, 4 };\n"
        "\n//Register loops\n"
        "Line Loop(1) = { 0, 1, 2, 3 };\nLine Loop(2) = { 4, 5, 6 };\n"
        "\n//Define surface\n"
        "Plane Surface(1) = { 2, 1 };\n"
        "\n// Define Mesh Settings\nMesh.ElementOrder = 1;\nMesh.Algorithm  = 1;\n"
        "Mesh.CharacteristicLengthMin = 0;\nMesh.CharacteristicLengthMax = 0.3;\nMesh 2;\n")
    three = geometry.geo_text([sq, hole, [Vertex(5, 1), Vertex(6, 1), Vertex(6, 2)]], 10.0, 30.0)
    assert "Plane Surface(1) = { 1, 2, 3 };\n" in three and "Line(9) = { 9, 7 };\n" in three
    assert "Mesh.CharacteristicLengthMin = 10;\nMesh.CharacteristicLengthMax = 30;\n" in three
    assert geometry.geo_text([sq], 0.0, 1.5).count("Line Loop") == 1 and "Plane Surface(1) = { 1 };\n" in geometry.geo_text([sq], 0.0, 1.5)
    with pytest.raises(MagnetiteError, match="no geometry"):
        geometry.geo_text([], 0.0, 1.0)


def test_compute_mesh_runs_gmsh_and_reports_a_missing_binary(tmp_path, monkeypatch):
    """compute_mesh (mesher.rs:481-519) against a stand-in gmsh that re-reads the .geo: the script parses, its
    loops close, geom.geo is removed afterwards; without a gmsh on PATH the error is the reference's."""
    monkeypatch.chdir(tmp_path)
    outer = [Vertex(0, 0), Vertex(40, 0), Vertex(40, 20), Vertex(0, 20)]
    hole = [Vertex(10, 5), Vertex(20, 5), Vertex(20, 15), Vertex(10, 15)]
    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    with pytest.raises(MagnetiteError, match=r"^Mesher error: Gmsh failed: No such file or directory \(os error 2\)$"):
        geometry.compute_mesh([outer, hole], "geom.msh", 0.0, 4.0, quiet=True)
    _install_fake_gmsh(tmp_path, monkeypatch)
    geometry.compute_mesh([outer, hole], "geom.msh", 0.0, 4.0, quiet=True)
    assert not (tmp_path / "geom.geo").exists()
    nodes, elements = geometry.parse_mesh("geom.msh")
    assert len(nodes) > 50 and len(elements) > 60
    cx = np.array([np.mean([nodes[i].vertex.x for i in e.nodes]) for e in elements])
    cy = np.array([np.mean([nodes[i].vertex.y for i in e.nodes]) for e in elements])
    assert not ((cx > 10) & (cx < 20) & (cy > 5) & (cy < 15)).any()          # the hole stayed a hole


def test_mesher_run_strings_the_input_side_together(tmp_path, monkeypatch, capsys):
    """mesher::run (mesher.rs:939-974) on the reference's tensile example (CSV outline + input.json) with the
    stand-in gmsh; check_ccw's areas come from the oracle here (on a GPU box they come from mag_element_area)."""
    import shutil
    from magnetite_b200 import mesher, solver
    monkeypatch.chdir(tmp_path)
    _install_fake_gmsh(tmp_path, monkeypatch)
    monkeypatch.setattr(solver, "element_areas", lambda mesh, ctx=None: O.element_area(O.Mesh(mesh)))
    src = ROOT / "tests" / "golden"
    shutil.copy(src / "tensile_input.json", tmp_path / "input.json")
    xs = [-12, 12, 12, -12]; ys = [-3, -3, 3, 3]
    (tmp_path / "vertices.csv").write_text("x,y\n" + "".join(f"{a},{b}\n" for a, b in zip(xs, ys)))
    nodes, elements, meta = mesher.run(["vertices.csv"], "input.json")
    out = capsys.readouterr().out
    assert "info: building .geo for Gmsh with 0.000< CL < 0.300" in out and "info: running gmsh..." in out
    assert f"info: loaded {len(nodes)} nodes and {len(elements)} elements" in out and "info: loaded 2 boundary rules" in out
    assert not (tmp_path / "geom.msh").exists() and not (tmp_path / "geom.geo").exists()
    assert (meta.youngs_modulus, meta.poisson_ratio, meta.part_thickness) == (69e9, 0.33, 0.5)
    clamped = [n for n in nodes if -12 < n.vertex.x < -10]
    pulled = [n for n in nodes if 10 < n.vertex.x < 12]
    on_the_edge = [n for n in nodes if abs(n.vertex.x) == 12]               # strict > / < (mesher.rs:915-918): not selected
    assert on_the_edge and all((n.ux, n.uy, n.fx, n.fy) == (None, None, 0.0, 0.0) for n in on_the_edge)
    assert clamped and all((n.ux, n.uy, n.fx, n.fy) == (0.0, 0.0, None, None) for n in clamped)
    assert pulled and all((n.ux, n.uy, n.fx, n.fy) == (3.0, None, None, 0.0) for n in pulled)
    free = [n for n in nodes if -10 < n.vertex.x < 10]
    assert free and all((n.ux, n.uy, n.fx, n.fy) == (None, None, 0.0, 0.0) for n in free)
    areas = O.element_area(O.Mesh(MeshSoA.from_aos(nodes, elements)))
    assert (areas < 0).all()                               # CL 0.3: every triangle is below 1.0 and got reversed
    with pytest.raises(MagnetiteError, match="Unrecognized geometry filetype outline.dxf"):
        mesher.run(["outline.dxf"], "input.json", quiet=True)


def test_cpp_outline_path_agrees_with_the_python_mirror(tmp_path, monkeypatch):
    """The C++ host layer's parse_csv / build_geo / compute_mesh (host/magnetite_io.cpp) against the Python
    mirror: byte-identical .geo scripts on random outlines, the same mesh through the stand-in gmsh, the same
    errors.  No GPU: `--geo` and `--mesh` stop before the solver."""
    import subprocess
    subprocess.run(["make", "-C", str(ROOT / "host")], check=True, capture_output=True)
    exe = str(ROOT / "host" / "magnetite_b200")
    inp = str(GOLDEN / "tensile_input.json")
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(8)
    for case in range(12):
        files, containers = [], []
        for c in range(int(rng.integers(1, 5))):
            n = int(rng.integers(3, 9))
            xs = rng.normal(size=n) * 10.0 ** int(rng.integers(-2, 4)); ys = np.round(rng.normal(size=n) * 50) / 4
            swap = bool(rng.integers(0, 2))                      # column order and padding are free (mesher.rs:274-284)
            name = f"c{case}_{c}.csv"
            rows = [f" {b!r} ,{a!r}" if swap else f"{a!r},  {b!r}" for a, b in zip(xs.tolist(), ys.tolist())]
            (tmp_path / name).write_text(("y, x" if swap else "x,y") + "\n" + "\n".join(rows) + "\n\n")
            files.append(name)
            containers.append(geometry.parse_csv(name))
            assert [(v.x, v.y) for v in containers[-1]] == list(zip(xs.tolist(), ys.tolist()))
        r = subprocess.run([exe, "--geo", f"out{case}.geo", inp, *files], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert (tmp_path / f"out{case}.geo").read_text() == geometry.geo_text(containers, 0.0, float(np.float32(0.3)))
    # errors of parse_csv (mesher.rs:253-299)
    (tmp_path / "nohdr.csv").write_text("a,b\n1,2\n")
    (tmp_path / "bad.csv").write_text("x,y\n1,two\n")
    for name, msg in (("nohdr.csv", "Error in csv file: Missing x and/or y field"), ("bad.csv", "Non-float value in csv points"),
                      ("missing.csv", "Unable to open csv file missing.csv")):
        r = subprocess.run([exe, "--geo", "e.geo", inp, name], capture_output=True, text=True)
        assert r.returncode == 1 and r.stderr.strip() == f"Received error: Input error: {msg}"
        with pytest.raises(MagnetiteError, match=msg):
            geometry.parse_csv(name)
    # gmsh: not installed -> the reference's Mesher error; installed (stand-in) -> the same mesh as the Python mirror gets
    (tmp_path / "outer.csv").write_text("x,y\n0,0\n40,0\n40,20\n0,20\n")
    (tmp_path / "hole.csv").write_text("x,y\n10,5\n20,5\n20,15\n10,15\n")
    (tmp_path / "input.json").write_text(Path(inp).read_text().replace('"characteristic_length_max": 0.3', '"characteristic_length_max": 4'))
    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    r = subprocess.run([exe, "--mesh", "cpp.msh", "input.json", "outer.csv", "hole.csv"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.strip() == "Received error: Mesher error: Gmsh failed: No such file or directory (os error 2)"
    _install_fake_gmsh(tmp_path, monkeypatch)
    r = subprocess.run([exe, "--mesh", "cpp.msh", "input.json", "outer.csv", "hole.csv"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "info: building .geo for Gmsh with 0.000< CL < 4.000" in r.stdout and "info: running gmsh..." in r.stdout
    assert not (tmp_path / "geom.geo").exists()
    geometry.compute_mesh([geometry.parse_csv("outer.csv"), geometry.parse_csv("hole.csv")], "py.msh", 0.0, 4.0, quiet=True)
    assert (tmp_path / "cpp.msh").read_bytes() == (tmp_path / "py.msh").read_bytes()
    nodes, elements = geometry.parse_mesh("py.msh")
    assert r.stdout.splitlines()[-1] == f"nodes {len(nodes)} elements {len(elements)}"


def _oracle_backed_solver(monkeypatch):
    """Stand-ins for the two GPU-backed functions the entry point reaches, so its plumbing can run on the CPU."""
    from magnetite_b200 import solver

    def fake_run(nodes, elements, meta, options=None, quiet=False, reorder=False):
        res = O.run(O.Mesh(MeshSoA.from_aos(nodes, elements)), meta, O.cg_options(), dense=False)
        for i, nd in enumerate(nodes):
            nd.ux, nd.uy, nd.fx, nd.fy = (float(res[k][i]) for k in ("ux", "uy", "fx", "fy"))
        for i, el in enumerate(elements):
            el.stress = float(res["stress"][i])
        fake_run.calls.append({"reorder": reorder, "n": len(nodes)})

    fake_run.calls = []
    monkeypatch.setattr(solver, "element_areas", lambda mesh, ctx=None: O.element_area(O.Mesh(mesh)))
    monkeypatch.setattr(solver, "run", fake_run)
    return fake_run


def test_python_entry_point_follows_main_rs(tmp_path, monkeypatch, capsys):
    """`python -m magnetite_b200` (main.rs:21-76): the gmsh route, the .msh route and the stand-in route all end in
    nodes.csv / elements.csv; errors print `Received error: ...` and return 1 (main.rs:43-51)."""
    from magnetite_b200 import __main__ as cli
    monkeypatch.chdir(tmp_path)
    fake = _oracle_backed_solver(monkeypatch)
    _install_fake_gmsh(tmp_path, monkeypatch)
    (tmp_path / "outer.csv").write_text("x,y\n-12,-3\n12,-3\n12,3\n-12,3\n")
    inp = tmp_path / "input.json"
    inp.write_text((GOLDEN / "tensile_input.json").read_text().replace('"characteristic_length_max": 0.3', '"characteristic_length_max": 1.5'))
    # 1. outlines through gmsh (stand-in binary), flags after the positionals
    assert cli.main(["input.json", "outer.csv", "--skip", "--reorder"]) == 0
    out = capsys.readouterr().out
    assert "info: running gmsh..." in out and "info: wrote output to nodes.csv and elements.csv" in out
    assert fake.calls[-1]["reorder"] is True
    n_rows = (tmp_path / "nodes.csv").read_text().splitlines()
    assert n_rows[0] == "x,y,ux,uy" and len(n_rows) == fake.calls[-1]["n"] + 1
    ux = np.array([float(r.split(",")[2]) for r in n_rows[1:]])
    # both rules arrived; the end faces themselves (x = +-12, on the rules' edges, hence free) overshoot a little
    assert (ux == 3.0).sum() >= 2 and (ux == 0.0).sum() >= 2 and -0.5 < ux.min() <= 0.0 and 3.0 <= ux.max() < 3.5
    assert (tmp_path / "elements.csv").read_text().splitlines()[0] == "n0,n1,n2,stress"
    assert not (tmp_path / "geom.msh").exists()
    first = (tmp_path / "nodes.csv").read_bytes()
    # 2. the same mesh handed over as a .msh file: same CSVs, and the file is left alone
    geometry.compute_mesh([geometry.parse_csv("outer.csv")], "kept.msh", 0.0, 1.5, quiet=True)
    assert cli.main(["input.json", "kept.msh"]) == 0
    assert (tmp_path / "nodes.csv").read_bytes() == first and (tmp_path / "kept.msh").exists() and fake.calls[-1]["reorder"] is False
    # 3. no gmsh at all: the built-in mesher
    monkeypatch.setenv("PATH", str(tmp_path / "nowhere"))
    capsys.readouterr()
    assert cli.main(["input.json", "--standin", "outer.csv"]) == 0
    assert (tmp_path / "nodes.csv").read_bytes() == first                      # the stand-in gmsh IS the built-in mesher
    # errors: main.rs:43-51
    assert cli.main(["input.json", "outer.csv"]) == 1
    assert capsys.readouterr().err.strip() == "Received error: Mesher error: Gmsh failed: No such file or directory (os error 2)"
    assert cli.main(["input.json", "outline.dxf"]) == 1
    assert capsys.readouterr().err.strip() == "Received error: Input error: Unrecognized geometry filetype outline.dxf"
    assert cli.main(["nope.json", "outer.csv"]) == 1
    assert capsys.readouterr().err.strip() == "Received error: Input error: Unable to open input file nope.json"


SVG_TRICKY = """<?xml version='1.0' encoding="UTF-8"?>
<!DOCTYPE svg PUBLIC "-//W3C//DTD SVG 1.1//EN" "http://www.w3.org/Graphics/SVG/1.1/DTD/svg11.dtd" [ <!ENTITY unused "x"> ]>
<!-- a comment with a <polygon id="OUTER" points="0 0 1 1 2 2"/> inside must not count -->
<svg:svg xmlns:svg="http://www.w3.org/2000/svg" viewBox="0 0 200 100">
  <svg:defs><svg:style><![CDATA[ .a > .b { fill: #fff; } <rect id="OUTER" width="1" height="1"/> ]]></svg:style></svg:defs>
  <svg:g id=' OUTER&#45;shell'>
    <svg:polygon class="a &amp; b"
        points='0 0 200 0
 200 100 200.2 100 0 100 0 0'/>
  </svg:g>
  <svg:polyline id="INNER&#x2d;1" points="20 20  40 20 40 40 20 40"/>
  <svg:g id="INNER-2"><svg:rect width="10" height="12.5"/></svg:g>
  <svg:rect id="note" x="1" y="1" width="2" height="2"/>
  <svg:polygon id = "INNER-3" points="100 50 120 50 110 70"></svg:polygon>
</svg:svg>
"""


def test_cpp_parse_svg_agrees_with_the_python_mirror(tmp_path, monkeypatch):
    """host/magnetite_io.cpp parse_svg (own small XML reader) against magnetite_b200.geometry.parse_svg
    (ElementTree), through the .geo script both write: the repo's test SVG, a deliberately awkward document, and
    every rejection of mesher.rs:26-244."""
    import subprocess
    subprocess.run(["make", "-C", str(ROOT / "host")], check=True, capture_output=True)
    exe = str(ROOT / "host" / "magnetite_b200")
    monkeypatch.chdir(tmp_path)
    inp = tmp_path / "input.json"
    for k, (text, cl_min) in enumerate(((SVG, 0.5), (SVG, 0.0), (SVG_TRICKY, 0.5), (SVG_TRICKY, 0.0))):
        (tmp_path / f"g{k}.svg").write_text(text)
        inp.write_text((GOLDEN / "tensile_input.json").read_text().replace('"characteristic_length_min": 0', f'"characteristic_length_min": {cl_min}'))
        containers = geometry.parse_svg(f"g{k}.svg", cl_min)
        r = subprocess.run([exe, "--geo", f"g{k}.geo", "input.json", "ignored.csv", f"g{k}.svg", "never-read.dxf"], capture_output=True, text=True)
        assert r.returncode == 1 and "Unable to open csv file ignored.csv" in r.stderr         # files before the .svg are still read
        r = subprocess.run([exe, "--geo", f"g{k}.geo", "input.json", f"g{k}.svg", "never-read.dxf"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr                                                     # the .svg ends the list (mesher.rs:948-950)
        assert (tmp_path / f"g{k}.geo").read_text() == geometry.geo_text(containers, cl_min, float(np.float32(0.3)))
    tricky = geometry.parse_svg("g3.svg", 0.0)
    assert [len(c) for c in tricky] == [5, 4, 3, 4]              # OUTER (closing repeat dropped), polyline, polygon, then the rect
    assert len(geometry.parse_svg("g2.svg", 0.5)[0]) == 4       # the vertex 0.2 away from its predecessor is skipped at CL 0.5
    head = '<svg xmlns="http://www.w3.org/2000/svg">'
    bad = {
        "no_outer": (head + '<polygon id="INNER" points="0 0 1 0 1 1"/></svg>', "No OUTER geometry"),
        "two_outer": (head + '<rect id="OUTER" width="5" height="5"/><polygon id="OUTER-2" points="0 0 1 0 1 1"/></svg>', "Multiple OUTER geometries in SVG"),
        "no_id": (head + '<polygon points="0 0 1 0 1 1"/></svg>', "Error in svg file. Missing id field on polyline"),
        "no_points": (head + '<polygon id="OUTER"/></svg>', "Error in svg file. No points in polyline element"),
        "no_width": (head + '<rect id="OUTER" height="5"/></svg>', "Error in svg file. No width/height definition in rectangle."),
        "not_float": (head + '<polygon id="OUTER" points="0 0 1 zero 1 1"/></svg>', "Non-float value in svg points"),
    }
    inp.write_text((GOLDEN / "tensile_input.json").read_text())
    for name, (text, msg) in bad.items():
        (tmp_path / f"{name}.svg").write_text(text)
        with pytest.raises(MagnetiteError, match=re.escape(msg)):
            geometry.parse_svg(f"{name}.svg", 0.0)
        r = subprocess.run([exe, "--geo", "e.geo", "input.json", f"{name}.svg"], capture_output=True, text=True)
        assert r.returncode == 1 and r.stderr.strip().startswith(f"Received error: Input error: {msg}"), (name, r.stderr)
    r = subprocess.run([exe, "--geo", "e.geo", "input.json", "absent.svg"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.strip() == "Received error: Input error: Unable to open svg file absent.svg"
