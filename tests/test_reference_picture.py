"""The numeric outputs of its own solver the reference publishes: the TRUE-SCALE deformed shapes of its three
examples — examples/linkedin-logo/output.png (readme.md:28-30), media/tensilve-results.png (the tensile example,
BASELINE config 1) and examples/cover-eample/output.png (readme.md:1) — drawn by scripts/plot.py:143-147 at
(x + ux, y + uy), the first two beside the undeformed mesh.

tests/golden/measure_reference_picture.py read those pictures in the build container and committed what they show —
the intervals the solved and the initial model cover along horizontal and vertical lines, to one pixel (0.73 units
of the 628-unit logo; 0.016 x 0.023 units of the 22 x 9 tensile bar) — as tests/golden/reference_*_picture.json.
Here the same examples (same outlines, same input.json, the stand-in mesher's triangulation instead of gmsh's:
tests/golden/example_*.npz) are solved by the oracle (CPU) and by the library (GPU) and the deformed outline is laid
over the reference's.

What this pins, to 1-2 % of the displacements: the boundary-rule semantics (which nodes are held, which are
pulled), plane STRESS (plane strain would contract the waist by 0.49 instead of 0.33 of the stretch), the assembly
and the solve — for the tensile example on the all-clockwise mesh check_ccw produces (mesher.rs:522-526: a negative
definite K) — signs and axes, and that nothing is magnified; the face colours pin the ORDER of the element
stresses and the sign rule of solver.rs:524-530.  It cannot pin rounding-level arithmetic — the
oracle header's "parity unpinned" stays true for that — but these are outputs of the reference itself.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from magnetite_b200 import meshgen
from magnetite_b200.datatypes import MeshSoA

GOLDEN = Path(__file__).resolve().parent / "golden"
NAMES = ["linkedin", "tensile", "cover"]
PICTURES = {n: json.loads((GOLDEN / f"reference_{n}_picture.json").read_text()) for n in NAMES}
# Distances are measured in PIXELS of the picture (the tensile plot's axes are not to the same scale).  An edge in
# the picture: half a pixel of anti-aliasing counted as model + half a pixel of sampling + the two triangulations'
# different boundary vertices on curved edges.  Measured: 1.41 (logo) / 1.28 (bar) for the undeformed outlines.
TOL_PX = 2.5
INSET_PX = 0.6
TEETH_PX = 3.5


def example(name):
    g = np.load(GOLDEN / f"example_{name}.npz")
    mesh = MeshSoA(g["x"], g["y"], g["n0"], g["n1"], g["n2"], g["bc_ux"], g["bc_uy"], g["bc_fx"], g["bc_fy"], g["known"])
    meta = meshgen.EXAMPLE_MATERIAL.__class__(*g["material"])
    return g, mesh, meta


def outline_edges(tri):
    """Edges that belong to exactly one triangle: the outline (outer boundary and holes)."""
    e = np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]])
    e.sort(axis=1)
    uniq, count = np.unique(e, axis=0, return_counts=True)
    return uniq[count == 1]


def picture_points(panel):
    """Every point where a line of the picture enters or leaves the model: (x, y) on the reference's outline.
    The reader counts every partly covered (anti-aliased) pixel as model, so an interval starts up to a pixel early
    and ends up to a pixel late: its ends are moved INSET_PX inwards (calibrated on the undeformed panels, whose
    geometry is known: their mean distance to the outline falls from 0.78 / 0.89 px to 0.45 / 0.53 px)."""
    dx, dy = INSET_PX / panel["pixels_per_unit_x"], INSET_PX / panel["pixels_per_unit_y"]
    pts = []
    for line in panel["along_y"]:
        pts += [(v + sign * dx, line["at"]) for iv in line["intervals"] for v, sign in zip(iv, (1, -1))]
    for line in panel["along_x"]:
        pts += [(line["at"], v + sign * dy) for iv in line["intervals"] for v, sign in zip(iv, (1, -1))]
    return np.array(pts)


def misfit(panel, px, py, tri):
    """(number of picture points, their largest, mean and 99th-percentile distance to our outline) in pixels."""
    scale = np.array([panel["pixels_per_unit_x"], panel["pixels_per_unit_y"]])
    edges = outline_edges(tri)
    a = (np.stack([px[edges[:, 0]], py[edges[:, 0]]], 1) * scale)[None]        # (1, S, 2)
    b = (np.stack([px[edges[:, 1]], py[edges[:, 1]]], 1) * scale)[None]
    p = (picture_points(panel) * scale)[:, None, :]                            # (P, 1, 2)
    ab = b - a
    t = np.clip(((p - a) * ab).sum(2) / np.maximum((ab * ab).sum(2), 1e-300), 0.0, 1.0)
    d = np.linalg.norm(p - (a + t[..., None] * ab), axis=2).min(1)
    return len(d), float(d.max()), float(d.mean()), float(np.percentile(d, 99))


def check(name, ux, uy, g):
    pic = PICTURES[name]["panels"]
    tri = np.stack([g["n0"], g["n1"], g["n2"]], 1).astype(np.int64)
    x, y = g["x"], g["y"]
    if name == "cover":
        return check_cover(pic["solved"], x, y, ux, uy, tri)
    # the geometry first: the undeformed outline is the picture's "Initial Model"
    n0, worst0, mean0, _ = misfit(pic["initial"], x, y, tri)
    assert n0 >= 200 and worst0 <= TOL_PX, (n0, worst0, mean0)
    # the solution: every point of the reference's deformed outline lies on ours, about as closely as the undeformed
    # one (measured, pixels: logo 254 points, worst 1.49 / mean 0.64 against 1.41 / 0.44 for the geometry alone;
    # bar 210 points, 2.10 / 0.74 against 1.28 / 0.53)
    n, worst, mean, _ = misfit(pic["solved"], x + ux, y + uy, tri)
    assert n >= 200 and worst <= TOL_PX and mean <= mean0 + 0.35, (n, worst, mean)
    # and the comparison has teeth: no deformation across the pull, plane strain's contraction (0.49 / 0.33 of plane
    # stress's), a 2 % error of the stretch or a magnified plot do not fit
    across = (0.0, 1.0) if name == "linkedin" else (1.0, 0.0)                 # the logo is pulled in y, the bar in x
    strain = (0.49 / 0.33, 1.0) if name == "linkedin" else (1.0, 0.49 / 0.33)
    along = (1.0, 0.98) if name == "linkedin" else (0.98, 1.0)
    for fx, fy in (across, strain, along, (1.02, 1.02)):
        assert misfit(pic["solved"], x + fx * ux, y + fy * uy, tri)[1] > TEETH_PX, (fx, fy)
    # sharper: scale our displacement along the pull by f — the picture is explained best by f = 1.00 +- 0.01 (the
    # mean distance is a V around it: logo 0.99 / 0.64 / 0.69 px at f = 0.99 / 1.00 / 1.01, bar 1.05 / 0.74 / 1.09); for
    # the logo the same holds for the lateral contraction (the bar's is 5 pixels in all and too small to weigh)
    factors = [0.97, 0.98, 0.99, 1.0, 1.01, 1.02, 1.03]
    pulls = {"linkedin": [(0, 1), (1, 0)], "tensile": [(1, 0)]}[name]
    for ax, ay in pulls:
        means = [misfit(pic["solved"], x + (f if ax else 1.0) * ux, y + (f if ay else 1.0) * uy, tri)[2] for f in factors]
        assert 0.99 <= factors[int(np.argmin(means))] <= 1.01, (ax, ay, means)


def check_cover(solved, x, y, ux, uy, tri):
    """The cover picture (readme.md:1) is cropped to the solved panel: no undeformed panel separates what the two
    meshers make of the letters' outlines from the deformation, and 2 of its 960 points sit 7 pixels off such a
    feature — so the 99th percentile stands in for the worst point (measured 1.23 px, mean 0.47), and the pull (10
    units = 41 pixels, lateral motion up to 5 units) is resolved to a few per cent: a weaker pin than the other two
    pictures, of a third geometry."""
    n, worst, mean, p99 = misfit(solved, x + ux, y + uy, tri)
    assert n >= 900 and p99 <= TOL_PX and mean <= 0.8 and worst <= 9.0, (n, worst, mean, p99)
    for fx, fy in ((1.0, 0.9), (1.0, 1.1), (0.0, 1.0), (1.5, 1.0), (1.0, 0.0)):
        assert misfit(solved, x + fx * ux, y + fy * uy, tri)[3] > TEETH_PX, (fx, fy)
    factors = [0.96, 0.97, 0.98, 0.99, 1.0, 1.01, 1.02, 1.03, 1.04]          # the stretch: best at 1.01 (0.451 px; 0.468 at 1.00)
    means = [misfit(solved, x + ux, y + f * uy, tri)[2] for f in factors]
    assert 0.98 <= factors[int(np.argmin(means))] <= 1.02, means


def elements_at(points, px, py, tri):
    """Index of the triangle that holds each point (-1: none)."""
    ax, ay, bx, by, cx, cy = px[tri[:, 0]], py[tri[:, 0]], px[tri[:, 1]], py[tri[:, 1]], px[tri[:, 2]], py[tri[:, 2]]
    det = (bx - ax) * (cy - ay) - (cx - ax) * (by - ay)
    out = np.full(len(points), -1)
    for i, (X, Y) in enumerate(points):
        l1 = ((bx - X) * (cy - Y) - (cx - X) * (by - Y)) / det
        l2 = ((cx - X) * (ay - Y) - (ax - X) * (cy - Y)) / det
        m = np.minimum(np.minimum(l1, l2), 1.0 - l1 - l2)
        k = int(np.argmax(m))
        if m[k] >= -1e-9:
            out[i] = k
    return out


def check_stress(name, stress, ux, uy, g):
    """The pictures colour every element by its `stress` (solver.rs:524-533: sqrt(sx^2 + sy^2), negative where
    sx + sy < 1) through a matplotlib colormap between the smallest and the largest value.  Those two depend on the
    single most stressed corner triangle, i.e. on the triangulation, so only the ORDER is comparable: the rank
    correlation between the face colour on a grid of points (a monotone function of the plotted value: red minus
    blue for the logo's "coolwarm", red for the dark-to-bright maps of the other two) and our stress in the element
    under each point.  Measured: logo 0.980 (0.947 if the sign is dropped, 0.942 for a von Mises stress), bar 0.979,
    cover 0.975 (0.958 without the sign)."""
    from scipy.stats import spearmanr
    rgb = np.array(PICTURES[name]["panels"]["solved"]["face_rgb"])
    value = rgb[:, 2] - rgb[:, 4] if name == "linkedin" else rgb[:, 2]
    tri = np.stack([g["n0"], g["n1"], g["n2"]], 1).astype(np.int64)
    el = elements_at(rgb[:, :2], g["x"] + ux, g["y"] + uy, tri)
    inside = el >= 0
    assert len(rgb) >= 900 and inside.mean() >= 0.99
    rho = spearmanr(value[inside], stress[el[inside]]).statistic
    assert rho >= 0.965, rho
    if name != "tensile":                         # the bar is in tension nearly everywhere: 5 % negative elements
        assert spearmanr(value[inside], np.abs(stress[el[inside]])).statistic <= rho - 0.01     # the sign rule is visible


@pytest.mark.parametrize("name", NAMES)
def test_oracle_solution_lies_on_the_reference_picture(name):
    from oracle import oracle as O
    g, mesh, meta = example(name)
    res = O.run(O.Mesh(mesh), meta, O.cg_options(), dense=False)       # the reference's solver semantics
    check(name, res["ux"], res["uy"], g)
    check_stress(name, res["stress"], res["ux"], res["uy"], g)
    # ... and the committed fixture is that solution
    assert np.linalg.norm(res["ux"] - g["ux"]) <= 1e-9 * np.linalg.norm(g["ux"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_solution_lies_on_the_reference_picture(ctx, name):
    from magnetite_b200 import _lib, solver
    g, mesh, meta = example(name)
    sol = solver.solve_soa(mesh, meta, ctx, _lib.default_options(compat=1))
    check(name, sol.ux, sol.uy, g)
    check_stress(name, sol.stress, sol.ux, sol.uy, g)
    sol = solver.solve_soa(mesh, meta, ctx, _lib.default_options())   # the library's default solver, same picture
    check(name, sol.ux, sol.uy, g)
