"""CPU: the C-ABI library loads and exports every symbol include/magnetite_b200.h declares, the
product fails loudly without a GPU (no CPU fallback), and the host-side mirrors of the
reference interface (datatypes, error, mesher BC rules, csv_output) behave like the reference."""
import ctypes as C
import json
import re
from pathlib import Path

import numpy as np
import pytest

from magnetite_b200 import _lib, mesher, meshgen, post_processor
from magnetite_b200.datatypes import Element, MeshSoA, Node, Vertex
from magnetite_b200.error import MagnetiteError

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"


def header_symbols():
    text = (ROOT / "include" / "magnetite_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mag_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.declared_symbols()) == syms, "ctypes prototypes drifted from the header"
    assert lib.mag_abi_version() == _lib.ABI_VERSION == 3


def test_struct_layouts_match_header(built):
    # sizes computed by hand from include/magnetite_b200.h (LP64)
    assert C.sizeof(_lib.MagMesh) == 2 * 8 + 10 * 8 + 8
    assert C.sizeof(_lib.MagMaterial) == 24
    assert C.sizeof(_lib.MagOptions) == 8 + 8 + 8 + 12 * 4 + 8
    assert C.sizeof(_lib.MagResult) == 6 * 8 + 8
    o = _lib.default_options()
    assert (o.rel_tol, o.abs_tol, o.max_iter, o.precond, o.compat, o.drop_exact_zeros) == (1e-9, 1e-4, 10_000_000, 3, 0, 1)


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(MagnetiteError) as ei:
        _lib.Context(0)
    assert ei.value.kind == "Solver" and ei.value.code == _lib.MAG_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|magnetite_oracle|liboracle|orc_[a-z]+\(", re.M)
    files = [p for p in (ROOT / "magnetite_b200").rglob("*") if p.suffix in (".py", ".cu", ".cuh", ".h")]
    assert len(files) > 10
    for path in files:
        assert not pat.search(path.read_text(errors="ignore")), f"{path} references the oracle"


def test_partition_nodes(built):
    lib = _lib.load()
    lo, hi = C.c_uint64(), C.c_uint64()
    covered = []
    for r in range(8):
        assert lib.mag_partition_nodes(8006001, 8, r, C.byref(lo), C.byref(hi)) == 0
        covered.append((lo.value, hi.value))
    assert covered[0][0] == 0 and covered[-1][1] == 8006001
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    sizes = [b - a for a, b in covered]
    assert max(sizes) - min(sizes) <= 1
    assert lib.mag_partition_nodes(10, 0, 0, C.byref(lo), C.byref(hi)) == _lib.MAG_ERR_BAD_ARG


def test_error_display_matches_reference():
    assert str(MagnetiteError.Solver("boom")) == "Solver error: boom"
    assert str(MagnetiteError.PostProcessor("x")) == "Post Processor error: x"
    assert str(MagnetiteError.Input("x")) == "Input error: x"
    assert str(MagnetiteError.Mesher("x")) == "Mesher error: x"


def test_rust_display_formatting():
    f = post_processor.rust_f64_display
    assert f(3.0) == "3" and f(-0.0) == "-0" and f(0.0) == "0"
    assert f(69e9) == "69000000000" and f(1e-7) == "0.0000001"
    assert f(0.1 + 0.2) == "0.30000000000000004" and f(-4.5) == "-4.5"
    assert f(1e21) == "1000000000000000000000" and f(1.5e-10) == "0.00000000015"
    assert f(1.2345678901234567e25) == "12345678901234566000000000"      # shortest digits, zero padded (not the exact expansion)
    assert f(float("nan")) == "NaN" and f(float("inf")) == "inf" and f(float("-inf")) == "-inf"
    for v in np.random.default_rng(0).normal(size=200) * 10.0 ** np.random.default_rng(1).integers(-12, 12, 200):
        assert float(f(v)) == v and "e" not in f(v)


def test_csv_output_format(tmp_path):
    nodes = [Node(Vertex(0.0, 4.5), 0.0, -0.0, 1.0, 2.0), Node(Vertex(-11.0, 0.1), 3.0, 1e-7, 0.0, 0.0)]
    els = [Element([1, 0, 1], -345000000.0)]
    a, b = tmp_path / "nodes.csv", tmp_path / "elements.csv"
    post_processor.csv_output(els, nodes, str(a), str(b), quiet=True)
    assert a.read_bytes() == b"x,y,ux,uy\n0,4.5,0,-0\n-11,0.1,3,0.0000001\n"
    assert b.read_bytes() == b"n0,n1,n2,stress\n1,0,1,-345000000\n"
    with pytest.raises(MagnetiteError):
        post_processor.csv_output(els, nodes, str(tmp_path / "nope" / "n.csv"), str(b), quiet=True)
    with pytest.raises(MagnetiteError):
        post_processor.csv_output([Element([0, 1, 0])], nodes, str(a), str(b), quiet=True)


def test_boundary_rules_follow_the_code_not_the_docs():
    data = mesher.load_input_file(str(GOLDEN / "tensile_input.json"))
    meta = mesher.parse_input_metadata(data)
    assert (meta.youngs_modulus, meta.poisson_ratio, meta.part_thickness) == (69e9, 0.33, 0.5)
    xs = [-11.0, -10.0, -9.0, 0.0, 10.0, 10.5, 12.0]
    nodes = mesher.default_nodes(xs, [0.0] * len(xs))
    assert nodes[0].ux is None and nodes[0].fx == 0.0                    # mesher.rs:615-624
    mesher.apply_boundary_conditions(data, nodes)
    assert (nodes[0].ux, nodes[0].uy, nodes[0].fx, nodes[0].fy) == (0.0, 0.0, None, None)
    assert nodes[1].ux is None and nodes[1].fx == 0.0                    # x == x_max: strict <, not selected
    assert nodes[4].ux is None                                           # x == x_min: strict >
    assert (nodes[5].ux, nodes[5].uy, nodes[5].fx, nodes[5].fy) == (3.0, None, None, 0.0)
    assert nodes[6].ux is None
    soa = MeshSoA.from_aos(nodes, [])
    assert list(soa.known) == [3, 12, 12, 12, 12, 9, 12]
    n2, _ = soa.to_aos()
    assert n2[5].ux == 3.0 and n2[5].uy is None


def test_boundary_rule_validation(tmp_path):
    base = json.loads((GOLDEN / "tensile_input.json").read_text())
    bad = json.loads(json.dumps(base)); bad["boundary_conditions"]["load"]["targets"]["fx"] = 1.0
    with pytest.raises(MagnetiteError, match="over-constrained in x-axis"):
        mesher.parse_boundary_rules(bad)
    bad = json.loads(json.dumps(base)); bad["boundary_conditions"]["load"]["targets"]["fy"] = None
    with pytest.raises(MagnetiteError, match="under-constrained in y-axis"):
        mesher.parse_boundary_rules(bad)
    bad = json.loads(json.dumps(base)); bad["boundary_conditions"]["load"]["region"]["x_target_min"] = 99
    with pytest.raises(MagnetiteError, match="x_target_min greater"):
        mesher.parse_boundary_rules(bad)
    p = tmp_path / "x.json"; p.write_text("{\"metadata\": {}}")
    with pytest.raises(MagnetiteError, match="boundary_conditions"):
        mesher.load_input_file(str(p))
    with pytest.raises(MagnetiteError, match="Unable to open"):
        mesher.load_input_file(str(tmp_path / "missing.json"))


def test_meshgen_invariants():
    m = meshgen.plate(1000, 500)
    assert m.n_elems == 1_000_000 and m.n_nodes == 501_501
    m = meshgen.plate(7, 5)
    x, y = m.x, m.y
    a = 0.5 * (x[m.n0] * (y[m.n1] - y[m.n2]) + x[m.n1] * (y[m.n2] - y[m.n0]) + x[m.n2] * (y[m.n0] - y[m.n1]))
    assert (a == 2.0).all()                                              # CCW, area >= 1: check_ccw no-op
    p = meshgen.perforated_plate(64, 32, pitch=16, radius=4)
    used = np.zeros(p.n_nodes, bool); used[p.n0] = used[p.n1] = used[p.n2] = True
    assert used.all() and p.n_elems < 2 * 64 * 32
    j = meshgen.jitter(m)
    assert (j.x != m.x).any() and np.array_equal(j.known, m.known)


def test_cpp_host_layer_formats_like_rust(built):
    """host/plate_demo (C++ mirror of solver::run / csv_output) — the float formatter and the error
    Display need no GPU."""
    import subprocess
    exe = ROOT / "host" / "plate_demo"
    subprocess.run(["make", "-C", str(ROOT / "host")], check=True, capture_output=True)
    r = subprocess.run([str(exe), "--format-selftest"], capture_output=True, text=True)
    assert r.returncode == 0 and "FORMAT_OK" in r.stdout, r.stdout + r.stderr
    # the C++ formatter is its own implementation: random bit patterns (every exponent, subnormals, both zeros,
    # NaN and the infinities) against the Python mirror
    rng = np.random.default_rng(17)
    bits = np.concatenate([rng.integers(0, 2 ** 64, 4000, dtype=np.uint64),
                           np.array([0, 1 << 63, 1, 0x7FF0000000000000, 0xFFF0000000000000, 0x7FF8000000000000,
                                     0x7FEFFFFFFFFFFFFF, 0x0010000000000000, 0x000FFFFFFFFFFFFF], np.uint64)])
    r = subprocess.run([str(exe), "--format-stdin"], input="".join(f"{int(b):016x}\n" for b in bits), capture_output=True, text=True)
    got = r.stdout.splitlines()
    assert r.returncode == 0 and len(got) == len(bits)
    for text, v in zip(got, bits.view(np.float64)):
        assert text == post_processor.rust_f64_display(float(v)), (hex(int(np.float64(v).view(np.uint64))), text)


def test_stress_strain_matrix_matches_oracle(built):
    """solver::compute_stress_strain_matrix is host arithmetic in the library: bit-identical to the oracle."""
    from magnetite_b200 import solver
    from oracle import oracle as O
    for nu, E in ((0.33, 69e9), (0.25, 210e9), (0.0, 1.0), (0.49, 3e6)):
        assert np.array_equal(solver.compute_stress_strain_matrix(nu, E), O.stress_strain(nu, E))


def test_cpp_cli_parses_input_json_like_the_python_mirror(built):
    """host/magnetite_b200 --dump-rules: the C++ mirror of load_input_file / parse_input_metadata /
    parse_boundary_rules agrees with magnetite_b200.mesher on the tensile example (no GPU needed)."""
    import subprocess
    subprocess.run(["make", "-C", str(ROOT / "host")], check=True, capture_output=True)
    path = str(GOLDEN / "tensile_input.json")
    r = subprocess.run([str(ROOT / "host" / "magnetite_b200"), "--dump-rules", path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    f = post_processor.rust_f64_display
    data = mesher.load_input_file(path)
    meta = mesher.parse_input_metadata(data)
    want = [f"metadata {f(meta.youngs_modulus)} {f(meta.poisson_ratio)} {f(meta.part_thickness)} "
            f"{f(meta.characteristic_length_min)} {f(meta.characteristic_length_max)}"]
    o = lambda v: "None" if v is None else f(v)
    for rule in mesher.parse_boundary_rules(data):
        g, tg = rule.region, rule.target
        want.append(f"rule {rule.name} region {f(g.x_min)} {f(g.x_max)} {f(g.y_min)} {f(g.y_max)} "
                    f"targets {o(tg.ux)} {o(tg.uy)} {o(tg.fx)} {o(tg.fy)}")
    assert r.stdout.splitlines() == want
    bad = subprocess.run([str(ROOT / "host" / "magnetite_b200"), path, "outline.dxf"], capture_output=True, text=True)
    assert bad.returncode == 1 and bad.stderr.strip() == "Received error: Input error: Unrecognized geometry filetype outline.dxf"
    bad = subprocess.run([str(ROOT / "host" / "magnetite_b200"), path, "absent.svg"], capture_output=True, text=True)
    assert bad.returncode == 1 and bad.stderr.strip() == "Received error: Input error: Unable to open svg file absent.svg"


def test_library_csv_writer_matches_python_writer(tmp_path, built):
    """mag_csv_output (host-only code in the library) writes byte-identical files to the Python writer,
    including awkward floats; mag_format_f64 agrees with the Rust-style formatter on random values."""
    import ctypes as C
    rng = np.random.default_rng(11)
    n, e = 2000, 3000
    mag = 10.0 ** rng.integers(-14, 14, n)
    x = rng.normal(size=n) * mag; y = np.round(rng.normal(size=n) * 10) ; ux = rng.normal(size=n) * 1e-3
    uy = np.where(rng.random(n) < 0.1, 0.0, rng.normal(size=n)); uy[:3] = [-0.0, 3.0, 1e-7]
    n0, n1, n2 = (rng.integers(0, n, e).astype(np.uint32) for _ in range(3))
    stress = rng.normal(size=e) * 1e8; stress[:4] = [float("nan"), float("inf"), -float("inf"), 69e9]
    post_processor.write_csv_arrays(x, y, ux, uy, n0, n1, n2, stress, str(tmp_path / "n_py.csv"), str(tmp_path / "e_py.csv"))
    post_processor.write_csv_fast(x, y, ux, uy, n0, n1, n2, stress, str(tmp_path / "n_c.csv"), str(tmp_path / "e_c.csv"))
    assert (tmp_path / "n_py.csv").read_bytes() == (tmp_path / "n_c.csv").read_bytes()
    assert (tmp_path / "e_py.csv").read_bytes() == (tmp_path / "e_c.csv").read_bytes()
    lib = _lib.load()
    buf = C.create_string_buffer(512)
    for v in list(rng.normal(size=300) * 10.0 ** rng.integers(-300, 300, 300)) + [5e-324, 1.7976931348623157e308, -0.0]:
        lib.mag_format_f64(float(v), buf)
        assert buf.value.decode() == post_processor.rust_f64_display(v)
    # post_processor.rs:24-39: the io::Error's Display inside MagnetiteError::Solver, from both writers
    want = r"^Solver error: Failed to create nodes.csv: No such file or directory \(os error 2\)$"
    for writer in (post_processor.write_csv_fast, post_processor.write_csv_arrays):
        with pytest.raises(MagnetiteError, match=want):
            writer(x, y, ux, uy, n0, n1, n2, stress, str(tmp_path / "no" / "n.csv"), str(tmp_path / "e.csv"))
        with pytest.raises(MagnetiteError, match=want.replace("nodes", "elements")):
            writer(x, y, ux, uy, n0, n1, n2, stress, str(tmp_path / "n.csv"), str(tmp_path / "no" / "e.csv"))


def test_format_f64_properties_over_all_finite_doubles(built):
    """Rust's `{}` for f64 as the CSV contract needs it (post_processor.rs:44-75), fuzzed over the whole double
    range: the library's formatter agrees with the Python mirror, the text parses back to the same bits,
    never uses an exponent, never ends in '.0', and carries the sign of negative zero."""
    import ctypes as C
    import math
    import struct
    from hypothesis import given, settings, strategies as st
    lib = _lib.load()
    buf = C.create_string_buffer(512)

    @settings(max_examples=3000, deadline=None)
    @given(st.floats(allow_nan=False, allow_infinity=False, width=64))
    def check(v):
        n = lib.mag_format_f64(v, buf)
        text = buf.value.decode()
        assert n == len(text) and text == post_processor.rust_f64_display(v)
        assert struct.pack("<d", float(text)) == struct.pack("<d", v)
        assert "e" not in text.lower() and not text.endswith(".0") and not text.endswith(".")
        assert text.startswith("-") == (math.copysign(1.0, v) < 0)
        assert n <= 330                       # "-0." + 323 zeros + 17 digits at most: inside the 400-byte contract

    check()
    # powers of ten and their neighbours: where digit counts and the decimal point position change
    for e in range(-323, 309):
        for v in (float(f"1e{e}"), np.nextafter(float(f"1e{e}"), np.inf), np.nextafter(float(f"1e{e}"), -np.inf)):
            lib.mag_format_f64(float(v), buf)
            assert buf.value.decode() == post_processor.rust_f64_display(float(v)) and float(buf.value) == float(v)


def _random_input_json(rng):
    """A random input.json: usually valid, sometimes broken in one of the ways mesher.rs:713-930 rejects."""
    def num():
        kind = int(rng.integers(0, 3))
        if kind == 0:
            return float(rng.integers(-50, 50))                                   # "12.0"
        if kind == 1:
            return round(float(rng.normal()) * 10.0 ** int(rng.integers(-3, 4)), 6)   # "-0.004217", "1830.25"
        return int(rng.integers(-5, 5))                                           # a JSON integer
    md = {"part_thickness": abs(num()) + 0.1, "material_elasticity": 69e9, "poisson_ratio": 0.33,
          "characteristic_length_min": 0, "characteristic_length_max": 0.3}
    rules = {}
    for i in range(int(rng.integers(0, 4))):
        region = {}
        for axis in "xy":
            lo, hi = sorted([num(), num()])
            if rng.random() < 0.7: region[f"{axis}_target_min"] = lo
            if rng.random() < 0.7: region[f"{axis}_target_max"] = hi
        tx = {"ux": num(), "fx": None} if rng.random() < 0.5 else {"ux": None, "fx": num()}
        ty = {"uy": num(), "fy": None} if rng.random() < 0.5 else {"uy": None, "fy": num()}
        rules[f"rule{i}"] = {"region": region, "targets": {**tx, **ty}}
    data = {"metadata": md, "boundary_conditions": rules}
    fault = rng.integers(0, 14)
    first = next(iter(rules), None)
    if fault == 0: del data["metadata"]
    elif fault == 1: del data["boundary_conditions"]
    elif fault == 2: del md[rng.choice(["part_thickness", "material_elasticity", "poisson_ratio"])]
    elif fault == 3: md[rng.choice(["characteristic_length_min", "characteristic_length_max"])] = "fine"
    elif fault == 4 and first: del rules[first]["region"]
    elif fault == 5 and first: del rules[first]["targets"]
    elif fault == 6 and first: rules[first]["region"]["x_target_min"] = "left"
    elif fault == 7 and first: rules[first]["region"].update(y_target_min=5, y_target_max=-5)
    elif fault == 8 and first: rules[first]["targets"].update(ux=1.0, fx=2.0)
    elif fault == 9 and first: rules[first]["targets"].update(uy=None, fy=None)
    return data


def test_cpp_and_python_input_semantics_agree_on_random_files(tmp_path, built):
    """Differential test of the two host mirrors of load_input_file / parse_input_metadata / the boundary-rule
    validation (mesher.rs:713-930): same parsed values on valid files, same error text on rejected ones."""
    import json
    import subprocess
    subprocess.run(["make", "-C", str(ROOT / "host")], check=True, capture_output=True)
    exe = str(ROOT / "host" / "magnetite_b200")
    f = post_processor.rust_f64_display
    o = lambda v: "None" if v is None else f(v)          # noqa: E731
    rng = np.random.default_rng(2024)
    outcomes = set()
    for case in range(80):
        path = tmp_path / f"in{case}.json"
        path.write_text(json.dumps(_random_input_json(rng)))
        try:
            data = mesher.load_input_file(str(path))
            meta = mesher.parse_input_metadata(data)
            want = [f"metadata {f(meta.youngs_modulus)} {f(meta.poisson_ratio)} {f(meta.part_thickness)} "
                    f"{f(meta.characteristic_length_min)} {f(meta.characteristic_length_max)}"]
            for rule in mesher.parse_boundary_rules(data):
                g, tg = rule.region, rule.target
                want.append(f"rule {rule.name} region {f(g.x_min)} {f(g.x_max)} {f(g.y_min)} {f(g.y_max)} "
                            f"targets {o(tg.ux)} {o(tg.uy)} {o(tg.fx)} {o(tg.fy)}")
            error = None
        except MagnetiteError as err:
            error = str(err)
        r = subprocess.run([exe, "--dump-rules", str(path)], capture_output=True, text=True)
        if error is None:
            assert r.returncode == 0 and r.stdout.splitlines() == want, (case, path.read_text(), r.stdout, r.stderr)
            outcomes.add("ok")
        else:
            assert r.returncode == 1 and r.stderr.strip() == f"Received error: {error}", (case, path.read_text(), r.stderr, error)
            outcomes.add(re.sub(r"rule\d|'[^']*'|\b[xy]_target_\w+", "*", error))
    assert "ok" in outcomes and len(outcomes) >= 7, outcomes    # valid files and at least six different rejections
