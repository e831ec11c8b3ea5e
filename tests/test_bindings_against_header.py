"""CPU: both foreign-function bindings of the drop-in boundary against include/magnetite_b200.h, mechanically.

The ctypes binding (magnetite_b200/_lib.py — what the GPU tests and the benchmark call through): every Structure
field and every prototype's argument classes against the header.

The Rust side (rust/) cannot be compiled in this image (no rustc / cargo), so what can be checked without a
compiler is checked here:

* rust/magnetite-b200-sys/src/lib.rs against include/magnetite_b200.h, mechanically: every struct (field names,
  order, types), every prototype (name, argument types, return type), the constants, and the struct sizes it
  asserts against the ctypes binding the GPU tests drive;
* rust/reference-integration/solver_b200.rs only calls functions the sys crate declares, with the declared number
  of arguments, and keeps the reference's signatures (src/solver.rs:543-547, src/post_processor.rs:18-23);
* rust/reference-integration/apply.sh edits a Magnetite checkout the way INTEGRATION.md says (a miniature tree
  always; the real reference when /root/reference exists — never on the GPU box).
"""
import ctypes as C
import re
import shutil
import subprocess
from pathlib import Path

import pytest

from magnetite_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "magnetite_b200.h"
SYS_RS = ROOT / "rust" / "magnetite-b200-sys" / "src" / "lib.rs"
WRAP_RS = ROOT / "rust" / "reference-integration" / "solver_b200.rs"
APPLY = ROOT / "rust" / "reference-integration" / "apply.sh"

BASE = {"uint64_t": "u64", "uint32_t": "u32", "int32_t": "i32", "int64_t": "i64", "uint8_t": "u8", "double": "f64",
        "float": "f32", "int": "c_int", "size_t": "usize", "char": "c_char", "void": "c_void",
        "mag_mesh": "mag_mesh", "mag_material": "mag_material", "mag_options": "mag_options",
        "mag_result": "mag_result", "mag_stats": "mag_stats", "mag_ctx": "mag_ctx", "mag_system": "mag_system",
        "mag_devmesh": "mag_devmesh"}


def strip_c_comments(text):
    return re.sub(r"/\*.*?\*/", " ", text, flags=re.S)


def rust_type(base, const, stars, array=None):
    t = BASE[base]
    for level in range(stars):
        t = ("*const " if (const and level == 0) else "*mut ") + t
    if array:
        t = f"[{t}; {array}]"
    return t


def parse_c_decl(decl):
    """'const double *x, *y' -> [('x', '*const f64'), ('y', '*const f64')]"""
    decl = decl.strip()
    const = decl.startswith("const ")
    if const:
        decl = decl[6:].strip()
    base, rest = decl.split(None, 1) if " " in decl else (decl, "")
    out = []
    for d in rest.split(","):
        d = d.strip()
        stars = d.count("*")
        name = d.replace("*", "").strip()
        arr = None
        m = re.match(r"(\w+)\[(\d+)\]$", name)
        if m:
            name, arr = m.group(1), m.group(2)
        out.append((name, rust_type(base, const, stars, arr)))
    return out


def c_structs():
    text = strip_c_comments(HEADER.read_text())
    structs = {}
    for body, name in re.findall(r"typedef struct \{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in body.split(";"):
            if decl.strip():
                fields += parse_c_decl(" ".join(decl.split()))
        structs[name] = fields
    return structs


def c_functions():
    text = strip_c_comments(HEADER.read_text())
    text = re.sub(r"typedef struct \{.*?\}\s*\w+\s*;", " ", text, flags=re.S)
    text = " ".join(text.split())
    funcs = {}
    for ret, name, params in re.findall(r"(?:^|;|\{|\})\s*((?:const\s+)?\w+\s*\**)\s*\b(mag_\w+)\s*\(([^)]*)\)", text):
        ret = ret.strip()
        const = ret.startswith("const ")
        rbase = ret.replace("const ", "").replace("*", "").strip()
        rstars = ret.count("*")
        rtype = None if (rbase == "void" and rstars == 0) else rust_type(rbase, const, rstars)
        args = []
        if params.strip() != "void":
            for p in params.split(","):
                (pname, ptype), = parse_c_decl(" ".join(p.split()))
                args.append((pname, ptype))
        funcs[name] = (args, rtype)
    return funcs


def rust_structs():
    text = re.sub(r"//[^\n]*", "", SYS_RS.read_text())
    structs = {}
    for name, body in re.findall(r"#\[repr\(C\)\](?:\s*#\[derive\([^)]*\)\])?\s*pub struct (\w+)\s*\{(.*?)\n\}", text, flags=re.S):
        fields = re.findall(r"pub (\w+):\s*([^,\n]+),", body)
        structs[name] = [(n, " ".join(t.split())) for n, t in fields]
    return structs


def rust_functions():
    text = re.sub(r"//[^\n]*", "", SYS_RS.read_text())
    block = re.search(r'extern "C" \{(.*?)\n\}', text, flags=re.S).group(1)
    funcs = {}
    for name, params, ret in re.findall(r"pub fn (\w+)\(([^)]*)\)\s*(?:->\s*([^;]+))?;", block):
        args = []
        for p in params.split(","):
            if p.strip():
                n, t = p.split(":", 1)
                args.append((n.strip(), " ".join(t.split())))
        funcs[name] = (args, ret.strip() or None)
    return funcs


def test_rust_structs_mirror_the_header():
    cs, rs = c_structs(), rust_structs()
    assert set(cs) == {"mag_mesh", "mag_material", "mag_options", "mag_result", "mag_stats"}
    for name, fields in cs.items():
        assert rs[name] == fields, f"{name}: rust/magnetite-b200-sys drifted from the header"
    for opaque in ("mag_ctx", "mag_system", "mag_devmesh"):
        assert rs[opaque] == []                     # `_private: [u8; 0]` is not pub: no visible field


def test_rust_prototypes_mirror_the_header():
    cf, rf = c_functions(), rust_functions()
    assert sorted(cf) == sorted(_lib.declared_symbols()), "header parser of this test lost a prototype"
    assert sorted(rf) == sorted(cf), "the sys crate does not declare exactly the header's functions"
    keyword = {"in": "input"}                        # `in` is a Rust keyword
    for name, (args, ret) in cf.items():
        rargs, rret = rf[name]
        assert rret == ret, f"{name}: return type"
        assert [t for _, t in rargs] == [t for _, t in args], f"{name}: argument types"
        assert [n for n, _ in rargs] == [keyword.get(n, n) for n, _ in args], f"{name}: argument names"


def test_rust_constants_and_sizes():
    text = SYS_RS.read_text()
    consts = dict(re.findall(r"pub const (\w+): \w+ = (-?[\w.]+(?:-\d+)?);", text))
    header = strip_c_comments(HEADER.read_text())
    assert int(consts["MAG_ABI_VERSION"]) == _lib.ABI_VERSION == int(re.search(r"#define MAG_ABI_VERSION (\d+)", header).group(1))
    enum = re.search(r"enum \{(.*?)\};", header, flags=re.S).group(1)
    for name, value in re.findall(r"(MAG_\w+) = (-?\d+)", enum):
        assert int(consts[name]) == int(value), name
    for name, value in re.findall(r"#define (MAG_KNOWN_\w+) (\d+)u", header):
        assert int(consts[name]) == int(value), name
    assert int(consts["MAG_DOF"]) == 2 and int(consts["MAG_MAX_CG_ITER"].replace("_", "")) == 10_000_000
    assert float(consts["MAG_TARGET_CG_COST"]) == 1e-4
    sizes = {"SIZEOF_MAG_MESH": _lib.MagMesh, "SIZEOF_MAG_MATERIAL": _lib.MagMaterial, "SIZEOF_MAG_OPTIONS": _lib.MagOptions,
             "SIZEOF_MAG_RESULT": _lib.MagResult, "SIZEOF_MAG_STATS": _lib.MagStats}
    for name, struct in sizes.items():
        assert int(consts[name]) == C.sizeof(struct), name


def split_top_level(args):
    out, depth, cur = [], 0, ""
    for ch in args:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def test_wrapper_calls_only_declared_functions_with_their_arity():
    rf = rust_functions()
    text = re.sub(r"//[^\n]*", "", WRAP_RS.read_text())
    calls = list(re.finditer(r"sys::(mag_\w+)\s*\(", text))
    assert {m.group(1) for m in calls} >= {"mag_solve", "mag_ctx_create", "mag_ctx_destroy", "mag_options_default",
                                           "mag_csv_output", "mag_reorder_rcm", "mag_element_area", "mag_last_error"}
    for m in calls:
        name = m.group(1)
        if name in ("mag_mesh", "mag_material", "mag_result", "mag_stats", "mag_options", "mag_ctx"):
            continue
        assert name in rf, f"solver_b200.rs calls undeclared {name}"
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        n_args = len(split_top_level(text[m.end():i - 1]))
        assert n_args == len(rf[name][0]), f"{name}: {n_args} arguments passed, {len(rf[name][0])} declared"
    # the reference's signatures (src/solver.rs:543-547, src/post_processor.rs:18-23)
    flat = " ".join(text.split())
    assert "pub fn run( nodes: &mut Vec<Node>, elements: &mut Vec<Element>, model_metadata: &ModelMetadata, ) -> Result<(), MagnetiteError>" in flat
    assert "pub fn csv_output( elements: &Vec<Element>, nodes: &Vec<Node>, nodes_output: &str, elements_output: &str, ) -> Result<(), MagnetiteError>" in flat
    assert "use crate::datatypes::{Element, ModelMetadata, Node};" in text and "use crate::error::MagnetiteError;" in text
    assert "opt.compat = 1" in text                   # reference solver semantics through the drop-in
    for line in ("info: building element stiffness matrices...", "info: building total stiffness matrix...", "info: solving...",
                 "info: finished conjugate gradient approximation in {} iterations", "info: solved system in {:.3} seconds",
                 "info: solve complete"):
        assert line in text                           # solver.rs:551,570,437,101-104,441,484
    for opener, closer in ("()", "{}", "[]"):
        assert text.count(opener) == text.count(closer), f"unbalanced {opener}{closer} in solver_b200.rs"


MINI_MAIN = """use error::MagnetiteError;
mod datatypes;
mod error;
mod mesher;
mod post_processor;
mod solver;

fn entry() -> Result<(), MagnetiteError> {
    let (mut nodes, mut elements, model_metadata) = mesher::run(vec![], "input.json")?;
    solver::run(&mut nodes, &mut elements, &model_metadata)?;
    post_processor::csv_output(&elements, &nodes, "nodes.csv", "elements.csv")?;
    Ok(())
}
"""


def check_applied(tree):
    main = (tree / "src" / "main.rs").read_text()
    assert "mod solver;\nmod solver_b200;\n" in main
    assert "solver_b200::run(&mut nodes, &mut elements, &model_metadata)?;" in main
    assert "solver_b200::csv_output(&elements, &nodes," in main
    assert " solver::run(" not in main and "post_processor::csv_output(" not in main
    toml = (tree / "Cargo.toml").read_text()
    assert f'[dependencies]\nmagnetite-b200-sys = {{ path = "{ROOT / "rust" / "magnetite-b200-sys"}" }}\n' in toml
    assert (tree / "src" / "solver_b200.rs").read_bytes() == WRAP_RS.read_bytes()


def test_apply_script_on_a_miniature_checkout(tmp_path):
    tree = tmp_path / "Magnetite"
    (tree / "src").mkdir(parents=True)
    (tree / "Cargo.toml").write_text('[package]\nname = "magnetite"\n\n[dependencies]\nclap = "4"\n')
    (tree / "src" / "main.rs").write_text(MINI_MAIN)
    for f in ("datatypes.rs", "error.rs"):
        (tree / "src" / f).write_text("")
    r = subprocess.run(["bash", str(APPLY), str(tree)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    check_applied(tree)
    again = subprocess.run(["bash", str(APPLY), str(tree)], capture_output=True, text=True)
    assert again.returncode != 0 and "already applied" in again.stderr
    stranger = subprocess.run(["bash", str(APPLY), str(tmp_path)], capture_output=True, text=True)
    assert stranger.returncode != 0 and "not a Magnetite checkout" in stranger.stderr


def test_apply_script_on_the_reference_checkout(tmp_path):
    ref = Path("/root/reference")
    if not (ref / "src" / "main.rs").exists():
        pytest.skip("/root/reference is not on this box")
    tree = tmp_path / "Magnetite"
    shutil.copytree(ref, tree)
    r = subprocess.run(["bash", str(APPLY), str(tree)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    check_applied(tree)
    # nothing else changed
    for p in ref.rglob("*"):
        if p.is_file() and p.name not in ("main.rs", "Cargo.toml"):
            assert (tree / p.relative_to(ref)).read_bytes() == p.read_bytes()


# ---- the ctypes binding --------------------------------------------------------------------------------------
CTYPES_SCALAR = {"u64": C.c_uint64, "u32": C.c_uint32, "i32": C.c_int32, "i64": C.c_int64, "u8": C.c_uint8,
                 "f64": C.c_double, "f32": C.c_float, "c_int": C.c_int, "usize": C.c_size_t}


def ctypes_matches(ctype, rust):
    """Is the ctypes type `ctype` ABI-compatible with the header type written in Rust notation?"""
    if rust.startswith("*"):
        return ctype in (C.c_void_p, C.c_char_p) or (isinstance(ctype, type) and issubclass(ctype, C._Pointer))
    m = re.match(r"\[(\w+); (\d+)\]$", rust)
    if m:
        return issubclass(ctype, C.Array) and ctype._length_ == int(m.group(2)) and ctype._type_ is CTYPES_SCALAR[m.group(1)]
    want = CTYPES_SCALAR[rust]
    return ctype is want or (C.sizeof(ctype) == C.sizeof(want) and {ctype, want} <= {C.c_int, C.c_int32})


def test_ctypes_structures_mirror_the_header():
    pairs = {"mag_mesh": _lib.MagMesh, "mag_material": _lib.MagMaterial, "mag_options": _lib.MagOptions,
             "mag_result": _lib.MagResult, "mag_stats": _lib.MagStats}
    for name, fields in c_structs().items():
        got = pairs[name]._fields_
        assert [f[0] for f in got] == [n for n, _ in fields], f"{name}: field names / order"
        for (fname, ctype), (_, rust) in zip(got, fields):
            assert ctypes_matches(ctype, rust), f"{name}.{fname}: {ctype} vs {rust}"


def test_ctypes_prototypes_mirror_the_header():
    for name, (args, ret) in c_functions().items():
        res, argtypes = _lib._SIGNATURES[name]
        assert len(argtypes) == len(args), f"{name}: {len(argtypes)} ctypes arguments, {len(args)} in the header"
        for i, (ctype, (aname, rust)) in enumerate(zip(argtypes, args)):
            assert ctypes_matches(ctype, rust), f"{name} argument {i} ({aname}): {ctype} vs {rust}"
        if ret is None:
            assert res is None, name
        else:
            assert ctypes_matches(res, ret), f"{name}: return type {res} vs {ret}"
