"""Mirror of the reference's output stage, src/post_processor.rs:18-83 (`csv_output`).

nodes.csv:    header "x,y,ux,uy",        one "{x},{y},{ux},{uy}" row per node, node order
elements.csv: header "n0,n1,n2,stress",  one "{n0},{n1},{n2},{stress}" row per element
"\\n" line ends.  Floats use Rust's `Display` for f64: the shortest decimal that round-trips,
never scientific notation, integral values without ".0" ("3", "-0", "69000000000"),
"NaN" / "inf" / "-inf" for non-finite values.  `pyplot` (post_processor.rs:90-123) is
visualisation glue and out of scope.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

from .datatypes import Element, Node
from .error import MagnetiteError


def rust_f64_display(v: float) -> str:
    """Format like Rust's `format!("{}", v)` for f64."""
    v = float(v)
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"
    return np.format_float_positional(v, unique=True, trim="-")


def write_csv_fast(x, y, ux, uy, n0, n1, n2, stress, nodes_output: str, elements_output: str) -> None:
    """The same files through the library's host-side writer (mag_csv_output, csrc/csv.cpp): tens of
    millions of rows per minute instead of Python string formatting."""
    from . import _lib
    c = np.ascontiguousarray
    x, y, ux, uy, stress = (c(a, np.float64) for a in (x, y, ux, uy, stress))
    n0, n1, n2 = (c(a, np.uint32) for a in (n0, n1, n2))
    lib = _lib.load()
    rc = lib.mag_csv_output(nodes_output.encode(), elements_output.encode(), x.shape[0], _lib.ptr(x), _lib.ptr(y),
                            _lib.ptr(ux), _lib.ptr(uy), n0.shape[0], _lib.ptr(n0), _lib.ptr(n1), _lib.ptr(n2),
                            _lib.ptr(stress))
    if rc != 0:
        raise MagnetiteError.Solver((lib.mag_host_last_error() or b"").decode(), code=rc)


def _os_error(err: OSError) -> str:
    """Rust's Display of std::io::Error, which the reference embeds in its message (post_processor.rs:27-29)."""
    return f"{err.strerror} (os error {err.errno})"


def write_csv_arrays(x, y, ux, uy, n0, n1, n2, stress, nodes_output: str, elements_output: str) -> None:
    """Array-level writer in pure Python (buffered; the reference issues one unbuffered write per row)."""
    try:
        nf = open(nodes_output, "w", newline="")
    except OSError as err:                                   # post_processor.rs:24-31
        raise MagnetiteError.Solver(f"Failed to create nodes.csv: {_os_error(err)}")
    try:
        ef = open(elements_output, "w", newline="")
    except OSError as err:                                   # post_processor.rs:32-39
        nf.close()
        raise MagnetiteError.Solver(f"Failed to create elements.csv: {_os_error(err)}")
    f = rust_f64_display
    with nf, ef:
        nf.write("x,y,ux,uy\n")                              # post_processor.rs:42
        nf.writelines(f"{f(a)},{f(b)},{f(c)},{f(d)}\n" for a, b, c, d in zip(x, y, ux, uy))
        ef.write("n0,n1,n2,stress\n")                        # post_processor.rs:60
        ef.writelines(f"{int(a)},{int(b)},{int(c)},{f(s)}\n" for a, b, c, s in zip(n0, n1, n2, stress))


def csv_output(elements: Sequence[Element], nodes: Sequence[Node], nodes_output: str,
               elements_output: str, quiet: bool = False) -> None:
    """post_processor::csv_output (post_processor.rs:18-83) — same argument order."""
    for nd in nodes:
        if nd.ux is None or nd.uy is None:                   # the reference unwrap()s: :50-51
            raise MagnetiteError.PostProcessor("node without a displacement: run the solver first")
    for el in elements:
        if el.stress is None:                                # :70
            raise MagnetiteError.PostProcessor("element without a stress: run the solver first")
    # the library's host-side writer (csrc/csv.cpp): the same bytes as write_csv_arrays, formatted in C++
    write_csv_fast([n.vertex.x for n in nodes], [n.vertex.y for n in nodes],
                   [n.ux for n in nodes], [n.uy for n in nodes],
                   [e.nodes[0] for e in elements], [e.nodes[1] for e in elements],
                   [e.nodes[2] for e in elements], [e.stress for e in elements],
                   nodes_output, elements_output)
    if not quiet:
        print(f"info: wrote output to {nodes_output} and {elements_output}")   # :77-80
