"""magnetite_b200 — B200-native (sm_100a, fp64) replacement for the numerical core of
kyle-tennison/Magnetite: CST element stiffness, deterministic sort-and-reduce assembly,
in-kernel Dirichlet elimination, (Jacobi-)CG and stress recovery, behind the reference's own
`solver::run` / `post_processor::csv_output` interface.  See DESIGN.md and INTEGRATION.md.
"""
from . import datatypes, error, geometry, mesher, meshgen, post_processor, reorder, solver  # noqa: F401
from .datatypes import Element, MeshSoA, ModelMetadata, Node, Vertex  # noqa: F401
from .error import MagnetiteError  # noqa: F401

__all__ = ["datatypes", "error", "geometry", "mesher", "meshgen", "post_processor", "reorder", "solver",
           "Element", "MeshSoA", "ModelMetadata", "Node", "Vertex", "MagnetiteError"]
