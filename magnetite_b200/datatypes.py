"""Boundary types of the drop-in — mirrors of the reference's src/datatypes.rs:2-52.

`Node` / `Element` keep the reference's AoS shape with Option<f64> -> Optional[float];
`MeshSoA` is the flattened structure-of-arrays view that crosses the C ABI
(include/magnetite_b200.h: mag_mesh).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

KNOWN_UX, KNOWN_UY, KNOWN_FX, KNOWN_FY = 1, 2, 4, 8


@dataclass
class Vertex:                     # datatypes.rs:2-5
    x: float
    y: float


@dataclass
class Node:                       # datatypes.rs:8-14
    vertex: Vertex
    ux: Optional[float] = None
    uy: Optional[float] = None
    fx: Optional[float] = None
    fy: Optional[float] = None


@dataclass
class Element:                    # datatypes.rs:17-20
    nodes: List[int]
    stress: Optional[float] = None


@dataclass
class ModelMetadata:              # datatypes.rs:23-29
    youngs_modulus: float
    poisson_ratio: float
    part_thickness: float
    characteristic_length_min: float = 0.0
    characteristic_length_max: float = 0.0


@dataclass
class BoundaryRegion:             # datatypes.rs:32-37
    x_min: float
    x_max: float
    y_min: float
    y_max: float


@dataclass
class BoundaryTarget:             # datatypes.rs:40-45
    ux: Optional[float]
    uy: Optional[float]
    fx: Optional[float]
    fy: Optional[float]


@dataclass
class BoundaryRule:               # datatypes.rs:48-52
    name: str
    region: BoundaryRegion
    target: BoundaryTarget


@dataclass
class MeshSoA:
    """Vec<Node> + Vec<Element> flattened: what mag_mesh points at."""
    x: np.ndarray
    y: np.ndarray
    n0: np.ndarray
    n1: np.ndarray
    n2: np.ndarray
    ux: np.ndarray
    uy: np.ndarray
    fx: np.ndarray
    fy: np.ndarray
    known: np.ndarray
    meta: dict = field(default_factory=dict)

    @property
    def n_nodes(self) -> int:
        return int(self.x.shape[0])

    @property
    def n_elems(self) -> int:
        return int(self.n0.shape[0])

    def normalised(self) -> "MeshSoA":
        c = np.ascontiguousarray
        return MeshSoA(c(self.x, np.float64), c(self.y, np.float64), c(self.n0, np.uint32),
                       c(self.n1, np.uint32), c(self.n2, np.uint32), c(self.ux, np.float64),
                       c(self.uy, np.float64), c(self.fx, np.float64), c(self.fy, np.float64),
                       c(self.known, np.uint8), dict(self.meta))

    def copy(self) -> "MeshSoA":
        m = self.normalised()
        return MeshSoA(*(getattr(m, k).copy() for k in ("x", "y", "n0", "n1", "n2", "ux", "uy", "fx", "fy", "known")),
                       dict(self.meta))

    # ---- AoS <-> SoA (what the Rust shim does around the FFI call) ---------
    @staticmethod
    def from_aos(nodes: Sequence[Node], elements: Sequence[Element]) -> "MeshSoA":
        n, e = len(nodes), len(elements)
        x = np.fromiter((nd.vertex.x for nd in nodes), np.float64, n)
        y = np.fromiter((nd.vertex.y for nd in nodes), np.float64, n)
        known = np.zeros(n, np.uint8)
        payload = []
        for attr, bit in (("ux", KNOWN_UX), ("uy", KNOWN_UY), ("fx", KNOWN_FX), ("fy", KNOWN_FY)):
            opt = [getattr(nd, attr) for nd in nodes]                     # Option<f64>: None or a float
            some = np.fromiter((v is not None for v in opt), np.bool_, n)
            known |= some.astype(np.uint8) * np.uint8(bit)
            payload.append(np.fromiter((0.0 if v is None else v for v in opt), np.float64, n))
        ux, uy, fx, fy = payload
        conn = (np.array([el.nodes for el in elements], np.int64).reshape(e, 3) if e else np.empty((0, 3), np.int64))
        if e and (conn.min() < 0 or conn.max() >= 2 ** 32):
            raise ValueError("element node index does not fit usize/u32")
        conn = conn.astype(np.uint32)
        return MeshSoA(x, y, conn[:, 0].copy(), conn[:, 1].copy(), conn[:, 2].copy(),
                       ux, uy, fx, fy, known)

    def to_aos(self):
        nodes = []
        for i in range(self.n_nodes):
            k = int(self.known[i])
            nodes.append(Node(Vertex(float(self.x[i]), float(self.y[i])),
                              float(self.ux[i]) if k & KNOWN_UX else None,
                              float(self.uy[i]) if k & KNOWN_UY else None,
                              float(self.fx[i]) if k & KNOWN_FX else None,
                              float(self.fy[i]) if k & KNOWN_FY else None))
        elements = [Element([int(self.n0[i]), int(self.n1[i]), int(self.n2[i])])
                    for i in range(self.n_elems)]
        return nodes, elements
