"""CPU restatement (numpy / scipy) of the preconditioned CG the GPU path runs by default — TEST INFRASTRUCTURE.

Unlike the rest of oracle/, this file does not restate the reference (the reference has no preconditioner: argmin's
plain ConjugateGradient, solver.rs:141-157).  It restates THIS repository's extension, from its definition, so that
the CUDA implementation has a second, independent statement of the same mathematics to be compared with:

  * Jacobi-PCG (magnetite_b200/csrc/pcg.cuh): z = D^-1 r with D^-1 = 1/d (1 where d == 0), x0 = 0, stop when
    r.r <= rel_tol^2 b.b (the recursive residual), checked after the x / r update of an iteration;
  * the two-level preconditioner (magnetite_b200/csrc/coarse.cuh, solve.cuh: setup_coarse):
    M^-1 = D^-1 + P Ac^-1 P^T, Ac = P^T K_ff P, P = the three rigid-body modes (x, y, rotation about the box centre,
    the rotation coefficient scaled by the box size and ROUNDED TO fp32 as the device stores it) of every box of a
    regular nbx x nby grid over the bounding box; the grid is chosen as solve.cuh does it (target = n_free / 4096
    aggregates, at most 2048, split by the aspect ratio); empty modes get a unit diagonal.

The matrix, right-hand side and DOF map come from the oracle proper (oracle.py: assemble_sparse + partition), so the
only things restated here are the preconditioner and the recurrence.  What it is used for: iteration counts.  The
recurrences are the same to rounding, and the counts agree EXACTLY with the B200's wherever both have run (400 x 200
plate: 2801 Jacobi / 450 two-level; 1000 x 500: 424 two-level; tests/test_oracle_two_level.py reads the B200's
counts from the committed bench line under profiles/).  Only tests/ and bench.py's documentation refer to it; the
product never imports it."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import oracle as O

COARSE_MAX_AGG = 2048          # coarse.cuh: kCoarseMaxAgg
AUTO_MIN_ROWS = 20000          # solve.cuh: kAutoTwoLevelMinRows (precond 3)


@dataclass
class Reduced:
    A: sp.csr_matrix
    rhs: np.ndarray
    node: np.ndarray           # per reduced column: mesh node
    axis: np.ndarray           # per reduced column: 0 = x, 1 = y
    x: np.ndarray
    y: np.ndarray
    box: tuple                 # (xmin, xmax, ymin, ymax) of ALL mesh nodes


def reduced_system(mesh, meta) -> Reduced:
    om = O.Mesh(mesh)
    (rp, col, val), rhs, fmap = O.partition(om, O.assemble_sparse(om, O.element_stiffness(om, meta)), dense=False)
    n = len(rhs)
    A = sp.csr_matrix((val, col, rp), shape=(n, n))
    dofs = np.nonzero(fmap >= 0)[0]
    dofs = dofs[np.argsort(fmap[dofs])]
    node, axis = dofs // 2, dofs % 2
    x, y = np.asarray(mesh.x, float), np.asarray(mesh.y, float)
    return Reduced(A, np.asarray(rhs, float), node, axis, x[node], y[node], (x.min(), x.max(), y.min(), y.max()))


def coarse_grid(n_free: int, box, coarse_aggregates: int = 0):
    """(nbx, nby, hx, hy) as setup_coarse picks them (solve.cuh: 'bounding box' stage)."""
    xmin, xmax, ymin, ymax = box
    wd, hd = max(xmax - xmin, 1e-300), max(ymax - ymin, 1e-300)
    target = coarse_aggregates if coarse_aggregates > 0 else min(COARSE_MAX_AGG, max(4, n_free // 4096))
    target = min(target, COARSE_MAX_AGG)
    nbx = max(1, int(math.floor(math.sqrt(target * wd / hd) + 0.5)))        # std::lround
    nby = max(1, int(math.floor(target / nbx + 0.5)))
    while nbx * nby > COARSE_MAX_AGG:
        if nbx >= nby:
            nbx -= 1
        else:
            nby -= 1
    return nbx, nby, wd / nbx * (1.0 + 1e-12), hd / nby * (1.0 + 1e-12)


def prolongation(S: Reduced, coarse_aggregates: int = 0) -> sp.csr_matrix:
    """P (n_free x 3*nbx*nby): coarse.cuh coarse_colinfo_kernel.  The numbering of the boxes does not matter here."""
    n = S.A.shape[0]
    nbx, nby, hx, hy = coarse_grid(n, S.box, coarse_aggregates)
    xmin, _, ymin, _ = S.box
    bx = np.clip(np.floor((S.x - xmin) / hx), 0, nbx - 1).astype(np.int64)
    by = np.clip(np.floor((S.y - ymin) / hy), 0, nby - 1).astype(np.int64)
    agg = by * nbx + bx
    xc, yc = xmin + (bx + 0.5) * hx, ymin + (by + 0.5) * hy
    rot = np.where(S.axis == 1, (S.x - xc) / hx, -(S.y - yc) / hy).astype(np.float32).astype(np.float64)
    rows = np.concatenate([np.arange(n), np.arange(n)])
    cols = np.concatenate([3 * agg + S.axis, 3 * agg + 2])
    return sp.csr_matrix((np.concatenate([np.ones(n), rot]), (rows, cols)), shape=(n, 3 * nbx * nby))


def preconditioner(S: Reduced, kind: int, coarse_aggregates: int = 0):
    """kind 0: identity, 1: Jacobi, 2: two-level.  Returns (apply, info); raises ValueError when Ac is not SPD (the
    device falls back to Jacobi then)."""
    d = S.A.diagonal()
    dinv = np.where(d != 0.0, 1.0 / np.where(d != 0.0, d, 1.0), 1.0)
    if kind == 0:
        return (lambda r: r.copy()), {"precond": 0}
    if kind == 1:
        return (lambda r: r * dinv), {"precond": 1}
    P = prolongation(S, coarse_aggregates)
    Ac = (P.T @ S.A @ P).tocsc()
    dc = Ac.diagonal()
    Ac = (Ac + sp.diags(np.where(dc == 0.0, 1.0, 0.0))).tocsc()
    lu = spla.splu(Ac)
    if Ac.shape[0] <= 1024:                                   # small enough to ask: does Ac have a Cholesky factor?
        try:
            np.linalg.cholesky(Ac.toarray())
        except np.linalg.LinAlgError as e:
            raise ValueError("Ac is not positive definite") from e
    PT = P.T.tocsr()
    return (lambda r: r * dinv + P @ lu.solve(PT @ r)), {"precond": 2, "n_coarse": P.shape[1]}


def pcg(S: Reduced, kind: int, rel_tol: float = 1e-9, max_iter: int = 10_000_000, coarse_aggregates: int = 0):
    """(x, iterations).  The recurrence of pcg.cuh: alpha = r.z / p.Ap, x += alpha p, r -= alpha Ap, stop test,
    z = M^-1 r, beta = r.z_new / r.z_old, p = z + beta p."""
    M, _ = preconditioner(S, kind, coarse_aggregates)
    A, b = S.A, S.rhs
    x = np.zeros_like(b)
    r = b.copy()
    bb = float(b @ b)
    if bb == 0.0:
        return x, 0
    z = M(r)
    p = z.copy()
    rz = float(r @ z)
    for it in range(1, max_iter + 1):
        q = A @ p
        alpha = rz / float(p @ q)
        x += alpha * p
        r -= alpha * q
        if float(r @ r) <= rel_tol * rel_tol * bb:
            return x, it
        z = M(r)
        rz_new = float(r @ z)
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, max_iter


def default_kind(n_free: int) -> int:
    """mag_options.precond = 3 (auto): two-level from 20 000 unknowns."""
    return 2 if n_free >= AUTO_MIN_ROWS else 1
