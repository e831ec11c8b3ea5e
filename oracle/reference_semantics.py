"""Independent numpy restatement of reference src/solver.rs for SMALL meshes (dense, pure
Python loops) — a second opinion on the C oracle, written separately from it, plus a direct
sparse solve as ground truth for displacements.  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED.
"""
from __future__ import annotations

import numpy as np


def area(x, y, el):                                   # solver.rs:187-193
    a, b, c = el
    return 0.5 * (x[a] * (y[b] - y[c]) + x[b] * (y[c] - y[a]) + x[c] * (y[a] - y[b]))


def B_matrix(x, y, el, A):                            # solver.rs:204-230
    a, b, c = el
    b1, b2, b3 = y[b] - y[c], y[c] - y[a], y[a] - y[b]
    g1, g2, g3 = x[c] - x[b], x[a] - x[c], x[b] - x[a]
    B = np.array([[b1, 0, b2, 0, b3, 0], [0, g1, 0, g2, 0, g3], [g1, b1, g2, b2, g3, b3]], float)
    return B / (2.0 * A)


def D_matrix(nu, E):                                  # solver.rs:240-250
    D = np.array([[1.0, nu, 0.0], [nu, 1.0, 0.0], [0.0, 0.0, (1.0 - nu) / 2.0]])
    return D * (E / (1.0 - nu * nu))


def _mm(a, b):
    """Matrix product with the k-ascending, round-every-operation order of nalgebra's small gemm."""
    n, kk = a.shape
    m = b.shape[1]
    out = np.zeros((n, m))
    for i in range(n):
        for j in range(m):
            s = a[i, 0] * b[0, j]
            for k in range(1, kk):
                s = a[i, k] * b[k, j] + s
            out[i, j] = s
    return out


def element_stiffness(x, y, el, nu, E, t):            # solver.rs:263-278
    A = area(x, y, el)
    B = B_matrix(x, y, el, A)
    return _mm(_mm(B.T.copy(), D_matrix(nu, E)), B) * A * t


def assemble(x, y, conn, nu, E, t):                   # solver.rs:290-331
    n = 2 * len(x)
    K = np.zeros((n, n))
    for el in conn:
        ke = element_stiffness(x, y, el, nu, E, t)
        for lr, nr in enumerate(el):
            for lc, nc in enumerate(el):
                K[2 * nr:2 * nr + 2, 2 * nc:2 * nc + 2] += ke[2 * lr:2 * lr + 2, 2 * lc:2 * lc + 2]
    return K


def solve_direct(x, y, conn, known, ux, uy, fx, fy, nu, E, t):
    """Dense assembly + partition + LU solve (ground truth, no CG): returns U (2N), F (2N), K_ff, rhs."""
    n = 2 * len(x)
    K = assemble(x, y, conn, nu, E, t)
    U = np.zeros(n); F = np.zeros(n)
    uk = np.zeros(n, bool); fk = np.zeros(n, bool)
    for i in range(len(x)):
        uk[2 * i], uk[2 * i + 1] = bool(known[i] & 1), bool(known[i] & 2)
        fk[2 * i], fk[2 * i + 1] = bool(known[i] & 4), bool(known[i] & 8)
        U[2 * i], U[2 * i + 1] = ux[i], uy[i]
        F[2 * i], F[2 * i + 1] = fx[i], fy[i]
    rows = np.flatnonzero(fk); free = np.flatnonzero(~uk); fixed = np.flatnonzero(uk)
    Kff = K[np.ix_(rows, free)]
    rhs = F[rows] - K[np.ix_(rows, fixed)] @ U[fixed]
    sol = np.linalg.solve(Kff, rhs)
    U[free] = sol
    Fall = K @ U
    F[~fk] = Fall[~fk]
    return U, F, Kff, rhs


def stress(x, y, conn, U, nu, E):                     # solver.rs:496-535
    out = np.zeros(len(conn)); sig = np.zeros((len(conn), 3))
    for e, el in enumerate(conn):
        ue = np.array([U[2 * el[0]], U[2 * el[0] + 1], U[2 * el[1]], U[2 * el[1] + 1], U[2 * el[2]], U[2 * el[2] + 1]])
        s = _mm(_mm(D_matrix(nu, E), B_matrix(x, y, el, area(x, y, el))), ue.reshape(6, 1)).ravel()
        sign = -1.0 if s[0] + s[1] < 1.0 else 1.0
        out[e] = np.sqrt(s[0] * s[0] + s[1] * s[1]) * sign
        sig[e] = s
    return out, sig
