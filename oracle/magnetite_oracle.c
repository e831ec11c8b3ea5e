/*
 * magnetite_oracle.c — CPU restatement of Magnetite's solver.rs / the crates it
 * calls.  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see magnetite_oracle.h).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
 * Every function cites the reference lines it restates (paths relative to the
 * reference checkout).  No code is copied: the reference is Rust on nalgebra /
 * argmin; this is plain C with explicit loops in the operation order those
 * crates use at the reference's call sites.
 */
#include "magnetite_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static __thread char g_err[256];
const char *orc_last_error(void) { return g_err; }
static int fail(int code, const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}
static size_t nz1(size_t v) { return v ? v : 1; }
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ------------------------------------------------------------------------ */
/* element level                                                             */
/* ------------------------------------------------------------------------ */

/* solver.rs:187-193 — signed area, evaluated left to right. */
double orc_element_area(const orc_mesh *m, uint64_t e) {
    const uint32_t a = m->n0[e], b = m->n1[e], c = m->n2[e];
    const double x0 = m->x[a], y0 = m->y[a];
    const double x1 = m->x[b], y1 = m->y[b];
    const double x2 = m->x[c], y2 = m->y[c];
    return 0.5 * (x0 * (y1 - y2) + x1 * (y2 - y0) + x2 * (y0 - y1));
}

/* solver.rs:204-230 — B (3x6, row-major here), every entry divided by 2A. */
void orc_strain_displacement(const orc_mesh *m, uint64_t e, double area, double B[18]) {
    const uint32_t a = m->n0[e], b = m->n1[e], c = m->n2[e];
    const double x0 = m->x[a], y0 = m->y[a];
    const double x1 = m->x[b], y1 = m->y[b];
    const double x2 = m->x[c], y2 = m->y[c];
    const double b1 = y1 - y2, b2 = y2 - y0, b3 = y0 - y1;   /* :213-215 */
    const double g1 = x2 - x1, g2 = x0 - x2, g3 = x1 - x0;   /* :217-219 */
    const double raw[18] = {                                  /* :221-225 */
        b1, 0., b2, 0., b3, 0.,
        0., g1, 0., g2, 0., g3,
        g1, b1, g2, b2, g3, b3};
    const double den = 2.0 * area;                            /* :227 */
    for (int i = 0; i < 18; ++i) B[i] = raw[i] / den;
}

/* solver.rs:240-250 — plane-stress D, every entry times E/(1-nu^2). */
void orc_stress_strain(double nu, double E, double D[9]) {
    const double base[9] = {1.0, nu, 0.0, nu, 1.0, 0.0, 0.0, 0.0, (1.0 - nu) / 2.0};
    const double c = E / (1.0 - nu * nu);                     /* powi(nu,2) = nu*nu */
    for (int i = 0; i < 9; ++i) D[i] = base[i] * c;
}

/* nalgebra static Mul (gemm -> per-column gemv -> axcpy): out[i][j] starts as
 * a[i][0]*b[0][j] and then adds a[i][k]*b[k][j] for k = 1.., each product and
 * each sum rounded separately.  Row-major operands here. */
static void matmul_seq(const double *a, const double *b, double *out, int n, int kk, int mcols) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < mcols; ++j) {
            double s = a[i * kk + 0] * b[0 * mcols + j];
            for (int k = 1; k < kk; ++k) s = a[i * kk + k] * b[k * mcols + j] + s;
            out[i * mcols + j] = s;
        }
}

/* solver.rs:263-278 — K_e = ((B^T * D) * B) * A * t, local DOF order
 * [0x,0y,1x,1y,2x,2y]; Ke row-major 6x6. */
void orc_element_stiffness_one(const orc_mesh *m, uint64_t e, const orc_material *mat,
                               double Ke[36]) {
    double B[18], D[9], Bt[18], BtD[18];
    const double area = orc_element_area(m, e);               /* :270 */
    orc_stress_strain(mat->poisson_ratio, mat->youngs_modulus, D); /* :271 */
    orc_strain_displacement(m, e, area, B);                   /* :272 */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 6; ++j) Bt[j * 3 + i] = B[i * 6 + j];
    matmul_seq(Bt, D, BtD, 6, 3, 3);                          /* :274 */
    matmul_seq(BtD, B, Ke, 6, 3, 6);                          /* :275 */
    for (int i = 0; i < 36; ++i) Ke[i] = Ke[i] * area;        /* :276 */
    for (int i = 0; i < 36; ++i) Ke[i] = Ke[i] * mat->part_thickness; /* :277 */
}

/* solver.rs:553-563 */
void orc_element_stiffness(const orc_mesh *m, const orc_material *mat, double *Ke) {
    for (uint64_t e = 0; e < m->n_elems; ++e) orc_element_stiffness_one(m, e, mat, Ke + 36 * e);
}

/* ------------------------------------------------------------------------ */
/* assembly                                                                  */
/* ------------------------------------------------------------------------ */

/* solver.rs:290-331 — dense (2N)^2, column-major like DMatrix; per element in
 * order, 3x3 node pairs, four += per pair.  Global DOF = 2*node + axis. */
void orc_assemble_dense(const orc_mesh *m, const double *Ke, double *K) {
    const size_t n = (size_t)ORC_DOF * m->n_nodes;
    memset(K, 0, n * n * sizeof(double));                     /* :295-296 */
    for (uint64_t e = 0; e < m->n_elems; ++e) {
        const uint32_t nd[3] = {m->n0[e], m->n1[e], m->n2[e]};
        const double *k = Ke + 36 * e;
        for (int lr = 0; lr < 3; ++lr)
            for (int lc = 0; lc < 3; ++lc) {                  /* :304-305 */
                const size_t gr = (size_t)nd[lr] * 2, gc = (size_t)nd[lc] * 2;
                const int r = lr * 2, c = lc * 2;
                K[gr + gc * n] += k[r * 6 + c];               /* :312 */
                K[gr + (gc + 1) * n] += k[r * 6 + c + 1];     /* :315 */
                K[(gr + 1) + gc * n] += k[(r + 1) * 6 + c];   /* :318 */
                K[(gr + 1) + (gc + 1) * n] += k[(r + 1) * 6 + c + 1]; /* :321 */
            }
    }
}

static int cmp_u32(const void *a, const void *b) {
    const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

void orc_csr_free(orc_csr *A) {
    if (!A) return;
    free(A->rowptr); free(A->col); free(A->val);
    memset(A, 0, sizeof *A);
}

/* Same accumulation as orc_assemble_dense (every entry receives its element
 * contributions in ascending element index, starting from +0.0), stored as the
 * structural CSR of the full K: every (row dof, col dof) touched by at least
 * one element is stored, including entries that sum (or are) exactly 0.0. */
int orc_assemble_sparse(const orc_mesh *m, const double *Ke, orc_csr *K) {
    const uint64_t N = m->n_nodes, E = m->n_elems;
    memset(K, 0, sizeof *K);
    /* node -> neighbour nodes (with duplicates), then sort+unique per node */
    int64_t *cnt = calloc(N + 1, sizeof(int64_t));
    if (!cnt) return fail(ORC_ERR_OOM, "oom");
    for (uint64_t e = 0; e < E; ++e) {
        const uint32_t nd[3] = {m->n0[e], m->n1[e], m->n2[e]};
        for (int i = 0; i < 3; ++i) {
            if (nd[i] >= N) { free(cnt); return fail(ORC_ERR_BAD_INDEX, "element node index out of range"); }
            cnt[nd[i] + 1] += 3;
        }
    }
    for (uint64_t i = 0; i < N; ++i) cnt[i + 1] += cnt[i];
    uint32_t *adj = malloc((size_t)(cnt[N] ? cnt[N] : 1) * sizeof(uint32_t));
    int64_t *fill = malloc((N + 1) * sizeof(int64_t));
    int64_t *nptr = malloc((N + 1) * sizeof(int64_t));
    if (!adj || !fill || !nptr) { free(cnt); free(adj); free(fill); free(nptr); return fail(ORC_ERR_OOM, "oom"); }
    memcpy(fill, cnt, (N + 1) * sizeof(int64_t));
    for (uint64_t e = 0; e < E; ++e) {
        const uint32_t nd[3] = {m->n0[e], m->n1[e], m->n2[e]};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) adj[fill[nd[i]]++] = nd[j];
    }
    nptr[0] = 0;
    for (uint64_t i = 0; i < N; ++i) {
        uint32_t *a = adj + cnt[i];
        const int64_t len = cnt[i + 1] - cnt[i];
        qsort(a, (size_t)len, sizeof(uint32_t), cmp_u32);
        int64_t u = 0;
        for (int64_t k = 0; k < len; ++k)
            if (k == 0 || a[k] != a[k - 1]) a[u++] = a[k];
        fill[i] = u;                       /* unique neighbour count */
        nptr[i + 1] = nptr[i] + u;
    }
    const uint64_t n = 2 * N, nnz = 4 * (uint64_t)nptr[N];
    K->n_rows = K->n_cols = n; K->nnz = nnz;
    K->rowptr = malloc((n + 1) * sizeof(int64_t));
    K->col = malloc((size_t)(nnz ? nnz : 1) * sizeof(int32_t));
    K->val = calloc((size_t)(nnz ? nnz : 1), sizeof(double));
    if (!K->rowptr || !K->col || !K->val) {
        free(cnt); free(adj); free(fill); free(nptr); orc_csr_free(K);
        return fail(ORC_ERR_OOM, "oom");
    }
    for (uint64_t i = 0; i < N; ++i) {
        const int64_t u = fill[i];
        const int64_t base = 4 * nptr[i];
        K->rowptr[2 * i] = base;
        K->rowptr[2 * i + 1] = base + 2 * u;
        const uint32_t *a = adj + cnt[i];
        for (int64_t k = 0; k < u; ++k)
            for (int ax = 0; ax < 2; ++ax) {
                K->col[base + 2 * k + ax] = (int32_t)(2 * a[k] + ax);
                K->col[base + 2 * u + 2 * k + ax] = (int32_t)(2 * a[k] + ax);
            }
    }
    K->rowptr[n] = (int64_t)nnz;
    /* accumulate, element-ascending (solver.rs:299-323) */
    for (uint64_t e = 0; e < E; ++e) {
        const uint32_t nd[3] = {m->n0[e], m->n1[e], m->n2[e]};
        const double *k = Ke + 36 * e;
        for (int lr = 0; lr < 3; ++lr) {
            const uint32_t rn = nd[lr];
            const uint32_t *a = adj + cnt[rn];
            const int64_t u = fill[rn];
            for (int lc = 0; lc < 3; ++lc) {
                const uint32_t *hit = bsearch(&nd[lc], a, (size_t)u, sizeof(uint32_t), cmp_u32);
                const int64_t pos = hit - a;
                const int64_t r0 = K->rowptr[2 * rn] + 2 * pos;
                const int64_t r1 = K->rowptr[2 * rn + 1] + 2 * pos;
                const int r = lr * 2, c = lc * 2;
                K->val[r0] += k[r * 6 + c];
                K->val[r0 + 1] += k[r * 6 + c + 1];
                K->val[r1] += k[(r + 1) * 6 + c];
                K->val[r1 + 1] += k[(r + 1) * 6 + c + 1];
            }
        }
    }
    free(cnt); free(adj); free(fill); free(nptr);
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* partition + rhs                                                           */
/* ------------------------------------------------------------------------ */

/* solver.rs:340-354 — F and U as Option vectors: returns flags + payloads. */
static void col_vecs(const orc_mesh *m, uint8_t *f_known, uint8_t *u_known, double *F, double *U) {
    for (uint64_t i = 0; i < m->n_nodes; ++i) {
        const uint8_t k = m->known[i];
        f_known[2 * i] = (k & ORC_KNOWN_FX) != 0;  F[2 * i] = m->fx ? m->fx[i] : 0.0;
        f_known[2 * i + 1] = (k & ORC_KNOWN_FY) != 0;  F[2 * i + 1] = m->fy ? m->fy[i] : 0.0;
        u_known[2 * i] = (k & ORC_KNOWN_UX) != 0;  U[2 * i] = m->ux ? m->ux[i] : 0.0;
        u_known[2 * i + 1] = (k & ORC_KNOWN_UY) != 0;  U[2 * i + 1] = m->uy ? m->uy[i] : 0.0;
    }
}

static int check_bc_counts(const uint8_t *f_known, const uint8_t *u_known, uint64_t n,
                           uint64_t *n_free, uint64_t *n_known) {
    uint64_t rows = 0, uk = 0;
    for (uint64_t i = 0; i < n; ++i) { rows += f_known[i]; uk += u_known[i]; }
    *n_known = uk; *n_free = n - uk;
    /* The reference sizes both blocks from the displacement vector
     * (solver.rs:370-376) but walks rows by force (solver.rs:380-383); a
     * mismatch indexes out of bounds and panics there. */
    if (rows != n - uk) return fail(ORC_ERR_BAD_BC, "rows with known force != unknown displacements");
    return ORC_OK;
}

static int csr_alloc(orc_csr *A, uint64_t rows, uint64_t cols, uint64_t nnz) {
    A->n_rows = rows; A->n_cols = cols; A->nnz = nnz;
    A->rowptr = malloc((rows + 1) * sizeof(int64_t));
    A->col = malloc((size_t)(nnz ? nnz : 1) * sizeof(int32_t));
    A->val = malloc((size_t)(nnz ? nnz : 1) * sizeof(double));
    if (!A->rowptr || !A->col || !A->val) { orc_csr_free(A); return fail(ORC_ERR_OOM, "oom"); }
    return ORC_OK;
}

/* solver.rs:365-404 (dense known/unknown blocks), :427-432 (rhs) and
 * :126-137 (row-major scan of the dense K_ff keeping k != 0.0). */
int orc_partition_dense(const orc_mesh *m, const double *K, orc_csr *Kff, double *rhs,
                        int64_t *free_map, uint64_t *n_free_out) {
    const size_t n = 2 * (size_t)m->n_nodes;
    memset(Kff, 0, sizeof *Kff);
    uint8_t *fk = malloc(n ? n : 1), *uk = malloc(n ? n : 1);
    double *F = malloc((n ? n : 1) * sizeof(double)), *U = malloc((n ? n : 1) * sizeof(double));
    if (!fk || !uk || !F || !U) { free(fk); free(uk); free(F); free(U); return fail(ORC_ERR_OOM, "oom"); }
    col_vecs(m, fk, uk, F, U);
    uint64_t nf, nk;
    int rc = check_bc_counts(fk, uk, n, &nf, &nk);
    if (rc) { free(fk); free(uk); free(F); free(U); return rc; }
    *n_free_out = nf;
    {
        int64_t c = 0;
        for (size_t i = 0; i < n; ++i) free_map[i] = uk[i] ? -1 : c++;
    }
    double *known = calloc(nz1(nf * nk), sizeof(double));   /* :373-374 */
    double *unknown = calloc(nz1(nf * nf), sizeof(double)); /* :375-376 */
    if (!known || !unknown) { free(known); free(unknown); free(fk); free(uk); free(F); free(U); return fail(ORC_ERR_OOM, "oom (dense partition)"); }
    size_t lr = 0;
    for (size_t row = 0; row < n; ++row) {                    /* :380 */
        if (!fk[row]) continue;
        size_t ki = 0, ui = 0;
        for (size_t col = 0; col < n; ++col) {                /* :388 */
            if (uk[col]) { known[lr + ki * nf] = K[row + col * n] * U[col]; ++ki; }
            else { unknown[lr + ui * nf] = K[row + col * n]; ++ui; }
        }
        ++lr;
    }
    for (size_t i = 0; i < nf * nk; ++i) known[i] *= -1.0;    /* :402 */
    /* column_sum(): out = 0; out += column j for j ascending (:427) */
    for (size_t i = 0; i < nf; ++i) rhs[i] = 0.0;
    for (size_t j = 0; j < nk; ++j)
        for (size_t i = 0; i < nf; ++i) rhs[i] += known[i + j * nf];
    {
        size_t i = 0;                                         /* :428-432 */
        for (size_t d = 0; d < n; ++d) if (fk[d]) { rhs[i] += F[d]; ++i; }
    }
    /* :126-136 dense -> COO -> CSR, row-major scan, k != 0.0 */
    uint64_t nnz = 0;
    for (size_t r = 0; r < nf; ++r)
        for (size_t c = 0; c < nf; ++c) if (unknown[r + c * nf] != 0.0) ++nnz;
    rc = csr_alloc(Kff, nf, nf, nnz);
    if (!rc) {
        int64_t p = 0;
        for (size_t r = 0; r < nf; ++r) {
            Kff->rowptr[r] = p;
            for (size_t c = 0; c < nf; ++c) {
                const double k = unknown[r + c * nf];
                if (k != 0.0) { Kff->col[p] = (int32_t)c; Kff->val[p] = k; ++p; }
            }
        }
        Kff->rowptr[nf] = p;
    }
    free(known); free(unknown); free(fk); free(uk); free(F); free(U);
    return rc;
}

/* Identical arithmetic on the structural CSR: entries absent from the CSR are
 * exact +0.0 in the dense matrix and contribute -0.0 to the row sum, which
 * never changes it (s + -0.0 == s, and the sum starts at +0.0). */
int orc_partition_sparse(const orc_mesh *m, const orc_csr *K, orc_csr *Kff, double *rhs,
                         int64_t *free_map, uint64_t *n_free_out) {
    const size_t n = 2 * (size_t)m->n_nodes;
    memset(Kff, 0, sizeof *Kff);
    uint8_t *fk = malloc(n ? n : 1), *uk = malloc(n ? n : 1);
    double *F = malloc((n ? n : 1) * sizeof(double)), *U = malloc((n ? n : 1) * sizeof(double));
    if (!fk || !uk || !F || !U) { free(fk); free(uk); free(F); free(U); return fail(ORC_ERR_OOM, "oom"); }
    col_vecs(m, fk, uk, F, U);
    uint64_t nf, nk;
    int rc = check_bc_counts(fk, uk, n, &nf, &nk);
    if (rc) { free(fk); free(uk); free(F); free(U); return rc; }
    *n_free_out = nf;
    {
        int64_t c = 0;
        for (size_t i = 0; i < n; ++i) free_map[i] = uk[i] ? -1 : c++;
    }
    uint64_t nnz = 0;
    for (size_t row = 0; row < n; ++row) {
        if (!fk[row]) continue;
        for (int64_t p = K->rowptr[row]; p < K->rowptr[row + 1]; ++p)
            if (!uk[K->col[p]] && K->val[p] != 0.0) ++nnz;
    }
    rc = csr_alloc(Kff, nf, nf, nnz);
    if (!rc) {
        int64_t q = 0; size_t lr = 0;
        for (size_t row = 0; row < n; ++row) {
            if (!fk[row]) continue;
            Kff->rowptr[lr] = q;
            double s = 0.0;
            for (int64_t p = K->rowptr[row]; p < K->rowptr[row + 1]; ++p) {
                const int32_t col = K->col[p];
                const double k = K->val[p];
                if (uk[col]) s += (k * U[col]) * -1.0;
                else if (k != 0.0) { Kff->col[q] = (int32_t)free_map[col]; Kff->val[q] = k; ++q; }
            }
            rhs[lr] = s + F[row];
            ++lr;
        }
        Kff->rowptr[nf] = q;
    }
    free(fk); free(uk); free(F); free(U);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* CG                                                                        */
/* ------------------------------------------------------------------------ */

/* solver.rs:31-36 — `&CsrMatrix * DVector` (nalgebra-sparse spmm_csr_dense):
 * per row, dot = 0; dot += a_ik * x_k in stored (ascending-column) order. */
void orc_spmv(const orc_csr *A, const double *x, double *y) {
    for (uint64_t i = 0; i < A->n_rows; ++i) {
        double dot = 0.0;
        for (int64_t p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) dot += A->val[p] * x[A->col[p]];
        y[i] = dot;
    }
}

/* argmin-math Vec<f64> dot: zip, multiply, sum from 0.0 left to right. */
static double dot_seq(const double *a, const double *b, uint64_t n) {
    double s = 0.0;
    for (uint64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* solver.rs:141-157, 167-176 with argmin 0.10 ConjugateGradient + Executor +
 * IterState restated:
 *   init : r0 = -(b - A x0), p0 = -r0, rtr = r0.r0, cost = sqrt(rtr)
 *   iter : q = A p; alpha = rtr/(p.q); x = x + alpha p; r = r + alpha q;
 *          rtr' = r.r; beta = rtr'/rtr; p = (-1*r) + beta p; cost = sqrt(r.r)
 *   loop : stop when iter >= max_iters or best_cost <= target_cost, checked
 *          before every iteration; best_param := param whenever cost < best.
 * Returns best_param (solver.rs:167).  opt->jacobi / opt->rel_tol select the
 * north-star port-mode PCG instead (not the reference algorithm). */
int orc_cg(const orc_csr *A, const double *b, double *x, const orc_cg_options *opt,
           uint64_t *iters_out, double *final_cost) {
    const uint64_t n = A->n_rows;
    double *r = malloc((n ? n : 1) * sizeof(double)), *p = malloc((n ? n : 1) * sizeof(double));
    double *q = malloc((n ? n : 1) * sizeof(double)), *xc = calloc((n ? n : 1), sizeof(double));
    double *dinv = NULL, *z = NULL;
    if (!r || !p || !q || !xc) { free(r); free(p); free(q); free(xc); return fail(ORC_ERR_OOM, "oom"); }
    uint64_t it = 0;
    if (opt->jacobi || opt->rel_tol > 0.0) {
        /* ---- port mode: (Jacobi-)PCG to a relative residual ---------------- */
        dinv = malloc((n ? n : 1) * sizeof(double)); z = malloc((n ? n : 1) * sizeof(double));
        if (!dinv || !z) { free(r); free(p); free(q); free(xc); free(dinv); free(z); return fail(ORC_ERR_OOM, "oom"); }
        for (uint64_t i = 0; i < n; ++i) {
            double d = 1.0;
            if (opt->jacobi) {
                d = 0.0;
                for (int64_t k = A->rowptr[i]; k < A->rowptr[i + 1]; ++k) if ((uint64_t)A->col[k] == i) d = A->val[k];
                if (d == 0.0) d = 1.0;
            }
            dinv[i] = 1.0 / d;
        }
        for (uint64_t i = 0; i < n; ++i) { x[i] = 0.0; r[i] = b[i]; z[i] = r[i] * dinv[i]; p[i] = z[i]; }
        const double bb = dot_seq(b, b, n);
        const double thr2 = (opt->rel_tol > 0.0) ? opt->rel_tol * opt->rel_tol * bb
                                                 : opt->target_cost * opt->target_cost;
        double rz = dot_seq(r, z, n), rr = bb;
        while (it < opt->max_iter && rr > thr2) {
            orc_spmv(A, p, q);
            const double alpha = rz / dot_seq(p, q, n);
            for (uint64_t i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * q[i]; z[i] = r[i] * dinv[i]; }
            const double rz_n = dot_seq(r, z, n);
            rr = dot_seq(r, r, n);
            const double beta = rz_n / rz;
            rz = rz_n;
            for (uint64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
            ++it;
        }
        *iters_out = it; *final_cost = sqrt(rr);
        free(r); free(p); free(q); free(xc); free(dinv); free(z);
        return ORC_OK;
    }
    /* ---- reference mode ---------------------------------------------------- */
    for (uint64_t i = 0; i < n; ++i) xc[i] = 0.0;             /* solver.rs:143 */
    orc_spmv(A, xc, q);
    for (uint64_t i = 0; i < n; ++i) { r[i] = (b[i] - q[i]) * -1.0; p[i] = r[i] * -1.0; }
    double rtr = dot_seq(r, r, n);
    double cost = (opt->cost_kind == ORC_COST_SQ) ? rtr : sqrt(rtr);
    double best = INFINITY;
    if (cost < best) { best = cost; memcpy(x, xc, n * sizeof(double)); }
    for (;;) {
        if (it >= opt->max_iter) break;                       /* solver.rs:153 */
        if (best <= opt->target_cost) break;                  /* solver.rs:154 */
        orc_spmv(A, p, q);
        const double alpha = rtr / dot_seq(p, q, n);
        for (uint64_t i = 0; i < n; ++i) xc[i] = xc[i] + alpha * p[i];
        for (uint64_t i = 0; i < n; ++i) r[i] = r[i] + alpha * q[i];
        const double rtr_n = dot_seq(r, r, n);
        const double beta = rtr_n / rtr;
        rtr = rtr_n;
        for (uint64_t i = 0; i < n; ++i) p[i] = (r[i] * -1.0) + beta * p[i];
        cost = (opt->cost_kind == ORC_COST_SQ) ? rtr : sqrt(rtr);
        ++it;
        if (cost < best) { best = cost; memcpy(x, xc, n * sizeof(double)); }
        if (!(cost == cost)) break;                           /* NaN: nothing can improve */
    }
    *iters_out = it; *final_cost = best;
    free(r); free(p); free(q); free(xc);
    return ORC_OK;
}

/* ------------------------------------------------------------------------ */
/* reactions, stress                                                         */
/* ------------------------------------------------------------------------ */

/* solver.rs:457-469 — for DOFs whose force was None: f = sum over ALL columns
 * (ascending) of K[i,col]*u[col], starting from 0.0. */
void orc_reactions_dense(const orc_mesh *m, const double *K, const double *u, double *f) {
    const size_t n = 2 * (size_t)m->n_nodes;
    for (size_t i = 0; i < n; ++i) {
        const uint8_t k = m->known[i / 2];
        if (k & ((i & 1) ? ORC_KNOWN_FY : ORC_KNOWN_FX)) continue;
        double s = 0.0;
        for (size_t col = 0; col < n; ++col) s += K[i + col * n] * u[col];
        f[i] = s;
    }
}

void orc_reactions_sparse(const orc_mesh *m, const orc_csr *K, const double *u, double *f) {
    const size_t n = 2 * (size_t)m->n_nodes;
    for (size_t i = 0; i < n; ++i) {
        const uint8_t k = m->known[i / 2];
        if (k & ((i & 1) ? ORC_KNOWN_FY : ORC_KNOWN_FX)) continue;
        double s = 0.0;
        for (int64_t p = K->rowptr[i]; p < K->rowptr[i + 1]; ++p) s += K->val[p] * u[K->col[p]];
        f[i] = s;
    }
}

/* solver.rs:496-535 — sigma = (D*B)*u_e, sign = -1 iff sx+sy < 1.0,
 * stress = sqrt(sx^2+sy^2)*sign.  txy is computed and ignored. */
void orc_stress(const orc_mesh *m, const orc_material *mat, const double *ux, const double *uy,
                double *stress, double *sigma3) {
    double D[9];
    for (uint64_t e = 0; e < m->n_elems; ++e) {
        const uint32_t nd[3] = {m->n0[e], m->n1[e], m->n2[e]};
        const double ue[6] = {ux[nd[0]], uy[nd[0]], ux[nd[1]], uy[nd[1]], ux[nd[2]], uy[nd[2]]};
        double B[18], DB[18], s[3];
        orc_stress_strain(mat->poisson_ratio, mat->youngs_modulus, D);   /* :516 */
        orc_strain_displacement(m, e, orc_element_area(m, e), B);        /* :517-521 */
        matmul_seq(D, B, DB, 3, 3, 6);
        matmul_seq(DB, ue, s, 3, 6, 1);                                   /* :522 */
        const int sign = (s[0] + s[1] < 1.0) ? -1 : 1;                    /* :524-530 */
        stress[e] = sqrt(s[0] * s[0] + s[1] * s[1]) * (double)sign;       /* :532-533 */
        if (sigma3) { sigma3[3 * e] = s[0]; sigma3[3 * e + 1] = s[1]; sigma3[3 * e + 2] = s[2]; }
    }
}

/* ------------------------------------------------------------------------ */
/* pipeline                                                                  */
/* ------------------------------------------------------------------------ */

/* solver.rs:543-586 (run) and :412-487 (solve). */
int orc_run(const orc_mesh *m, const orc_material *mat, const orc_cg_options *opt, int dense,
            orc_result *out, orc_stats *st) {
    const uint64_t N = m->n_nodes, E = m->n_elems;
    const size_t n = 2 * (size_t)N;
    orc_stats local; if (!st) st = &local;
    memset(st, 0, sizeof *st);
    for (uint64_t e = 0; e < E; ++e)
        if (m->n0[e] >= N || m->n1[e] >= N || m->n2[e] >= N)
            return fail(ORC_ERR_BAD_INDEX, "element node index out of range");
    double t0 = now_s();
    double *Ke = malloc((size_t)(E ? E : 1) * 36 * sizeof(double));
    if (!Ke) return fail(ORC_ERR_OOM, "oom");
    orc_element_stiffness(m, mat, Ke);                        /* :553-563 */
    st->t_elem = now_s() - t0;

    t0 = now_s();
    double *Kd = NULL; orc_csr Ks; memset(&Ks, 0, sizeof Ks);
    int rc = ORC_OK;
    if (dense) {
        Kd = malloc(nz1(n * n) * sizeof(double));
        if (!Kd) { free(Ke); return fail(ORC_ERR_OOM, "oom (dense K)"); }
        orc_assemble_dense(m, Ke, Kd);                        /* :571-572 */
    } else {
        rc = orc_assemble_sparse(m, Ke, &Ks);
        if (rc) { free(Ke); return rc; }
        st->nnz_structural = Ks.nnz;
    }
    free(Ke);
    st->t_asm = now_s() - t0;

    t0 = now_s();
    orc_csr Kff; memset(&Kff, 0, sizeof Kff);
    double *rhs = malloc((n ? n : 1) * sizeof(double));
    int64_t *free_map = malloc((n ? n : 1) * sizeof(int64_t));
    double *U = malloc((n ? n : 1) * sizeof(double)), *F = malloc((n ? n : 1) * sizeof(double));
    double *xs = NULL;
    uint64_t nf = 0;
    if (!rhs || !free_map || !U || !F) { rc = fail(ORC_ERR_OOM, "oom"); goto done; }
    rc = dense ? orc_partition_dense(m, Kd, &Kff, rhs, free_map, &nf)
               : orc_partition_sparse(m, &Ks, &Kff, rhs, free_map, &nf);
    if (rc) goto done;
    st->n_free = nf; st->n_constrained = n - nf; st->nnz_ff = Kff.nnz;
    st->t_part = now_s() - t0;

    t0 = now_s();                                             /* :435 timer (CG only here; the
                                                                 dense->CSR scan is in t_part) */
    xs = malloc((nf ? nf : 1) * sizeof(double));
    if (!xs) { rc = fail(ORC_ERR_OOM, "oom"); goto done; }
    st->b_norm = sqrt(dot_seq(rhs, rhs, nf));
    rc = orc_cg(&Kff, rhs, xs, opt, &st->iters, &st->final_cost);
    if (rc) goto done;
    st->t_solve = now_s() - t0;

    t0 = now_s();
    for (size_t d = 0; d < n; ++d) {                          /* :444-454 */
        const uint8_t k = m->known[d / 2];
        const int uknown = (k & ((d & 1) ? ORC_KNOWN_UY : ORC_KNOWN_UX)) != 0;
        const int fknown = (k & ((d & 1) ? ORC_KNOWN_FY : ORC_KNOWN_FX)) != 0;
        U[d] = uknown ? ((d & 1) ? m->uy[d / 2] : m->ux[d / 2]) : xs[free_map[d]];
        F[d] = fknown ? ((d & 1) ? m->fy[d / 2] : m->fx[d / 2]) : 0.0;
    }
    if (dense) orc_reactions_dense(m, Kd, U, F); else orc_reactions_sparse(m, &Ks, U, F);
    for (uint64_t i = 0; i < N; ++i) {                        /* :476-482 */
        out->ux[i] = U[2 * i]; out->uy[i] = U[2 * i + 1];
        out->fx[i] = F[2 * i]; out->fy[i] = F[2 * i + 1];
    }
    st->t_react = now_s() - t0;

    t0 = now_s();
    orc_stress(m, mat, out->ux, out->uy, out->stress, NULL);  /* :578-583 */
    st->t_stress = now_s() - t0;
done:
    free(Kd); orc_csr_free(&Ks); orc_csr_free(&Kff);
    free(rhs); free(free_map); free(U); free(F); free(xs);
    return rc;
}
