/*
 * magnetite_oracle.h — CPU restatement of Magnetite's numerical core.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker / the timed
 * CPU baseline.  The product path (magnetite_b200/, libmagnetite_b200.so)
 * never links, imports or executes it.
 *
 * PARITY UNPINNED: the reference (kyle-tennison/Magnetite, Rust) ships no
 * tests, no golden vectors and no numeric fixtures, and neither rustc/cargo
 * nor the crates holding the arithmetic (nalgebra ^0.32.4, nalgebra-sparse
 * ^0.9.0, argmin ^0.10.0, argmin-math ^0.4.0; Cargo.toml:14-20) exist in this
 * image, so the reference itself cannot be run here.  This oracle follows the
 * reference *source* line by line (citations below are file:line under the
 * reference checkout) and restates the published algorithms of those crates
 * at the reference's call sites.  It is pinned only against analytic
 * known-answer tests (patch test, rigid-body modes), the closed forms of the
 * reference's own documentation (under-the-hood.md), an independent
 * numpy/scipy restatement (oracle/reference_semantics.py) and — to the
 * resolution of a picture, about 1 % of the displacement, nowhere near
 * rounding level — the outputs of its own solver the reference publishes:
 * the true-scale deformed outlines of examples/linkedin-logo/output.png,
 * media/tensilve-results.png and examples/cover-eample/output.png
 * (tests/test_reference_picture.py,
 * tests/golden/measure_reference_picture.py).
 *
 * All arithmetic is fp64, compiled with -ffp-contract=off so that every
 * multiply and add is rounded separately, as rustc does.
 */
#ifndef MAGNETITE_ORACLE_H
#define MAGNETITE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/solver.rs:17-19 */
#define ORC_DOF 2
#define ORC_MAX_CG_ITER 10000000ull
#define ORC_TARGET_CG_COST 1e-4

/* bit flags of orc_mesh.known[node]: which Option<f64> fields of
 * datatypes.rs:8-14 are Some(..) on input */
#define ORC_KNOWN_UX 1u
#define ORC_KNOWN_UY 2u
#define ORC_KNOWN_FX 4u
#define ORC_KNOWN_FY 8u

/* Flattened view of Vec<Node>, Vec<Element> (datatypes.rs:2-20). */
typedef struct {
    uint64_t n_nodes, n_elems;
    const double *x, *y;            /* node.vertex.{x,y}                     */
    const uint32_t *n0, *n1, *n2;   /* element.nodes[0..3]                   */
    const double *ux, *uy, *fx, *fy;/* payload of the Some(..) fields        */
    const uint8_t *known;           /* ORC_KNOWN_* per node                  */
} orc_mesh;

typedef struct {                     /* datatypes.rs:23-29 (CL fields unused) */
    double youngs_modulus, poisson_ratio, part_thickness;
} orc_material;

typedef struct {                     /* CSR with 64-bit row pointers          */
    uint64_t n_rows, n_cols, nnz;
    int64_t *rowptr;                 /* n_rows+1                              */
    int32_t *col;                    /* nnz, ascending per row                */
    double *val;                     /* nnz                                   */
} orc_csr;

enum { ORC_COST_NORM = 0, ORC_COST_SQ = 1 };

typedef struct {
    uint64_t max_iter;     /* solver.rs:18 default 1e7                        */
    double target_cost;    /* solver.rs:19 default 1e-4 (absolute)            */
    int cost_kind;         /* argmin cost: ||r||_2 (default) or r.r           */
    /* port-mode PCG (north-star semantics; not the reference algorithm):    */
    int jacobi;            /* 0 = reference CG, 1 = Jacobi-PCG                */
    double rel_tol;        /* if >0: stop at ||r||_2 <= rel_tol*||b||_2       */
} orc_cg_options;

typedef struct {
    uint64_t iters;
    double final_cost;     /* in the unit of cost_kind / ||r||_2 for rel_tol  */
    double b_norm;
    uint64_t n_free, n_constrained, nnz_ff, nnz_structural;
    double t_elem, t_asm, t_part, t_solve, t_react, t_stress; /* seconds      */
} orc_stats;

typedef struct {
    double *ux, *uy, *fx, *fy;       /* n_nodes each, caller-allocated        */
    double *stress;                  /* n_elems, caller-allocated             */
} orc_result;

/* error codes */
enum { ORC_OK = 0, ORC_ERR_OOM = -1, ORC_ERR_BAD_BC = -2, ORC_ERR_BAD_INDEX = -3,
       ORC_ERR_CG = -4 };

/* --- element level (solver.rs:187-278) --------------------------------- */
double orc_element_area(const orc_mesh *m, uint64_t e);
void orc_strain_displacement(const orc_mesh *m, uint64_t e, double area, double B[18]);
void orc_stress_strain(double nu, double E, double D[9]);
void orc_element_stiffness_one(const orc_mesh *m, uint64_t e, const orc_material *mat,
                               double Ke[36]);
void orc_element_stiffness(const orc_mesh *m, const orc_material *mat, double *Ke /*E*36*/);

/* --- assembly (solver.rs:290-331) --------------------------------------- */
/* faithful-dense: K is (2N)x(2N) column-major (nalgebra DMatrix), zeroed here */
void orc_assemble_dense(const orc_mesh *m, const double *Ke, double *K);
/* sparse: same per-entry accumulation order, structural CSR of the full K  */
int orc_assemble_sparse(const orc_mesh *m, const double *Ke, orc_csr *K);

/* --- partition + rhs (solver.rs:340-432) -------------------------------- */
/* free_map[dof] = reduced index or -1; returns n_free in *n_free            */
int orc_partition_dense(const orc_mesh *m, const double *K, orc_csr *Kff, double *rhs,
                        int64_t *free_map, uint64_t *n_free);
int orc_partition_sparse(const orc_mesh *m, const orc_csr *K, orc_csr *Kff, double *rhs,
                         int64_t *free_map, uint64_t *n_free);

/* --- CG (solver.rs:119-177 + argmin ConjugateGradient/Executor) -------- */
void orc_spmv(const orc_csr *A, const double *x, double *y);
int orc_cg(const orc_csr *A, const double *b, double *x, const orc_cg_options *opt,
           uint64_t *iters, double *final_cost);

/* --- reactions, stress (solver.rs:457-473, 496-535) --------------------- */
void orc_reactions_dense(const orc_mesh *m, const double *K, const double *u /*2N*/,
                         double *f /*2N in/out*/);
void orc_reactions_sparse(const orc_mesh *m, const orc_csr *K, const double *u, double *f);
void orc_stress(const orc_mesh *m, const orc_material *mat, const double *ux, const double *uy,
                double *stress, double *sigma3 /* optional E*3: sx,sy,txy or NULL */);

/* --- whole pipeline (solver.rs:543-586) --------------------------------- */
/* dense=1: the reference's data structures (dense (2N)^2 + dense partition);
 * dense=0: identical arithmetic with CSR storage ("ref-sparse").           */
int orc_run(const orc_mesh *m, const orc_material *mat, const orc_cg_options *opt, int dense,
            orc_result *out, orc_stats *stats);

void orc_csr_free(orc_csr *A);
const char *orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
