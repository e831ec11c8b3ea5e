/*
 * magnetite_b200.h — C ABI of the B200-native replacement for Magnetite's
 * numerical core (libmagnetite_b200.so, CUDA sm_100a, fp64).
 *
 * This is the drop-in boundary for everything inside the reference's
 *     solver::run(&mut Vec<Node>, &mut Vec<Element>, &ModelMetadata)
 *         -> Result<(), MagnetiteError>                  (src/solver.rs:543-586)
 * plus the pieces of it that other reference modules call directly
 * (solver::compute_element_area, used by mesher::check_ccw, src/mesher.rs:522-526).
 * The Rust side keeps its own structs; a thin shim (rust/)
 * flattens Vec<Node>/Vec<Element> (src/datatypes.rs:2-20) into the SoA arrays
 * below, calls mag_solve and writes the results back as Some(..).
 *
 * Conventions
 *  - plain pointers and sizes only; the caller owns every buffer it passes and
 *    the library keeps no pointer past the call (device scratch lives in the
 *    mag_ctx / mag_system handles and is freed with them);
 *  - every function returns 0 (MAG_OK) or a negative MAG_ERR_* code and never
 *    aborts or throws across the boundary; mag_last_error() gives the
 *    thread-local message, which the shim wraps in MagnetiteError::Solver
 *    (src/error.rs:4-22; src/solver.rs:160-164);
 *  - there is no CPU fallback: without a CUDA device every compute entry
 *    point fails with MAG_ERR_CUDA.
 */
#ifndef MAGNETITE_B200_H
#define MAGNETITE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAG_ABI_VERSION 3   /* 3: mag_options.assembly renumbered (0 = gather, now the default; 1 = sorted COO keys), precond 3,
                             *    result_scope; mag_stats.precond_used; mag_system_residual */

/* src/solver.rs:17-19 */
#define MAG_DOF 2
#define MAG_MAX_CG_ITER 10000000ull
#define MAG_TARGET_CG_COST 1e-4

enum {
    MAG_OK = 0,
    MAG_ERR_CUDA = -1,          /* CUDA runtime failure / no device            */
    MAG_ERR_OOM = -2,           /* device or host allocation failed            */
    MAG_ERR_BAD_BC = -3,        /* #rows with known force != #unknown displ.   *
                                 * (the reference panics: solver.rs:380-396)   */
    MAG_ERR_BAD_INDEX = -4,     /* element references a node >= n_nodes        */
    MAG_ERR_INDEFINITE = -5,    /* p.Ap == 0 or NaN during CG                  */
    MAG_ERR_NOT_CONVERGED = -6, /* max_iter reached (results still written)    */
    MAG_ERR_NCCL = -7,
    MAG_ERR_BAD_ARG = -8
};

/* which Option<f64> fields of Node (src/datatypes.rs:8-14) are Some(..) */
#define MAG_KNOWN_UX 1u
#define MAG_KNOWN_UY 2u
#define MAG_KNOWN_FX 4u
#define MAG_KNOWN_FY 8u

/* Vec<Node> + Vec<Element> flattened to SoA.  Replaces the `nodes`/`elements`
 * arguments of solver::run (src/solver.rs:544-545). */
typedef struct {
    uint64_t n_nodes, n_elems;
    const double *x, *y;               /* node.vertex.{x,y}                      */
    const uint32_t *n0, *n1, *n2;      /* element.nodes[0..3], 0-based           */
    const double *ux, *uy, *fx, *fy;   /* payload of Some(..) (ignored if None)  */
    const uint8_t *known;              /* MAG_KNOWN_* per node                   */
    int32_t on_device;                 /* 0: host pointers; 1: device pointers   */
} mag_mesh;

/* ModelMetadata (src/datatypes.rs:23-29; CL fields are unused by the solver). */
typedef struct {
    double youngs_modulus, poisson_ratio, part_thickness;
} mag_material;

/* Solver knobs.  The reference has compile-time constants only
 * (src/solver.rs:18-19, 143, 153-154); compat=1 reproduces them. */
typedef struct {
    double rel_tol;          /* stop at ||r||_2 <= rel_tol*||b||_2   (default 1e-9)   */
    double abs_tol;          /* compat target cost                   (default 1e-4)   */
    uint64_t max_iter;       /* default 1e7                                            */
    int32_t precond;         /* 0 none, 1 Jacobi, 2 Jacobi + aggregation coarse space (falls back to Jacobi when the
                              * system is not SPD), 3 (default) auto: 2 for systems of >= 20 000 unknowns, else 1 */
    int32_t compat;          /* 1: reference semantics — plain CG, x0=0, absolute cost */
    int32_t cost_kind;       /* compat cost: 0 = ||r||_2 (default), 1 = r.r            */
    int32_t drop_exact_zeros;/* 1 (default): K_ff keeps k != 0.0 only (solver.rs:132)  */
    int32_t check_every;     /* iterations per CUDA-graph chunk between residual polls */
    int32_t spmv_format;     /* 0/2 SELL-32, packed 16-bit column offsets when the band is < 32768 (default); 1 scalar CSR; 3 SELL-32 with 32-bit columns */
    int32_t want_sigma;      /* also return sx,sy,txy per element                      */
    int32_t allreduce;       /* multi-GPU dot products: 0 peer-memory mailbox (default), 1 NCCL  */
    int32_t coarse_aggregates; /* precond 2: number of aggregates (0 = auto, at most 2048)       */
    int32_t assembly;        /* how the full K (2x2-block CSR) is built; both modes leave the same K, K_ff, rhs bit for bit.
                              * 0 (default): gather — 3 (node, incidence) pairs per triangle sorted by node, then one
                              *    thread per node row adds the recomputed K_e rows of its incident triangles;
                              * 1: the north star's wording — K_e per triangle, 9 COO keys per triangle, stable sort,
                              *    warp-shuffle segmented reduction.
                              * (ABI 2 had them the other way round: 0 sorted keys, 1 gather.) */
    int32_t result_scope;    /* multi-GPU: 0 (default) every rank returns the complete result arrays; 1: a rank writes only
                              * its slice of them — nodes [N*r/R, N*(r+1)/R) (mag_partition_nodes) and elements
                              * [E*r/R, E*(r+1)/R) — so the job reads the result back over PCIe once, not once per rank */
    int32_t reserved0;
    void *stream;            /* cudaStream_t to run on, or NULL for the ctx's own      */
} mag_options;

/* Outputs of solver::run: every node gets ux,uy,fx,fy = Some (solver.rs:476-482),
 * every element gets stress = Some (solver.rs:532-533).  Caller-allocated. */
typedef struct {
    double *ux, *uy, *fx, *fy;         /* n_nodes each                            */
    double *stress;                    /* n_elems                                 */
    double *sigma;                     /* optional n_elems*3 (sx,sy,txy) or NULL  */
    int32_t on_device;
} mag_result;

typedef struct {
    uint64_t n_nodes, n_elems, n_dof, n_free, n_constrained;
    uint64_t nnz_structural;           /* stored entries of the full K (2x2 BSR)  */
    uint64_t nnz;                      /* stored entries of K_ff                   */
    uint64_t sell_entries;             /* padded entries of the SELL-32 copy       */
    uint64_t iters;                    /* CG iterations (solver.rs:101-104)        */
    double final_residual;             /* ||r||_2 at exit (recursive residual)     */
    double b_norm;                     /* ||b||_2                                  */
    int32_t converged;
    int32_t negative_definite;         /* all-clockwise mesh           */
    /* device time per phase, milliseconds (CUDA events on the run stream) */
    float ms_upload, ms_elem, ms_sort, ms_reduce, ms_bc, ms_format, ms_solve,
          ms_post, ms_download, ms_total;
    uint64_t kernel_launches;          /* launches issued by this call             */
    uint64_t spmv_bytes;               /* algorithmic bytes of one SpMV            */
    /* device-side timeline of one CG iteration, ns, averaged over the solve (%globaltimer):
     * 0 spmv.start - update_p.start(prev)   1 spmv duration       2 update_xr.start - spmv.end
     * 3 update_xr wait for the global p.q   4 update_xr duration  5 update_p.start - update_xr.end
     * 6 update_p wait for the global r.z    7 iterations timed                                  */
    double prof[8];
    float ms_coarse_setup;             /* precond 2: building P, Ac and Ac^-1 (inside ms_solve, first solve only) */
    uint32_t n_coarse;                 /* precond 2: coarse unknowns (3 per aggregate)                            */
    uint32_t sell_index_bits;          /* 16: column offsets from the row (band < 32768), 32: absolute columns    */
    uint32_t precond_used;             /* what the solve ran with: 0 plain CG, 1 Jacobi, 2 two-level             */
} mag_stats;

typedef struct mag_ctx mag_ctx;        /* device + stream + memory pool + comm    */
typedef struct mag_system mag_system;  /* assembled K, K_ff, rhs, maps on device  */

/* ---- lifecycle ----------------------------------------------------------- */
int mag_abi_version(void);
const char *mag_last_error(void);
int mag_device_count(int *count);
int mag_ctx_create(mag_ctx **ctx, int device);
void mag_ctx_destroy(mag_ctx *ctx);       /* collective on a context that has a communicator (mag_comm_init) */
void mag_options_default(mag_options *opt);

/* ---- the drop-in call: replaces solver::run (src/solver.rs:543-586) ------ */
int mag_solve(mag_ctx *ctx, const mag_mesh *mesh, const mag_material *mat,
              const mag_options *opt, mag_result *out, mag_stats *stats);

/* ---- phase-level entry points (bench + parity) --------------------------- */
/* element stiffness + COO sort + segmented reduce + BC elimination; replaces
 * solver.rs:549-572 and :420-432 (build_col_vecs .. rhs). */
int mag_assemble(mag_ctx *ctx, const mag_mesh *mesh, const mag_material *mat,
                 const mag_options *opt, mag_system **sys, mag_stats *stats);
/* CG on the assembled system + scatter + reactions + stress; replaces
 * solver.rs:435-482 and :578-583. */
int mag_system_solve(mag_system *sys, const mag_options *opt, mag_result *out, mag_stats *stats);
void mag_system_free(mag_system *sys);
int mag_system_info(const mag_system *sys, mag_stats *stats);

/* parity exports (host buffers, caller-allocated from mag_system_info sizes) */
int mag_system_export_kff(const mag_system *sys, int64_t *rowptr /*n_free+1*/, int32_t *col,
                          double *val, double *rhs /*n_free*/, int64_t *free_map /*n_dof*/);
int mag_system_export_full(const mag_system *sys, int64_t *rowptr /*n_dof+1*/, int32_t *col,
                           double *val /*nnz_structural*/);
/* K_e for every element, row-major 6x6 (solver.rs:263-278); host output. */
int mag_element_stiffness(mag_ctx *ctx, const mag_mesh *mesh, const mag_material *mat,
                          double *ke /* n_elems*36 */);
/* signed areas (solver.rs:187-193, pub: used by mesher::check_ccw). */
int mag_element_area(mag_ctx *ctx, const mag_mesh *mesh, double *area /* n_elems */);
/* B of every element, 3x6 row-major: solver::compute_strain_displacement_matrix (solver.rs:204-230, pub). */
int mag_strain_displacement(mag_ctx *ctx, const mag_mesh *mesh, double *b /* n_elems*18 */);
/* D, 3x3 row-major: solver::compute_stress_strain_matrix (solver.rs:240-250, pub). Host only. */
int mag_stress_strain(double poisson_ratio, double youngs_modulus, double *d /* 9 */);
/* stress recovery only (solver.rs:496-535). */
int mag_stress(mag_ctx *ctx, const mag_mesh *mesh, const mag_material *mat, const double *ux,
               const double *uy, double *stress, double *sigma /*or NULL*/);
/* y = K_ff x on host vectors (parity) and timed device loop (roofline). */
int mag_system_spmv(mag_system *sys, int format, const double *x, double *y);
int mag_system_spmv_bench(mag_system *sys, int format, int reps, float *ms_per_spmv,
                          uint64_t *algorithmic_bytes);
/* True residual of a returned displacement field over the rows this rank owns:
 * *rr_owned = sum (b - K_ff x)^2, *bb_owned = sum b^2, x rebuilt from (ux, uy) through the free-DOF map, the row
 * products in the reference's order (src/solver.rs:31-36).  ux, uy: n_nodes each, host (on_device = 0) or device
 * pointers.  Not collective: a multi-GPU caller adds both sums over the ranks.  The check a caller runs on the
 * result of solver::run (src/solver.rs:412-487), which only ever sees CG's recursive residual. */
int mag_system_residual(mag_system *sys, const double *ux, const double *uy, int on_device,
                        double *rr_owned, double *bb_owned);

/* ---- output stage: post_processor::csv_output (src/post_processor.rs:18-83), host only ---------
 * nodes.csv "x,y,ux,uy", elements.csv "n0,n1,n2,stress", "\n" line ends, f64 printed like Rust's `{}`
 * (shortest round-trip digits, never scientific, no ".0").  Buffered.  Errors: mag_host_last_error(). */
int mag_csv_output(const char *nodes_path, const char *elements_path, uint64_t n_nodes, const double *x,
                   const double *y, const double *ux, const double *uy, uint64_t n_elems,
                   const uint32_t *n0, const uint32_t *n1, const uint32_t *n2, const double *stress);
size_t mag_format_f64(double v, char *out /* >= 400 bytes, NUL-terminated */);
/* message of the last failed HOST-ONLY entry point on this thread (mag_csv_output, mag_reorder_rcm,
 * mag_mesh_band); the device entry points report through mag_last_error(). */
const char *mag_host_last_error(void);

/* ---- node renumbering for meshes without locality in their ids (gmsh order; the reference keeps
 * gmsh's tags as node ids, src/mesher.rs:663-671), host only — SURVEY §8(e) ----------------------
 * Reverse Cuthill-McKee on the node graph of the elements.  new_of_old[i] = new id of node i; nodes
 * no element references go last.  The caller permutes the node arrays and the connectivity (element
 * order and local node order unchanged), calls mag_solve, and reads result j of node i at
 * new_of_old[i]: the drop-in boundary keeps the original ids.  band_* (optional) = max |a-b| over
 * the node pairs of every element, before and after.  Deterministic. */
int mag_reorder_rcm(uint64_t n_nodes, uint64_t n_elems, const uint32_t *n0, const uint32_t *n1,
                    const uint32_t *n2, uint32_t *new_of_old /* n_nodes */, uint64_t *band_before,
                    uint64_t *band_after);
int mag_mesh_band(uint64_t n_nodes, uint64_t n_elems, const uint32_t *n0, const uint32_t *n1,
                  const uint32_t *n2, uint64_t *band);

/* ---- synthetic meshes generated on the device ------------- */
/* Plate(nx,ny,h): node (i,j) -> id j*(nx+1)+i at (i*h, j*h); cell -> [a,b,d],[a,d,c];
 * left edge clamped, right edge ux=ux_right, fy=0.  All outputs are device
 * buffers owned by the returned mesh handle. */
typedef struct mag_devmesh mag_devmesh;
int mag_devmesh_plate(mag_ctx *ctx, uint32_t nx, uint32_t ny, double h, double ux_right,
                      mag_devmesh **out);
/* Plate(nx,ny,h) minus the cells whose centre lies within radius*h of the lattice points
 * ((k+0.5)*pitch*h, (l+0.5)*pitch*h); unreferenced nodes dropped, survivors renumbered row-major. */
int mag_devmesh_perforated(mag_ctx *ctx, uint32_t nx, uint32_t ny, double h, uint32_t pitch,
                           uint32_t radius, double ux_right, mag_devmesh **out);
int mag_devmesh_download(const mag_devmesh *dm, double *x, double *y, uint32_t *n0, uint32_t *n1,
                         uint32_t *n2, double *ux, double *uy, double *fx, double *fy, uint8_t *known);
int mag_devmesh_view(const mag_devmesh *dm, mag_mesh *view);
void mag_devmesh_free(mag_devmesh *dm);

/* ---- multi-GPU: one process per GPU, contiguous row blocks --------------- */
int mag_comm_unique_id(void *id128 /* 128 bytes out */);
int mag_comm_init(mag_ctx *ctx, int rank, int nranks, const void *id128);
int mag_comm_rank(const mag_ctx *ctx, int *rank, int *nranks);
/* node range [lo,hi) owned by `rank` of `nranks` for a mesh of n_nodes */
int mag_partition_nodes(uint64_t n_nodes, int nranks, int rank, uint64_t *lo, uint64_t *hi);
/* halo plan of `rank` (host logic, no device needed): given every rank's reduced-row
 * boundaries row_lo[0..nranks] and the column extent [ext_lo[r], ext_hi[r]) its rows touch,
 * lists the index ranges [seg_lo,seg_hi) of `rank`'s own rows that rank seg_dst reads as halo. */
int mag_halo_plan(int nranks, int rank, const uint32_t *row_lo, const uint32_t *ext_lo,
                  const uint32_t *ext_hi, uint32_t *seg_lo, uint32_t *seg_hi, int32_t *seg_dst,
                  int32_t capacity, int32_t *n_segs);

/* ---- debug entry points used by the GPU unit tests ----------------------- */
int mag_debug_sort_pairs(mag_ctx *ctx, uint64_t *keys, uint32_t *payload, uint64_t n, int key_bits);
int mag_debug_exclusive_scan(mag_ctx *ctx, const uint32_t *in, uint32_t *out /*n+1*/, uint64_t n);
/* nranks-way row-block solve emulated on ONE GPU: the same partitioning, kernels and halo
 * stores as the multi-process path, with the allreduce replaced by a kernel. */
int mag_debug_virtual_solve(mag_ctx *ctx, const mag_mesh *mesh, const mag_material *mat,
                            const mag_options *opt, int nranks, mag_result *out, mag_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
