// api.cu — the extern "C" boundary of libmagnetite_b200.so and the host-side
// orchestration of the device pipeline.  See include/magnetite_b200.h for the
// contract and the reference lines each entry point replaces.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>

#include "assembly.cuh"
#include "bc.cuh"
#include "common.cuh"
#include "element.cuh"
#include "meshgen.cuh"
#include "pcg.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "spmv.cuh"

using namespace mag;

static thread_local std::string g_last_error;

// Runs `body`, mapping internal failures to an ABI error code.
template <class F>
static int guarded(F &&body) {
    try {
        body();
        return MAG_OK;
    } catch (const Failure &f) {
        g_last_error = f.msg;
        cudaGetLastError();   // clear a sticky-free error so the next call starts clean
        return f.code;
    } catch (const std::bad_alloc &) {
        g_last_error = "host allocation failed";
        return MAG_ERR_OOM;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return MAG_ERR_BAD_ARG;
    }
}

// Binds the context to the calling thread and selects the stream of the call.
struct CallScope {
    mag_ctx *ctx;
    CallScope(mag_ctx *c, const mag_options *opt) : ctx(c) {
        if (!c) fail(MAG_ERR_BAD_ARG, "null context");
        MAG_CUDA(cudaSetDevice(c->device));
        // the legacy NULL stream cannot be captured into a graph: use our own
        c->stream = (opt && opt->stream) ? (cudaStream_t)opt->stream : c->own_stream;
        c->launches = 0;
    }
    ~CallScope() { ctx->stream = ctx->own_stream; }
};

// Everything mag_assemble leaves on the device.
struct mag_system {
    mag_ctx *ctx = nullptr;
    uint64_t n_nodes = 0, n_elems = 0;
    uint32_t node_lo = 0, node_hi = 0;
    mag_material mat{};
    DevBuf<double2> xy;
    DevBuf<uint32_t> n0, n1, n2;
    DevBuf<uint8_t> known;
    DevBuf<double> bc_ux, bc_uy, bc_fx, bc_fy;
    BsrMatrix K;
    DevBuf<uint32_t> rowmap, colmap;     // n_dof+1 each (last = total)
    uint32_t n_free = 0;
    CsrMatrix Kff;
    DevBuf<double> rhs, diag;
    SellMatrix sell;
    mag_stats stats{};
};

// ---------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------
extern "C" int mag_abi_version(void) { return MAG_ABI_VERSION; }
extern "C" const char *mag_last_error(void) { return g_last_error.c_str(); }

extern "C" int mag_device_count(int *count) {
    return guarded([&] {
        if (!count) fail(MAG_ERR_BAD_ARG, "null count");
        *count = 0;
        cudaError_t e = cudaGetDeviceCount(count);
        if (e != cudaSuccess) {
            *count = 0;
            fail(MAG_ERR_CUDA, "no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
        }
    });
}

extern "C" int mag_ctx_create(mag_ctx **out, int device) {
    return guarded([&] {
        if (!out) fail(MAG_ERR_BAD_ARG, "null ctx out");
        *out = nullptr;
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            fail(MAG_ERR_CUDA, "no CUDA device available: %s (magnetite_b200 has no CPU fallback)",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        if (device < 0 || device >= n) fail(MAG_ERR_BAD_ARG, "device %d out of range [0,%d)", device, n);
        std::unique_ptr<mag_ctx> c(new mag_ctx);
        c->device = device;
        MAG_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        MAG_CUDA(cudaGetDeviceProperties(&prop, device));
        c->sm_count = prop.multiProcessorCount;
        MAG_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        MAG_CUDA(cudaDeviceGetDefaultMemPool(&c->pool, device));
        uint64_t keep = ~0ull;   // keep freed blocks in the pool between solves
        MAG_CUDA(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep));
        MAG_CUDA(cudaMallocHost((void **)&c->h_scal, 64 * sizeof(double)));
        *out = c.release();
    });
}

extern "C" void mag_ctx_destroy(mag_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->own_stream);
    if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" void mag_options_default(mag_options *o) {
    if (!o) return;
    std::memset(o, 0, sizeof *o);
    o->rel_tol = 1e-9;
    o->abs_tol = MAG_TARGET_CG_COST;
    o->max_iter = MAG_MAX_CG_ITER;
    o->precond = 1;
    o->compat = 0;
    o->cost_kind = 0;
    o->drop_exact_zeros = 1;
    o->check_every = 50;
    o->spmv_format = 0;
    o->want_sigma = 0;
    o->stream = nullptr;
}

// ---------------------------------------------------------------------------
// mesh upload
// ---------------------------------------------------------------------------
static void check_mesh_args(const mag_mesh *m) {
    if (!m) fail(MAG_ERR_BAD_ARG, "null mesh");
    if (m->n_nodes >= (1ull << 31)) fail(MAG_ERR_BAD_ARG, "n_nodes must be < 2^31");
    if (m->n_elems * 9 >= (1ull << 32)) fail(MAG_ERR_BAD_ARG, "n_elems*9 must be < 2^32 per GPU");
    if (m->n_nodes && (!m->x || !m->y || !m->known)) fail(MAG_ERR_BAD_ARG, "mesh: x, y and known are required");
    if (m->n_elems && (!m->n0 || !m->n1 || !m->n2)) fail(MAG_ERR_BAD_ARG, "mesh: n0, n1, n2 are required");
}

template <class T>
static void upload_or_zero(mag_ctx *ctx, DevBuf<T> &dst, const T *src, size_t n, bool on_device) {
    dst.alloc(ctx, n);
    if (src) copy_to_device(ctx, dst.p, src, n, on_device);
    else dst.zero();
}

// geometry + connectivity only (enough for K_e, area, stress)
static void upload_geometry(mag_ctx *ctx, const mag_mesh *m, DevBuf<double2> &xy, DevBuf<uint32_t> &n0,
                            DevBuf<uint32_t> &n1, DevBuf<uint32_t> &n2) {
    const size_t N = m->n_nodes, E = m->n_elems;
    const bool dev = m->on_device != 0;
    xy.alloc(ctx, N);
    if (N) {
        if (dev) {
            MAG_LAUNCH(ctx, pack_xy_kernel, cdiv(N, 256), 256, 0, m->x, m->y, xy.p, N);
        } else {
            DevBuf<double> tx(ctx, N), ty(ctx, N);
            copy_to_device(ctx, tx.p, m->x, N, false);
            copy_to_device(ctx, ty.p, m->y, N, false);
            MAG_LAUNCH(ctx, pack_xy_kernel, cdiv(N, 256), 256, 0, (const double *)tx.p,
                       (const double *)ty.p, xy.p, N);
        }
    }
    upload_or_zero(ctx, n0, m->n0, E, dev);
    upload_or_zero(ctx, n1, m->n1, E, dev);
    upload_or_zero(ctx, n2, m->n2, E, dev);
    if (E) {
        DevBuf<int> bad(ctx, 1);
        bad.zero();
        MAG_LAUNCH(ctx, validate_conn_kernel, cdiv(E, 256), 256, 0, (const uint32_t *)n0.p,
                   (const uint32_t *)n1.p, (const uint32_t *)n2.p, E, (uint32_t)N, bad.p);
        int h_bad = 0;
        MAG_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        if (h_bad) fail(MAG_ERR_BAD_INDEX, "an element references a node index >= n_nodes (%llu)",
                        (unsigned long long)N);
    }
}

static uint32_t read_u32(mag_ctx *ctx, const uint32_t *dptr) {
    uint32_t v = 0;
    MAG_CUDA(cudaMemcpyAsync(&v, dptr, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    return v;
}

// ---------------------------------------------------------------------------
// assembly
// ---------------------------------------------------------------------------
static void assemble_impl(mag_ctx *ctx, const mag_mesh *m, const mag_material *mat,
                          const mag_options *opt, mag_system *S) {
    check_mesh_args(m);
    if (!mat) fail(MAG_ERR_BAD_ARG, "null material");
    const size_t N = m->n_nodes, E = m->n_elems, n_dof = 2 * N;
    const bool dev = m->on_device != 0;
    mag_stats &st = S->stats;
    std::memset(&st, 0, sizeof st);
    st.n_nodes = N; st.n_elems = E; st.n_dof = n_dof;
    S->ctx = ctx; S->n_nodes = N; S->n_elems = E; S->mat = *mat;
    S->node_lo = 0; S->node_hi = (uint32_t)N;
    EventTimer total(ctx->stream), phase(ctx->stream);
    total.start();

    // ---- upload -------------------------------------------------------------
    phase.start();
    upload_geometry(ctx, m, S->xy, S->n0, S->n1, S->n2);
    upload_or_zero(ctx, S->known, m->known, N, dev);
    upload_or_zero(ctx, S->bc_ux, m->ux, N, dev);
    upload_or_zero(ctx, S->bc_uy, m->uy, N, dev);
    upload_or_zero(ctx, S->bc_fx, m->fx, N, dev);
    upload_or_zero(ctx, S->bc_fy, m->fy, N, dev);
    upload_material(ctx, *mat);
    st.ms_upload = phase.stop();

    // ---- element stiffness (solver.rs:553-563) ------------------------------
    phase.start();
    DevBuf<double> kblk(ctx, E * 36);
    if (E)
        MAG_LAUNCH(ctx, element_stiffness_kernel, cdiv(E, kElemThreads), kElemThreads, 0,
                   (const double2 *)S->xy.p, (const uint32_t *)S->n0.p, (const uint32_t *)S->n1.p,
                   (const uint32_t *)S->n2.p, (const uint32_t *)nullptr, E, 1, kblk.p);
    st.ms_elem = phase.stop();

    // ---- COO keys + stable sort ----------------------------------------------
    phase.start();
    const size_t n_keys = E * 9;
    const int bits = bits_for(N + 1);
    DevBuf<uint64_t> keys(ctx, n_keys), keys_alt(ctx, n_keys);
    DevBuf<uint32_t> pay(ctx, n_keys), pay_alt(ctx, n_keys);
    if (E) {
        MAG_LAUNCH(ctx, emit_keys_kernel, cdiv(E, 256), 256, 0, (const uint32_t *)S->n0.p,
                   (const uint32_t *)S->n1.p, (const uint32_t *)S->n2.p, (const uint32_t *)nullptr, E,
                   bits, S->node_lo, S->node_hi, keys.p, pay.p);
        radix_sort_pairs(ctx, keys.p, pay.p, keys_alt.p, pay_alt.p, n_keys, 2 * bits);
    }
    keys_alt.release();
    pay_alt.release();
    st.ms_sort = phase.stop();

    // ---- segmented reduction into BSR ------------------------------------------
    phase.start();
    BsrMatrix &K = S->K;
    K.node_lo = S->node_lo; K.node_hi = S->node_hi;
    const uint32_t n_own = K.node_hi - K.node_lo;
    K.browptr.alloc(ctx, (size_t)n_own + 1);
    K.browptr.zero();
    {
        DevBuf<uint32_t> head(ctx, n_keys + 1);
        if (n_keys) {
            MAG_LAUNCH(ctx, mark_heads_kernel, cdiv(n_keys, 256), 256, 0, (const uint64_t *)keys.p,
                       n_keys, bits, K.node_lo, head.p, K.browptr.p);
        }
        // head -> uid (exclusive scan; uid[n_keys] = number of blocks)
        DevBuf<uint32_t> uid(ctx, n_keys + 1);
        exclusive_scan_u32(ctx, head.p, n_keys, uid.p, n_keys + 1);
        exclusive_scan_u32(ctx, K.browptr.p, n_own, K.browptr.p, (size_t)n_own + 1);
        K.n_blocks = read_u32(ctx, uid.p + n_keys);
        K.bcol.alloc(ctx, K.n_blocks);
        K.bval.alloc(ctx, (size_t)K.n_blocks * 4);
        if (n_keys)
            MAG_LAUNCH(ctx, segment_reduce_kernel, cdiv(n_keys, 256), 256, 0, (const uint64_t *)keys.p,
                       (const uint32_t *)pay.p, (const uint32_t *)head.p, (const uint32_t *)uid.p,
                       n_keys, bits, (const double *)kblk.p, K.bcol.p, K.bval.p);
    }
    keys.release();
    pay.release();
    kblk.release();
    st.nnz_structural = (uint64_t)K.n_blocks * 4;
    st.ms_reduce = phase.stop();

    // ---- Dirichlet elimination (solver.rs:340-432, 126-137) --------------------
    phase.start();
    S->rowmap.alloc(ctx, n_dof + 1);
    S->colmap.alloc(ctx, n_dof + 1);
    if (n_dof)
        MAG_LAUNCH(ctx, dof_flags_kernel, cdiv(n_dof, 256), 256, 0, (const uint8_t *)S->known.p, n_dof,
                   S->rowmap.p, S->colmap.p);
    exclusive_scan_u32(ctx, S->rowmap.p, n_dof, S->rowmap.p, n_dof + 1);
    exclusive_scan_u32(ctx, S->colmap.p, n_dof, S->colmap.p, n_dof + 1);
    const uint32_t n_rows = read_u32(ctx, S->rowmap.p + n_dof);
    const uint32_t n_cols = read_u32(ctx, S->colmap.p + n_dof);
    if (n_rows != n_cols)
        fail(MAG_ERR_BAD_BC,
             "inconsistent boundary conditions: %u DOFs have a known force but %u have an unknown "
             "displacement (the reference panics here, solver.rs:380-396)", n_rows, n_cols);
    S->n_free = n_cols;
    st.n_free = n_cols; st.n_constrained = n_dof - n_cols;
    CsrMatrix &A = S->Kff;
    A.n_rows = n_rows; A.row_lo = 0; A.n_cols = n_cols;
    A.rowptr.alloc(ctx, (size_t)n_rows + 1);
    S->rhs.alloc(ctx, n_rows);
    S->diag.alloc(ctx, n_rows);
    const int drop = opt ? opt->drop_exact_zeros : 1;
    const uint32_t n_owned_dof = 2 * n_own;
    if (n_owned_dof) {
        MAG_LAUNCH(ctx, eliminate_kernel<0>, cdiv(n_owned_dof, 256), 256, 0,
                   (const uint32_t *)K.browptr.p, (const uint32_t *)K.bcol.p, (const double *)K.bval.p,
                   K.node_lo, n_owned_dof, (const uint8_t *)S->known.p, (const uint32_t *)S->rowmap.p,
                   (const uint32_t *)S->colmap.p, (const double *)S->bc_ux.p, (const double *)S->bc_uy.p,
                   (const double *)S->bc_fx.p, (const double *)S->bc_fy.p, drop, A.row_lo,
                   A.rowptr.p, (const uint32_t *)nullptr, (int32_t *)nullptr, (double *)nullptr,
                   (double *)nullptr, (double *)nullptr);
    }
    exclusive_scan_u32(ctx, A.rowptr.p, n_rows, A.rowptr.p, (size_t)n_rows + 1);
    A.nnz = read_u32(ctx, A.rowptr.p + n_rows);
    A.col.alloc(ctx, A.nnz);
    A.val.alloc(ctx, A.nnz);
    if (n_owned_dof) {
        MAG_LAUNCH(ctx, eliminate_kernel<1>, cdiv(n_owned_dof, 256), 256, 0,
                   (const uint32_t *)K.browptr.p, (const uint32_t *)K.bcol.p, (const double *)K.bval.p,
                   K.node_lo, n_owned_dof, (const uint8_t *)S->known.p, (const uint32_t *)S->rowmap.p,
                   (const uint32_t *)S->colmap.p, (const double *)S->bc_ux.p, (const double *)S->bc_uy.p,
                   (const double *)S->bc_fx.p, (const double *)S->bc_fy.p, drop, A.row_lo,
                   (uint32_t *)nullptr, (const uint32_t *)A.rowptr.p, A.col.p, A.val.p, S->rhs.p,
                   S->diag.p);
    }
    st.nnz = A.nnz;
    st.ms_bc = phase.stop();

    // ---- solver format -------------------------------------------------------------
    phase.start();
    build_sell(ctx, A, S->sell);
    st.sell_entries = S->sell.entries;
    st.ms_format = phase.stop();
    st.spmv_bytes = A.nnz * 12ull + (uint64_t)n_rows * 16ull + ((uint64_t)n_rows + 1) * 4ull;
    st.ms_total = total.stop();
    st.kernel_launches = ctx->launches;
}

extern "C" int mag_assemble(mag_ctx *ctx, const mag_mesh *mesh, const mag_material *mat,
                            const mag_options *opt, mag_system **sys, mag_stats *stats) {
    return guarded([&] {
        if (!sys) fail(MAG_ERR_BAD_ARG, "null sys out");
        *sys = nullptr;
        CallScope scope(ctx, opt);
        std::unique_ptr<mag_system> S(new mag_system);
        try {
            assemble_impl(ctx, mesh, mat, opt, S.get());
        } catch (...) {
            cudaStreamSynchronize(ctx->stream);
            throw;
        }
        if (stats) *stats = S->stats;
        *sys = S.release();
    });
}

extern "C" void mag_system_free(mag_system *sys) {
    if (!sys) return;
    mag_ctx *ctx = sys->ctx;
    if (ctx) {
        cudaSetDevice(ctx->device);
        ctx->stream = ctx->own_stream;
    }
    delete sys;
    if (ctx) cudaStreamSynchronize(ctx->own_stream);
}

extern "C" int mag_system_info(const mag_system *sys, mag_stats *stats) {
    return guarded([&] {
        if (!sys || !stats) fail(MAG_ERR_BAD_ARG, "null argument");
        *stats = sys->stats;
    });
}

// ---------------------------------------------------------------------------
// solve
// ---------------------------------------------------------------------------
static void pcg_alloc(mag_ctx *ctx, mag_system *S, PcgWork &W) {
    const uint32_t n = S->Kff.n_rows;
    W.n = n; W.row_lo = 0;
    W.x.alloc(ctx, n); W.r.alloc(ctx, n); W.q.alloc(ctx, n); W.dinv.alloc(ctx, n);
    W.p_store.alloc(ctx, (size_t)S->n_free + 32);
    W.p_store.zero();
    W.p = W.p_store.p;
    const unsigned cap = (unsigned)ctx->sm_count * 8u;
    W.grid_vec = std::max(1u, std::min(cdiv(n, 256), cap));
    W.grid_spmv = sell_grid(ctx, S->sell.n_slices);
    W.partials.alloc(ctx, 2 * (size_t)std::max(W.grid_vec, std::max(W.grid_spmv, cap)));
    W.scal.alloc(ctx, 1);
    W.scal.zero();
}

static void enqueue_iteration(mag_ctx *ctx, mag_system *S, PcgWork &W, int parity, int format) {
    const SellMatrix &L = S->sell;
    const CsrMatrix &A = S->Kff;
    if (format == 1) {
        MAG_LAUNCH(ctx, pcg_spmv_csr_kernel, W.grid_vec, 256, 0, (const uint32_t *)A.rowptr.p,
                   (const int32_t *)A.col.p, (const double *)A.val.p, (const double *)W.p, W.q.p, W.n,
                   W.row_lo, W.partials.p, W.scal.p);
    } else {
        MAG_LAUNCH(ctx, pcg_spmv_kernel, W.grid_spmv, 256, 0, (const uint32_t *)L.slice_off.p,
                   (const int32_t *)L.col.p, (const double *)L.val.p, (const double *)W.p, W.q.p, W.n,
                   L.n_slices, W.row_lo, W.partials.p, W.scal.p);
    }
    MAG_LAUNCH(ctx, pcg_update_xr_kernel, W.grid_vec, 256, 0, W.x.p, W.r.p, (const double *)W.p,
               (const double *)W.q.p, (const double *)W.dinv.p, W.n, W.row_lo, parity, W.partials.p,
               W.scal.p);
    MAG_LAUNCH(ctx, pcg_update_p_kernel, W.grid_vec, 256, 0, W.p, (const double *)W.r.p,
               (const double *)W.dinv.p, W.n, W.row_lo, parity, (const PcgScalars *)W.scal.p);
}

static void solve_impl(mag_system *S, const mag_options *opt_in, mag_result *out, mag_stats *stats_out) {
    mag_ctx *ctx = S->ctx;
    mag_options opt;
    if (opt_in) opt = *opt_in; else mag_options_default(&opt);
    if (!out) fail(MAG_ERR_BAD_ARG, "null result");
    const size_t N = S->n_nodes, E = S->n_elems;
    if (N && (!out->ux || !out->uy || !out->fx || !out->fy)) fail(MAG_ERR_BAD_ARG, "result: ux, uy, fx, fy are required");
    if (E && !out->stress) fail(MAG_ERR_BAD_ARG, "result: stress is required");
    mag_stats st = S->stats;
    const uint64_t launches_before = ctx->launches;
    EventTimer phase(ctx->stream);
    const uint32_t n = S->Kff.n_rows;
    const int format = opt.spmv_format == 1 ? 1 : 2;
    const bool compat = opt.compat != 0;
    const int jacobi = compat ? 0 : (opt.precond != 0);
    int chunk = opt.check_every > 0 ? opt.check_every : 50;
    chunk += chunk & 1;   // iteration parity is baked into the graph: even chunk length

    // ---- CG ------------------------------------------------------------------------
    phase.start();
    PcgWork W;
    pcg_alloc(ctx, S, W);
    PcgScalars hs;
    std::memset(&hs, 0, sizeof hs);
    if (n) {
        MAG_LAUNCH(ctx, pcg_init_kernel, W.grid_vec, 256, 0, W.x.p, W.r.p, W.p, W.dinv.p,
                   (const double *)S->rhs.p, (const double *)S->diag.p, jacobi, n, W.row_lo,
                   W.partials.p, W.scal.p);
        MAG_CUDA(cudaMemcpyAsync(&hs, W.scal.p, sizeof hs, cudaMemcpyDeviceToHost, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    const double bb = hs.rr;
    st.b_norm = std::sqrt(bb);
    if (compat) hs.thr2 = opt.cost_kind == 1 ? opt.abs_tol : opt.abs_tol * opt.abs_tol;
    else hs.thr2 = opt.rel_tol * opt.rel_tol * bb;
    hs.max_iter = opt.max_iter;
    hs.iter = 0;
    hs.stop = 0;
    if (!(bb == bb)) hs.stop = 3;                       // NaN right-hand side
    else if (bb <= hs.thr2) hs.stop = 1;                // argmin: init cost already <= target
    else if (opt.max_iter == 0) hs.stop = 2;
    hs.ticket_a = hs.ticket_b = 0;
    if (n) {
        MAG_CUDA(cudaMemcpyAsync(W.scal.p, &hs, sizeof hs, cudaMemcpyHostToDevice, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (n && !hs.stop) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const uint64_t l0 = ctx->launches;
        MAG_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
        try {
            for (int i = 0; i < chunk; ++i) enqueue_iteration(ctx, S, W, i & 1, format);
        } catch (...) {
            cudaStreamEndCapture(ctx->stream, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
        }
        MAG_CUDA(cudaStreamEndCapture(ctx->stream, &graph));
        const uint64_t per_chunk = ctx->launches - l0;
        ctx->launches = l0;
        MAG_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        // pinned mirror of {stop} polled one chunk behind the queue
        PcgScalars *h_poll = reinterpret_cast<PcgScalars *>(ctx->h_scal);
        static_assert(sizeof(PcgScalars) <= 32 * sizeof(double), "pinned scratch too small");
        cudaEvent_t ev[2];
        MAG_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
        MAG_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
        PcgScalars *slot[2] = {h_poll, h_poll + 1};
        static_assert(2 * sizeof(PcgScalars) <= 64 * sizeof(double), "pinned scratch too small");
        int stop = 0;
        uint64_t queued = 0;
        try {
            for (uint64_t c = 0;; ++c) {
                const int s = (int)(c & 1);
                MAG_CUDA(cudaGraphLaunch(exec, ctx->stream));
                ctx->launches += per_chunk;
                MAG_CUDA(cudaMemcpyAsync(slot[s], W.scal.p, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream));
                MAG_CUDA(cudaEventRecord(ev[s], ctx->stream));
                queued += (uint64_t)chunk;
                if (c > 0) {           // look at the previous chunk while this one runs
                    MAG_CUDA(cudaEventSynchronize(ev[s ^ 1]));
                    stop = slot[s ^ 1]->stop;
                    if (stop) break;
                }
                if (queued >= opt.max_iter + (uint64_t)chunk) break;
            }
            MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        } catch (...) {
            cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
            cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
            throw;
        }
        cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
        MAG_CUDA(cudaMemcpyAsync(&hs, W.scal.p, sizeof hs, cudaMemcpyDeviceToHost, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    st.iters = hs.iter;
    st.final_residual = std::sqrt(hs.rr);
    st.converged = (hs.stop == 1) || n == 0;
    st.negative_definite = hs.first_pq < 0.0;
    st.ms_solve = phase.stop();
    if (hs.stop == 3)
        fail(MAG_ERR_INDEFINITE, "conjugate gradient broke down at iteration %llu (p.Ap = %g, r.r = %g)",
             (unsigned long long)hs.iter, hs.pq, hs.rr);

    // ---- scatter, reactions, stress (solver.rs:444-482, 496-535) -------------------------
    phase.start();
    DevBuf<double> ux(ctx, N), uy(ctx, N), fx(ctx, N), fy(ctx, N), stress(ctx, E), sigma;
    const bool want_sigma = out->sigma != nullptr;
    if (want_sigma) sigma.alloc(ctx, E * 3);
    if (N) {
        MAG_LAUNCH(ctx, scatter_solution_kernel, cdiv(N, 256), 256, 0, (const uint8_t *)S->known.p,
                   (const uint32_t *)S->colmap.p, (const double *)S->bc_ux.p, (const double *)S->bc_uy.p,
                   (const double *)W.x.p, N, ux.p, uy.p);
        const uint32_t n_owned_dof = 2 * (S->K.node_hi - S->K.node_lo);
        MAG_LAUNCH(ctx, reactions_kernel, cdiv(n_owned_dof, 256), 256, 0, (const uint32_t *)S->K.browptr.p,
                   (const uint32_t *)S->K.bcol.p, (const double *)S->K.bval.p, S->K.node_lo, n_owned_dof,
                   (const uint8_t *)S->known.p, (const double *)S->bc_fx.p, (const double *)S->bc_fy.p,
                   (const double *)ux.p, (const double *)uy.p, fx.p, fy.p);
    }
    if (E) {
        upload_material(ctx, S->mat);
        MAG_LAUNCH(ctx, stress_kernel, cdiv(E, 256), 256, 0, (const double2 *)S->xy.p,
                   (const uint32_t *)S->n0.p, (const uint32_t *)S->n1.p, (const uint32_t *)S->n2.p, E,
                   (const double *)ux.p, (const double *)uy.p, stress.p, want_sigma ? sigma.p : nullptr);
    }
    st.ms_post = phase.stop();

    phase.start();
    const bool odev = out->on_device != 0;
    copy_from_device(ctx, out->ux, ux.p, N, odev);
    copy_from_device(ctx, out->uy, uy.p, N, odev);
    copy_from_device(ctx, out->fx, fx.p, N, odev);
    copy_from_device(ctx, out->fy, fy.p, N, odev);
    copy_from_device(ctx, out->stress, stress.p, E, odev);
    if (want_sigma) copy_from_device(ctx, out->sigma, sigma.p, E * 3, odev);
    st.ms_download = phase.stop();
    st.kernel_launches = ctx->launches - launches_before;
    S->stats.iters = st.iters;
    if (stats_out) *stats_out = st;
    if (hs.stop == 2)
        fail(MAG_ERR_NOT_CONVERGED, "conjugate gradient stopped at max_iter = %llu with ||r|| = %g",
             (unsigned long long)hs.iter, st.final_residual);
}

extern "C" int mag_system_solve(mag_system *sys, const mag_options *opt, mag_result *out,
                                mag_stats *stats) {
    return guarded([&] {
        if (!sys) fail(MAG_ERR_BAD_ARG, "null system");
        CallScope scope(sys->ctx, opt);
        try {
            solve_impl(sys, opt, out, stats);
        } catch (...) {
            cudaStreamSynchronize(sys->ctx->stream);
            throw;
        }
    });
}

extern "C" int mag_solve(mag_ctx *ctx, const mag_mesh *mesh, const mag_material *mat,
                         const mag_options *opt, mag_result *out, mag_stats *stats) {
    mag_system *sys = nullptr;
    mag_stats a{};
    int rc = mag_assemble(ctx, mesh, mat, opt, &sys, &a);
    if (rc != MAG_OK) return rc;
    mag_stats s{};
    rc = mag_system_solve(sys, opt, out, &s);
    if (stats) {
        *stats = s;
        stats->kernel_launches = a.kernel_launches + s.kernel_launches;
        stats->ms_total = a.ms_total + s.ms_solve + s.ms_post + s.ms_download;
    }
    mag_system_free(sys);
    return rc;
}

// ---------------------------------------------------------------------------
// parity exports
// ---------------------------------------------------------------------------
extern "C" int mag_system_export_kff(const mag_system *sys, int64_t *rowptr, int32_t *col, double *val,
                                     double *rhs, int64_t *free_map) {
    return guarded([&] {
        if (!sys) fail(MAG_ERR_BAD_ARG, "null system");
        mag_ctx *ctx = sys->ctx;
        CallScope scope(ctx, nullptr);
        const CsrMatrix &A = sys->Kff;
        const size_t n = A.n_rows, n_dof = 2 * sys->n_nodes;
        std::vector<uint32_t> tmp(std::max(n + 1, n_dof + 1));
        if (rowptr) {
            MAG_CUDA(cudaMemcpyAsync(tmp.data(), A.rowptr.p, (n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            MAG_CUDA(cudaStreamSynchronize(ctx->stream));
            for (size_t i = 0; i <= n; ++i) rowptr[i] = (int64_t)tmp[i];
        }
        if (col) copy_from_device(ctx, col, (const int32_t *)A.col.p, A.nnz, false);
        if (val) copy_from_device(ctx, val, (const double *)A.val.p, A.nnz, false);
        if (rhs) copy_from_device(ctx, rhs, (const double *)sys->rhs.p, n, false);
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        if (free_map) {
            std::vector<uint8_t> known(sys->n_nodes);
            MAG_CUDA(cudaMemcpyAsync(tmp.data(), sys->colmap.p, (n_dof + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            MAG_CUDA(cudaMemcpyAsync(known.data(), sys->known.p, sys->n_nodes, cudaMemcpyDeviceToHost, ctx->stream));
            MAG_CUDA(cudaStreamSynchronize(ctx->stream));
            for (size_t d = 0; d < n_dof; ++d) {
                const bool uknown = (known[d >> 1] >> (d & 1)) & 1u;
                free_map[d] = uknown ? -1 : (int64_t)tmp[d];
            }
        }
    });
}

extern "C" int mag_system_export_full(const mag_system *sys, int64_t *rowptr, int32_t *col, double *val) {
    return guarded([&] {
        if (!sys || !rowptr || !col || !val) fail(MAG_ERR_BAD_ARG, "null argument");
        mag_ctx *ctx = sys->ctx;
        CallScope scope(ctx, nullptr);
        const BsrMatrix &K = sys->K;
        const uint32_t n_own = K.node_hi - K.node_lo;
        const size_t nnz = (size_t)K.n_blocks * 4, n_rows = 2 * (size_t)n_own;
        DevBuf<int64_t> d_rowptr(ctx, n_rows + 1);
        DevBuf<int32_t> d_col(ctx, nnz);
        DevBuf<double> d_val(ctx, nnz);
        d_rowptr.zero();
        if (n_rows)
            MAG_LAUNCH(ctx, bsr_to_csr_kernel, cdiv(n_rows, 256), 256, 0, (const uint32_t *)K.browptr.p,
                       (const uint32_t *)K.bcol.p, (const double *)K.bval.p, n_own, d_rowptr.p, d_col.p, d_val.p);
        copy_from_device(ctx, rowptr, (const int64_t *)d_rowptr.p, n_rows + 1, false);
        copy_from_device(ctx, col, (const int32_t *)d_col.p, nnz, false);
        copy_from_device(ctx, val, (const double *)d_val.p, nnz, false);
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

extern "C" int mag_element_stiffness(mag_ctx *ctx, const mag_mesh *m, const mag_material *mat, double *ke) {
    return guarded([&] {
        CallScope scope(ctx, nullptr);
        check_mesh_args(m);
        if (!mat || (m->n_elems && !ke)) fail(MAG_ERR_BAD_ARG, "null argument");
        DevBuf<double2> xy; DevBuf<uint32_t> n0, n1, n2;
        upload_geometry(ctx, m, xy, n0, n1, n2);
        upload_material(ctx, *mat);
        const size_t E = m->n_elems;
        DevBuf<double> out(ctx, E * 36);
        if (E)
            MAG_LAUNCH(ctx, element_stiffness_kernel, cdiv(E, kElemThreads), kElemThreads, 0,
                       (const double2 *)xy.p, (const uint32_t *)n0.p, (const uint32_t *)n1.p,
                       (const uint32_t *)n2.p, (const uint32_t *)nullptr, E, 0, out.p);
        copy_from_device(ctx, ke, (const double *)out.p, E * 36, false);
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

extern "C" int mag_element_area(mag_ctx *ctx, const mag_mesh *m, double *area) {
    return guarded([&] {
        CallScope scope(ctx, nullptr);
        check_mesh_args(m);
        if (m->n_elems && !area) fail(MAG_ERR_BAD_ARG, "null argument");
        DevBuf<double2> xy; DevBuf<uint32_t> n0, n1, n2;
        upload_geometry(ctx, m, xy, n0, n1, n2);
        const size_t E = m->n_elems;
        DevBuf<double> out(ctx, E);
        if (E)
            MAG_LAUNCH(ctx, element_area_kernel, cdiv(E, 256), 256, 0, (const double2 *)xy.p,
                       (const uint32_t *)n0.p, (const uint32_t *)n1.p, (const uint32_t *)n2.p, E, out.p);
        copy_from_device(ctx, area, (const double *)out.p, E, false);
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

extern "C" int mag_stress(mag_ctx *ctx, const mag_mesh *m, const mag_material *mat, const double *ux,
                          const double *uy, double *stress, double *sigma) {
    return guarded([&] {
        CallScope scope(ctx, nullptr);
        check_mesh_args(m);
        if (!mat || (m->n_nodes && (!ux || !uy)) || (m->n_elems && !stress)) fail(MAG_ERR_BAD_ARG, "null argument");
        DevBuf<double2> xy; DevBuf<uint32_t> n0, n1, n2;
        upload_geometry(ctx, m, xy, n0, n1, n2);
        upload_material(ctx, *mat);
        const size_t N = m->n_nodes, E = m->n_elems;
        DevBuf<double> dux(ctx, N), duy(ctx, N), ds(ctx, E), dsig;
        copy_to_device(ctx, dux.p, ux, N, false);
        copy_to_device(ctx, duy.p, uy, N, false);
        if (sigma) dsig.alloc(ctx, E * 3);
        if (E)
            MAG_LAUNCH(ctx, stress_kernel, cdiv(E, 256), 256, 0, (const double2 *)xy.p, (const uint32_t *)n0.p,
                       (const uint32_t *)n1.p, (const uint32_t *)n2.p, E, (const double *)dux.p,
                       (const double *)duy.p, ds.p, sigma ? dsig.p : nullptr);
        copy_from_device(ctx, stress, (const double *)ds.p, E, false);
        if (sigma) copy_from_device(ctx, sigma, (const double *)dsig.p, E * 3, false);
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

// ---------------------------------------------------------------------------
// SpMV on its own
// ---------------------------------------------------------------------------
static void launch_spmv(mag_ctx *ctx, mag_system *S, int format, const double *x, double *y) {
    const CsrMatrix &A = S->Kff;
    const SellMatrix &L = S->sell;
    if (!A.n_rows) return;
    if (format == 1) {
        MAG_LAUNCH(ctx, spmv_csr_kernel, cdiv(A.n_rows, 256), 256, 0, (const uint32_t *)A.rowptr.p,
                   (const int32_t *)A.col.p, (const double *)A.val.p, x, y, A.n_rows);
    } else {
        MAG_LAUNCH(ctx, spmv_sell_kernel, sell_grid(ctx, L.n_slices), 256, 0,
                   (const uint32_t *)L.slice_off.p, (const int32_t *)L.col.p, (const double *)L.val.p, x, y,
                   L.n_rows, L.n_slices, L.row_lo);
    }
}

extern "C" int mag_system_spmv(mag_system *sys, int format, const double *x, double *y) {
    return guarded([&] {
        if (!sys || !x || !y) fail(MAG_ERR_BAD_ARG, "null argument");
        mag_ctx *ctx = sys->ctx;
        CallScope scope(ctx, nullptr);
        const size_t n = sys->Kff.n_rows;
        DevBuf<double> dx(ctx, n + 32), dy(ctx, n);
        dx.zero();
        copy_to_device(ctx, dx.p, x, n, false);
        launch_spmv(ctx, sys, format, dx.p, dy.p);
        copy_from_device(ctx, y, (const double *)dy.p, n, false);
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

extern "C" int mag_system_spmv_bench(mag_system *sys, int format, int reps, float *ms_per_spmv,
                                     uint64_t *algorithmic_bytes) {
    return guarded([&] {
        if (!sys || reps <= 0 || !ms_per_spmv) fail(MAG_ERR_BAD_ARG, "bad argument");
        mag_ctx *ctx = sys->ctx;
        CallScope scope(ctx, nullptr);
        const size_t n = sys->Kff.n_rows;
        DevBuf<double> dx(ctx, n + 32), dy(ctx, n);
        MAG_CUDA(cudaMemsetAsync(dx.p, 0, (n + 32) * sizeof(double), ctx->stream));
        copy_to_device(ctx, dx.p, (const double *)sys->rhs.p, n, true);
        // format 2 times the kernel CG actually runs (SELL SpMV + fused p.q); 3 = SELL without the dot
        const SellMatrix &L = sys->sell;
        const unsigned grid = sell_grid(ctx, L.n_slices);
        DevBuf<double> partials(ctx, 2 * (size_t)grid);
        DevBuf<PcgScalars> scal(ctx, 1);
        scal.zero();
        auto once = [&] {
            if (format == 2 && n)
                MAG_LAUNCH(ctx, pcg_spmv_kernel, grid, 256, 0, (const uint32_t *)L.slice_off.p,
                           (const int32_t *)L.col.p, (const double *)L.val.p, (const double *)dx.p, dy.p,
                           L.n_rows, L.n_slices, L.row_lo, partials.p, scal.p);
            else
                launch_spmv(ctx, sys, format, dx.p, dy.p);
        };
        for (int i = 0; i < 3; ++i) once();
        EventTimer t(ctx->stream);
        t.start();
        for (int i = 0; i < reps; ++i) once();
        *ms_per_spmv = t.stop() / (float)reps;
        if (algorithmic_bytes) {
            if (format == 1) *algorithmic_bytes = sys->stats.spmv_bytes;
            else *algorithmic_bytes = sys->sell.entries * 12ull + (uint64_t)n * 16ull +
                                      ((uint64_t)sys->sell.n_slices + 1) * 4ull;
        }
    });
}

// ---------------------------------------------------------------------------
// synthetic device meshes
// ---------------------------------------------------------------------------
extern "C" int mag_devmesh_plate(mag_ctx *ctx, uint32_t nx, uint32_t ny, double h, double ux_right,
                                 mag_devmesh **out) {
    return guarded([&] {
        if (!out || nx == 0 || ny == 0) fail(MAG_ERR_BAD_ARG, "bad argument");
        *out = nullptr;
        CallScope scope(ctx, nullptr);
        std::unique_ptr<mag_devmesh> dm(new mag_devmesh);
        dm->ctx = ctx;
        const size_t N = (size_t)(nx + 1) * (ny + 1), E = 2 * (size_t)nx * ny;
        dm->n_nodes = N; dm->n_elems = E;
        dm->x.alloc(ctx, N); dm->y.alloc(ctx, N);
        dm->ux.alloc(ctx, N); dm->uy.alloc(ctx, N); dm->fx.alloc(ctx, N); dm->fy.alloc(ctx, N);
        dm->known.alloc(ctx, N);
        dm->n0.alloc(ctx, E); dm->n1.alloc(ctx, E); dm->n2.alloc(ctx, E);
        MAG_LAUNCH(ctx, plate_nodes_kernel, cdiv(N, 256), 256, 0, nx, ny, h, ux_right, dm->x.p, dm->y.p,
                   dm->ux.p, dm->uy.p, dm->fx.p, dm->fy.p, dm->known.p);
        MAG_LAUNCH(ctx, plate_elems_kernel, cdiv((size_t)nx * ny, 256), 256, 0, nx, ny, dm->n0.p, dm->n1.p,
                   dm->n2.p);
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        *out = dm.release();
    });
}

extern "C" int mag_devmesh_view(const mag_devmesh *dm, mag_mesh *v) {
    return guarded([&] {
        if (!dm || !v) fail(MAG_ERR_BAD_ARG, "null argument");
        v->n_nodes = dm->n_nodes; v->n_elems = dm->n_elems;
        v->x = dm->x.p; v->y = dm->y.p;
        v->n0 = dm->n0.p; v->n1 = dm->n1.p; v->n2 = dm->n2.p;
        v->ux = dm->ux.p; v->uy = dm->uy.p; v->fx = dm->fx.p; v->fy = dm->fy.p;
        v->known = dm->known.p;
        v->on_device = 1;
    });
}

extern "C" void mag_devmesh_free(mag_devmesh *dm) {
    if (!dm) return;
    mag_ctx *ctx = dm->ctx;
    if (ctx) {
        cudaSetDevice(ctx->device);
        ctx->stream = ctx->own_stream;
    }
    delete dm;
    if (ctx) cudaStreamSynchronize(ctx->own_stream);
}

// ---------------------------------------------------------------------------
// debug entry points (GPU unit tests of the building blocks)
// ---------------------------------------------------------------------------
extern "C" int mag_debug_sort_pairs(mag_ctx *ctx, uint64_t *keys, uint32_t *payload, uint64_t n, int key_bits) {
    return guarded([&] {
        CallScope scope(ctx, nullptr);
        if (n && (!keys || !payload)) fail(MAG_ERR_BAD_ARG, "null argument");
        DevBuf<uint64_t> k(ctx, n), ka(ctx, n);
        DevBuf<uint32_t> p(ctx, n), pa(ctx, n);
        copy_to_device(ctx, k.p, (const uint64_t *)keys, n, false);
        copy_to_device(ctx, p.p, (const uint32_t *)payload, n, false);
        radix_sort_pairs(ctx, k.p, p.p, ka.p, pa.p, n, key_bits);
        copy_from_device(ctx, keys, (const uint64_t *)k.p, n, false);
        copy_from_device(ctx, payload, (const uint32_t *)p.p, n, false);
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

extern "C" int mag_debug_exclusive_scan(mag_ctx *ctx, const uint32_t *in, uint32_t *out, uint64_t n) {
    return guarded([&] {
        CallScope scope(ctx, nullptr);
        if (!out || (n && !in)) fail(MAG_ERR_BAD_ARG, "null argument");
        DevBuf<uint32_t> d(ctx, n + 1);
        copy_to_device(ctx, d.p, in, n, false);
        exclusive_scan_u32(ctx, d.p, n, d.p, n + 1);
        copy_from_device(ctx, out, (const uint32_t *)d.p, n + 1, false);
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

// ---------------------------------------------------------------------------
// partition helper (host logic, no device needed)
// ---------------------------------------------------------------------------
extern "C" int mag_partition_nodes(uint64_t n_nodes, int nranks, int rank, uint64_t *lo, uint64_t *hi) {
    return guarded([&] {
        if (nranks <= 0 || rank < 0 || rank >= nranks || !lo || !hi) fail(MAG_ERR_BAD_ARG, "bad partition request");
        const uint64_t base = n_nodes / (uint64_t)nranks, rem = n_nodes % (uint64_t)nranks;
        const uint64_t r = (uint64_t)rank;
        *lo = r * base + std::min(r, rem);
        *hi = *lo + base + (r < rem ? 1 : 0);
    });
}

// multi-GPU entry points: dist.cu
