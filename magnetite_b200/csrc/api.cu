// api.cu — the extern "C" boundary of libmagnetite_b200.so and the host-side
// orchestration of the device pipeline.  See include/magnetite_b200.h for the
// contract and the reference lines each entry point replaces.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "assembly.cuh"
#include "bc.cuh"
#include "comm.cuh"
#include "common.cuh"
#include "element.cuh"
#include "meshgen.cuh"
#include "pcg.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "solve.cuh"
#include "spmv.cuh"
#include "system.cuh"

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
        MAG_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
        MAG_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        MAG_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        MAG_CUDA(cudaMallocHost((void **)&c->h_scal, 128 * sizeof(double)));
        if (const char *t = std::getenv("MAG_TUNE")) c->tune = std::atoi(t);
        *out = c.release();
    });
}

extern "C" void mag_ctx_destroy(mag_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->own_stream);
    if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
    if (ctx->comm) {
        Comm *c = ctx->comm;
        // every rank unmaps its peers' slabs, then all ranks meet, and only then does anyone free the buffer it
        // exported (freeing memory another process still maps is undefined)
        for (int r = 0; r < (int)c->peer_slab.size(); ++r)
            if (r != c->rank && c->peer_slab[r]) cudaIpcCloseMemHandle(c->peer_slab[r]);
        if (c->slab && c->nccl) {
            int *token = nullptr;
            if (cudaMalloc((void **)&token, sizeof(int)) == cudaSuccess) {
                cudaMemset(token, 0, sizeof(int));
                if (ncclAllReduce(token, token, 1, ncclInt, ncclSum, c->nccl, ctx->own_stream) == ncclSuccess)
                    cudaStreamSynchronize(ctx->own_stream);
                cudaFree(token);
            }
        }
        if (c->slab) cudaFree(c->slab);
        if (c->nccl) ncclCommDestroy(c->nccl);
        delete c;
    }
    ctx->heap.destroy();
    if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    delete ctx;
}

extern "C" void mag_options_default(mag_options *o) {
    if (!o) return;
    std::memset(o, 0, sizeof *o);
    o->rel_tol = 1e-9;
    o->abs_tol = MAG_TARGET_CG_COST;
    o->max_iter = MAG_MAX_CG_ITER;
    o->precond = 3;
    o->compat = 0;
    o->cost_kind = 0;
    o->drop_exact_zeros = 1;
    o->check_every = 50;
    o->spmv_format = 0;
    o->want_sigma = 0;
    o->stream = nullptr;
}

// ---------------------------------------------------------------------------
// assembly
// ---------------------------------------------------------------------------
extern "C" int mag_assemble(mag_ctx *ctx, const mag_mesh *mesh, const mag_material *mat,
                            const mag_options *opt, mag_system **sys, mag_stats *stats) {
    return guarded([&] {
        if (!sys) fail(MAG_ERR_BAD_ARG, "null sys out");
        *sys = nullptr;
        CallScope scope(ctx, opt);
        std::unique_ptr<mag_system> S(new mag_system);
        try {
            const int rank = ctx->comm ? ctx->comm->rank : 0, nranks = ctx->comm ? ctx->comm->nranks : 1;
            assemble_impl(ctx, mesh, mat, opt, S.get(), rank, nranks);
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
    // (the halo slab and its IPC mappings belong to the communicator: nothing collective happens here)
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

// B (3x6 row-major) of every element: solver::compute_strain_displacement_matrix (solver.rs:204-230).
extern "C" int mag_strain_displacement(mag_ctx *ctx, const mag_mesh *m, double *b_out) {
    return guarded([&] {
        CallScope scope(ctx, nullptr);
        check_mesh_args(m);
        if (m->n_elems && !b_out) fail(MAG_ERR_BAD_ARG, "null argument");
        DevBuf<double2> xy; DevBuf<uint32_t> n0, n1, n2;
        upload_geometry(ctx, m, xy, n0, n1, n2);
        const size_t E = m->n_elems;
        DevBuf<double> out(ctx, E * 18);
        if (E)
            MAG_LAUNCH(ctx, strain_displacement_kernel, cdiv(E, 256), 256, 0, (const double2 *)xy.p,
                       (const uint32_t *)n0.p, (const uint32_t *)n1.p, (const uint32_t *)n2.p, E, out.p);
        copy_from_device(ctx, b_out, (const double *)out.p, E * 18, false);
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

// D (3x3 row-major): solver::compute_stress_strain_matrix (solver.rs:240-250) — the constant the
// kernels keep in __constant__ memory.  Host arithmetic, no device needed.
extern "C" int mag_stress_strain(double poisson_ratio, double youngs_modulus, double *d_out) {
    return guarded([&] {
        if (!d_out) fail(MAG_ERR_BAD_ARG, "null argument");
        const mag_material m{youngs_modulus, poisson_ratio, 1.0};
        const MaterialConst c = make_material(m);
        for (int i = 0; i < 9; ++i) d_out[i] = c.D[i];
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
    } else if (L.narrow) {
        MAG_LAUNCH(ctx, spmv_sell_kernel<int16_t>, sell_grid(ctx, L.n_slices), 256, 0,
                   (const uint32_t *)L.slice_off.p, (const int16_t *)L.pcol.p, (const double *)L.val.p, x, y,
                   L.n_rows, L.n_slices, L.row_lo);
    } else {
        MAG_LAUNCH(ctx, spmv_sell_kernel<int32_t>, sell_grid(ctx, L.n_slices), 256, 0,
                   (const uint32_t *)L.slice_off.p, (const int32_t *)L.col.p, (const double *)L.val.p, x, y,
                   L.n_rows, L.n_slices, L.row_lo);
    }
}

extern "C" int mag_system_spmv(mag_system *sys, int format, const double *x, double *y) {
    return guarded([&] {
        if (!sys || !x || !y) fail(MAG_ERR_BAD_ARG, "null argument");
        mag_ctx *ctx = sys->ctx;
        CallScope scope(ctx, nullptr);
        const size_t n = sys->Kff.n_rows, nx = sys->n_free;    // x: all unknowns, y: owned rows
        DevBuf<double> dx(ctx, nx + 32), dy(ctx, n);
        dx.zero();
        copy_to_device(ctx, dx.p, x, nx, false);
        launch_spmv(ctx, sys, format, dx.p, dy.p);
        copy_from_device(ctx, y, (const double *)dy.p, n, false);
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

// ||b - K_ff x||^2 and ||b||^2 over the rows this rank owns, for the displacement field (ux, uy) a solve
// returned: x is rebuilt from the nodal values through the column map, the products run through the
// ordered CSR kernel.  Not collective: a multi-GPU caller adds the two sums over the ranks.
extern "C" int mag_system_residual(mag_system *sys, const double *ux, const double *uy, int on_device,
                                   double *rr_owned, double *bb_owned) {
    return guarded([&] {
        if (!sys || !rr_owned || !bb_owned || (sys->n_nodes && (!ux || !uy))) fail(MAG_ERR_BAD_ARG, "null argument");
        mag_ctx *ctx = sys->ctx;
        CallScope scope(ctx, nullptr);
        const size_t N = sys->n_nodes, n = sys->Kff.n_rows, nx = sys->n_free;
        DevBuf<double> dux, duy, dx(ctx, nx + 32);
        const double *pux = ux, *puy = uy;
        if (!on_device) {
            dux.alloc(ctx, N); duy.alloc(ctx, N);
            copy_to_device(ctx, dux.p, ux, N, false);
            copy_to_device(ctx, duy.p, uy, N, false);
            pux = dux.p; puy = duy.p;
        }
        dx.zero();
        if (N)
            MAG_LAUNCH(ctx, gather_solution_kernel, cdiv(N, 256), 256, 0, (const uint8_t *)sys->known.p,
                       (const uint32_t *)sys->colmap.p, pux, puy, N, dx.p);
        const unsigned grid = std::max(1u, std::min(cdiv(n, 256), (unsigned)ctx->sm_count * 8u));
        DevBuf<double> partials(ctx, 2 * (size_t)grid), out(ctx, 2);
        DevBuf<unsigned> ticket(ctx, 1);
        ticket.zero();
        out.zero();
        const CsrMatrix &A = sys->Kff;
        MAG_LAUNCH(ctx, true_residual_kernel, grid, 256, 0, (const uint32_t *)A.rowptr.p, (const int32_t *)A.col.p,
                   (const double *)A.val.p, (const double *)dx.p, (const double *)sys->rhs.p, A.n_rows, partials.p,
                   ticket.p, out.p);
        double h[2] = {0.0, 0.0};
        MAG_CUDA(cudaMemcpyAsync(h, out.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        *rr_owned = h[0];
        *bb_owned = h[1];
    });
}

extern "C" int mag_system_spmv_bench(mag_system *sys, int format, int reps, float *ms_per_spmv,
                                     uint64_t *algorithmic_bytes) {
    return guarded([&] {
        if (!sys || reps <= 0 || !ms_per_spmv) fail(MAG_ERR_BAD_ARG, "bad argument");
        mag_ctx *ctx = sys->ctx;
        CallScope scope(ctx, nullptr);
        const size_t n = sys->Kff.n_rows, nx = sys->n_free;
        DevBuf<double> dx(ctx, nx + 32), dy(ctx, n);
        MAG_CUDA(cudaMemsetAsync(dx.p, 0, (nx + 32) * sizeof(double), ctx->stream));
        copy_to_device(ctx, dx.p + sys->row_lo, (const double *)sys->rhs.p, n, true);
        // format 2 times the kernel CG actually runs (SELL SpMV + fused p.q); 3 = SELL without the dot
        const SellMatrix &L = sys->sell;
        const unsigned grid = sell_grid(ctx, L.n_slices);
        DevBuf<double> partials(ctx, 2 * (size_t)grid);
        DevBuf<PcgScalars> scal(ctx, 1);
        scal.zero();
        auto once = [&] {
            if (format == 2 && n && L.narrow)
                MAG_LAUNCH(ctx, pcg_spmv_kernel<int16_t>, grid, 256, 0, (const uint32_t *)L.slice_off.p,
                           (const int16_t *)L.pcol.p, (const double *)L.val.p, (const double *)dx.p, dy.p,
                           L.n_rows, L.n_slices, L.row_lo, 0, PeerLinks{}, partials.p, scal.p, &scal.p->pq);
            else if (format == 2 && n)
                MAG_LAUNCH(ctx, pcg_spmv_kernel<int32_t>, grid, 256, 0, (const uint32_t *)L.slice_off.p,
                           (const int32_t *)L.col.p, (const double *)L.val.p, (const double *)dx.p, dy.p,
                           L.n_rows, L.n_slices, L.row_lo, 0, PeerLinks{}, partials.p, scal.p, &scal.p->pq);
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
            else *algorithmic_bytes = sys->sell.entries * 8ull + sys->sell.index_bytes() + (uint64_t)n * 16ull +
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

extern "C" int mag_devmesh_perforated(mag_ctx *ctx, uint32_t nx, uint32_t ny, double h, uint32_t pitch,
                                      uint32_t radius, double ux_right, mag_devmesh **out) {
    return guarded([&] {
        if (!out || nx == 0 || ny == 0 || pitch == 0) fail(MAG_ERR_BAD_ARG, "bad argument");
        *out = nullptr;
        CallScope scope(ctx, nullptr);
        std::unique_ptr<mag_devmesh> dm(new mag_devmesh);
        dm->ctx = ctx;
        const size_t Ng = (size_t)(nx + 1) * (ny + 1), Cg = (size_t)nx * ny;
        if (Ng >= (1ull << 31)) fail(MAG_ERR_BAD_ARG, "grid too large");
        const double P = (double)pitch * h, R = (double)radius * h;
        DevBuf<uint32_t> cell_pos(ctx, Cg + 1), node_pos(ctx, Ng + 1);
        MAG_LAUNCH(ctx, perforated_flags_kernel, cdiv(Ng, 256), 256, 0, nx, ny, h, P, R * R, cell_pos.p, node_pos.p);
        exclusive_scan_u32(ctx, cell_pos.p, Cg, cell_pos.p, Cg + 1);
        exclusive_scan_u32(ctx, node_pos.p, Ng, node_pos.p, Ng + 1);
        const size_t N = read_u32(ctx, node_pos.p + Ng), E = 2 * (size_t)read_u32(ctx, cell_pos.p + Cg);
        dm->n_nodes = N; dm->n_elems = E;
        dm->x.alloc(ctx, N); dm->y.alloc(ctx, N);
        dm->ux.alloc(ctx, N); dm->uy.alloc(ctx, N); dm->fx.alloc(ctx, N); dm->fy.alloc(ctx, N);
        dm->known.alloc(ctx, N);
        dm->n0.alloc(ctx, E); dm->n1.alloc(ctx, E); dm->n2.alloc(ctx, E);
        MAG_LAUNCH(ctx, perforated_nodes_kernel, cdiv(Ng, 256), 256, 0, nx, ny, h, ux_right,
                   (const uint32_t *)node_pos.p, dm->x.p, dm->y.p, dm->ux.p, dm->uy.p, dm->fx.p, dm->fy.p, dm->known.p);
        MAG_LAUNCH(ctx, perforated_elems_kernel, cdiv(Cg, 256), 256, 0, nx, ny, (const uint32_t *)cell_pos.p,
                   (const uint32_t *)node_pos.p, dm->n0.p, dm->n1.p, dm->n2.p);
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        *out = dm.release();
    });
}

// Copies a device mesh into caller-allocated host arrays (sizes from mag_devmesh_view).
extern "C" int mag_devmesh_download(const mag_devmesh *dm, double *x, double *y, uint32_t *n0, uint32_t *n1,
                                    uint32_t *n2, double *ux, double *uy, double *fx, double *fy, uint8_t *known) {
    return guarded([&] {
        if (!dm) fail(MAG_ERR_BAD_ARG, "null mesh");
        mag_ctx *ctx = dm->ctx;
        CallScope scope(ctx, nullptr);
        const size_t N = dm->n_nodes, E = dm->n_elems;
        if (x) copy_from_device(ctx, x, (const double *)dm->x.p, N, false);
        if (y) copy_from_device(ctx, y, (const double *)dm->y.p, N, false);
        if (n0) copy_from_device(ctx, n0, (const uint32_t *)dm->n0.p, E, false);
        if (n1) copy_from_device(ctx, n1, (const uint32_t *)dm->n1.p, E, false);
        if (n2) copy_from_device(ctx, n2, (const uint32_t *)dm->n2.p, E, false);
        if (ux) copy_from_device(ctx, ux, (const double *)dm->ux.p, N, false);
        if (uy) copy_from_device(ctx, uy, (const double *)dm->uy.p, N, false);
        if (fx) copy_from_device(ctx, fx, (const double *)dm->fx.p, N, false);
        if (fy) copy_from_device(ctx, fy, (const double *)dm->fy.p, N, false);
        if (known) copy_from_device(ctx, known, (const uint8_t *)dm->known.p, N, false);
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
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
        partition_nodes(n_nodes, nranks, rank, lo, hi);
    });
}

extern "C" int mag_halo_plan(int nranks, int rank, const uint32_t *row_lo, const uint32_t *ext_lo,
                             const uint32_t *ext_hi, uint32_t *seg_lo, uint32_t *seg_hi, int32_t *seg_dst,
                             int32_t capacity, int32_t *n_segs) {
    return guarded([&] {
        if (nranks <= 0 || rank < 0 || rank >= nranks || !row_lo || !ext_lo || !ext_hi || !n_segs)
            fail(MAG_ERR_BAD_ARG, "bad halo plan request");
        const std::vector<HaloSeg> segs = halo_plan(nranks, rank, row_lo, ext_lo, ext_hi);
        *n_segs = (int32_t)segs.size();
        if ((int32_t)segs.size() > capacity) fail(MAG_ERR_BAD_ARG, "halo plan needs %zu segments", segs.size());
        for (size_t i = 0; i < segs.size(); ++i) {
            seg_lo[i] = segs[i].lo; seg_hi[i] = segs[i].hi; seg_dst[i] = segs[i].dst;
        }
    });
}

// ---------------------------------------------------------------------------
// multi-GPU: one process per GPU
// ---------------------------------------------------------------------------
extern "C" int mag_comm_unique_id(void *id128) {
    return guarded([&] {
        if (!id128) fail(MAG_ERR_BAD_ARG, "null id");
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
        ncclUniqueId id;
        MAG_NCCL(ncclGetUniqueId(&id));
        std::memcpy(id128, &id, sizeof id);
    });
}

extern "C" int mag_comm_init(mag_ctx *ctx, int rank, int nranks, const void *id128) {
    return guarded([&] {
        if (!ctx || !id128 || nranks <= 0 || rank < 0 || rank >= nranks) fail(MAG_ERR_BAD_ARG, "bad communicator request");
        if (ctx->comm) fail(MAG_ERR_BAD_ARG, "the context already has a communicator");
        if (nranks > kMaxRanks) fail(MAG_ERR_BAD_ARG, "at most %d ranks (mailboxes and push segments are sized for one NVSwitch box)", kMaxRanks);
        MAG_CUDA(cudaSetDevice(ctx->device));
        ncclUniqueId id;
        std::memcpy(&id, id128, sizeof id);
        std::unique_ptr<Comm> c(new Comm);
        c->rank = rank; c->nranks = nranks;
        MAG_NCCL(ncclCommInitRank(&c->nccl, nranks, id, rank));
        ctx->comm = c.release();
        // NCCL builds its channels on the first collective of each kind (seconds at 8 ranks): pay for that
        // here, once per process, not inside the first solve.
        {
            CallScope scope(ctx, nullptr);
            DevBuf<double> warm(ctx, (size_t)1 << 18);      // 2 MB
            warm.zero();
            MAG_NCCL(ncclAllReduce(warm.p, warm.p, 1, ncclDouble, ncclSum, ctx->comm->nccl, ctx->stream));
            MAG_NCCL(ncclAllReduce(warm.p, warm.p, (size_t)1 << 18, ncclDouble, ncclSum, ctx->comm->nccl, ctx->stream));
            MAG_NCCL(ncclBroadcast(warm.p, warm.p, (size_t)1 << 18, ncclDouble, 0, ctx->comm->nccl, ctx->stream));
            MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    });
}

extern "C" int mag_comm_rank(const mag_ctx *ctx, int *rank, int *nranks) {
    return guarded([&] {
        if (!ctx || !rank || !nranks) fail(MAG_ERR_BAD_ARG, "null argument");
        *rank = ctx->comm ? ctx->comm->rank : 0;
        *nranks = ctx->comm ? ctx->comm->nranks : 1;
    });
}

// Single-GPU emulation of an nranks-way row-block solve (same partition code, same
// kernels and halo stores; the allreduce is a tiny kernel).  Used by the GPU tests.
extern "C" int mag_debug_virtual_solve(mag_ctx *ctx, const mag_mesh *mesh, const mag_material *mat,
                                       const mag_options *opt, int nranks, mag_result *out, mag_stats *stats) {
    return guarded([&] {
        CallScope scope(ctx, opt);
        try {
            virtual_solve_impl(ctx, mesh, mat, opt, nranks, out, stats);
        } catch (...) {
            cudaStreamSynchronize(ctx->stream);
            throw;
        }
    });
}
