// system.cuh — mag_system (everything an assembly leaves in HBM) and the host
// orchestration of the assembly pipeline for one rank's row block.
//
// Partitioning (multi-GPU): rank g owns the contiguous node range
// [node_lo, node_hi) and therefore the DOFs 2*node+axis of those nodes and the
// reduced rows [row_lo, row_hi).  It computes K_e for every element touching an
// owned node (boundary elements are computed on both sides) and emits only the
// keys whose row node it owns, so assembly needs no communication and keeps the
// ascending-element accumulation order.  Node ids, DOF ids and reduced column
// ids stay GLOBAL everywhere (solver.rs:306-309 numbering at the boundary).
#pragma once
#include <cstring>
#include <memory>

#include "assembly.cuh"
#include "bc.cuh"
#include "coarse.cuh"
#include "comm.cuh"
#include "common.cuh"
#include "element.cuh"
#include "gather.cuh"
#include "pcg.cuh"
#include "spmv.cuh"

struct mag_system {
    mag_ctx *ctx = nullptr;
    int rank = 0, nranks = 1;
    uint64_t n_nodes = 0, n_elems = 0;
    uint32_t node_lo = 0, node_hi = 0;       // owned nodes
    uint32_t row_lo = 0, row_hi = 0;         // owned reduced rows
    uint32_t ext_lo = 0, ext_hi = 0;         // [min col, max col] referenced by the owned rows
    uint32_t n_free = 0;                     // global number of unknowns
    mag_material mat{};
    mag::DevBuf<double2> xy;
    mag::DevBuf<uint32_t> n0, n1, n2;
    mag::DevBuf<uint8_t> known;
    mag::DevBuf<double> bc_ux, bc_uy, bc_fx, bc_fy;
    mag::BsrMatrix K;                        // full K, owned node rows
    mag::DevBuf<uint32_t> rowmap, colmap;    // n_dof+1 each (global; last = total)
    mag::DevBuf<uint32_t> colid;             // n_dof: reduced column, or ~0 where the displacement is prescribed
    mag::CsrMatrix Kff;                      // owned rows x global cols
    mag::DevBuf<double> rhs, diag;           // owned rows
    mag::SellMatrix sell;
    // halo buffers other ranks store into (plain cudaMalloc: exported through CUDA IPC)
    double *shared_slab = nullptr;           // [ Dinv (global-indexed) | mailbox | coarse partials | halo buffer of r ]
    bool owns_slab = false;                  // virtual ranks own theirs; production systems borrow the communicator's
    unsigned long long slab_generation = 0;  // Comm::generation the pointers below were taken from
    mag::CoarseSpace coarse;                 // two-level preconditioner (built on first use)
    mag::PushSegs push;
    mag::PeerLinks links;                    // peer mailboxes (production multi-rank only)
    unsigned long long solve_epoch = 0;
    bool push_ready = false;
    bool bc_paired = true;                   // every DOF has exactly one of {displacement, force} known
    std::vector<uint32_t> all_row_lo, all_node_lo;   // nranks+1
    mag_stats stats{};

    ~mag_system() {
        mag::aux_drain(ctx);                 // side-stream copies into / out of this system's buffers (error paths)
        if (shared_slab && owns_slab) cudaFree(shared_slab);
    }
};

namespace mag {

static void check_mesh_args(const mag_mesh *m) {
    if (!m) fail(MAG_ERR_BAD_ARG, "null mesh");
    if (m->n_nodes >= (1ull << 31)) fail(MAG_ERR_BAD_ARG, "n_nodes must be < 2^31");
    if (m->n_elems >= (1ull << 32)) fail(MAG_ERR_BAD_ARG, "n_elems must be < 2^32");
    if (m->n_nodes && (!m->x || !m->y || !m->known)) fail(MAG_ERR_BAD_ARG, "mesh: x, y and known are required");
    if (m->n_elems && (!m->n0 || !m->n1 || !m->n2)) fail(MAG_ERR_BAD_ARG, "mesh: n0, n1, n2 are required");
}

template <class T>
static void upload_or_zero(mag_ctx *ctx, DevBuf<T> &dst, const T *src, size_t n, bool on_device) {
    dst.alloc(ctx, n);
    if (src) copy_to_device(ctx, dst.p, src, n, on_device);
    else dst.zero();
}

// Multi-GPU upload of a HOST array every rank holds: rank r copies only its 1/R slice over PCIe and the ranks
// exchange the slices over NVLink (in-place NCCL allgather), so the job moves the array over PCIe ONCE instead
// of once per rank (8 ranks x 584 MB through one host's memory cost 60 ms per step at 16 M DOF).  Collective.
template <class T>
static void upload_shared(mag_ctx *ctx, DevBuf<T> &dst, const T *src, size_t n, bool on_device, bool split) {
    Comm *c = ctx->comm;
    if (!split || !src || on_device || !c || c->nranks == 1 || n < (size_t)c->nranks * 1024) {
        upload_or_zero(ctx, dst, src, n, on_device);
        return;
    }
    const size_t R = (size_t)c->nranks, chunk = (n + R - 1) / R;
    dst.alloc(ctx, chunk * R);                                   // padded: equal slices for the allgather
    dst.n = n;
    const size_t lo = std::min(n, chunk * (size_t)c->rank), hi = std::min(n, lo + chunk);
    if (hi > lo) copy_to_device(ctx, dst.p + lo, src + lo, hi - lo, false);
    MAG_NCCL(ncclAllGather(dst.p + chunk * (size_t)c->rank, dst.p, chunk * sizeof(T), ncclChar, c->nccl, ctx->stream));
}

static uint32_t read_u32(mag_ctx *ctx, const uint32_t *dptr) {
    uint32_t v = 0;
    MAG_CUDA(cudaMemcpyAsync(&v, dptr, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    return v;
}

// geometry + connectivity only (enough for K_e, area, stress)
// split: every rank of the communicator makes this call with the same host mesh (mag_assemble): upload_shared.
static void upload_geometry(mag_ctx *ctx, const mag_mesh *m, DevBuf<double2> &xy, DevBuf<uint32_t> &n0,
                            DevBuf<uint32_t> &n1, DevBuf<uint32_t> &n2, bool split = false) {
    const size_t N = m->n_nodes, E = m->n_elems;
    const bool dev = m->on_device != 0;
    xy.alloc(ctx, N);
    if (N) {
        if (dev) {
            MAG_LAUNCH(ctx, pack_xy_kernel, cdiv(N, 256), 256, 0, m->x, m->y, xy.p, N);
        } else {
            DevBuf<double> tx, ty;
            upload_shared(ctx, tx, m->x, N, false, split);
            upload_shared(ctx, ty, m->y, N, false, split);
            MAG_LAUNCH(ctx, pack_xy_kernel, cdiv(N, 256), 256, 0, (const double *)tx.p,
                       (const double *)ty.p, xy.p, N);
        }
    }
    upload_shared(ctx, n0, m->n0, E, dev, split);
    upload_shared(ctx, n1, m->n1, E, dev, split);
    upload_shared(ctx, n2, m->n2, E, dev, split);
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

// flag[e] = 1 iff element e touches a node of [lo, hi)
__global__ void flag_elements_kernel(const uint32_t *__restrict__ n0, const uint32_t *__restrict__ n1,
                                     const uint32_t *__restrict__ n2, size_t n_elems, uint32_t lo,
                                     uint32_t hi, uint32_t *__restrict__ flag) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    const uint32_t a = n0[e], b = n1[e], c = n2[e];
    flag[e] = ((a >= lo && a < hi) || (b >= lo && b < hi) || (c >= lo && c < hi)) ? 1u : 0u;
}
// elist[pos[e]] = e for flagged elements (pos = exclusive scan of flag): ascending element ids
__global__ void compact_elements_kernel(const uint32_t *__restrict__ pos, size_t n_elems,
                                        uint32_t *__restrict__ elist) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_elems) return;
    if (pos[e + 1] != pos[e]) elist[pos[e]] = (uint32_t)e;
}

inline void partition_nodes(uint64_t n_nodes, int nranks, int rank, uint64_t *lo, uint64_t *hi) {
    const uint64_t base = n_nodes / (uint64_t)nranks, rem = n_nodes % (uint64_t)nranks;
    const uint64_t r = (uint64_t)rank;
    *lo = r * base + std::min(r, rem);
    *hi = *lo + base + (r < rem ? 1 : 0);
}

// Assembles the row block of `rank` (of `nranks`) into S.  No communication.
static void assemble_impl(mag_ctx *ctx, const mag_mesh *m, const mag_material *mat,
                          const mag_options *opt, mag_system *S, int rank, int nranks) {
    check_mesh_args(m);
    if (!mat) fail(MAG_ERR_BAD_ARG, "null material");
    const size_t N = m->n_nodes, E = m->n_elems, n_dof = 2 * N;
    const bool dev = m->on_device != 0;
    mag_stats &st = S->stats;
    std::memset(&st, 0, sizeof st);
    st.n_nodes = N; st.n_elems = E; st.n_dof = n_dof;
    S->ctx = ctx; S->n_nodes = N; S->n_elems = E; S->mat = *mat;
    S->rank = rank; S->nranks = nranks;
    S->all_node_lo.resize(nranks + 1);
    for (int r = 0; r <= nranks; ++r) {
        uint64_t lo, hi;
        if (r < nranks) partition_nodes(N, nranks, r, &lo, &hi); else lo = N;
        S->all_node_lo[r] = (uint32_t)lo;
    }
    S->node_lo = S->all_node_lo[rank]; S->node_hi = S->all_node_lo[rank + 1];
    EventTimer total(ctx->stream), phase(ctx->stream);
    total.start();

    // ---- upload -------------------------------------------------------------
    phase.start();
    // production multi-rank (this process is one rank of the communicator): every array crosses PCIe once per job
    const bool split = ctx->comm && ctx->comm->nranks == nranks && nranks > 1 && !(ctx->tune & 256);
    upload_geometry(ctx, m, S->xy, S->n0, S->n1, S->n2, split);
    upload_shared(ctx, S->known, m->known, N, dev, split);
    // One rank, host arrays: the prescribed values (4 x 8 bytes per node, 45 % of the upload) are first needed by the
    // Dirichlet elimination, so they cross PCIe on the side stream while the element / sort / gather kernels run.
    const bool side = !dev && !split && N >= (1u << 16) && !(ctx->tune & 2048);
    if (side) {
        const double *src[4] = {m->ux, m->uy, m->fx, m->fy};
        DevBuf<double> *dst[4] = {&S->bc_ux, &S->bc_uy, &S->bc_fx, &S->bc_fy};
        for (int a = 0; a < 4; ++a) {
            dst[a]->alloc(ctx, N);
            if (!src[a]) dst[a]->zero();
        }
        aux_fork(ctx);
        for (int a = 0; a < 4; ++a)
            if (src[a]) MAG_CUDA(cudaMemcpyAsync(dst[a]->p, src[a], N * sizeof(double), cudaMemcpyHostToDevice, ctx->aux_stream));
    } else {
        upload_shared(ctx, S->bc_ux, m->ux, N, dev, split);
        upload_shared(ctx, S->bc_uy, m->uy, N, dev, split);
        upload_shared(ctx, S->bc_fx, m->fx, N, dev, split);
        upload_shared(ctx, S->bc_fy, m->fy, N, dev, split);
    }
    upload_material(ctx, *mat);
    st.ms_upload = phase.stop();

    // ---- element stiffness (solver.rs:553-563) for the elements this rank needs ----
    phase.start();
    DevBuf<uint32_t> elist;
    size_t El = E;
    if (nranks > 1 && E) {
        DevBuf<uint32_t> flag(ctx, E + 1);
        MAG_LAUNCH(ctx, flag_elements_kernel, cdiv(E, 256), 256, 0, (const uint32_t *)S->n0.p,
                   (const uint32_t *)S->n1.p, (const uint32_t *)S->n2.p, E, S->node_lo, S->node_hi, flag.p);
        exclusive_scan_u32(ctx, flag.p, E, flag.p, E + 1);
        El = read_u32(ctx, flag.p + E);
        elist.alloc(ctx, El);
        MAG_LAUNCH(ctx, compact_elements_kernel, cdiv(E, 256), 256, 0, (const uint32_t *)flag.p, E, elist.p);
    }
    if ((uint64_t)El * 9 >= (1ull << 32))
        fail(MAG_ERR_BAD_ARG, "this rank would assemble %zu elements; 9 COO keys each must stay below 2^32: use more GPUs", El);
    const uint32_t *elist_p = (nranks > 1) ? elist.p : nullptr;
    BsrMatrix &K = S->K;
    K.node_lo = S->node_lo; K.node_hi = S->node_hi;
    const uint32_t n_own = K.node_hi - K.node_lo;
    const int asm_mode = opt ? opt->assembly : 0;       // 0 (default) gather, 1 sorted COO keys
    if (asm_mode < 0 || asm_mode > 1) fail(MAG_ERR_BAD_ARG, "mag_options.assembly must be 0 (gather) or 1 (sorted COO keys)");
    const bool use_gather = asm_mode == 0;
    if (use_gather) {
        // gather assembly (gather.cuh): no K_e in memory, 3E incidences sorted by node instead of 9E COO keys
        st.ms_elem = phase.stop();                  // only the rank's element list: K_e rows are recomputed in the gather
        assemble_gather(ctx, S->xy, S->n0, S->n1, S->n2, elist_p, El, N, K, &st.ms_sort, &st.ms_reduce);
        elist.release();
    } else {
        DevBuf<double> kblk(ctx, El * 36);
        if (El)
            MAG_LAUNCH(ctx, element_stiffness_kernel, cdiv(El, kElemThreads), kElemThreads, 0,
                       (const double2 *)S->xy.p, (const uint32_t *)S->n0.p, (const uint32_t *)S->n1.p,
                       (const uint32_t *)S->n2.p, elist_p, El, 1, kblk.p);
        st.ms_elem = phase.stop();

        // ---- COO keys + stable sort ----------------------------------------------
        phase.start();
        const size_t n_keys = El * 9;
        const int bits = bits_for(N + 1);
        DevBuf<uint64_t> keys(ctx, n_keys), keys_alt(ctx, n_keys);
        DevBuf<uint32_t> pay(ctx, n_keys), pay_alt(ctx, n_keys);
        if (El) {
            MAG_LAUNCH(ctx, emit_keys_kernel, cdiv(El, 256), 256, 0, (const uint32_t *)S->n0.p,
                       (const uint32_t *)S->n1.p, (const uint32_t *)S->n2.p, elist_p, El, bits, S->node_lo,
                       S->node_hi, keys.p, pay.p);
            radix_sort_pairs(ctx, keys.p, pay.p, keys_alt.p, pay_alt.p, n_keys, 2 * bits);
        }
        keys_alt.release();
        pay_alt.release();
        elist.release();
        st.ms_sort = phase.stop();

        // ---- segmented reduction into BSR ------------------------------------------
        phase.start();
        K.browptr.alloc(ctx, (size_t)n_own + 1);
        K.browptr.zero();
        {
            DevBuf<uint32_t> head(ctx, n_keys + 1);
            if (n_keys)
                MAG_LAUNCH(ctx, mark_heads_kernel, cdiv(n_keys, 256), 256, 0, (const uint64_t *)keys.p,
                           n_keys, bits, K.node_lo, head.p, K.browptr.p);
            DevBuf<uint32_t> uid(ctx, n_keys + 1);      // exclusive scan of head; uid[n_keys] = #blocks
            exclusive_scan_u32(ctx, head.p, n_keys, uid.p, n_keys + 1);
            exclusive_scan_u32(ctx, K.browptr.p, n_own, K.browptr.p, (size_t)n_own + 1);
            K.n_blocks = read_u32(ctx, uid.p + n_keys);
            K.brow.alloc(ctx, K.n_blocks);
            K.bcol.alloc(ctx, K.n_blocks);
            K.bval.alloc(ctx, (size_t)K.n_blocks * 4);
            if (n_keys)
                MAG_LAUNCH(ctx, segment_reduce_kernel, cdiv(n_keys, 256), 256, 0, (const uint64_t *)keys.p,
                           (const uint32_t *)pay.p, (const uint32_t *)head.p, (const uint32_t *)uid.p,
                           n_keys, bits, K.node_lo, (const double *)kblk.p, K.brow.p, K.bcol.p, K.bval.p);
        }
        keys.release();
        pay.release();
        kblk.release();
        st.ms_reduce = phase.stop();
    }
    st.nnz_structural = (uint64_t)K.n_blocks * 4;

    // ---- Dirichlet elimination (solver.rs:340-432, 126-137) --------------------
    phase.start();
    if (side) aux_join(ctx);                 // the prescribed values have arrived
    S->rowmap.alloc(ctx, n_dof + 1);
    S->colmap.alloc(ctx, n_dof + 1);
    DevBuf<int> unpaired(ctx, 1);
    unpaired.zero();
    if (n_dof)
        MAG_LAUNCH(ctx, dof_flags_kernel, cdiv(n_dof, 256), 256, 0, (const uint8_t *)S->known.p, n_dof,
                   S->rowmap.p, S->colmap.p, unpaired.p);
    {
        int h_unpaired = 0;
        MAG_CUDA(cudaMemcpyAsync(&h_unpaired, unpaired.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        S->bc_paired = h_unpaired == 0;
    }
    exclusive_scan_u32(ctx, S->rowmap.p, n_dof, S->rowmap.p, n_dof + 1);
    exclusive_scan_u32(ctx, S->colmap.p, n_dof, S->colmap.p, n_dof + 1);
    const uint32_t n_rows_glob = read_u32(ctx, S->rowmap.p + n_dof);
    const uint32_t n_cols = read_u32(ctx, S->colmap.p + n_dof);
    if (n_rows_glob != n_cols)
        fail(MAG_ERR_BAD_BC,
             "inconsistent boundary conditions: %u DOFs have a known force but %u have an unknown "
             "displacement (the reference panics here, solver.rs:380-396)", n_rows_glob, n_cols);
    S->n_free = n_cols;
    st.n_free = n_cols; st.n_constrained = n_dof - n_cols;
    // reduced-row boundaries of every rank (rowmap is global, so no exchange is needed)
    S->all_row_lo.resize(nranks + 1);
    for (int r = 0; r <= nranks; ++r)
        S->all_row_lo[r] = (nranks == 1) ? (r ? n_rows_glob : 0u)
                                         : read_u32(ctx, S->rowmap.p + 2 * (size_t)S->all_node_lo[r]);
    S->row_lo = S->all_row_lo[rank]; S->row_hi = S->all_row_lo[rank + 1];
    const uint32_t n_rows = S->row_hi - S->row_lo;
    CsrMatrix &A = S->Kff;
    A.n_rows = n_rows; A.row_lo = S->row_lo; A.n_cols = n_cols;
    A.rowptr.alloc(ctx, (size_t)n_rows + 1);
    S->rhs.alloc(ctx, n_rows);
    S->diag.alloc(ctx, n_rows);
    const int drop = opt ? opt->drop_exact_zeros : 1;
    const uint32_t n_owned_dof = 2 * n_own;
    S->colid.alloc(ctx, n_dof + 2);
    if (n_dof)
        MAG_LAUNCH(ctx, col_ids_kernel, cdiv(n_dof, 256), 256, 0, (const uint8_t *)S->known.p,
                   (const uint32_t *)S->colmap.p, n_dof, S->colid.p);
    const uint2 *colid2 = reinterpret_cast<const uint2 *>(S->colid.p);
    A.rowptr.zero();
    if (n_owned_dof)
        MAG_LAUNCH(ctx, eliminate_rows_kernel, cdiv(n_owned_dof, 256), 256, 0,
                   (const uint32_t *)K.browptr.p, (const uint32_t *)K.bcol.p, (const double *)K.bval.p,
                   K.node_lo, n_owned_dof, (const uint8_t *)S->known.p, (const uint32_t *)S->rowmap.p, colid2,
                   (const double *)S->bc_ux.p, (const double *)S->bc_uy.p, (const double *)S->bc_fx.p,
                   (const double *)S->bc_fy.p, drop, A.row_lo, A.rowptr.p, S->rhs.p, S->diag.p);
    exclusive_scan_u32(ctx, A.rowptr.p, n_rows, A.rowptr.p, (size_t)n_rows + 1);
    A.nnz = read_u32(ctx, A.rowptr.p + n_rows);
    A.col.alloc(ctx, A.nnz);
    A.val.alloc(ctx, A.nnz);
    if (K.n_blocks)
        MAG_LAUNCH(ctx, eliminate_fill_blocks_kernel, cdiv(K.n_blocks, 256), 256, 0,
                   (const uint32_t *)K.browptr.p, (const uint32_t *)K.brow.p, (const uint32_t *)K.bcol.p,
                   (const double *)K.bval.p, K.n_blocks, K.node_lo, (const uint8_t *)S->known.p,
                   (const uint32_t *)S->rowmap.p, colid2, drop, A.row_lo, (const uint32_t *)A.rowptr.p, A.col.p, A.val.p);
    st.nnz = A.nnz;
    // halo extent: the columns the owned rows touch
    S->ext_lo = S->row_lo; S->ext_hi = S->row_hi;
    if (nranks > 1 && A.nnz) {
        DevBuf<int> mm(ctx, 2);
        const int init[2] = {0x7fffffff, -1};
        MAG_CUDA(cudaMemcpyAsync(mm.p, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
        MAG_LAUNCH(ctx, col_range_kernel, std::min(cdiv(A.nnz, 256), 1024u), 256, 0,
                   (const int32_t *)A.col.p, (size_t)A.nnz, mm.p, mm.p + 1);
        int h[2];
        MAG_CUDA(cudaMemcpyAsync(h, mm.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        S->ext_lo = std::min<uint32_t>(S->row_lo, (uint32_t)h[0]);
        S->ext_hi = std::max<uint32_t>(S->row_hi, (uint32_t)h[1] + 1);
    }
    st.ms_bc = phase.stop();

    // ---- solver format -------------------------------------------------------------
    phase.start();
    build_sell(ctx, A, S->sell, !(opt && opt->spmv_format == 3));
    st.sell_entries = S->sell.entries;
    st.sell_index_bits = S->sell.narrow ? 16 : 32;
    st.ms_format = phase.stop();
    st.spmv_bytes = A.nnz * 12ull + (uint64_t)n_rows * 16ull + ((uint64_t)n_rows + 1) * 4ull;
    st.ms_total = total.stop();
    st.kernel_launches = ctx->launches;
}

}  // namespace mag
