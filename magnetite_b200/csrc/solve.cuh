// solve.cuh — host orchestration of the PCG solve and the post-processing, for
// one rank per process (production; NCCL + CUDA IPC between processes) or for
// several "virtual ranks" inside one process (single-GPU emulation used by the
// tests: same partition, same kernels, same halo stores, allreduce emulated by a
// tiny kernel).
#pragma once
#include <chrono>
#include <cmath>

#include "small.cuh"
#include "system.cuh"

namespace mag {

struct RankState {
    mag_system *S = nullptr;
    uint32_t n = 0;                  // owned rows
    DevBuf<double> x, q;             // local
    DevBuf<double> p_ext, r_store;   // global-indexed
    DevBuf<double> dinv_store;       // Dinv when the system has no shared slab (single rank)
    double *r_ext = nullptr, *dinv_ext = nullptr;
    HaloView halo;                   // my halo buffer (null for a single rank)
    DevBuf<double> partials;
    DevBuf<PcgScalars> scal;
    unsigned grid_vec = 1, grid_spmv = 1, grid_ext = 1;
};

inline size_t ext_len(const mag_system *S) { return (((size_t)S->n_free + 32) + 15) & ~(size_t)15; }   // keeps the LL words 16-byte aligned
inline size_t halo_count(const mag_system *S) {
    return (size_t)(S->row_lo - S->ext_lo) + (size_t)(S->ext_hi - S->row_hi);
}
// Slab other ranks store into: [ Dinv (global-indexed) | mailbox | coarse w partials | halo buffer of r (LL words) ]
inline size_t slab_bytes(const mag_system *S) {
    return ext_len(S) * sizeof(double) + kMailSlots * sizeof(MailSlot) + kCoarseWbufWords * sizeof(LLWord) +
           (halo_count(S) + 1) * sizeof(LLWord);
}
inline MailSlot *slab_mailbox(const mag_system *S, double *slab) {
    return reinterpret_cast<MailSlot *>(slab + ext_len(S));
}
inline LLWord *slab_wbuf(const mag_system *S, double *slab) {
    return reinterpret_cast<LLWord *>(slab_mailbox(S, slab) + kMailSlots);
}
inline LLWord *slab_halo(const mag_system *S, double *slab) { return slab_wbuf(S, slab) + kCoarseWbufWords; }

// Plain cudaMalloc so it can be exported through CUDA IPC.  Mailbox and halo buffer start
// zeroed and are never reset afterwards (sequence numbers only move forward).
static void ensure_shared_slab(mag_system *S) {      // virtual ranks (one process): every system owns its slab
    if (S->shared_slab) return;
    MAG_CUDA(cudaMalloc((void **)&S->shared_slab, slab_bytes(S)));
    MAG_CUDA(cudaMemset(S->shared_slab, 0, slab_bytes(S)));
    MAG_CUDA(cudaDeviceSynchronize());
    S->owns_slab = true;
}

// All ranks meet on the stream (a one-word NCCL allreduce).
static void comm_barrier(mag_ctx *ctx) {
    DevBuf<int> token(ctx, 1);
    token.zero();
    MAG_NCCL(ncclAllReduce(token.p, token.p, 1, ncclInt, ncclSum, ctx->comm->nccl, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
}

// Collective: the communicator's slab holds at least `bytes` afterwards (the same `bytes` on every rank).
static void comm_ensure_slab(mag_ctx *ctx, size_t bytes) {
    Comm *c = ctx->comm;
    if (c->slab && c->slab_bytes >= bytes) return;
    const int R = c->nranks;
    if (R > kMaxRanks) fail(MAG_ERR_BAD_ARG, "at most %d ranks", kMaxRanks);
    if (c->slab) {                          // nobody may free a buffer a peer still has mapped
        for (int r = 0; r < R; ++r)
            if (r != c->rank && c->peer_slab[r]) cudaIpcCloseMemHandle(c->peer_slab[r]);
        comm_barrier(ctx);
        cudaFree(c->slab);
        c->slab = nullptr; c->slab_bytes = 0;
    }
    const size_t want = ((bytes + bytes / 4) + ((size_t)64 << 20) - 1) & ~(((size_t)64 << 20) - 1);   // headroom, 64 MiB steps
    MAG_CUDA(cudaMalloc((void **)&c->slab, want));
    MAG_CUDA(cudaMemset(c->slab, 0, want));
    MAG_CUDA(cudaDeviceSynchronize());
    c->slab_bytes = want;
    cudaIpcMemHandle_t h;
    MAG_CUDA(cudaIpcGetMemHandle(&h, c->slab));
    std::vector<cudaIpcMemHandle_t> hs(R);
    allgather_bytes(ctx, &h, sizeof h, hs.data());
    c->peer_slab.assign(R, nullptr);
    for (int r = 0; r < R; ++r) {
        if (r == c->rank) { c->peer_slab[r] = c->slab; continue; }
        void *p = nullptr;      // every rank is mapped: the mailbox allreduce posts to all of them
        MAG_CUDA(cudaIpcOpenMemHandle(&p, hs[r], cudaIpcMemLazyEnablePeerAccess));
        c->peer_slab[r] = static_cast<double *>(p);
    }
    ++c->generation;
}

// Which of MY rows do the other ranks need?  Pure host logic (also exported as
// mag_halo_plan for the CPU tests): rank r reads [ext_lo[r], row_lo[r]) and
// [row_lo[r+1], ext_hi[r]) from whoever owns those rows; the part inside my block
// [row_lo[me], row_lo[me+1]) is what I store into r's buffers.
struct HaloSeg { uint32_t lo, hi; int dst; };
static std::vector<HaloSeg> halo_plan(int nranks, int me, const uint32_t *row_lo, const uint32_t *ext_lo,
                                      const uint32_t *ext_hi) {
    std::vector<HaloSeg> segs;
    for (int r = 0; r < nranks; ++r) {
        if (r == me) continue;
        const uint32_t need[2][2] = {{ext_lo[r], row_lo[r]}, {row_lo[r + 1], ext_hi[r]}};
        for (int k = 0; k < 2; ++k) {
            const uint32_t lo = std::max(need[k][0], row_lo[me]), hi = std::min(need[k][1], row_lo[me + 1]);
            if (lo < hi) segs.push_back({lo, hi, r});
        }
    }
    return segs;
}

// `peer_slab[r]` is rank r's shared slab as seen from this process.
static void build_push_segments(mag_system *S, const std::vector<uint32_t> &ext_lo,
                                const std::vector<uint32_t> &ext_hi,
                                const std::vector<double *> &peer_slab) {
    PushSegs &ps = S->push;
    ps.n = 0;
    const std::vector<uint32_t> &row = S->all_row_lo;
    for (const HaloSeg &g : halo_plan(S->nranks, S->rank, row.data(), ext_lo.data(), ext_hi.data())) {
        if (ps.n == kMaxPush) fail(MAG_ERR_BAD_ARG, "halo pattern needs more than %d push segments", kMaxPush);
        if (!peer_slab[g.dst]) fail(MAG_ERR_BAD_ARG, "no mapping of rank %d's halo buffer", g.dst);
        // slot of row g.lo in the destination's compact halo buffer: lower halo first, then upper
        const uint32_t d_ext_lo = ext_lo[g.dst], d_row_lo = row[g.dst], d_row_hi = row[g.dst + 1];
        const size_t slot = g.hi <= d_row_lo ? (size_t)(g.lo - d_ext_lo)
                                             : (size_t)(g.lo - d_row_hi) + (size_t)(d_row_lo - d_ext_lo);
        ps.lo[ps.n] = g.lo; ps.hi[ps.n] = g.hi;
        ps.ll_dst[ps.n] = slab_halo(S, peer_slab[g.dst]) + slot;
        ps.dinv_dst[ps.n] = peer_slab[g.dst];
        ++ps.n;
    }
    // where the partial restrictions of the two-level preconditioner go: every rank's buffer
    CoarseLinks &cl = S->coarse.links;
    cl.n = S->nranks; cl.me = S->rank;
    for (int r = 0; r < S->nranks; ++r) cl.wbuf[r] = slab_wbuf(S, peer_slab[r]);
    S->push_ready = true;
}

// Production: exchange the halo extents with the other processes and point this system at the communicator's slab.
static void setup_halo_ipc(mag_ctx *ctx, mag_system *S) {
    Comm *c = ctx->comm;
    if (S->push_ready && S->slab_generation == c->generation) return;
    const int R = S->nranks;
    std::vector<uint32_t> mine = {S->ext_lo, S->ext_hi}, all(2 * (size_t)R);
    allgather_u32(ctx, mine.data(), 2, all.data());
    std::vector<uint32_t> elo(R), ehi(R);
    size_t halo_max = 0;
    for (int r = 0; r < R; ++r) {
        elo[r] = all[2 * r]; ehi[r] = all[2 * r + 1];
        halo_max = std::max(halo_max, (size_t)(S->all_row_lo[r] - elo[r]) + (size_t)(ehi[r] - S->all_row_lo[r + 1]));
    }
    // the same number on every rank: the layout depends on n_free only, the halo part is sized for the widest rank
    const size_t need = slab_bytes(S) - (halo_count(S) + 1) * sizeof(LLWord) + (halo_max + 1) * sizeof(LLWord);
    comm_ensure_slab(ctx, need);
    if (c->layout_rows != ext_len(S)) {
        // another layout than the last system's: what were plain doubles (Dinv) may now lie where self-validating
        // words are read.  Start from zeros (never a valid sequence number); all ranks, before anyone writes.
        comm_barrier(ctx);
        MAG_CUDA(cudaMemsetAsync(c->slab, 0, c->slab_bytes, ctx->stream));
        comm_barrier(ctx);
        c->layout_rows = ext_len(S);
    }
    S->shared_slab = c->slab;
    S->owns_slab = false;
    S->slab_generation = c->generation;
    S->links.n = R; S->links.me = S->rank;
    for (int r = 0; r < R; ++r) S->links.box[r] = slab_mailbox(S, c->peer_slab[r]);
    build_push_segments(S, elo, ehi, c->peer_slab);
}

static void rank_alloc(mag_ctx *ctx, RankState &W, mag_system *S) {
    W.S = S;
    W.n = S->Kff.n_rows;
    W.x.alloc(ctx, W.n); W.q.alloc(ctx, W.n);
    W.p_ext.alloc(ctx, ext_len(S));
    W.p_ext.zero();
    W.r_store.alloc(ctx, ext_len(S));
    W.r_ext = W.r_store.p;
    if (S->shared_slab) {
        W.dinv_ext = S->shared_slab;
        W.halo.ll = slab_halo(S, S->shared_slab);
        W.halo.ext_lo = S->ext_lo; W.halo.row_lo = S->row_lo; W.halo.row_hi = S->row_hi;
    } else {
        W.dinv_store.alloc(ctx, ext_len(S));
        W.dinv_ext = W.dinv_store.p;
    }
    const unsigned cap = (unsigned)ctx->sm_count * 8u;
    W.grid_vec = std::max(1u, std::min(cdiv(W.n, 256), cap));
    W.grid_ext = std::max(1u, std::min(cdiv(S->ext_hi - S->ext_lo, 256), cap));
    W.grid_spmv = sell_grid(ctx, S->sell.n_slices);
    W.partials.alloc(ctx, 2 * (size_t)std::max(cap, W.grid_spmv));
    W.scal.alloc(ctx, 1);
    W.scal.zero();
}

// How the per-rank partial sums become global sums.
enum class Reduce { kNone, kMailbox, kNccl, kEmulated };
constexpr size_t kOffPair0 = offsetof(PcgScalars, pair) / sizeof(double);
constexpr size_t kOffPq = offsetof(PcgScalars, pq) / sizeof(double);
constexpr size_t kOffLocPair = offsetof(PcgScalars, loc_pair) / sizeof(double);
constexpr size_t kOffLocPq = offsetof(PcgScalars, loc_pq) / sizeof(double);

constexpr uint32_t kAutoTwoLevelMinRows = 20000;    // precond 3: below this Jacobi-PCG is faster than the coarse setup

struct SolveMode {
    Reduce reduce = Reduce::kNone;
    int format = 2;
    bool two_level = false;
    bool pdl = true;          // programmatic dependent launches between the kernels of the loop (MAG_TUNE=512: off)
};

inline CoarseView coarse_view(RankState &W, const SolveMode &m) {
    CoarseView cv;
    if (m.two_level) { cv.mode = W.S->coarse.mode.p; cv.rot = W.S->coarse.rot.p; cv.y = W.S->coarse.y.p; }
    return cv;
}

inline double *scal_field(RankState &W, size_t off_doubles) {
    return reinterpret_cast<double *>(W.scal.p) + off_doubles;
}
// Where a kernel writes its sums: the global slot when nothing follows, the local slot otherwise.
inline double *pq_target(RankState &W, const SolveMode &m) {
    return scal_field(W, (m.reduce == Reduce::kNccl || m.reduce == Reduce::kEmulated) ? kOffLocPq : kOffPq);
}
inline double *pair_target(RankState &W, const SolveMode &m, int slot) {
    return scal_field(W, (m.reduce == Reduce::kNccl || m.reduce == Reduce::kEmulated) ? kOffLocPair
                                                                                      : kOffPair0 + 2 * (size_t)slot);
}
inline PeerLinks links_of(RankState &W, const SolveMode &m) {
    PeerLinks none = {};
    return m.reduce == Reduce::kMailbox ? W.S->links : none;
}

// loc -> global (count doubles) for the NCCL and emulated modes; a no-op otherwise.
static void reduce_scalars(mag_ctx *ctx, std::vector<RankState> &ranks, const SolveMode &m, size_t src_off,
                           size_t dst_off, int count) {
    if (m.reduce == Reduce::kEmulated) {
        ScalPtrs sp;
        sp.n = (int)ranks.size();
        for (int r = 0; r < sp.n; ++r) sp.p[r] = ranks[r].scal.p;
        MAG_LAUNCH(ctx, emulated_allreduce_kernel, 1, 32, 0, sp, (int)src_off, (int)dst_off, count);
    } else if (m.reduce == Reduce::kNccl) {
        Comm *c = ctx->comm;
        MAG_NCCL(ncclAllReduce(scal_field(ranks[0], src_off), scal_field(ranks[0], dst_off), (size_t)count,
                               ncclDouble, ncclSum, c->nccl, ctx->stream));
    }
}

// Sum over ranks of a device vector each rank holds (in place).  No-op for a single rank.
static void reduce_vector(mag_ctx *ctx, std::vector<RankState> &ranks, const SolveMode &m,
                          const std::vector<double *> &vec, size_t count) {
    if (ranks.size() > 1) {
        VecPtrs vp;
        vp.n = (int)ranks.size();
        for (int r = 0; r < vp.n; ++r) vp.p[r] = vec[r];
        MAG_LAUNCH(ctx, emulated_vec_allreduce_kernel, std::min(cdiv(count, 256), 1024u), 256, 0, vp, count);
    } else if (ranks[0].S->nranks > 1) {
        MAG_NCCL(ncclAllReduce(vec[0], vec[0], count, ncclDouble, ncclSum, ctx->comm->nccl, ctx->stream));
    }
    (void)m;
}

constexpr size_t kOffWy = offsetof(PcgScalars, wy) / sizeof(double);
constexpr size_t kOffLocWy = offsetof(PcgScalars, loc_wy) / sizeof(double);

// w = P^T r (per rank, straight into every rank's buffer) -> y = (this rank's rows of Ac^-1) w and its share of w.y
// step < 0: the initial residual.
static void enqueue_coarse_solve(mag_ctx *ctx, std::vector<RankState> &ranks, const SolveMode &m, int step) {
    for (RankState &W : ranks) {
        CoarseSpace &C = W.S->coarse;
        if (C.n_lagg)
            MAG_LAUNCH_DEP(ctx, m.pdl, coarse_restrict_kernel, C.n_lagg, kRestrictThreads, 0, (const uint32_t *)C.lagg.p,
                       (const uint32_t *)C.agg_ptr.p, (const uint32_t *)C.perm_ax.p, (const float *)C.rot_perm.p,
                       (const double *)W.r_ext, W.S->row_lo, step, C.links, (const PcgScalars *)W.scal.p);
    }
    const bool local_sum = m.reduce == Reduce::kNccl || m.reduce == Reduce::kEmulated;
    for (RankState &W : ranks) {
        CoarseSpace &C = W.S->coarse;
        MAG_LAUNCH_DEP(ctx, m.pdl, coarse_gather_w_kernel, cdiv(C.nc, 256), 256, 0, (const uint16_t *)C.touch.p, C.nc, step, C.links,
                   C.w.p, W.scal.p);
        // enough rows to fill the machine with one warp each (a single GPU): warp per row; else a CTA per row
        const unsigned warp_ctas = cdiv((size_t)C.m * 32, 256);
        if (warp_ctas >= (unsigned)ctx->sm_count * 4u) {
            MAG_LAUNCH_DEP(ctx, m.pdl, coarse_apply_warp_kernel, std::min(warp_ctas, (unsigned)ctx->sm_count * 8u), 256, 0,
                       (const double *)C.Ainv.p, (const uint32_t *)C.crow.p, (const uint8_t *)C.wy_mine.p,
                       (const double *)C.w.p, C.m, C.nc, step, links_of(W, m), C.y.p, C.partials.p, C.ticket.p,
                       W.scal.p, scal_field(W, local_sum ? kOffLocWy : kOffWy));
        } else {
            const unsigned grid = std::max(1u, std::min(C.m, (unsigned)ctx->sm_count * 8u));
            MAG_LAUNCH_DEP(ctx, m.pdl, coarse_apply_kernel, grid, 256, 0, (const double *)C.Ainv.p, (const uint32_t *)C.crow.p,
                       (const uint8_t *)C.wy_mine.p, (const double *)C.w.p, C.m, C.nc, step, links_of(W, m), C.y.p,
                       C.partials.p, C.ticket.p, W.scal.p, scal_field(W, local_sum ? kOffLocWy : kOffWy));
        }
    }
    reduce_scalars(ctx, ranks, m, kOffLocWy, kOffWy, 1);
}

static void enqueue_iteration(mag_ctx *ctx, std::vector<RankState> &ranks, int step, const SolveMode &m) {
    const int parity = step & 1;
    for (RankState &W : ranks) {
        const SellMatrix &L = W.S->sell;
        const CsrMatrix &A = W.S->Kff;
        if (m.format == 1)
            MAG_LAUNCH_DEP(ctx, m.pdl, pcg_spmv_csr_kernel, W.grid_vec, 256, 0, (const uint32_t *)A.rowptr.p,
                       (const int32_t *)A.col.p, (const double *)A.val.p, (const double *)W.p_ext.p, W.q.p,
                       W.n, A.row_lo, step, links_of(W, m), W.partials.p, W.scal.p, pq_target(W, m));
        else if (L.narrow)
            MAG_LAUNCH_DEP(ctx, m.pdl, pcg_spmv_kernel<int16_t>, W.grid_spmv, 256, 0, (const uint32_t *)L.slice_off.p,
                       (const int16_t *)L.pcol.p, (const double *)L.val.p, (const double *)W.p_ext.p, W.q.p,
                       W.n, L.n_slices, L.row_lo, step, links_of(W, m), W.partials.p, W.scal.p, pq_target(W, m));
        else
            MAG_LAUNCH_DEP(ctx, m.pdl, pcg_spmv_kernel<int32_t>, W.grid_spmv, 256, 0, (const uint32_t *)L.slice_off.p,
                       (const int32_t *)L.col.p, (const double *)L.val.p, (const double *)W.p_ext.p, W.q.p,
                       W.n, L.n_slices, L.row_lo, step, links_of(W, m), W.partials.p, W.scal.p, pq_target(W, m));
    }
    reduce_scalars(ctx, ranks, m, kOffLocPq, kOffPq, 1);
    for (RankState &W : ranks)
        MAG_LAUNCH_DEP(ctx, m.pdl, pcg_update_xr_kernel, W.grid_vec, 256, 0, W.x.p, W.r_ext, (const double *)W.p_ext.p,
                   (const double *)W.q.p, (const double *)W.dinv_ext, W.n, W.S->row_lo, step, W.S->push,
                   links_of(W, m), W.partials.p, W.scal.p, pair_target(W, m, parity ^ 1));
    reduce_scalars(ctx, ranks, m, kOffLocPair, kOffPair0 + 2 * (size_t)(parity ^ 1), 2);
    if (m.two_level) enqueue_coarse_solve(ctx, ranks, m, step);
    for (RankState &W : ranks)
        MAG_LAUNCH_DEP(ctx, m.pdl, pcg_update_p_kernel, W.grid_ext, 256, 0, W.p_ext.p, (const double *)W.r_ext,
                   (const double *)W.dinv_ext, W.S->ext_lo, W.S->ext_hi, step, W.halo, coarse_view(W, m),
                   links_of(W, m), W.scal.p);
}

// Small host-side exchanges of the setup: one byte string per rank -> all of them on every rank.
static std::vector<std::vector<uint8_t>> exchange_bytes(mag_ctx *ctx, std::vector<RankState> &ranks,
                                                         const std::vector<std::vector<uint8_t>> &mine, size_t bytes) {
    const int R = ranks.size() > 1 ? (int)ranks.size() : ranks[0].S->nranks;
    std::vector<std::vector<uint8_t>> all(R);
    if (ranks.size() > 1) {
        for (int r = 0; r < R; ++r) all[r] = mine[r];
    } else if (R == 1) {
        all[0] = mine[0];
    } else {
        std::vector<uint8_t> flat((size_t)R * bytes);
        allgather_bytes(ctx, mine[0].data(), bytes, flat.data());
        for (int r = 0; r < R; ++r) all[r].assign(flat.begin() + (size_t)r * bytes, flat.begin() + (size_t)(r + 1) * bytes);
    }
    return all;
}

// Builds the aggregation coarse space of every rank's system (once per system).  Returns false when Ac is not
// positive definite (the system is not SPD, e.g. an all-clockwise mesh): the caller falls back to Jacobi.
static bool setup_coarse(mag_ctx *ctx, std::vector<RankState> &ranks, const SolveMode &m, const mag_options &opt) {
    const bool trace = (ctx->tune & 32) != 0;          // MAG_TUNE=32: wall-clock of the setup stages on stderr
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        cudaStreamSynchronize(ctx->stream);
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[coarse setup] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    bool all_ready = true, any_failed = false;
    for (RankState &W : ranks) { all_ready = all_ready && W.S->coarse.ready; any_failed = any_failed || W.S->coarse.failed; }
    if (any_failed) return false;
    if (all_ready) return true;
    const int R = ranks.size() > 1 ? (int)ranks.size() : ranks[0].S->nranks;
    std::vector<double *> mats;
    std::vector<std::vector<uint8_t>> has_rows(ranks.size()), need(ranks.size());
    bool any_far = false;
    for (size_t q = 0; q < ranks.size(); ++q) {
        RankState &W = ranks[q];
        mag_system *S = W.S;
        CoarseSpace &C = S->coarse;
        const size_t N = S->n_nodes, n_dof = 2 * N;
        // bounding box (global mesh: identical on every rank)
        const unsigned bgrid = std::max(1u, std::min(cdiv(N, 256), 256u));
        DevBuf<double> bb(ctx, (size_t)bgrid * 4);
        MAG_LAUNCH(ctx, bbox_kernel, bgrid, 256, 0, (const double2 *)S->xy.p, N, bb.p);
        std::vector<double> hb((size_t)bgrid * 4);
        MAG_CUDA(cudaMemcpyAsync(hb.data(), bb.p, hb.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
        for (unsigned b = 0; b < bgrid; ++b) {
            xmin = std::min(xmin, hb[4 * b]); xmax = std::max(xmax, hb[4 * b + 1]);
            ymin = std::min(ymin, hb[4 * b + 2]); ymax = std::max(ymax, hb[4 * b + 3]);
        }
        const double Wd = std::max(xmax - xmin, 1e-300), Hd = std::max(ymax - ymin, 1e-300);
        uint32_t target = opt.coarse_aggregates > 0 ? (uint32_t)opt.coarse_aggregates
                                                    : (uint32_t)std::min<uint64_t>(kCoarseMaxAgg, std::max<uint64_t>(4, S->n_free / 4096));
        target = std::min(target, kCoarseMaxAgg);
        CoarseGrid &g = C.grid;
        g.nbx = std::max(1u, (uint32_t)std::lround(std::sqrt((double)target * Wd / Hd)));
        g.nby = std::max(1u, (uint32_t)std::lround((double)target / g.nbx));
        while ((uint64_t)g.nbx * g.nby > kCoarseMaxAgg) { if (g.nbx >= g.nby) --g.nbx; else --g.nby; }
        g.x_fast = g.nbx <= g.nby ? 1 : 0;          // number along the shorter side: narrow band
        g.x0 = xmin; g.y0 = ymin;
        g.hx = Wd / g.nbx * (1.0 + 1e-12); g.hy = Hd / g.nby * (1.0 + 1e-12);
        C.n_agg = g.nbx * g.nby;
        C.nc = 3 * C.n_agg;
        C.hb = std::min(C.nc - 1, 3 * (std::min(g.nbx, g.nby) + 1) + 2);
        lap("bounding box");
        C.mode.alloc(ctx, ext_len(S)); C.rot.alloc(ctx, ext_len(S));
        C.mode.zero(); C.rot.zero();
        MAG_LAUNCH(ctx, coarse_colinfo_kernel, cdiv(n_dof, 256), 256, 0, (const double2 *)S->xy.p,
                   (const uint8_t *)S->known.p, (const uint32_t *)S->colmap.p, n_dof, g, C.mode.p, C.rot.p);
        lap("column info");
        // local rows sorted by aggregate
        const uint32_t n = S->Kff.n_rows;
        C.perm_ax.alloc(ctx, n); C.rot_perm.alloc(ctx, n);
        C.agg_ptr.alloc(ctx, (size_t)C.n_agg + 1);
        {
            DevBuf<uint64_t> keys(ctx, n), keys_alt(ctx, n);
            DevBuf<uint32_t> perm(ctx, n), pay_alt(ctx, n);
            if (n) {
                MAG_LAUNCH(ctx, coarse_rowkeys_kernel, cdiv(n, 256), 256, 0, (const uint32_t *)C.mode.p, n, S->row_lo,
                           keys.p, perm.p);
                radix_sort_pairs(ctx, keys.p, perm.p, keys_alt.p, pay_alt.p, n, bits_for((uint64_t)C.n_agg + 1));
                MAG_LAUNCH(ctx, coarse_pack_rows_kernel, cdiv(n, 256), 256, 0, (const uint32_t *)perm.p,
                           (const uint32_t *)C.mode.p, (const float *)C.rot.p, n, S->row_lo, C.perm_ax.p, C.rot_perm.p);
            }
            MAG_LAUNCH(ctx, coarse_segments_kernel, cdiv((size_t)C.n_agg + 1, 256), 256, 0, (const uint64_t *)keys.p, n,
                       C.n_agg, C.agg_ptr.p);
        }
        lap("sort rows by aggregate");
        // which aggregates have rows of this rank, and which ones its rows and halo touch
        std::vector<uint32_t> h_ptr((size_t)C.n_agg + 1);
        DevBuf<uint8_t> d_need(ctx, C.n_agg);
        d_need.zero();
        if (S->ext_hi > S->ext_lo)
            MAG_LAUNCH(ctx, coarse_needed_kernel, cdiv(S->ext_hi - S->ext_lo, 256), 256, 0, (const uint32_t *)C.mode.p,
                       S->ext_lo, S->ext_hi, d_need.p);
        need[q].resize(C.n_agg);
        MAG_CUDA(cudaMemcpyAsync(h_ptr.data(), C.agg_ptr.p, h_ptr.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        MAG_CUDA(cudaMemcpyAsync(need[q].data(), d_need.p, C.n_agg, cudaMemcpyDeviceToHost, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        has_rows[q].resize(C.n_agg);
        std::vector<uint32_t> h_lagg;
        for (uint32_t a = 0; a < C.n_agg; ++a) {
            has_rows[q][a] = h_ptr[a + 1] > h_ptr[a] ? 1 : 0;
            if (has_rows[q][a]) h_lagg.push_back(a);
        }
        C.n_lagg = (uint32_t)h_lagg.size();
        C.lagg.alloc(ctx, h_lagg.size());
        if (!h_lagg.empty())
            MAG_CUDA(cudaMemcpyAsync(C.lagg.p, h_lagg.data(), h_lagg.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        // Galerkin product of the local rows
        C.Ac_compact.alloc(ctx, (size_t)C.n_agg * 81);
        DevBuf<int> far(ctx, 1);
        far.zero();
        MAG_LAUNCH(ctx, coarse_galerkin_kernel, C.n_agg, kGalerkinWarps * 32, 0, (const uint32_t *)C.agg_ptr.p, (const uint32_t *)C.perm_ax.p,
                   (const uint32_t *)S->Kff.rowptr.p, (const int32_t *)S->Kff.col.p, (const double *)S->Kff.val.p,
                   (const uint32_t *)C.mode.p, (const float *)C.rot.p, S->row_lo, g, C.Ac_compact.p, far.p);
        int h_far = 0;
        MAG_CUDA(cudaMemcpyAsync(&h_far, far.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        if (h_far) any_far = true;       // an element spans non-adjacent boxes: this grid cannot carry the 9-point coarse stencil
        C.y.alloc(ctx, C.nc); C.w.alloc(ctx, C.nc);
        C.y.zero(); C.w.zero();
        C.partials.alloc(ctx, 2 * (size_t)ctx->sm_count * 8);
        C.ticket.alloc(ctx, 1);
        C.ticket.zero();
        mats.push_back(C.Ac_compact.p);
        lap("Galerkin product");
    }
    reduce_vector(ctx, ranks, m, mats, (size_t)ranks[0].S->coarse.n_agg * 81);      // block rows, summed over ranks
    lap("sum over ranks");
    {
        // boxes smaller than an element somewhere (on any rank): no coarse space, the caller runs Jacobi
        std::vector<std::vector<uint8_t>> flag(ranks.size(), std::vector<uint8_t>(1, any_far ? 1 : 0));
        for (const auto &f : exchange_bytes(ctx, ranks, flag, 1)) any_far = any_far || f[0];
        if (any_far) {
            for (RankState &W : ranks) { W.S->coarse.failed = true; W.S->coarse.ready = false; }
            return false;
        }
    }
    const uint32_t n_agg = ranks[0].S->coarse.n_agg;
    const std::vector<std::vector<uint8_t>> all_has = exchange_bytes(ctx, ranks, has_rows, n_agg);
    std::vector<uint16_t> touch(n_agg, 0);
    for (uint32_t a = 0; a < n_agg; ++a)
        for (int r = 0; r < R; ++r)
            if (all_has[r][a]) touch[a] |= (uint16_t)(1u << r);
    bool spd = true;
    for (size_t q = 0; q < ranks.size(); ++q) {
        RankState &W = ranks[q];
        mag_system *S = W.S;
        CoarseSpace &C = S->coarse;
        const int me = S->rank;
        const uint32_t nc = C.nc, hb = C.hb, Wb = hb + 1;
        // the rows of Ac^-1 this rank applies: 3 per aggregate its rows or halo touch; w.y is added by the lowest
        // rank that has rows in the aggregate
        std::vector<uint32_t> h_crow;
        std::vector<uint8_t> h_mine;
        for (uint32_t a = 0; a < n_agg; ++a) {
            if (!need[q][a] && !all_has[me][a]) continue;
            const bool mine = touch[a] != 0 && (touch[a] & (uint16_t)((1u << me) - 1u)) == 0 && ((touch[a] >> me) & 1u);
            for (uint32_t j = 0; j < 3; ++j) { h_crow.push_back(3 * a + j); h_mine.push_back(mine ? 1 : 0); }
        }
        C.m = (uint32_t)h_crow.size();
        C.crow.alloc(ctx, C.m); C.wy_mine.alloc(ctx, C.m); C.touch.alloc(ctx, n_agg);
        if (C.m) {
            MAG_CUDA(cudaMemcpyAsync(C.crow.p, h_crow.data(), (size_t)C.m * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
            MAG_CUDA(cudaMemcpyAsync(C.wy_mine.p, h_mine.data(), C.m, cudaMemcpyHostToDevice, ctx->stream));
        }
        MAG_CUDA(cudaMemcpyAsync(C.touch.p, touch.data(), (size_t)n_agg * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        // banded Cholesky of Ac (every rank, ~2 ms), then only this rank's rows of the inverse
        DevBuf<double> lower(ctx, (size_t)nc * Wb), upper(ctx, (size_t)nc * Wb), invd(ctx, nc);
        DevBuf<int> bad(ctx, 1);
        lower.zero(); bad.zero();
        MAG_LAUNCH(ctx, coarse_band_kernel, C.n_agg, 96, 0, (const double *)C.Ac_compact.p, C.grid, hb, lower.p);
        C.Ac_compact.release();
        MAG_LAUNCH(ctx, coarse_fix_diagonal_kernel, cdiv(nc, 256), 256, 0, lower.p, nc, hb);
        const size_t chol_smem = ((size_t)(hb + 3) * Wb + 3 * (size_t)(hb + 3)) * sizeof(double);
        if (chol_smem > 200 * 1024) fail(MAG_ERR_BAD_ARG, "two-level preconditioner: band of %u does not fit the factorisation window", hb);
        MAG_CUDA(cudaFuncSetAttribute(band_cholesky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem));
        MAG_LAUNCH(ctx, band_cholesky_kernel, 1, 1024, chol_smem, lower.p, nc, hb, invd.p, bad.p);
        int h_bad = 0;
        MAG_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        lap("banded Cholesky");
        if (h_bad) { spd = false; C.failed = true; continue; }
        MAG_LAUNCH(ctx, band_transpose_kernel, cdiv((size_t)nc * Wb, 256), 256, 0, (const double *)lower.p, nc, hb, upper.p);
        C.Ainv.alloc(ctx, (size_t)std::max(C.m, 1u) * nc);
        if (C.m) {
            if (hb > kInvMaxBand) fail(MAG_ERR_BAD_ARG, "two-level preconditioner: band of %u exceeds the substitution kernel's %u", hb, kInvMaxBand);
            const size_t inv_smem = 2 * (size_t)kInvChunk * Wb * sizeof(double);
            // few right-hand sides (a rank of a multi-GPU run): 8 warps per CTA — twice the CTAs, shorter steps
            if (C.m <= (uint32_t)ctx->sm_count * 8u) {
                MAG_CUDA(cudaFuncSetAttribute(band_inverse_rows_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)inv_smem));
                MAG_LAUNCH(ctx, band_inverse_rows_kernel<8>, cdiv(C.m, 8), 8 * 32, inv_smem, (const double *)lower.p,
                           (const double *)upper.p, (const double *)invd.p, nc, hb, (const uint32_t *)C.crow.p, C.m, C.Ainv.p);
            } else {
                MAG_CUDA(cudaFuncSetAttribute(band_inverse_rows_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)inv_smem));
                MAG_LAUNCH(ctx, band_inverse_rows_kernel<16>, cdiv(C.m, 16), 16 * 32, inv_smem, (const double *)lower.p,
                           (const double *)upper.p, (const double *)invd.p, nc, hb, (const uint32_t *)C.crow.p, C.m, C.Ainv.p);
            }
        }
        // single rank without a shared slab: its own buffer for the partial restrictions
        if (!S->shared_slab) {
            C.wbuf_local.alloc(ctx, kCoarseWbufWords);
            C.wbuf_local.zero();
            C.links.n = 1; C.links.me = 0;
            C.links.wbuf[0] = C.wbuf_local.p;
        }
        MAG_CUDA(cudaStreamSynchronize(ctx->stream));
        C.ready = true;
        lap("rows of the inverse");
    }
    if (!spd) {
        for (RankState &W : ranks) { W.S->coarse.failed = true; W.S->coarse.ready = false; }
        return false;
    }
    return true;
}

struct SolveOutcome {
    PcgScalars hs;
    double bb = 0.0;
    float ms_coarse_setup = 0.f;
    uint32_t n_coarse = 0;
    int precond_used = 0;
    bool rerun_best = false;
};

// Systems that fit one thread-block cluster's shared memory (small.cuh): the whole solve in one kernel.
// Returns false when the system does not fit (or the cluster cannot be launched): the caller takes the general path.
static bool small_cg_try(mag_ctx *ctx, RankState &W, const mag_options &opt, int jacobi, bool compat, SolveOutcome &out) {
    mag_system *S = W.S;
    const CsrMatrix &A = S->Kff;
    const uint32_t n = A.n_rows;
    if ((ctx->tune & 128) || n == 0 || n > 65535u || A.nnz > (1u << 22)) return false;
    std::vector<uint32_t> h_rowptr((size_t)n + 1);
    MAG_CUDA(cudaMemcpyAsync(h_rowptr.data(), A.rowptr.p, h_rowptr.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    SmallArgs a{};
    unsigned C = 0;
    size_t smem = 0;
    for (unsigned c : {8u, 16u}) {
        const size_t need = small_cg_smem(n, A.nnz, c, &a.rows_per_cta, &a.nnz_cap, &a.n_pad, h_rowptr);
        if (need && need <= 200 * 1024) { C = c; smem = need; break; }
    }
    if (!C) return false;
    a.rowptr = A.rowptr.p; a.col = A.col.p; a.val = A.val.p; a.b = S->rhs.p; a.diag = S->diag.p; a.x = W.x.p;
    a.n = n; a.jacobi = jacobi; a.compat = compat ? 1 : 0;
    a.rel_tol2 = opt.rel_tol * opt.rel_tol;
    a.abs_thr2 = opt.cost_kind == 1 ? opt.abs_tol : opt.abs_tol * opt.abs_tol;
    a.max_iter = opt.max_iter;
    a.out = W.scal.p;
    if (cudaFuncSetAttribute(small_cg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        (C > 8 && cudaFuncSetAttribute(small_cg_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)) {
        cudaGetLastError();
        return false;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(C); cfg.blockDim = dim3(kSmallThreads); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, small_cg_kernel, &cfg) != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return false;
    }
    if (cudaLaunchKernelEx(&cfg, small_cg_kernel, a) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    ctx->launches++;
    MAG_CUDA(cudaMemcpyAsync(&out.hs, W.scal.p, sizeof out.hs, cudaMemcpyDeviceToHost, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    out.bb = out.hs.rzc[0];
    out.precond_used = compat ? 0 : (jacobi ? 1 : 0);
    return true;
}

// Runs CG over `ranks` (size 1 in production).  All ranks see identical scalars.
static SolveOutcome pcg_drive(mag_ctx *ctx, std::vector<RankState> &ranks, const mag_options &opt) {
    SolveMode mode;
    mode.format = opt.spmv_format == 1 ? 1 : 2;
    mode.pdl = !(ctx->tune & 512);
    if (ranks.size() > 1) mode.reduce = Reduce::kEmulated;
    else if (ranks[0].S->nranks > 1) mode.reduce = opt.allreduce == 1 ? Reduce::kNccl : Reduce::kMailbox;
    const bool compat = opt.compat != 0;
    const int jacobi = compat ? 0 : (opt.precond != 0);
    // precond 3 (the default): two-level when the system is large enough to pay for the setup and turns out SPD
    const bool want_two = !compat && (opt.precond == 2 || (opt.precond == 3 && ranks[0].S->n_free >= kAutoTwoLevelMinRows));
    mode.two_level = want_two;
    int chunk = opt.check_every > 0 ? opt.check_every : 50;
    chunk += chunk & 1;   // iteration parity is baked into the graph: even chunk length
    SolveOutcome out;
    std::memset(&out.hs, 0, sizeof out.hs);
    PcgScalars &hs = out.hs;
    const uint32_t n_glob = ranks[0].S->n_free;
    if (n_glob == 0) { hs.stop = 1; return out; }
    // the sequence number of a peer message carries iteration + 1 in 24 bits (pcg.cuh: ll_seq)
    if (opt.max_iter >= (1ull << 24) - 2 && (ranks.size() > 1 || ranks[0].S->nranks > 1))
        fail(MAG_ERR_BAD_ARG, "multi-GPU solves take max_iter < 2^24 - 2 (got %llu)", (unsigned long long)opt.max_iter);
    // one rank, no coarse space, small enough for one cluster's shared memory: the single-kernel solve
    if (ranks.size() == 1 && ranks[0].S->nranks == 1 && !mode.two_level && mode.format != 1 &&
        small_cg_try(ctx, ranks[0], opt, jacobi, compat, out))
        return out;
    if (mode.two_level) {
        EventTimer t(ctx->stream);
        t.start();
        mode.two_level = setup_coarse(ctx, ranks, mode, opt);      // false: Ac not positive definite -> Jacobi
        out.ms_coarse_setup = t.stop();
        out.n_coarse = mode.two_level ? ranks[0].S->coarse.nc : 0;
    }
    out.precond_used = compat ? 0 : (mode.two_level ? 2 : (jacobi ? 1 : 0));

    for (RankState &W : ranks) {
        PcgScalars z;
        std::memset(&z, 0, sizeof z);
        // the same on every rank: solves are collective.  Production: the communicator counts (its slab outlives systems)
        Comm *cm = (ranks.size() == 1 && W.S->nranks > 1) ? ctx->comm : nullptr;
        if (cm) {
            if (++cm->epoch % 255ull == 0) {            // the 8-bit epoch tag wraps: forget every old message first
                comm_barrier(ctx);
                const size_t off = ext_len(W.S) * sizeof(double);
                MAG_CUDA(cudaMemsetAsync(reinterpret_cast<char *>(cm->slab) + off, 0, cm->slab_bytes - off, ctx->stream));
                comm_barrier(ctx);
            }
            W.S->solve_epoch = cm->epoch;
        } else {
            ++W.S->solve_epoch;
        }
        z.epoch = W.S->solve_epoch;
        z.tune = ctx->tune;
        MAG_CUDA(cudaMemcpyAsync(W.scal.p, &z, sizeof z, cudaMemcpyHostToDevice, ctx->stream));
    }
    for (RankState &W : ranks)
        MAG_LAUNCH(ctx, pcg_init_kernel, W.grid_vec, 256, 0, W.x.p, W.r_ext, W.dinv_ext,
                   (const double *)W.S->rhs.p, (const double *)W.S->diag.p, jacobi, W.n, W.S->row_lo,
                   W.S->push, links_of(W, mode), W.partials.p, W.scal.p, pair_target(W, mode, 0));
    // the reduction also orders the halo stores of r and Dinv before their readers
    reduce_scalars(ctx, ranks, mode, kOffLocPair, kOffPair0, 2);
    if (mode.two_level) enqueue_coarse_solve(ctx, ranks, mode, -1);
    for (RankState &W : ranks)
        MAG_LAUNCH(ctx, pcg_init_p_kernel, W.grid_ext, 256, 0, W.p_ext.p, (const double *)W.r_ext,
                   (const double *)W.dinv_ext, W.S->ext_lo, W.S->ext_hi, W.halo, coarse_view(W, mode),
                   links_of(W, mode), W.scal.p);
    MAG_CUDA(cudaMemcpyAsync(&hs, ranks[0].scal.p, sizeof hs, cudaMemcpyDeviceToHost, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    const double bb = hs.pair[0][1];
    out.bb = bb;
    const double thr2 = compat ? (opt.cost_kind == 1 ? opt.abs_tol : opt.abs_tol * opt.abs_tol)
                               : opt.rel_tol * opt.rel_tol * bb;
    int stop0 = 0;
    if (!(bb == bb)) stop0 = 3;                         // NaN right-hand side
    else if (bb <= thr2) stop0 = 1;                     // argmin: the initial cost is already <= target
    else if (opt.max_iter == 0) stop0 = 2;
    for (RankState &W : ranks) {
        PcgScalars init;
        std::memset(&init, 0, sizeof init);
        init.pair[0][0] = hs.pair[0][0]; init.pair[0][1] = hs.pair[0][1]; init.rzc[0] = hs.rzc[0];
        init.thr2 = thr2; init.max_iter = opt.max_iter; init.stop = stop0;
        init.best_rr = bb; init.best_iter = 0;
        init.epoch = W.S->solve_epoch;
        init.tune = ctx->tune;
        MAG_CUDA(cudaMemcpyAsync(W.scal.p, &init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    }
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    hs.thr2 = thr2; hs.stop = stop0; hs.iter = 0;
    hs.best_rr = bb; hs.best_iter = 0;
    if (stop0) return out;

    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const uint64_t l0 = ctx->launches;
    MAG_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    try {
        for (int i = 0; i < chunk; ++i) enqueue_iteration(ctx, ranks, i, mode);
        for (RankState &W : ranks) MAG_LAUNCH_DEP(ctx, mode.pdl, pcg_chunk_end_kernel, 1, 1, 0, chunk, W.scal.p);
    } catch (...) {
        cudaStreamEndCapture(ctx->stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
    }
    MAG_CUDA(cudaStreamEndCapture(ctx->stream, &graph));
    const uint64_t per_chunk = ctx->launches - l0;
    ctx->launches = l0;
    MAG_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    static_assert(2 * sizeof(PcgScalars) <= 128 * sizeof(double), "pinned scratch too small");
    PcgScalars *slot[2] = {reinterpret_cast<PcgScalars *>(ctx->h_scal),
                           reinterpret_cast<PcgScalars *>(ctx->h_scal) + 1};
    cudaEvent_t ev[2];
    MAG_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    MAG_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    uint64_t queued = 0;
    try {
        for (uint64_t c = 0;; ++c) {
            const int s = (int)(c & 1);
            MAG_CUDA(cudaGraphLaunch(exec, ctx->stream));
            ctx->launches += per_chunk;
            MAG_CUDA(cudaMemcpyAsync(slot[s], ranks[0].scal.p, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream));
            MAG_CUDA(cudaEventRecord(ev[s], ctx->stream));
            queued += (uint64_t)chunk;
            if (c > 0) {           // look at the previous chunk while this one runs
                MAG_CUDA(cudaEventSynchronize(ev[s ^ 1]));
                if (slot[s ^ 1]->stop) break;
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
    MAG_CUDA(cudaMemcpyAsync(&hs, ranks[0].scal.p, sizeof hs, cudaMemcpyDeviceToHost, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    return out;
}

// compat mode = the reference's Executor: it returns `best_param`, the iterate with the lowest cost
// (solver.rs:167-176).  When the target cost is reached that is the last iterate.  When max_iters ends
// the run it may be an earlier one: CG is deterministic here, so the solve is simply repeated up to that
// iteration instead of copying x aside on every improvement.
static SolveOutcome pcg_drive_best(mag_ctx *ctx, std::vector<RankState> &ranks, const mag_options &opt) {
    SolveOutcome o = pcg_drive(ctx, ranks, opt);
    if (opt.compat && o.hs.stop == 2 && o.hs.best_iter != o.hs.iter) {
        mag_options again = opt;
        again.max_iter = o.hs.best_iter;
        const SolveOutcome b = pcg_drive(ctx, ranks, again);
        o.hs.best_rr = b.hs.iter ? b.hs.pair[b.hs.iter & 1][1] : b.bb;
        o.rerun_best = true;
    }
    return o;
}

static void fill_solve_stats(mag_stats &st, const SolveOutcome &o, uint32_t n_glob) {
    const PcgScalars &hs = o.hs;
    const int last = (int)(hs.iter & 1);      // pair[] slot written by the last completed iteration
    st.iters = hs.iter;
    st.b_norm = std::sqrt(o.bb);
    st.final_residual = std::sqrt(o.rerun_best ? hs.best_rr : (hs.iter ? hs.pair[last][1] : o.bb));
    st.converged = (hs.stop == 1) || n_glob == 0;
    st.ms_coarse_setup = o.ms_coarse_setup;
    st.n_coarse = o.n_coarse;
    st.precond_used = (uint32_t)o.precond_used;
    for (int i = 0; i < 8; ++i) st.prof[i] = hs.prof[7] > 0 && i < 7 ? hs.prof[i] / hs.prof[7] : hs.prof[i];
    st.negative_definite = hs.first_pq < 0.0;
}

// Post-processing shared by both drivers: full displacement field from the full solution
// vector, reactions for the owned nodes, stresses for all elements (solver.rs:444-482, 496-535).
struct PostBuffers {
    DevBuf<double> xfull, ux, uy, fx, fy, stress, sigma;
};

// side_out (one rank, host result arrays): ux, uy leave on the side stream while the reactions are computed, and
// fx, fy while the stresses are; the caller joins the streams (aux_join) before it synchronises.
static void post_local(mag_ctx *ctx, mag_system *S, PostBuffers &B, bool want_sigma, mag_result *side_out = nullptr) {
    const size_t N = S->n_nodes, E = S->n_elems;
    if (N) {
        MAG_LAUNCH(ctx, scatter_solution_kernel, cdiv(N, 256), 256, 0, (const uint8_t *)S->known.p,
                   (const uint32_t *)S->colmap.p, (const double *)S->bc_ux.p, (const double *)S->bc_uy.p,
                   (const double *)B.xfull.p, N, B.ux.p, B.uy.p);
        if (side_out) {
            aux_fork(ctx);
            MAG_CUDA(cudaMemcpyAsync(side_out->ux, B.ux.p, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->aux_stream));
            MAG_CUDA(cudaMemcpyAsync(side_out->uy, B.uy.p, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->aux_stream));
        }
        const uint32_t n_owned_dof = 2 * (S->K.node_hi - S->K.node_lo);
        if (n_owned_dof)
            MAG_LAUNCH(ctx, reactions_kernel, cdiv(n_owned_dof, 256), 256, 0, (const uint32_t *)S->K.browptr.p,
                       (const uint32_t *)S->K.bcol.p, (const double *)S->K.bval.p, S->K.node_lo, n_owned_dof,
                       (const uint8_t *)S->known.p, (const double *)S->bc_fx.p, (const double *)S->bc_fy.p,
                       (const double *)B.ux.p, (const double *)B.uy.p, B.fx.p, B.fy.p);
        if (side_out) {
            aux_fork(ctx);
            MAG_CUDA(cudaMemcpyAsync(side_out->fx, B.fx.p, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->aux_stream));
            MAG_CUDA(cudaMemcpyAsync(side_out->fy, B.fy.p, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->aux_stream));
        }
    }
    if (E) {
        upload_material(ctx, S->mat);
        MAG_LAUNCH(ctx, stress_kernel, cdiv(E, 256), 256, 0, (const double2 *)S->xy.p,
                   (const uint32_t *)S->n0.p, (const uint32_t *)S->n1.p, (const uint32_t *)S->n2.p, E,
                   (const double *)B.ux.p, (const double *)B.uy.p, B.stress.p, want_sigma ? B.sigma.p : nullptr);
    }
}

static void check_result_args(const mag_system *S, const mag_result *out) {
    if (!out) fail(MAG_ERR_BAD_ARG, "null result");
    // The reduced row and column numberings must coincide DOF by DOF: the Jacobi diagonal, the solution
    // scatter and the row-block halo layout all rely on it.  The counts alone (what the reference compares,
    // solver.rs:380-396) do not guarantee it.
    if (!S->bc_paired)
        fail(MAG_ERR_BAD_BC, "inconsistent boundary conditions: some DOF has both its displacement and its force "
                             "known, or neither; the solve needs exactly one of the two per DOF "
                             "(the reference's mesher enforces this, mesher.rs:881-900)");
    if (S->n_nodes && (!out->ux || !out->uy || !out->fx || !out->fy)) fail(MAG_ERR_BAD_ARG, "result: ux, uy, fx, fy are required");
    if (S->n_elems && !out->stress) fail(MAG_ERR_BAD_ARG, "result: stress is required");
}

// scope 1 (multi-rank): only this rank's slice of every array — nodes by mag_partition_nodes, elements split evenly.
// nodal false: ux, uy, fx, fy have already left on the side stream (post_local).
static void download_result(mag_ctx *ctx, const mag_system *S, PostBuffers &B, mag_result *out, bool want_sigma,
                            int scope = 0, bool nodal = true) {
    const size_t N = S->n_nodes, E = S->n_elems;
    const bool odev = out->on_device != 0;
    size_t n0 = 0, n1 = N, e0 = 0, e1 = E;
    if (scope == 1 && S->nranks > 1) {
        uint64_t lo, hi;
        partition_nodes(N, S->nranks, S->rank, &lo, &hi);
        n0 = lo; n1 = hi;
        partition_nodes(E, S->nranks, S->rank, &lo, &hi);
        e0 = lo; e1 = hi;
    }
    if (nodal) {
        copy_from_device(ctx, out->ux + n0, (const double *)B.ux.p + n0, n1 - n0, odev);
        copy_from_device(ctx, out->uy + n0, (const double *)B.uy.p + n0, n1 - n0, odev);
        copy_from_device(ctx, out->fx + n0, (const double *)B.fx.p + n0, n1 - n0, odev);
        copy_from_device(ctx, out->fy + n0, (const double *)B.fy.p + n0, n1 - n0, odev);
    }
    copy_from_device(ctx, out->stress + e0, (const double *)B.stress.p + e0, e1 - e0, odev);
    if (want_sigma) copy_from_device(ctx, out->sigma + 3 * e0, (const double *)B.sigma.p + 3 * e0, (e1 - e0) * 3, odev);
}

static void raise_solver_status(const SolveOutcome &o, const mag_stats &st, bool compat = false) {
    if (o.hs.stop == 4)
        fail(MAG_ERR_NCCL, "a peer GPU never delivered its share of a dot product (iteration %llu): "
                           "a rank has crashed or the ranks are out of step", (unsigned long long)o.hs.iter);
    if (o.hs.stop == 3)
        fail(MAG_ERR_INDEFINITE, "conjugate gradient broke down at iteration %llu (p.Ap = %g, r.r = %g)",
             (unsigned long long)o.hs.iter, o.hs.pq, st.final_residual * st.final_residual);
    if (o.hs.stop == 2 && !compat)      // the reference returns Ok(best_param) when max_iters ends the run
        fail(MAG_ERR_NOT_CONVERGED, "conjugate gradient stopped at max_iter = %llu with ||r|| = %g",
             (unsigned long long)o.hs.iter, st.final_residual);
}

// Production solve: this process owns one rank of S->nranks.
static void solve_impl(mag_system *S, const mag_options *opt_in, mag_result *out, mag_stats *stats_out) {
    mag_ctx *ctx = S->ctx;
    mag_options opt;
    if (opt_in) opt = *opt_in; else mag_options_default(&opt);
    check_result_args(S, out);
    const size_t N = S->n_nodes, E = S->n_elems;
    mag_stats st = S->stats;
    const uint64_t launches_before = ctx->launches;
    EventTimer phase(ctx->stream);

    phase.start();
    if (S->nranks > 1) setup_halo_ipc(ctx, S);
    std::vector<RankState> ranks(1);
    rank_alloc(ctx, ranks[0], S);
    SolveOutcome o = pcg_drive_best(ctx, ranks, opt);
    fill_solve_stats(st, o, S->n_free);
    st.ms_solve = phase.stop();

    phase.start();
    const bool want_sigma = out->sigma != nullptr;
    PostBuffers B;
    struct AuxGuard { mag_ctx *c; ~AuxGuard() { aux_drain(c); } } aux_guard{ctx};   // error paths: before B is handed back
    B.xfull.alloc(ctx, ext_len(S));
    B.ux.alloc(ctx, N); B.uy.alloc(ctx, N); B.fx.alloc(ctx, N); B.fy.alloc(ctx, N);
    B.stress.alloc(ctx, E);
    if (want_sigma) B.sigma.alloc(ctx, E * 3);
    if (ranks[0].n)
        MAG_CUDA(cudaMemcpyAsync(B.xfull.p + S->row_lo, ranks[0].x.p, (size_t)ranks[0].n * sizeof(double),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
    allgather_slices(ctx, B.xfull.p, S->all_row_lo);
    // one rank, host result arrays: the nodal fields cross PCIe beside the remaining post-processing kernels
    const bool side = S->nranks == 1 && !out->on_device && N >= (1u << 16) && !(ctx->tune & 2048);
    post_local(ctx, S, B, want_sigma, side ? out : nullptr);
    if (S->nranks > 1) {          // every rank returns complete fx, fy
        allgather_slices(ctx, B.fx.p, S->all_node_lo);
        allgather_slices(ctx, B.fy.p, S->all_node_lo);
    }
    st.ms_post = phase.stop();

    phase.start();
    download_result(ctx, S, B, out, want_sigma, opt.result_scope, !side);
    if (side) aux_join(ctx);
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (side) ctx->aux_busy = false;
    st.ms_download = phase.stop();
    st.kernel_launches = ctx->launches - launches_before;
    S->stats.iters = st.iters;
    if (stats_out) *stats_out = st;
    raise_solver_status(o, st, opt.compat != 0);
}

// Single-process emulation of an R-rank solve on one GPU (tests): the mesh is assembled R
// times, once per row block, and the blocks are driven in lockstep.
static void virtual_solve_impl(mag_ctx *ctx, const mag_mesh *mesh, const mag_material *mat,
                               const mag_options *opt_in, int R, mag_result *out, mag_stats *stats_out) {
    if (R < 1 || R > 16) fail(MAG_ERR_BAD_ARG, "virtual ranks must be in [1,16]");
    mag_options opt;
    if (opt_in) opt = *opt_in; else mag_options_default(&opt);
    std::vector<std::unique_ptr<mag_system>> sys(R);
    uint64_t launches = 0;
    for (int r = 0; r < R; ++r) {
        sys[r].reset(new mag_system);
        assemble_impl(ctx, mesh, mat, &opt, sys[r].get(), r, R);
        launches += ctx->launches;
        ctx->launches = 0;
    }
    check_result_args(sys[0].get(), out);
    std::vector<uint32_t> elo(R), ehi(R);
    std::vector<double *> slabs(R);
    for (int r = 0; r < R; ++r) {
        ensure_shared_slab(sys[r].get());
        elo[r] = sys[r]->ext_lo; ehi[r] = sys[r]->ext_hi; slabs[r] = sys[r]->shared_slab;
    }
    for (int r = 0; r < R; ++r) build_push_segments(sys[r].get(), elo, ehi, slabs);
    std::vector<RankState> ranks(R);
    for (int r = 0; r < R; ++r) rank_alloc(ctx, ranks[r], sys[r].get());
    EventTimer phase(ctx->stream);
    phase.start();
    SolveOutcome o = pcg_drive_best(ctx, ranks, opt);
    mag_stats st = sys[0]->stats;
    fill_solve_stats(st, o, sys[0]->n_free);
    st.ms_solve = phase.stop();
    st.nnz = 0; st.nnz_structural = 0;
    for (int r = 0; r < R; ++r) { st.nnz += sys[r]->Kff.nnz; st.nnz_structural += (uint64_t)sys[r]->K.n_blocks * 4; }

    const size_t N = sys[0]->n_nodes, E = sys[0]->n_elems;
    const bool want_sigma = out->sigma != nullptr;
    PostBuffers B;
    B.xfull.alloc(ctx, ext_len(sys[0].get()));
    B.ux.alloc(ctx, N); B.uy.alloc(ctx, N); B.fx.alloc(ctx, N); B.fy.alloc(ctx, N);
    B.stress.alloc(ctx, E);
    if (want_sigma) B.sigma.alloc(ctx, E * 3);
    for (int r = 0; r < R; ++r)
        if (ranks[r].n)
            MAG_CUDA(cudaMemcpyAsync(B.xfull.p + sys[r]->row_lo, ranks[r].x.p, (size_t)ranks[r].n * sizeof(double),
                                     cudaMemcpyDeviceToDevice, ctx->stream));
    for (int r = 0; r < R; ++r) post_local(ctx, sys[r].get(), B, want_sigma);   // each fills its own fx, fy rows
    download_result(ctx, sys[0].get(), B, out, want_sigma);
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    st.kernel_launches = launches + ctx->launches;
    if (stats_out) *stats_out = st;
    raise_solver_status(o, st, opt.compat != 0);
}

}  // namespace mag
