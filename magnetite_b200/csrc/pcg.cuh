// pcg.cuh — conjugate gradients on the device (replaces argmin's
// ConjugateGradient + Executor at reference src/solver.rs:141-157), written so
// the same kernels run on one GPU and on a row-block partition over several.
//
// Three kernels per iteration; alpha, beta, the residual, the iteration count
// and the stop flag live in device memory, so an iteration needs no host round
// trip.  `check_every` iterations are captured in one CUDA graph and the host
// only polls the stop flag between graph launches; once the flag is set every
// kernel returns immediately, so the result is the iterate at the exact
// stopping iteration regardless of the chunk size.
//
//   A: q = K p  and  pq = p.q                      (SELL SpMV + fused dot)
//   B: alpha = rz/pq;  x += alpha p;  r -= alpha q;  {rz', rr} = {r.Dinv r, r.r}
//   C: iteration count + stop test; beta = rz'/rz;  p = Dinv r + beta p over the
//      owned rows AND the halo rows.
//
// Multi-GPU (one process per GPU, contiguous row blocks), all of it fused into
// these kernels over NVLink peer memory — no extra launch, no fence, no
// collective call inside the iteration:
//   * halo: B stores the boundary entries of the new r straight into the
//     neighbours' halo buffers; C reads them and updates ITS OWN copy of the halo
//     of p, so p is never sent;
//   * dot products: the last CTA of A / B stores the local sum into a mailbox slot
//     on EVERY rank; B / C poll their own mailbox and add the R slots in rank
//     order, so all ranks hold the bit-identical global value and stop together.
// Every word that crosses NVLink is self-validating ("LL" encoding: an 8-byte
// store carries 4 bytes of payload and a 4-byte sequence number derived from the
// solve epoch and the iteration), so no message needs a memory fence, nothing is
// ever reset, and a reader that finds a stale sequence number simply polls again.
// (A first version used value + flag with __threadfence_system() in between: each
// system-scope fence behind NVLink stores cost 8-16 us, more than NCCL's whole
// allreduce.)  A slot cannot be overwritten before it is consumed: the writer's next
// message to that slot depends on a message the reader sends after consuming it.
//
// Vectors p, r, Dinv are indexed by global reduced row (owned block + halo
// valid); x, q by local row.
//
// compat mode (reference semantics): no preconditioner, x0 = 0, stop when the
// cost (||r||_2, or r.r) is <= 1e-4 absolute or after 1e7 iterations
// (src/solver.rs:18-19, 143, 153-154).  argmin carries r with the opposite
// sign (r = A x - b, p = -r + beta p); the iterates are identical.
#pragma once
#include "common.cuh"
#include "spmv.cuh"

namespace mag {

struct PcgScalars {
    double pair[2][2];   // pair[parity] = {r.Dinv r, r.r} entering an iteration of that parity (global sums)
    double rzc[2];       // rzc[parity] = complete r.z entering an iteration of that parity: pair[parity][0], plus
                         // the coarse part wy with the two-level preconditioner.  Written by ONE thread of the
                         // p-update and read only by LATER launches (never input and output of the same launch).
    double pq;           // global p.q
    double wy;           // two-level preconditioner: (P^T r).(Ac^-1 P^T r), the coarse part of r.z
    double loc_pair[2];  // this rank's partial sums (send buffers of the NCCL fallback)
    double loc_pq;
    double loc_wy;       // this rank's share of w.y (NCCL / emulated reductions)
    double thr2;         // stop when r.r <= thr2
    double first_pq;     // its sign tells negative-definite systems
    unsigned long long iter, max_iter;
    unsigned long long chunk_base;   // iterations completed when the current graph launch started
    unsigned long long epoch;   // solve counter: makes sequence numbers unique across solves
    double best_rr;             // lowest r.r seen so far and the iteration that produced it (argmin's
    unsigned long long best_iter;   // best_param: IterState::update keeps the iterate with the lowest cost)
    int stop;            // 1: converged, 2: max_iter, 3: breakdown, 4: peer timeout
    unsigned ticket_a, ticket_b;
    int tune;            // debug switches (MAG_TUNE): 1 = no halo stores, 2 = no timeline, 4 = poll backoff
    // device-side timeline of the iteration (ns, %globaltimer), accumulated over iterations:
    // 0 A.start-C.start(prev)  1 A.duration  2 B.start-A.end  3 B gather wait  4 B.duration
    // 5 C.start-B.end          6 C gather wait              7 iterations timed
    unsigned long long t_mark;
    double prof[8];
};

__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// thread 0 of block 0 at kernel start: time since the previous mark
__device__ __forceinline__ void prof_start(PcgScalars *sc, int slot) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && !(sc->tune & 2)) {
        const unsigned long long t = gtime_ns();
        if (sc->t_mark) sc->prof[slot] += (double)(t - sc->t_mark);
        sc->t_mark = t;
    }
}
// thread 0 of the last CTA at kernel end
__device__ __forceinline__ void prof_end(PcgScalars *sc, int slot) {
    if (threadIdx.x == 0 && !(sc->tune & 2)) {
        const unsigned long long t = gtime_ns();
        sc->prof[slot] += (double)(t - sc->t_mark);
        sc->t_mark = t;
    }
}

// ---- self-validating 16-byte words over peer memory ------------------------------
struct LLWord { unsigned long long lo, hi; };     // {payload[31:0] | seq<<32, payload[63:32] | seq<<32}

__device__ __forceinline__ void ll_store(LLWord *dst, double v, uint32_t seq) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long s = (unsigned long long)seq << 32;
    const unsigned long long lo = (bits & 0xffffffffull) | s, hi = (bits >> 32) | s;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(lo), "l"(hi) : "memory");
}
// false until both halves carry `seq`
__device__ __forceinline__ bool ll_try_load(const LLWord *src, uint32_t seq, double &v) {
    unsigned long long lo, hi;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
    if ((uint32_t)(lo >> 32) != seq || (uint32_t)(hi >> 32) != seq) return false;
    v = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    return true;
}
// Spins until the word is valid; a peer that never answers (crashed rank) trips the
// timeout instead of hanging the GPU.
__device__ __forceinline__ double ll_wait(const LLWord *src, uint32_t seq, PcgScalars *sc) {
    double v = 0.0;
    const long long t0 = clock64();
    while (!ll_try_load(src, seq, v)) {
        if (clock64() - t0 > 60000000000ll) { sc->stop = 4; break; }    // ~30 s
        if (sc->tune & 4) __nanosleep(100);
    }
    return v;
}

// epoch in the top byte (never 0, so zero-initialised memory is never valid), iteration+1 below
__device__ __forceinline__ uint32_t ll_seq(const PcgScalars *sc, unsigned long long it_plus) {
    return (uint32_t)(((sc->epoch % 255ull) + 1ull) << 24) | (uint32_t)(it_plus & 0xffffffull);
}

// ---- allreduce over peer memory: mailboxes ----------------------------------------
constexpr int kMaxRanks = 16;
enum { kMailPq = 0, kMailPair = 1, kMailInit = 2, kMailWy = 3, kMailKinds = 4 };
struct MailSlot { LLWord w[2]; };
constexpr int kMailSlots = kMailKinds * 2 * kMaxRanks;     // [kind][parity][src]
struct PeerLinks {
    int n = 0, me = 0;                 // n == 0: single rank (or NCCL / emulated allreduce)
    MailSlot *box[kMaxRanks];          // box[r]: rank r's mailbox as mapped in this process
};

// Called by all threads of the CTA that holds the local sums (v0, v1 valid in thread 0):
// thread 2r+j stores value j into rank r's mailbox.
__device__ __forceinline__ void mailbox_post(const PeerLinks &L, int kind, int parity, double v0, double v1,
                                             uint32_t seq) {
    __shared__ double sv[2];
    if (threadIdx.x == 0) { sv[0] = v0; sv[1] = v1; }
    __syncthreads();
    if ((int)threadIdx.x < 2 * L.n) {
        const int r = threadIdx.x >> 1, j = threadIdx.x & 1;
        MailSlot *s = L.box[r] + (kind * 2 + parity) * kMaxRanks + L.me;
        ll_store(&s->w[j], sv[j], seq);
    }
}

// Called by all threads of a CTA; returns the global sums (added in rank order).
__device__ __forceinline__ double2 mailbox_gather(const PeerLinks &L, int kind, int parity, uint32_t seq,
                                                  PcgScalars *sc, int prof_slot) {
    __shared__ double sg[2][kMaxRanks];
    __shared__ double sum[2];
    const unsigned long long g0 = (blockIdx.x == 0 && threadIdx.x == 0) ? gtime_ns() : 0ull;
    if ((int)threadIdx.x < 2 * L.n) {
        const int r = threadIdx.x >> 1, j = threadIdx.x & 1;
        const MailSlot *mine = L.box[L.me] + (kind * 2 + parity) * kMaxRanks;
        sg[j][r] = ll_wait(&mine[r].w[j], seq, sc);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < L.n; ++r) { a += sg[0][r]; b += sg[1][r]; }
        sum[0] = a; sum[1] = b;
        if (blockIdx.x == 0 && prof_slot >= 0 && !(sc->tune & 2)) sc->prof[prof_slot] += (double)(gtime_ns() - g0);
    }
    __syncthreads();
    return make_double2(sum[0], sum[1]);
}

// ---- halo ---------------------------------------------------------------------------
constexpr int kMaxPush = 16;
// Index ranges [lo,hi) of MY rows (global reduced index) that other ranks read as halo.
// ll_dst[s][gi - lo[s]] is the slot of row gi in the destination rank's halo buffer;
// dinv_dst is the destination's global-indexed Dinv (filled once per solve).
struct PushSegs {
    int n = 0;
    uint32_t lo[kMaxPush], hi[kMaxPush];
    LLWord *ll_dst[kMaxPush];
    double *dinv_dst[kMaxPush];
};
// My own halo buffer: rows [ext_lo,row_lo) then [row_hi,ext_hi), compactly.
struct HaloView {
    const LLWord *ll = nullptr;
    uint32_t ext_lo = 0, row_lo = 0, row_hi = 0;
    __device__ __forceinline__ bool is_halo(uint32_t gi) const { return gi < row_lo || gi >= row_hi; }
    __device__ __forceinline__ const LLWord *slot(uint32_t gi) const {
        return ll + (gi < row_lo ? gi - ext_lo : (gi - row_hi) + (row_lo - ext_lo));
    }
};

__device__ __forceinline__ void push_r(const PushSegs &ps, uint32_t gi, double v, uint32_t seq) {
    for (int s = 0; s < ps.n; ++s)
        if (gi >= ps.lo[s] && gi < ps.hi[s]) ll_store(ps.ll_dst[s] + (gi - ps.lo[s]), v, seq);
}
__device__ __forceinline__ void push_dinv(const PushSegs &ps, uint32_t gi, double v) {
    for (int s = 0; s < ps.n; ++s)
        if (gi >= ps.lo[s] && gi < ps.hi[s]) ps.dinv_dst[s][gi] = v;
}

// ---- A ----------------------------------------------------------------------------------
template <class IDX>
__global__ void __launch_bounds__(256, 6)
pcg_spmv_kernel(const uint32_t *__restrict__ slice_off, const IDX *__restrict__ scol,
                const double *__restrict__ sval, const double *__restrict__ p,
                double *__restrict__ q, uint32_t n_rows, uint32_t n_slices, uint32_t row_lo,
                int step, PeerLinks links, double *__restrict__ partials, PcgScalars *sc,
                double *pq_out) {
    pdl_wait();
    if (sc->stop) return;
    prof_start(sc, 0);
    const int parity = step & 1;
    double v[1] = {sell_rows<true, IDX>(slice_off, scol, sval, p, q, n_rows, n_slices, row_lo)};
    double tot[1] = {0.0};
    const bool last = grid_sum_256<1>(v, partials, &sc->ticket_a, tot);
    if (links.n) {
        if (grid_is_last_cta()) mailbox_post(links, kMailPq, parity, tot[0], 0.0, ll_seq(sc, sc->chunk_base + step + 1));
    } else if (last) {
        *pq_out = tot[0];
    }
    if (grid_is_last_cta()) prof_end(sc, 1);
}

// same, scalar CSR (format comparison)
__global__ void __launch_bounds__(256)
pcg_spmv_csr_kernel(const uint32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                    const double *__restrict__ val, const double *__restrict__ p,
                    double *__restrict__ q, uint32_t n_rows, uint32_t row_lo, int step,
                    PeerLinks links, double *__restrict__ partials, PcgScalars *sc, double *pq_out) {
    pdl_wait();
    if (sc->stop) return;
    prof_start(sc, 0);
    const int parity = step & 1;
    double dot = 0.0;
    for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows;
         row += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (uint32_t k = rowptr[row]; k < rowptr[row + 1]; ++k) acc = fma(val[k], __ldg(p + col[k]), acc);
        q[row] = acc;
        dot = fma(__ldg(p + row_lo + row), acc, dot);
    }
    double v[1] = {dot};
    double tot[1] = {0.0};
    const bool last = grid_sum_256<1>(v, partials, &sc->ticket_a, tot);
    if (links.n) {
        if (grid_is_last_cta()) mailbox_post(links, kMailPq, parity, tot[0], 0.0, ll_seq(sc, sc->chunk_base + step + 1));
    } else if (last) {
        *pq_out = tot[0];
    }
    if (grid_is_last_cta()) prof_end(sc, 1);
}

// ---- B ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pcg_update_xr_kernel(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
                     const double *__restrict__ q, const double *__restrict__ dinv, uint32_t n,
                     uint32_t row_lo, int step, PushSegs push, PeerLinks links,
                     double *__restrict__ partials, PcgScalars *sc, double *pair_out) {
    pdl_wait();
    if (sc->stop) return;
    prof_start(sc, 2);
    const int parity = step & 1;
    double pq = sc->pq;
    const uint32_t seq = ll_seq(sc, sc->chunk_base + step + 1);
    if (links.n) {
        pq = mailbox_gather(links, kMailPq, parity, seq, sc, 3).x;
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->pq = pq;
    }
    const bool do_push = push.n && !(sc->tune & 1);
    const double alpha = sc->rzc[parity] / pq;
    double v[2] = {0.0, 0.0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t gi = row_lo + i;
        const double xi = fma(alpha, p[gi], x[i]);
        const double ri = fma(-alpha, q[i], r[gi]);
        x[i] = xi;
        r[gi] = ri;
        if (do_push) push_r(push, gi, ri, seq);
        v[0] = fma(ri * dinv[gi], ri, v[0]);   // r.z with z = Dinv r
        v[1] = fma(ri, ri, v[1]);
    }
    double tot[2] = {0.0, 0.0};
    const bool last = grid_sum_256<2>(v, partials, &sc->ticket_b, tot);
    if (links.n) {
        if (grid_is_last_cta()) mailbox_post(links, kMailPair, parity, tot[0], tot[1], seq);
    } else if (last) {
        pair_out[0] = tot[0];
        pair_out[1] = tot[1];
    }
    if (grid_is_last_cta()) prof_end(sc, 4);
}

// Two-level preconditioner as seen by the p-update: z = Dinv r + P y (coarse.cuh); mode == null: off.
struct CoarseView {
    const uint32_t *mode = nullptr;   // per global reduced row: 3*aggregate + axis
    const float *rot = nullptr;       // rotation-mode coefficient (fp32: see CoarseSpace::rot)
    const double *y = nullptr;        // Ac^-1 P^T r
    __device__ __forceinline__ double prolong(uint32_t gi) const {
        const uint32_t m = mode[gi];
        return __ldg(y + m) + (double)rot[gi] * __ldg(y + (m - m % 3u) + 2u);
    }
};

// ---- C ----------------------------------------------------------------------------------
// Rows [ext_lo, ext_hi) = owned block plus halo; halo rows take r from the halo buffer.
__global__ void __launch_bounds__(256)
pcg_update_p_kernel(double *__restrict__ p, const double *__restrict__ r,
                    const double *__restrict__ dinv, uint32_t ext_lo, uint32_t ext_hi, int step,
                    HaloView halo, CoarseView cv, PeerLinks links, PcgScalars *sc) {
    pdl_wait();
    if (sc->stop) return;
    prof_start(sc, 5);
    const int parity = step & 1;
    const uint32_t seq = ll_seq(sc, sc->chunk_base + step + 1);
    double rz_new = sc->pair[parity ^ 1][0], rr = sc->pair[parity ^ 1][1];
    if (links.n) {
        const double2 g = mailbox_gather(links, kMailPair, parity, seq, sc, 6);
        rz_new = g.x; rr = g.y;
    }
    const double g_rz = rz_new;
    if (cv.mode) {                          // r.z = r.Dinv r + (P^T r).(Ac^-1 P^T r)
        double wy = sc->wy;
        if (links.n) {                      // every rank's share of w.y, added in rank order
            wy = mailbox_gather(links, kMailWy, parity, seq, sc, -1).x;
        }
        rz_new += wy;
    }
    const double rz_old = sc->rzc[parity], pq = sc->pq;
    const bool use_halo = halo.ll != nullptr && !(sc->tune & 1);
    // One thread moves the iteration on.  Nothing another CTA of this launch still reads is
    // touched: sequence numbers come from chunk_base + step (constant during the launch), and a
    // CTA that starts late and already sees the stop flag just skips an update nobody needs.
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->rzc[parity ^ 1] = rz_new;       // a field of its own: pair[parity^1][0] is still being read by late CTAs
        if (links.n) { sc->pair[parity ^ 1][0] = g_rz; sc->pair[parity ^ 1][1] = rr; }   // mailbox mode: inputs came from the mailbox
        sc->prof[7] += 1.0;
        const unsigned long long it = sc->iter + 1;
        if (sc->iter == 0) sc->first_pq = pq;
        sc->iter = it;
        if (rr < sc->best_rr) { sc->best_rr = rr; sc->best_iter = it; }
        if (sc->stop == 4) {}                                   // a peer never answered
        else if (!(pq != 0.0) || !(rr == rr)) sc->stop = 3;     // breakdown / NaN
        else if (rr <= sc->thr2) sc->stop = 1;
        else if (it >= sc->max_iter) sc->stop = 2;
    }
    const double beta = rz_new / rz_old;
    const uint32_t n = ext_hi - ext_lo;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t gi = ext_lo + i;
        const double ri = (use_halo && halo.is_halo(gi)) ? ll_wait(halo.slot(gi), seq, sc) : r[gi];
        double z = ri * dinv[gi];
        if (cv.mode) z += cv.prolong(gi);
        p[gi] = fma(beta, p[gi], z);
    }
}

// Last node of every graph launch: the next launch continues the iteration numbering.
__global__ void pcg_chunk_end_kernel(int chunk, PcgScalars *sc) {
    pdl_wait();
    if (!sc->stop) sc->chunk_base += (unsigned long long)chunk;
}

// ---- init -------------------------------------------------------------------------------
// x = 0, r = b, Dinv from the diagonal (both pushed to the neighbours), {r.z, r.r}
__global__ void __launch_bounds__(256)
pcg_init_kernel(double *__restrict__ x, double *__restrict__ r, double *__restrict__ dinv,
                const double *__restrict__ b, const double *__restrict__ diag, int jacobi, uint32_t n,
                uint32_t row_lo, PushSegs push, PeerLinks links, double *__restrict__ partials,
                PcgScalars *sc, double *pair_out) {
    const uint32_t seq = ll_seq(sc, 0);
    double v[2] = {0.0, 0.0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t gi = row_lo + i;
        const double d = diag[i];
        const double di = (jacobi && d != 0.0) ? 1.0 / d : 1.0;
        const double bi = b[i];
        dinv[gi] = di;
        x[i] = 0.0;
        r[gi] = bi;
        if (push.n) { push_r(push, gi, bi, seq); push_dinv(push, gi, di); }
        v[0] = fma(bi * di, bi, v[0]);
        v[1] = fma(bi, bi, v[1]);
    }
    double tot[2] = {0.0, 0.0};
    // the Dinv halo is plain data: a system-scope fence orders it before the init message (once per solve)
    const bool last = grid_sum_256<2>(v, partials, &sc->ticket_b, tot, /*system_scope=*/true);
    if (links.n) {
        if (grid_is_last_cta()) mailbox_post(links, kMailInit, 0, tot[0], tot[1], seq);
    } else if (last) {
        pair_out[0] = tot[0];
        pair_out[1] = tot[1];
    }
}

// p = Dinv r over owned + halo rows
__global__ void __launch_bounds__(256)
pcg_init_p_kernel(double *__restrict__ p, const double *__restrict__ r, const double *dinv,
                  uint32_t ext_lo, uint32_t ext_hi, HaloView halo, CoarseView cv, PeerLinks links,
                  PcgScalars *sc) {
    const uint32_t seq = ll_seq(sc, 0);
    if (links.n) {
        const double2 g = mailbox_gather(links, kMailInit, 0, seq, sc, -1);
        double wy = 0.0;
        if (cv.mode) wy = mailbox_gather(links, kMailWy, 0, seq, sc, -1).x;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            sc->pair[0][0] = g.x; sc->pair[0][1] = g.y;
            sc->rzc[0] = g.x + wy;
        }
        __threadfence_system();        // acquire side of the Dinv halo
    } else if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->rzc[0] = sc->pair[0][0] + (cv.mode ? sc->wy : 0.0);
    }
    const uint32_t n = ext_hi - ext_lo;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t gi = ext_lo + i;
        const double ri = (halo.ll != nullptr && halo.is_halo(gi)) ? ll_wait(halo.slot(gi), seq, sc) : r[gi];
        double z = ri * __ldcv(dinv + gi);
        if (cv.mode) z += cv.prolong(gi);
        p[gi] = z;
    }
}

// Single-process emulation of an allreduce(sum) over R virtual ranks: sums `count`
// doubles at offset src of every rank's PcgScalars in rank order and writes the result
// to offset dst of all of them.  (Tests only.)
struct ScalPtrs { int n; PcgScalars *p[16]; };
__global__ void emulated_allreduce_kernel(ScalPtrs sp, int src_off, int dst_off, int count) {
    const int j = threadIdx.x;
    if (j >= count) return;
    double s = 0.0;
    for (int r = 0; r < sp.n; ++r) s += reinterpret_cast<double *>(sp.p[r])[src_off + j];
    for (int r = 0; r < sp.n; ++r) reinterpret_cast<double *>(sp.p[r])[dst_off + j] = s;
}

struct VecPtrs { int n; double *p[16]; };
__global__ void emulated_vec_allreduce_kernel(VecPtrs vp, size_t count) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += (size_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < vp.n; ++r) s += vp.p[r][j];
        for (int r = 0; r < vp.n; ++r) vp.p[r][j] = s;
    }
}

// [min, max] of the column indices of a CSR block (halo extent of a rank)
__global__ void col_range_kernel(const int32_t *__restrict__ col, size_t nnz, int *__restrict__ mn,
                                 int *__restrict__ mx) {
    int lo = 0x7fffffff, hi = -1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (size_t)gridDim.x * blockDim.x) {
        const int c = col[i];
        lo = min(lo, c); hi = max(hi, c);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(mn, lo); atomicMax(mx, hi); }
}

}  // namespace mag
