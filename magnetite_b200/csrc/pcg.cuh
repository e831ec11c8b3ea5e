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
//      [multi-GPU: allreduce pq]
//   B: alpha = rz/pq;  x += alpha p;  r -= alpha q;  {rz', rr} = {r.Dinv r, r.r}
//      r is stored by GLOBAL reduced row; boundary entries are also stored
//      straight into the neighbouring GPUs' copies over NVLink (peer pointers),
//      so the halo exchange costs no extra launch and no extra synchronisation:
//      the allreduce that follows orders it.
//      [multi-GPU: allreduce {rz', rr}]
//   C: iteration count + stop test; beta = rz'/rz;  p = Dinv r + beta p over the
//      owned rows AND the halo rows (each GPU updates its own copy of the halo
//      of p from the halo of r it was sent).
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
    double pair[2][2];   // pair[parity] = {r.z, r.r} entering an iteration of that parity (global sums)
    double pq;           // global p.q
    double loc_pair[2];  // this rank's partial sums (send buffers of the NCCL fallback)
    double loc_pq;
    double thr2;         // stop when r.r <= thr2
    double first_pq;     // its sign tells negative-definite systems
    unsigned long long iter, max_iter;
    unsigned long long epoch;   // solve counter: makes mailbox sequence numbers unique across solves
    int stop;            // 1: converged, 2: max_iter, 3: breakdown, 4: peer timeout
    unsigned ticket_a, ticket_b;
    int pad;
};

// ---- allreduce over peer memory ---------------------------------------------
// Every rank owns a mailbox (in its IPC-exported slab).  The last CTA of the kernel
// that finishes a local dot product stores {v0, v1, seq} into slot [kind][parity][me]
// of EVERY rank's mailbox (NVLink peer stores); the kernel that needs the global value
// polls the R slots of its OWN mailbox and adds them in rank order, so all ranks get
// the bit-identical sum.  seq = epoch<<32 | iteration+1 never repeats, nothing is ever
// reset, and a slot cannot be overwritten before it is consumed: the writer's next
// message of the same kind depends on a message the reader sends after consuming.
// One producer fence.sys + flag store orders the halo stores of the whole kernel.
constexpr int kMaxRanks = 16;
struct MailSlot { double v[2]; unsigned long long seq; unsigned long long pad; };
constexpr int kMailSlots = 2 * 2 * kMaxRanks;     // [kind][parity][src]
struct PeerLinks {
    int n = 0, me = 0;                 // n == 0: single rank (or NCCL / emulated allreduce)
    MailSlot *box[kMaxRanks];          // box[r]: rank r's mailbox as mapped in this process
};
enum { kMailPq = 0, kMailPair = 1 };

__device__ __forceinline__ unsigned long long mail_seq(const PcgScalars *sc, unsigned long long it_plus) {
    return (sc->epoch << 32) | it_plus;
}

// Called by all threads of the CTA that holds the local sums (v0, v1 valid in thread 0).
__device__ __forceinline__ void mailbox_post(const PeerLinks &L, int kind, int parity, double v0, double v1,
                                             unsigned long long seq) {
    __shared__ double sv[2];
    if (threadIdx.x == 0) { sv[0] = v0; sv[1] = v1; }
    __syncthreads();
    if ((int)threadIdx.x < L.n) {
        volatile MailSlot *s = L.box[threadIdx.x] + (kind * 2 + parity) * kMaxRanks + L.me;
        __threadfence_system();        // everything this kernel stored (incl. halo) before the flag
        s->v[0] = sv[0];
        s->v[1] = sv[1];
        __threadfence_system();
        s->seq = seq;
    }
}

// Called by all threads of a CTA; returns the global sums.  A peer that never answers
// (crashed rank) trips the timeout instead of hanging the GPU.
__device__ __forceinline__ double2 mailbox_gather(const PeerLinks &L, int kind, int parity,
                                                  unsigned long long seq, PcgScalars *sc) {
    __shared__ double sg[2];
    if (threadIdx.x == 0) {
        const volatile MailSlot *mine = L.box[L.me] + (kind * 2 + parity) * kMaxRanks;
        double a = 0.0, b = 0.0;
        const long long t0 = clock64();
        for (int r = 0; r < L.n; ++r) {
            while (mine[r].seq != seq) {
                if (clock64() - t0 > 60000000000ll) { sc->stop = 4; break; }   // ~30 s
            }
            __threadfence_system();
            a += mine[r].v[0];
            b += mine[r].v[1];
        }
        sg[0] = a; sg[1] = b;
    }
    __syncthreads();
    return make_double2(sg[0], sg[1]);
}

constexpr int kMaxPush = 16;
// Index ranges [lo,hi) of MY rows (global reduced index) that other ranks read as
// halo, and the base pointer of the destination rank's global-indexed arrays.
struct PushSegs {
    int n = 0;
    uint32_t lo[kMaxPush], hi[kMaxPush];
    double *r_dst[kMaxPush];
    double *dinv_dst[kMaxPush];
};

__device__ __forceinline__ void push_value(const PushSegs &ps, uint32_t gi, double v, bool dinv) {
    for (int s = 0; s < ps.n; ++s)
        if (gi >= ps.lo[s] && gi < ps.hi[s]) (dinv ? ps.dinv_dst[s] : ps.r_dst[s])[gi] = v;
}

__global__ void __launch_bounds__(256, 6)
pcg_spmv_kernel(const uint32_t *__restrict__ slice_off, const int32_t *__restrict__ scol,
                const double *__restrict__ sval, const double *__restrict__ p,
                double *__restrict__ q, uint32_t n_rows, uint32_t n_slices, uint32_t row_lo,
                int parity, PeerLinks links, double *__restrict__ partials, PcgScalars *sc,
                double *pq_out) {
    if (sc->stop) return;
    double v[1] = {sell_rows<true>(slice_off, scol, sval, p, q, n_rows, n_slices, row_lo)};
    double tot[1] = {0.0};
    const bool last = grid_sum_256<1>(v, partials, &sc->ticket_a, tot);
    if (links.n) {
        if (grid_is_last_cta()) mailbox_post(links, kMailPq, parity, tot[0], 0.0, mail_seq(sc, sc->iter + 1));
    } else if (last) {
        *pq_out = tot[0];
    }
}

// same, scalar CSR (format comparison)
__global__ void __launch_bounds__(256)
pcg_spmv_csr_kernel(const uint32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                    const double *__restrict__ val, const double *__restrict__ p,
                    double *__restrict__ q, uint32_t n_rows, uint32_t row_lo, int parity,
                    PeerLinks links, double *__restrict__ partials, PcgScalars *sc, double *pq_out) {
    if (sc->stop) return;
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
        if (grid_is_last_cta()) mailbox_post(links, kMailPq, parity, tot[0], 0.0, mail_seq(sc, sc->iter + 1));
    } else if (last) {
        *pq_out = tot[0];
    }
}

__global__ void __launch_bounds__(256)
pcg_update_xr_kernel(double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
                     const double *__restrict__ q, const double *__restrict__ dinv, uint32_t n,
                     uint32_t row_lo, int parity, PushSegs push, PeerLinks links,
                     double *__restrict__ partials, PcgScalars *sc, double *pair_out) {
    if (sc->stop) return;
    double pq = sc->pq;
    if (links.n) {
        pq = mailbox_gather(links, kMailPq, parity, mail_seq(sc, sc->iter + 1), sc).x;
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->pq = pq;
    }
    const double alpha = sc->pair[parity][0] / pq;
    double v[2] = {0.0, 0.0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t gi = row_lo + i;
        const double xi = fma(alpha, p[gi], x[i]);
        const double ri = fma(-alpha, q[i], r[gi]);
        x[i] = xi;
        r[gi] = ri;
        if (push.n) push_value(push, gi, ri, false);
        v[0] = fma(ri * dinv[gi], ri, v[0]);   // r.z with z = Dinv r
        v[1] = fma(ri, ri, v[1]);
    }
    double tot[2] = {0.0, 0.0};
    const bool last = grid_sum_256<2>(v, partials, &sc->ticket_b, tot);
    if (links.n) {
        if (grid_is_last_cta()) mailbox_post(links, kMailPair, parity, tot[0], tot[1], mail_seq(sc, sc->iter + 1));
    } else if (last) {
        pair_out[0] = tot[0];
        pair_out[1] = tot[1];
    }
}

// Rows [ext_lo, ext_hi) = owned block plus halo.
__global__ void __launch_bounds__(256)
pcg_update_p_kernel(double *__restrict__ p, const double *__restrict__ r,
                    const double *__restrict__ dinv, uint32_t ext_lo, uint32_t ext_hi, int parity,
                    PeerLinks links, PcgScalars *sc) {
    if (sc->stop) return;
    double rz_new = sc->pair[parity ^ 1][0], rr = sc->pair[parity ^ 1][1];
    if (links.n) {
        const double2 g = mailbox_gather(links, kMailPair, parity, mail_seq(sc, sc->iter + 1), sc);
        rz_new = g.x; rr = g.y;
    }
    const double rz_old = sc->pair[parity][0], pq = sc->pq;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->pair[parity ^ 1][0] = rz_new;
        sc->pair[parity ^ 1][1] = rr;
        const unsigned long long it = sc->iter + 1;
        if (sc->iter == 0) sc->first_pq = pq;
        sc->iter = it;
        if (sc->stop == 4) {}                                   // a peer never answered
        else if (!(pq != 0.0) || !(rr == rr)) sc->stop = 3;     // breakdown / NaN
        else if (rr <= sc->thr2) sc->stop = 1;
        else if (it >= sc->max_iter) sc->stop = 2;
    }
    const double beta = rz_new / rz_old;
    const uint32_t n = ext_hi - ext_lo;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t gi = ext_lo + i;
        p[gi] = fma(beta, p[gi], r[gi] * dinv[gi]);
    }
}

// x = 0, r = b, Dinv from the diagonal (both pushed to the neighbours), {r.z, r.r}
__global__ void __launch_bounds__(256)
pcg_init_kernel(double *__restrict__ x, double *__restrict__ r, double *__restrict__ dinv,
                const double *__restrict__ b, const double *__restrict__ diag, int jacobi, uint32_t n,
                uint32_t row_lo, PushSegs push, PeerLinks links, double *__restrict__ partials,
                PcgScalars *sc, double *pair_out) {
    double v[2] = {0.0, 0.0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t gi = row_lo + i;
        const double d = diag[i];
        const double di = (jacobi && d != 0.0) ? 1.0 / d : 1.0;
        const double bi = b[i];
        dinv[gi] = di;
        x[i] = 0.0;
        r[gi] = bi;
        if (push.n) { push_value(push, gi, bi, false); push_value(push, gi, di, true); }
        v[0] = fma(bi * di, bi, v[0]);
        v[1] = fma(bi, bi, v[1]);
    }
    double tot[2] = {0.0, 0.0};
    const bool last = grid_sum_256<2>(v, partials, &sc->ticket_b, tot);
    if (links.n) {
        if (grid_is_last_cta()) mailbox_post(links, kMailPair, 0, tot[0], tot[1], mail_seq(sc, 0));
    } else if (last) {
        pair_out[0] = tot[0];
        pair_out[1] = tot[1];
    }
}

// p = Dinv r over owned + halo rows (after the neighbours' pushes are visible)
__global__ void __launch_bounds__(256)
pcg_init_p_kernel(double *__restrict__ p, const double *__restrict__ r, const double *__restrict__ dinv,
                  uint32_t ext_lo, uint32_t ext_hi, PeerLinks links, PcgScalars *sc) {
    if (links.n) {
        const double2 g = mailbox_gather(links, kMailPair, 0, mail_seq(sc, 0), sc);
        if (blockIdx.x == 0 && threadIdx.x == 0) { sc->pair[0][0] = g.x; sc->pair[0][1] = g.y; }
    }
    const uint32_t n = ext_hi - ext_lo;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        p[ext_lo + i] = r[ext_lo + i] * dinv[ext_lo + i];
}

// Single-process emulation of an allreduce(sum) over R virtual ranks: sums `count`
// doubles at byte offset `off` of every rank's PcgScalars in rank order and
// writes the result to offset dst of all of them.  (Tests only.)
struct ScalPtrs { int n; PcgScalars *p[16]; };
__global__ void emulated_allreduce_kernel(ScalPtrs sp, int src_off, int dst_off, int count) {
    const int j = threadIdx.x;
    if (j >= count) return;
    double s = 0.0;
    for (int r = 0; r < sp.n; ++r) s += reinterpret_cast<double *>(sp.p[r])[src_off + j];
    for (int r = 0; r < sp.n; ++r) reinterpret_cast<double *>(sp.p[r])[dst_off + j] = s;
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
