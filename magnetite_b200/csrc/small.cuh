// small.cuh — conjugate gradients for systems that fit ONE thread-block cluster's shared memory.
//
// The reference's own examples (examples/linkedin-logo: ~6 k unknowns, K_ff ~1 MB; the only timing the
// reference publishes, readme.md:28) are far too small for the three-kernels-per-iteration loop of pcg.cuh:
// there an iteration costs 15.6 us of launch latency and moves 1 MB.  Here the whole solve is one kernel on
// one cluster of 8 (or 16) CTAs:
//   * every CTA keeps its block of rows of K_ff (CSR, 16-bit columns) in its shared memory for the whole solve,
//     and full copies of p and z = D^-1 r; x, r, q and D^-1 of its rows live in registers, one row per thread;
//   * q = K p reads shared memory only.  An iteration has TWO cluster barriers: one behind the partial sums of
//     p.q, one behind the partial sums of {r.z, r.r} AND the new entries of z, all of them stores into the other
//     CTAs' shared memory (distributed shared memory).  p = z + beta p is then recomputed by every CTA for the
//     whole vector from its own copies (a few entries per thread), so p itself is never exchanged;
//   * the sums are added in a fixed order (lanes, warps, then CTAs by rank), so every CTA holds the same bits,
//     takes the same stop decision, and runs are bit-identical.
// Same recurrence and stop rules as pcg.cuh (compat: plain CG, absolute cost; solver.rs:141-157).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "pcg.cuh"

namespace mag {

namespace cg = cooperative_groups;

constexpr int kSmallThreads = 1024;
constexpr int kSmallMaxCluster = 16;

struct SmallArgs {
    const uint32_t *rowptr;
    const int32_t *col;
    const double *val, *b, *diag;
    double *x;
    uint32_t n, rows_per_cta, nnz_cap, n_pad;
    int jacobi, compat;
    double rel_tol2, abs_thr2;          // stop when r.r <= rel_tol2 * b.b (compat: <= abs_thr2)
    unsigned long long max_iter;
    PcgScalars *out;                    // iter, stop, pair[iter&1][1] = r.r, first_pq, best_rr, best_iter, pair[0] of the start
};

// Sum of one value per thread over the whole cluster; every thread of every CTA gets the same bits.
// slot: which of the two exchange buffers (consecutive sums alternate, a barrier lies between reuse).
template <int NV>
__device__ __forceinline__ void cluster_sum(cg::cluster_group &cluster, double (&v)[NV], double (*xch)[3][kSmallMaxCluster],
                                            double (*wred)[32], int slot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
        if (lane == 0) wred[i][warp] = v[i];
    }
    __syncthreads();
    const unsigned rank = cluster.block_rank(), C = cluster.num_blocks();
    if (warp < NV) {                                                 // warp i adds the 32 warp sums of value i: fixed tree
        double s = wred[warp][lane];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if ((unsigned)lane < C) {                                    // lane c: my partial into CTA c's buffer
            double(*remote)[3][kSmallMaxCluster] = cluster.map_shared_rank(xch, (unsigned)lane);
            remote[slot][warp][rank] = s;
        }
    }
    cluster.sync();
    // the C partials of each value, added by a fixed shuffle tree over 16 lanes: the same bits in every CTA
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = (unsigned)(lane & 15) < C ? xch[slot][i][lane & 15] : 0.0;
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        v[i] = s;
    }
}

__global__ void __launch_bounds__(kSmallThreads, 1)
small_cg_kernel(SmallArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), C = cluster.num_blocks();
    extern __shared__ __align__(16) unsigned char small_smem[];
    double *p_full = reinterpret_cast<double *>(small_smem);                  // n_pad
    double *z_full = p_full + a.n_pad;                                        // n_pad
    double *sval = z_full + a.n_pad;                                          // nnz_cap
    uint32_t *srow = reinterpret_cast<uint32_t *>(sval + a.nnz_cap);          // rows_per_cta + 1 (+ pad)
    uint16_t *scol = reinterpret_cast<uint16_t *>(srow + a.rows_per_cta + 2); // nnz_cap
    __shared__ double xch[2][3][kSmallMaxCluster];
    __shared__ double wred[3][32];

    const uint32_t r0 = min(a.n, rank * a.rows_per_cta), r1 = min(a.n, r0 + a.rows_per_cta);
    const uint32_t i = r0 + threadIdx.x;
    const bool mine = i < r1;
    const uint32_t e0 = a.rowptr[r0], e1 = a.rowptr[r1];
    for (uint32_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        sval[e - e0] = a.val[e];
        scol[e - e0] = (uint16_t)a.col[e];
    }
    for (uint32_t t = threadIdx.x; t <= r1 - r0; t += blockDim.x) srow[t] = a.rowptr[r0 + t] - e0;
    cluster.sync();                     // every CTA of the cluster is running: its shared memory may be written now

    double xi = 0.0, ri = 0.0, di = 1.0, zi = 0.0;
    if (mine) {
        ri = a.b[i];
        const double d = a.diag[i];
        di = (a.jacobi && d != 0.0) ? 1.0 / d : 1.0;
        zi = ri * di;
        for (unsigned c = 0; c < C; ++c) cluster.map_shared_rank(z_full, c)[i] = zi;
    }
    int slot = 0;
    double v2[2] = {mine ? ri * zi : 0.0, mine ? ri * ri : 0.0};
    cluster_sum<2>(cluster, v2, xch, wred, slot); slot ^= 1;      // its barrier also delivers every CTA's z
    double rz = v2[0], rr = v2[1];
    const double bb = rr;
    const double thr2 = a.compat ? a.abs_thr2 : a.rel_tol2 * bb;
    for (uint32_t j = threadIdx.x; j < a.n; j += blockDim.x) p_full[j] = z_full[j];
    __syncthreads();

    unsigned long long iter = 0, best_iter = 0;
    double best_rr = bb, first_pq = 0.0, pq_last = 0.0;
    int stop = 0;
    if (!(bb == bb)) stop = 3;
    else if (bb <= thr2) stop = 1;
    else if (a.max_iter == 0) stop = 2;
    while (!stop) {
        double qi = 0.0, pi = 0.0;
        if (mine) {
            for (uint32_t k = srow[threadIdx.x]; k < srow[threadIdx.x + 1]; ++k) qi = fma(sval[k], p_full[scol[k]], qi);
            pi = p_full[i];
        }
        double v1[1] = {pi * qi};
        cluster_sum<1>(cluster, v1, xch, wred, slot); slot ^= 1;
        const double pq = v1[0];
        pq_last = pq;
        if (iter == 0) first_pq = pq;
        const double alpha = rz / pq;
        xi = fma(alpha, pi, xi);
        ri = fma(-alpha, qi, ri);
        zi = ri * di;
        if (mine)
            for (unsigned c = 0; c < C; ++c) cluster.map_shared_rank(z_full, c)[i] = zi;
        double w2[2] = {ri * zi, ri * ri};
        cluster_sum<2>(cluster, w2, xch, wred, slot); slot ^= 1;  // its barrier also delivers every CTA's new z
        const double rz_new = w2[0];
        rr = w2[1];
        ++iter;
        if (rr < best_rr) { best_rr = rr; best_iter = iter; }
        if (!(pq != 0.0) || !(rr == rr)) stop = 3;                   // breakdown / NaN
        else if (rr <= thr2) stop = 1;
        else if (iter >= a.max_iter) stop = 2;
        if (stop) break;                                             // the same decision in every CTA
        const double beta = rz_new / rz;
        rz = rz_new;
        // every CTA updates its whole copy of p (nobody writes z again before all CTAs have passed the next barrier)
        for (uint32_t j = threadIdx.x; j < a.n; j += blockDim.x) p_full[j] = fma(beta, p_full[j], z_full[j]);
        __syncthreads();
    }
    if (mine) a.x[i] = xi;
    if (rank == 0 && threadIdx.x == 0) {
        PcgScalars *o = a.out;
        o->iter = iter; o->stop = stop; o->first_pq = first_pq; o->pq = pq_last;
        o->best_rr = best_rr; o->best_iter = best_iter;
        o->pair[0][0] = 0.0; o->pair[0][1] = 0.0; o->pair[1][0] = 0.0; o->pair[1][1] = 0.0;
        o->pair[iter & 1][1] = rr;
        o->rzc[0] = bb;                                              // b.b for the caller (iter may be 0)
        o->thr2 = thr2;
    }
    cluster.sync();                                                  // nobody leaves while its shared memory may still be written
}

// Shared memory one CTA needs when the rows are spread over `C` CTAs; 0 when the system does not fit that shape.
inline size_t small_cg_smem(uint32_t n, uint64_t nnz, unsigned C, uint32_t *rows_per_cta, uint32_t *nnz_cap, uint32_t *n_pad,
                            const std::vector<uint32_t> &h_rowptr) {
    const uint32_t rpc = (n + C - 1) / C;
    if (rpc > (uint32_t)kSmallThreads || n > 65535u) return 0;
    uint32_t cap = 0;
    for (unsigned c = 0; c < C; ++c) {
        const uint32_t a = std::min(n, c * rpc), b = std::min(n, a + rpc);
        cap = std::max(cap, h_rowptr[b] - h_rowptr[a]);
    }
    cap = (cap + 7u) & ~7u;
    (void)nnz;
    *rows_per_cta = rpc; *nnz_cap = cap; *n_pad = (n + 1u) & ~1u;
    return (size_t)*n_pad * 16 + (size_t)cap * 8 + ((size_t)rpc + 2) * 4 + (size_t)cap * 2 + 16;
}

}  // namespace mag
