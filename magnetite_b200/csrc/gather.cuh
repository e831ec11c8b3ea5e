// gather.cuh — gather assembly (mag_options.assembly = 1): the full K as 2x2-block CSR without sorting
// COO keys and without materialising K_e.  Same output, bit for bit, as the sort-and-reduce path of
// assembly.cuh (reference src/solver.rs:290-331); see gather_core.h for the per-node algorithm and why
// the accumulation order is the reference's.
//
//   emit    3 (node, incidence) pairs per triangle, incidence = local_element*3 + corner; pairs whose node
//           another rank owns get the sentinel key; per-node counts by integer atomics (deterministic)
//   sort    the stable LSD radix sort of radix_sort.cuh by the log2(N)-bit node key: 3E pairs x 3 passes at
//           8 M nodes, against 9E pairs x 6 passes for the COO keys
//   count   one thread per owned node: distinct column nodes of its row -> browptr (scan)
//   fill    one thread per owned node: recompute the two K_e rows of every incident (element, corner) and
//           add the blocks column by column in ascending (element, corner) order -> bcol, bval
//
// HBM traffic per triangle (plate): 12 B connectivity + 3 x 12 B pairs per sort pass + ~10 x 12 B of
// connectivity re-reads and 16-byte coordinate gathers that mostly hit L1/L2 + 7/2 blocks x 36 B out,
// against 288 B of K_e written and read back plus 9 x 12 B pairs per pass on the sorted-key path.
#pragma once
#include "assembly.cuh"
#include "common.cuh"
#include "element.cuh"
#include "gather_core.h"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace mag {

constexpr int kGatherThreads = 128;

// One thread per local element.  cnt[node - node_lo] counts the incidences of every owned node.
__global__ void emit_incidence_kernel(const uint32_t *__restrict__ n0, const uint32_t *__restrict__ n1,
                                      const uint32_t *__restrict__ n2, const uint32_t *__restrict__ elist,
                                      size_t n_local, uint32_t node_lo, uint32_t node_hi,
                                      uint32_t *__restrict__ keys, uint32_t *__restrict__ payload,
                                      uint32_t *__restrict__ cnt) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    const gather::Conn conn{n0, n1, n2, elist};
    uint32_t nd[3];
    gather::corner_nodes(conn, (uint32_t)i, nd);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const bool mine = nd[k] >= node_lo && nd[k] < node_hi;
        keys[i * 3 + k] = mine ? nd[k] : 0xffffffffu;       // sentinel: sorts behind every node id
        payload[i * 3 + k] = (uint32_t)(i * 3 + k);
        if (mine) atomicAdd(&cnt[nd[k] - node_lo], 1u);          // integer: deterministic
    }
}

// nblk[r] = number of distinct column nodes of owned node row r.
__global__ void __launch_bounds__(kGatherThreads)
gather_count_kernel(const uint32_t *__restrict__ n0, const uint32_t *__restrict__ n1,
                    const uint32_t *__restrict__ n2, const uint32_t *__restrict__ elist,
                    const uint32_t *__restrict__ payload, const uint32_t *__restrict__ nptr, uint32_t n_own,
                    uint32_t *__restrict__ nblk) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_own) return;
    const gather::Conn conn{n0, n1, n2, elist};
    nblk[r] = gather::count_cols(conn, payload, nptr[r], nptr[r + 1]);
}

// Block row r of K: bcol/bval at browptr[r].
__global__ void __launch_bounds__(kGatherThreads)
gather_fill_kernel(const double2 *__restrict__ xy, const uint32_t *__restrict__ n0,
                   const uint32_t *__restrict__ n1, const uint32_t *__restrict__ n2,
                   const uint32_t *__restrict__ elist, const uint32_t *__restrict__ payload,
                   const uint32_t *__restrict__ nptr, uint32_t n_own, const uint32_t *__restrict__ browptr,
                   uint32_t *__restrict__ bcol, double *__restrict__ bval) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_own) return;
    const gather::Conn conn{n0, n1, n2, elist};
    const uint32_t b0 = browptr[r], b1 = browptr[r + 1];
    gather::fill_row(conn, xy, c_mat.D, c_mat.t, payload, nptr[r], nptr[r + 1], b1 - b0, bcol + b0,
                     bval + (size_t)b0 * 4);
}

// ---- fused assembly: node rows straight into K_ff (default, mag_options.assembly = 0) -----------------
// The block row of a node never reaches memory.  Phase 1, one thread per owned node: the row table (distinct
// column nodes + accumulated 2x2 blocks, gather_core.h build_row_table) is built in SHARED memory — the
// thread-local table of gather_fill_kernel lived in local memory and its spills were more than half of that
// kernel's DRAM writes (profiles/r2_gather_kernels_ncu.txt).  Phase 2, one warp per 32 nodes: for each of the
// warp's 64 DOF rows the lanes take one (column node, axis) candidate each, apply the Dirichlet elimination
// (bc.cuh: rows = known force, columns = unknown displacement, prescribed columns go to the rhs in ascending
// order as -(K u), exact zeros dropped: solver.rs:380-396, 427-432, 132), rank the kept entries with a ballot
// and write them as ONE coalesced run.  COUNT pass (FILL = 0): only row lengths; FILL pass: col, val, rhs, diag.
// K_e rows are computed twice — far cheaper than a round trip of K through HBM.
constexpr int kFusedThreads = 128;
constexpr int kAccStride = gather::kFastCols * 4 + 1;     // doubles per thread; odd: conflict-free 64-bit accesses
constexpr int kColStride = gather::kFastCols + 1;         // words per thread
constexpr size_t kFusedSmem = (size_t)kFusedThreads * (kAccStride * sizeof(double) + kColStride * sizeof(uint32_t));
static_assert(2 * gather::kFastCols <= 32, "one lane per (column node, axis) candidate");

struct ElimView {
    const uint8_t *known;
    const uint32_t *rowmap, *colmap;
    const double *ux, *uy, *fx, *fy;
    int drop_zeros;
    uint32_t row_lo, node_lo;
};

// A row with more columns than the table holds: one thread, no table (for_each_block_serial).
template <int FILL>
__device__ void fused_emit_serial(const gather::Conn &conn, const double2 *xy, const uint32_t *pay, uint32_t begin,
                                  uint32_t end, uint32_t node, const ElimView &E, uint32_t *row_nnz,
                                  const uint32_t *rowptr, int32_t *col, double *val, double *rhs, double *diag,
                                  uint32_t *n_cols_out) {
    const uint8_t kn = E.known[node];
    const bool row_on[2] = {(kn & MAG_KNOWN_FX) != 0, (kn & MAG_KNOWN_FY) != 0};
    uint32_t gr[2] = {0, 0}, w[2] = {0, 0}, cnt[2] = {0, 0}, ncols = 0;
    double s[2] = {0.0, 0.0}, dg[2] = {0.0, 0.0};
    for (int a = 0; a < 2; ++a)
        if (row_on[a]) { gr[a] = E.rowmap[2u * node + a]; if (FILL) w[a] = rowptr[gr[a] - E.row_lo]; }
    gather::for_each_block_serial(conn, xy, c_mat.D, c_mat.t, pay, begin, end,
                                  [&](uint32_t cn, double a0, double a1, double a2, double a3) {
        ++ncols;
        const double blk[4] = {a0, a1, a2, a3};
        const uint8_t knc = E.known[cn];
        for (int a = 0; a < 2; ++a) {
            if (!row_on[a]) continue;
            for (int b = 0; b < 2; ++b) {
                const double k = blk[a * 2 + b];
                if ((knc >> b) & 1u) {
                    if (FILL) s[a] = __dadd_rn(s[a], __dmul_rn(__dmul_rn(k, b ? E.uy[cn] : E.ux[cn]), -1.0));
                } else if (!E.drop_zeros || k != 0.0) {
                    if (FILL) {
                        const uint32_t c = E.colmap[2u * cn + b];
                        col[w[a]] = (int32_t)c; val[w[a]] = k;
                        if (c == gr[a]) dg[a] = k;
                        ++w[a];
                    }
                    ++cnt[a];
                }
            }
        }
    });
    for (int a = 0; a < 2; ++a) {
        if (!row_on[a]) continue;
        const uint32_t rr = gr[a] - E.row_lo;
        if (FILL) { rhs[rr] = __dadd_rn(s[a], a ? E.fy[node] : E.fx[node]); diag[rr] = dg[a]; }
        else row_nnz[rr] = cnt[a];
    }
    *n_cols_out = ncols;
}

template <int FILL>
__global__ void __launch_bounds__(kFusedThreads)
fused_rows_kernel(const double2 *__restrict__ xy, const uint32_t *__restrict__ n0, const uint32_t *__restrict__ n1,
                  const uint32_t *__restrict__ n2, const uint32_t *__restrict__ elist,
                  const uint32_t *__restrict__ payload, const uint32_t *__restrict__ nptr, uint32_t n_own, ElimView E,
                  uint32_t *__restrict__ row_nnz, const uint32_t *__restrict__ rowptr, int32_t *__restrict__ col,
                  double *__restrict__ val, double *__restrict__ rhs, double *__restrict__ diag,
                  unsigned long long *__restrict__ n_blocks_total) {
    extern __shared__ __align__(16) unsigned char fused_smem[];
    double *acc_all = reinterpret_cast<double *>(fused_smem);
    uint32_t *cols_all = reinterpret_cast<uint32_t *>(fused_smem + (size_t)kFusedThreads * kAccStride * sizeof(double));
    const int lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = r < n_own;
    const gather::Conn conn{n0, n1, n2, elist};
    double *acc = acc_all + (size_t)threadIdx.x * kAccStride;
    uint32_t *cols = cols_all + (size_t)threadIdx.x * kColStride;
    uint32_t p0 = 0, p1 = 0;
    int ncols = 0;
    if (active) {
        p0 = nptr[r]; p1 = nptr[r + 1];
        ncols = gather::build_row_table(conn, xy, c_mat.D, c_mat.t, payload, p0, p1, cols, acc);
    }
    __syncwarp();
    uint32_t blocks = ncols > 0 ? (uint32_t)ncols : 0u;
    const uint32_t wbase = threadIdx.x & ~31u;
    const uint32_t first = blockIdx.x * blockDim.x + wbase;          // first node row of this warp
#pragma unroll 1
    for (int t = 0; t < 32; ++t) {
        if (first + t >= n_own) break;                               // warp-uniform
        const int nc = __shfl_sync(0xffffffffu, ncols, t);
        const uint32_t node = E.node_lo + first + t;
        if (nc < 0) {                                                // more columns than the table holds
            if (lane == t) {
                uint32_t counted = 0;
                fused_emit_serial<FILL>(conn, xy, payload, p0, p1, node, E, row_nnz, rowptr, col, val, rhs, diag, &counted);
                blocks = counted;
            }
            __syncwarp();
            continue;
        }
        const uint8_t kn = __ldg(E.known + node);
        const double *acc_t = acc_all + (size_t)(wbase + t) * kAccStride;
        const uint32_t *cols_t = cols_all + (size_t)(wbase + t) * kColStride;
        const bool cand = lane < 2 * nc;
        const int j = lane >> 1, b = lane & 1;
        const uint32_t cn = cand ? cols_t[j] : 0u;
        const bool uk = cand && ((__ldg(E.known + cn) >> b) & 1u);
        uint32_t cmap = 0;
        if (FILL && cand && !uk) cmap = __ldg(E.colmap + 2u * cn + b);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            if (!((kn >> (2 + a)) & 1u)) continue;                   // force unknown: not a row of K_ff (warp-uniform)
            const uint32_t gr = __ldg(E.rowmap + 2u * node + a);
            const uint32_t rr = gr - E.row_lo;
            const double k = cand ? acc_t[j * 4 + a * 2 + b] : 0.0;
            const bool keep = cand && !uk && (!E.drop_zeros || k != 0.0);
            const uint32_t mask = __ballot_sync(0xffffffffu, keep);
            if (!FILL) {
                if (lane == 0) row_nnz[rr] = (uint32_t)__popc(mask);
                continue;
            }
            const uint32_t base = __ldg(rowptr + rr);
            if (keep) {
                const uint32_t pos = base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
                col[pos] = (int32_t)cmap;
                val[pos] = k;
            }
            const bool is_diag = keep && cmap == gr;
            const uint32_t dmask = __ballot_sync(0xffffffffu, is_diag);
            if (is_diag) diag[rr] = k;
            uint32_t umask = __ballot_sync(0xffffffffu, uk);
            if (lane == 0) {
                if (!dmask) diag[rr] = 0.0;
                double s = 0.0;                                      // prescribed columns, ascending (solver.rs:390-391, 427)
                while (umask) {
                    const int l2 = __ffs(umask) - 1;
                    umask &= umask - 1u;
                    const int j2 = l2 >> 1, b2 = l2 & 1;
                    const uint32_t c2 = cols_t[j2];
                    const double u = b2 ? E.uy[c2] : E.ux[c2];
                    s = __dadd_rn(s, __dmul_rn(__dmul_rn(acc_t[j2 * 4 + a * 2 + b2], u), -1.0));
                }
                rhs[rr] = __dadd_rn(s, a ? E.fy[node] : E.fx[node]);   // solver.rs:430-432
            }
        }
    }
    if (!FILL) {                                                     // structural size of K (mag_stats.nnz_structural)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) blocks += __shfl_xor_sync(0xffffffffu, blocks, off);
        if (lane == 0 && blocks) atomicAdd(n_blocks_total, (unsigned long long)blocks);     // integer: deterministic
    }
}

// Reaction forces without a stored K (solver.rs:457-473): the few nodes with an unknown force rebuild their
// block row (ascending columns) and multiply it with the displacements; known forces are copied.
__global__ void __launch_bounds__(128)
gather_reactions_kernel(const double2 *__restrict__ xy, const uint32_t *__restrict__ n0, const uint32_t *__restrict__ n1,
                        const uint32_t *__restrict__ n2, const uint32_t *__restrict__ elist,
                        const uint32_t *__restrict__ payload, const uint32_t *__restrict__ nptr, uint32_t n_own,
                        uint32_t node_lo, const uint8_t *__restrict__ known, const double *__restrict__ bc_fx,
                        const double *__restrict__ bc_fy, const double *__restrict__ ux, const double *__restrict__ uy,
                        double *__restrict__ fx, double *__restrict__ fy) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_own) return;
    const uint32_t node = node_lo + r;
    const uint8_t kn = known[node];
    const bool need_x = !(kn & MAG_KNOWN_FX), need_y = !(kn & MAG_KNOWN_FY);
    double f0 = bc_fx[node], f1 = bc_fy[node];
    if (need_x || need_y) {
        const gather::Conn conn{n0, n1, n2, elist};
        double s0 = 0.0, s1 = 0.0;
        gather::for_each_block_serial(conn, xy, c_mat.D, c_mat.t, payload, nptr[r], nptr[r + 1],
                                      [&](uint32_t cn, double a0, double a1, double a2, double a3) {
            const double ucx = ux[cn], ucy = uy[cn];
            s0 = __dadd_rn(s0, __dmul_rn(a0, ucx)); s0 = __dadd_rn(s0, __dmul_rn(a1, ucy));
            s1 = __dadd_rn(s1, __dmul_rn(a2, ucx)); s1 = __dadd_rn(s1, __dmul_rn(a3, ucy));
        });
        if (need_x) f0 = s0;
        if (need_y) f1 = s1;
    }
    fx[node] = f0;
    fy[node] = f1;
}

// node -> (local element, corner) lists of the owned nodes
struct Incidence {
    DevBuf<uint32_t> nptr;       // n_own + 1: offsets into pay
    DevBuf<uint32_t> pay;        // 3 * n_local payloads sorted by node (ascending incidence inside a node);
                                 // incidences of nodes another rank owns sort to the end
    bool ready = false;
};

// emit + stable sort by node id (32-bit keys) + per-node offsets
static void build_incidence(mag_ctx *ctx, const DevBuf<uint32_t> &n0, const DevBuf<uint32_t> &n1,
                            const DevBuf<uint32_t> &n2, const uint32_t *elist, size_t n_local, size_t n_nodes,
                            uint32_t node_lo, uint32_t node_hi, Incidence &I) {
    const uint32_t n_own = node_hi - node_lo;
    const size_t n_inc = n_local * 3;
    const int bits = bits_for(n_nodes + 1);          // the sentinel's low bits exceed every node id
    I.nptr.alloc(ctx, (size_t)n_own + 1);
    I.nptr.zero();
    I.pay.alloc(ctx, n_inc);
    DevBuf<uint32_t> keys(ctx, n_inc), keys_alt(ctx, n_inc), pay_alt(ctx, n_inc);
    if (n_local) {
        MAG_LAUNCH(ctx, emit_incidence_kernel, cdiv(n_local, 256), 256, 0, (const uint32_t *)n0.p,
                   (const uint32_t *)n1.p, (const uint32_t *)n2.p, elist, n_local, node_lo, node_hi, keys.p,
                   I.pay.p, I.nptr.p);
        radix_sort_pairs<uint32_t>(ctx, keys.p, I.pay.p, keys_alt.p, pay_alt.p, n_inc, bits);
    }
    exclusive_scan_u32(ctx, I.nptr.p, n_own, I.nptr.p, (size_t)n_own + 1);
    I.ready = true;
}

// K (owned node rows [K.node_lo, K.node_hi)) as 2x2-block CSR from the incidence lists.  The material must
// have been uploaded (upload_material).
static void build_bsr_from_incidence(mag_ctx *ctx, const DevBuf<double2> &xy, const DevBuf<uint32_t> &n0,
                                     const DevBuf<uint32_t> &n1, const DevBuf<uint32_t> &n2, const uint32_t *elist,
                                     const Incidence &I, BsrMatrix &K) {
    const uint32_t n_own = K.node_hi - K.node_lo;
    K.browptr.alloc(ctx, (size_t)n_own + 1);
    K.browptr.zero();
    if (n_own)
        MAG_LAUNCH(ctx, gather_count_kernel, cdiv(n_own, kGatherThreads), kGatherThreads, 0,
                   (const uint32_t *)n0.p, (const uint32_t *)n1.p, (const uint32_t *)n2.p, elist,
                   (const uint32_t *)I.pay.p, (const uint32_t *)I.nptr.p, n_own, K.browptr.p);
    exclusive_scan_u32(ctx, K.browptr.p, n_own, K.browptr.p, (size_t)n_own + 1);
    uint32_t n_blocks = 0;
    MAG_CUDA(cudaMemcpyAsync(&n_blocks, K.browptr.p + n_own, sizeof n_blocks, cudaMemcpyDeviceToHost, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    K.n_blocks = n_blocks;
    K.bcol.alloc(ctx, K.n_blocks);
    K.bval.alloc(ctx, (size_t)K.n_blocks * 4);
    if (n_own && K.n_blocks)
        MAG_LAUNCH(ctx, gather_fill_kernel, cdiv(n_own, kGatherThreads), kGatherThreads, 0,
                   (const double2 *)xy.p, (const uint32_t *)n0.p, (const uint32_t *)n1.p, (const uint32_t *)n2.p,
                   elist, (const uint32_t *)I.pay.p, (const uint32_t *)I.nptr.p, n_own, (const uint32_t *)K.browptr.p,
                   K.bcol.p, K.bval.p);
}

template <int FILL>
static void launch_fused_rows(mag_ctx *ctx, const DevBuf<double2> &xy, const DevBuf<uint32_t> &n0,
                              const DevBuf<uint32_t> &n1, const DevBuf<uint32_t> &n2, const uint32_t *elist,
                              const Incidence &I, uint32_t n_own, const ElimView &E, uint32_t *row_nnz,
                              const uint32_t *rowptr, int32_t *col, double *val, double *rhs, double *diag,
                              unsigned long long *n_blocks_total) {
    if (!n_own) return;
    if (!ctx->fused_attr_set) {
        MAG_CUDA(cudaFuncSetAttribute(fused_rows_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem));
        MAG_CUDA(cudaFuncSetAttribute(fused_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem));
        ctx->fused_attr_set = true;
    }
    MAG_LAUNCH(ctx, fused_rows_kernel<FILL>, cdiv(n_own, kFusedThreads), kFusedThreads, kFusedSmem,
               (const double2 *)xy.p, (const uint32_t *)n0.p, (const uint32_t *)n1.p, (const uint32_t *)n2.p, elist,
               (const uint32_t *)I.pay.p, (const uint32_t *)I.nptr.p, n_own, E, row_nnz, rowptr, col, val, rhs, diag,
               n_blocks_total);
}

}  // namespace mag
