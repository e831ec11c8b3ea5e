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
static_assert(gather::kFastCols <= 16, "half a warp per node row, one lane per column node");
static_assert(2 * kAccStride >= 8 * gather::kFastCols, "the tables of a node pair stage the pair's four rows");

struct ElimView {
    const uint8_t *known;
    const uint32_t *rowmap, *colmap;
    const uint2 *colid2;         // per node: reduced column of (node, x), (node, y); 0xffffffff where the displacement is prescribed
    const double *ux, *uy, *fx, *fy;
    int drop_zeros;
    uint32_t row_lo, node_lo;
};

// A row with more columns than the table holds: one thread, no table (for_each_block_serial).
template <int FILL>
__device__ void fused_emit_serial(const gather::Conn &conn, const double2 *xy, const uint32_t *pay, uint32_t begin,
                                  uint32_t end, uint32_t node, const ElimView &E, uint32_t *row_nnz,
                                  const uint32_t *rowptr, int32_t *col, double *val, double *rhs, double *diag,
                                  uint32_t *n_cols_out, uint32_t *cnt2 = nullptr) {
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
        else if (cnt2) cnt2[a] = cnt[a];            // one-pass kernel: the row lengths go to its shared-memory scan
        else row_nnz[rr] = cnt[a];
    }
    *n_cols_out = ncols;
}

template <int FILL>
__global__ void __launch_bounds__(kFusedThreads)
fused_rows_kernel(const double2 *__restrict__ xy, const uint32_t *__restrict__ n0, const uint32_t *__restrict__ n1,
                  const uint32_t *__restrict__ n2, const uint32_t *__restrict__ elist,
                  const double *__restrict__ kblk, const uint32_t *__restrict__ payload,
                  const uint32_t *__restrict__ nptr, uint32_t n_own, ElimView E,
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

    // ---- phase 1: one thread per node row --------------------------------------------------------------
    uint32_t p0 = 0, p1 = 0, gr0 = 0, gr1 = 0, first_base = 0;
    int ncols = 0;
    uint32_t kn = 0;
    if (active) {
        const uint32_t node = E.node_lo + r;
        p0 = nptr[r]; p1 = nptr[r + 1];
        kn = E.known[node];
        const uint2 g = *reinterpret_cast<const uint2 *>(E.rowmap + 2u * node);
        gr0 = g.x; gr1 = g.y;
        if (FILL) first_base = (kn & MAG_KNOWN_FX) ? rowptr[gr0 - E.row_lo] : ((kn & MAG_KNOWN_FY) ? rowptr[gr1 - E.row_lo] : 0u);
        ncols = gather::build_row_table_from_ke(conn, kblk, payload, p0, p1, cols, acc);
    }
    uint32_t blocks = ncols > 0 ? (uint32_t)ncols : 0u;
    if (ncols < 0) {                                                 // more columns than the table holds: table-free
        uint32_t counted = 0;
        fused_emit_serial<FILL>(conn, xy, payload, p0, p1, E.node_lo + r, E, row_nnz, rowptr, col, val, rhs, diag, &counted);
        blocks = counted;
    }
    __syncwarp();

    // ---- phase 2: half a warp per node, one lane per column node (its 2x2 block = up to 4 entries) --------
    const uint32_t wbase = threadIdx.x & ~31u;
    const uint32_t first = blockIdx.x * blockDim.x + wbase;          // first node row of this warp
    const int sub = lane >> 4, j = lane & 15, hshift = sub * 16;
    const uint32_t lt = (1u << j) - 1u;
#pragma unroll 2
    for (int it = 0; it < 16; ++it) {
        if (first + 2u * it >= n_own) break;                         // warp-uniform
        const int tt = 2 * it + sub;                                 // owner lane of my half's node
        const int nc = __shfl_sync(0xffffffffu, ncols, tt);
        const uint32_t knt = __shfl_sync(0xffffffffu, kn, tt);
        const uint32_t g0 = __shfl_sync(0xffffffffu, gr0, tt), g1 = __shfl_sync(0xffffffffu, gr1, tt);
        const uint32_t fb_a = __shfl_sync(0xffffffffu, first_base, 2 * it), fb_b = __shfl_sync(0xffffffffu, first_base, 2 * it + 1);
        const bool node_ok = first + tt < n_own && nc >= 0;          // nc < 0: the owner wrote its rows itself
        const bool ex0 = node_ok && (knt & MAG_KNOWN_FX), ex1 = node_ok && (knt & MAG_KNOWN_FY);   // rows of K_ff
        const bool cand = node_ok && j < nc;
        const double *acc_t = acc_all + (size_t)(wbase + tt) * kAccStride;
        const uint32_t *cols_t = cols_all + (size_t)(wbase + tt) * kColStride;
        const uint32_t node = E.node_lo + first + tt;
        uint32_t cn = 0;
        double k00 = 0.0, k01 = 0.0, k10 = 0.0, k11 = 0.0;
        uint2 cid = make_uint2(0xffffffffu, 0xffffffffu);
        if (cand) {
            cn = cols_t[j];
            k00 = acc_t[4 * j]; k01 = acc_t[4 * j + 1]; k10 = acc_t[4 * j + 2]; k11 = acc_t[4 * j + 3];
            cid = __ldg(E.colid2 + cn);                              // reduced columns of (cn, x), (cn, y); ~0: prescribed
        }
        const bool uk0 = cand && cid.x == 0xffffffffu, uk1 = cand && cid.y == 0xffffffffu;
        const bool keep00 = cand && ex0 && !uk0 && (!E.drop_zeros || k00 != 0.0);
        const bool keep01 = cand && ex0 && !uk1 && (!E.drop_zeros || k01 != 0.0);
        const bool keep10 = cand && ex1 && !uk0 && (!E.drop_zeros || k10 != 0.0);
        const bool keep11 = cand && ex1 && !uk1 && (!E.drop_zeros || k11 != 0.0);
        const uint32_t m00 = __ballot_sync(0xffffffffu, keep00), m01 = __ballot_sync(0xffffffffu, keep01);
        const uint32_t m10 = __ballot_sync(0xffffffffu, keep10), m11 = __ballot_sync(0xffffffffu, keep11);
        const uint32_t um0 = __ballot_sync(0xffffffffu, uk0), um1 = __ballot_sync(0xffffffffu, uk1);
        // entries of the four rows of this pair of nodes: (node a, x), (node a, y), (node b, x), (node b, y)
        const uint32_t na0 = __popc(m00 & 0xffffu) + __popc(m01 & 0xffffu), na1 = __popc(m10 & 0xffffu) + __popc(m11 & 0xffffu);
        const uint32_t nb0 = __popc(m00 >> 16) + __popc(m01 >> 16), nb1 = __popc(m10 >> 16) + __popc(m11 >> 16);
        const uint32_t my0 = sub ? nb0 : na0;
        const uint32_t off0 = (sub ? na0 + na1 : 0u) + __popc((m00 >> hshift) & lt) + __popc((m01 >> hshift) & lt);
        const uint32_t off1 = (sub ? na0 + na1 : 0u) + my0 + __popc((m10 >> hshift) & lt) + __popc((m11 >> hshift) & lt);
        if (!FILL) {
            if (j == 0) {
                if (ex0) row_nnz[g0 - E.row_lo] = my0;
                if (ex1) row_nnz[g1 - E.row_lo] = sub ? nb1 : na1;
            }
            continue;
        }
        // right-hand side: prescribed columns in ascending order, -(K u), then + f (solver.rs:390-391, 427-432)
        if (j == 0 && (ex0 || ex1)) {
            uint32_t pend = ((um0 | um1) >> hshift) & 0xffffu;
            double s0 = 0.0, s1 = 0.0;
            while (pend) {
                const int jj = __ffs(pend) - 1;
                pend &= pend - 1u;
                const uint32_t c2 = cols_t[jj];
                if ((um0 >> (hshift + jj)) & 1u) {
                    const double u = E.ux[c2];
                    s0 = __dadd_rn(s0, __dmul_rn(__dmul_rn(acc_t[4 * jj], u), -1.0));
                    s1 = __dadd_rn(s1, __dmul_rn(__dmul_rn(acc_t[4 * jj + 2], u), -1.0));
                }
                if ((um1 >> (hshift + jj)) & 1u) {
                    const double u = E.uy[c2];
                    s0 = __dadd_rn(s0, __dmul_rn(__dmul_rn(acc_t[4 * jj + 1], u), -1.0));
                    s1 = __dadd_rn(s1, __dmul_rn(__dmul_rn(acc_t[4 * jj + 3], u), -1.0));
                }
            }
            if (ex0) rhs[g0 - E.row_lo] = __dadd_rn(s0, E.fx[node]);
            if (ex1) rhs[g1 - E.row_lo] = __dadd_rn(s1, E.fy[node]);
        }
        __syncwarp();                                                // every lane holds its block: the two tables are dead
        const bool has_a = __shfl_sync(0xffffffffu, (int)(ex0 || ex1), 0) != 0;
        const uint32_t gbase = has_a ? fb_a : fb_b;                  // CSR position of the first row of the pair
        double *stage = acc_all + (size_t)(wbase + 2 * it) * kAccStride;        // 2 x 49 doubles >= 4 rows x 24 entries
        if (keep00) { stage[off0] = k00; col[gbase + off0] = (int32_t)cid.x; if (cid.x == g0) diag[g0 - E.row_lo] = k00; }
        if (keep01) { const uint32_t o = off0 + (keep00 ? 1u : 0u); stage[o] = k01; col[gbase + o] = (int32_t)cid.y; if (cid.y == g0) diag[g0 - E.row_lo] = k01; }
        if (keep10) { stage[off1] = k10; col[gbase + off1] = (int32_t)cid.x; if (cid.x == g1) diag[g1 - E.row_lo] = k10; }
        if (keep11) { const uint32_t o = off1 + (keep10 ? 1u : 0u); stage[o] = k11; col[gbase + o] = (int32_t)cid.y; if (cid.y == g1) diag[g1 - E.row_lo] = k11; }
        __syncwarp();
        const uint32_t total = na0 + na1 + nb0 + nb1;                // the pair's rows are consecutive in the CSR arrays
        for (uint32_t q = lane; q < total; q += 32) val[gbase + q] = stage[q];
    }
    if (!FILL) {                                                     // structural size of K (mag_stats.nnz_structural)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) blocks += __shfl_xor_sync(0xffffffffu, blocks, off);
        if (lane == 0 && blocks) atomicAdd(n_blocks_total, (unsigned long long)blocks);     // integer: deterministic
    }
}

// ---- the same in ONE pass: decoupled look-back instead of count -> scan -> fill ----------------------------
// A tile (CTA) of 128 node rows builds its tables once, counts the entries each of its rows keeps, learns where
// its rows start in the CSR arrays from the tiles before it (each tile publishes {flag, count} in one 64-bit
// word: first its own aggregate, then the inclusive prefix; a tile adds aggregates backwards until it meets a
// prefix), writes rowptr for its rows and then the rows themselves.  Tile ids come from a ticket counter, so a
// tile only ever waits for tiles that are already running.  The offsets are exact integer prefix sums: the
// result does not depend on the order in which tiles run.
constexpr unsigned long long kTileAggregate = 1ull << 62, kTilePrefix = 2ull << 62, kTileValueMask = (1ull << 62) - 1ull;

struct TileScan {
    unsigned long long *status;     // one word per tile, zeroed before the launch
    unsigned *ticket;               // zeroed before the launch
    unsigned long long *nnz_out;    // total entries (written by the last tile)
    int *error;                     // set when a predecessor never published (cannot happen; guards the GPU)
};

__global__ void __launch_bounds__(kFusedThreads)
fused_rows_onepass_kernel(const double2 *__restrict__ xy, const uint32_t *__restrict__ n0, const uint32_t *__restrict__ n1,
                          const uint32_t *__restrict__ n2, const uint32_t *__restrict__ elist,
                          const double *__restrict__ kblk, const uint32_t *__restrict__ payload,
                          const uint32_t *__restrict__ nptr, uint32_t n_own, uint32_t n_rows, ElimView E,
                          uint32_t *__restrict__ rowptr, int32_t *__restrict__ col, double *__restrict__ val,
                          double *__restrict__ rhs, double *__restrict__ diag,
                          unsigned long long *__restrict__ n_blocks_total, TileScan T) {
    extern __shared__ __align__(16) unsigned char fused_smem[];
    double *acc_all = reinterpret_cast<double *>(fused_smem);
    uint32_t *cols_all = reinterpret_cast<uint32_t *>(fused_smem + (size_t)kFusedThreads * kAccStride * sizeof(double));
    __shared__ uint32_t s_tile, s_base;
    __shared__ uint32_t s_rowcnt[2 * kFusedThreads];
    __shared__ uint32_t s_warp[kFusedThreads / 32];
    if (threadIdx.x == 0) s_tile = atomicAdd(T.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t r = tile * kFusedThreads + threadIdx.x;
    const bool active = r < n_own;
    const gather::Conn conn{n0, n1, n2, elist};
    double *acc = acc_all + (size_t)threadIdx.x * kAccStride;
    uint32_t *cols = cols_all + (size_t)threadIdx.x * kColStride;

    // ---- phase 1: one thread per node row: its table ------------------------------------------------------
    uint32_t p0 = 0, p1 = 0, gr0 = 0, gr1 = 0, kn = 0;
    int ncols = 0;
    s_rowcnt[2 * threadIdx.x] = 0; s_rowcnt[2 * threadIdx.x + 1] = 0;
    if (active) {
        const uint32_t node = E.node_lo + r;
        p0 = nptr[r]; p1 = nptr[r + 1];
        kn = E.known[node];
        const uint2 g = *reinterpret_cast<const uint2 *>(E.rowmap + 2u * node);
        gr0 = g.x; gr1 = g.y;
        ncols = gather::build_row_table_from_ke(conn, kblk, payload, p0, p1, cols, acc);
    }
    uint32_t blocks = ncols > 0 ? (uint32_t)ncols : 0u;
    if (ncols < 0) {                                                 // table overflow: count its rows without a table
        uint32_t counted = 0;
        fused_emit_serial<0>(conn, xy, payload, p0, p1, E.node_lo + r, E, nullptr, nullptr, nullptr, nullptr, nullptr,
                             nullptr, &counted, s_rowcnt + 2 * threadIdx.x);
        blocks = counted;
    }
    __syncwarp();

    const uint32_t wbase = threadIdx.x & ~31u;
    const uint32_t first = tile * kFusedThreads + wbase;             // first node row of this warp
    const int sub = lane >> 4, j = lane & 15, hshift = sub * 16;
    const uint32_t lt = (1u << j) - 1u;

    // what a lane sees of its half's node in pair `it`
    struct PairLane {
        bool ex0, ex1, cand, keep00, keep01, keep10, keep11;
        uint32_t m00, m01, m10, m11, um0, um1, g0, g1, node;
        double k00, k01, k10, k11;
        uint2 cid;
        const double *acc_t;
        const uint32_t *cols_t;
    };
    auto look = [&](int it) {
        PairLane L;
        const int tt = 2 * it + sub;
        const int nc = __shfl_sync(0xffffffffu, ncols, tt);
        const uint32_t knt = __shfl_sync(0xffffffffu, kn, tt);
        L.g0 = __shfl_sync(0xffffffffu, gr0, tt); L.g1 = __shfl_sync(0xffffffffu, gr1, tt);
        const bool node_ok = first + tt < n_own && nc >= 0;
        L.ex0 = node_ok && (knt & MAG_KNOWN_FX); L.ex1 = node_ok && (knt & MAG_KNOWN_FY);
        L.cand = node_ok && j < nc;
        L.acc_t = acc_all + (size_t)(wbase + tt) * kAccStride;
        L.cols_t = cols_all + (size_t)(wbase + tt) * kColStride;
        L.node = E.node_lo + first + tt;
        L.k00 = L.k01 = L.k10 = L.k11 = 0.0;
        L.cid = make_uint2(0xffffffffu, 0xffffffffu);
        if (L.cand) {
            const uint32_t cn = L.cols_t[j];
            L.k00 = L.acc_t[4 * j]; L.k01 = L.acc_t[4 * j + 1]; L.k10 = L.acc_t[4 * j + 2]; L.k11 = L.acc_t[4 * j + 3];
            L.cid = __ldg(E.colid2 + cn);
        }
        const bool uk0 = L.cand && L.cid.x == 0xffffffffu, uk1 = L.cand && L.cid.y == 0xffffffffu;
        L.keep00 = L.cand && L.ex0 && !uk0 && (!E.drop_zeros || L.k00 != 0.0);
        L.keep01 = L.cand && L.ex0 && !uk1 && (!E.drop_zeros || L.k01 != 0.0);
        L.keep10 = L.cand && L.ex1 && !uk0 && (!E.drop_zeros || L.k10 != 0.0);
        L.keep11 = L.cand && L.ex1 && !uk1 && (!E.drop_zeros || L.k11 != 0.0);
        L.m00 = __ballot_sync(0xffffffffu, L.keep00); L.m01 = __ballot_sync(0xffffffffu, L.keep01);
        L.m10 = __ballot_sync(0xffffffffu, L.keep10); L.m11 = __ballot_sync(0xffffffffu, L.keep11);
        L.um0 = __ballot_sync(0xffffffffu, uk0); L.um1 = __ballot_sync(0xffffffffu, uk1);
        return L;
    };

    // ---- phase 2a: entries kept per row --------------------------------------------------------------------
#pragma unroll 2
    for (int it = 0; it < 16; ++it) {
        if (first + 2u * it >= n_own) break;
        const PairLane L = look(it);
        if (j == 0) {
            const uint32_t c0 = __popc((L.m00 >> hshift) & 0xffffu) + __popc((L.m01 >> hshift) & 0xffffu);
            const uint32_t c1 = __popc((L.m10 >> hshift) & 0xffffu) + __popc((L.m11 >> hshift) & 0xffffu);
            if (L.ex0) s_rowcnt[2 * (wbase + 2 * it + sub)] = c0;
            if (L.ex1) s_rowcnt[2 * (wbase + 2 * it + sub) + 1] = c1;
        }
    }
    __syncthreads();

    // ---- where the tile's rows start: CTA scan + look-back over the tiles before this one -----------------------
    const uint32_t c0 = s_rowcnt[2 * threadIdx.x], c1 = s_rowcnt[2 * threadIdx.x + 1];
    uint32_t inc = c0 + c1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kFusedThreads / 32; ++w) {
        if (w < warp) woff += s_warp[w];
        total += s_warp[w];
    }
    if (threadIdx.x == 0) {
        unsigned long long excl = 0;
        volatile unsigned long long *st = T.status;
        if (tile == 0) {
            st[0] = kTilePrefix | (unsigned long long)total;
        } else {
            st[tile] = kTileAggregate | (unsigned long long)total;
            for (long long p = (long long)tile - 1; p >= 0; --p) {
                unsigned long long w;
                long long spins = 0;
                while (((w = st[p]) >> 62) == 0ull) {
                    if (++spins > (1ll << 28)) { *T.error = 1; break; }
                }
                excl += w & kTileValueMask;
                if ((w >> 62) != 1ull) break;                         // a prefix (or the guard tripped): done
            }
            st[tile] = kTilePrefix | (excl + (unsigned long long)total);
        }
        s_base = (uint32_t)excl;
        if ((size_t)(tile + 1) * kFusedThreads >= n_own) {           // the last tile closes the row pointer array
            rowptr[n_rows] = (uint32_t)(excl + total);
            *T.nnz_out = excl + total;
        }
    }
    __syncthreads();
    const uint32_t my_base = s_base + woff + inc - (c0 + c1);         // CSR position of this node's first kept entry
    if (active) {
        if (kn & MAG_KNOWN_FX) rowptr[gr0 - E.row_lo] = my_base;
        if (kn & MAG_KNOWN_FY) rowptr[gr1 - E.row_lo] = my_base + c0;
    }
    if (ncols < 0) {                                                 // table overflow: this thread writes its rows itself
        uint32_t counted = 0;
        fused_emit_serial<1>(conn, xy, payload, p0, p1, E.node_lo + r, E, nullptr, rowptr, col, val, rhs, diag, &counted);
    }

    // ---- phase 2b: the rows -----------------------------------------------------------------------------------
#pragma unroll 2
    for (int it = 0; it < 16; ++it) {
        if (first + 2u * it >= n_own) break;
        const PairLane L = look(it);
        const uint32_t na0 = __popc(L.m00 & 0xffffu) + __popc(L.m01 & 0xffffu), na1 = __popc(L.m10 & 0xffffu) + __popc(L.m11 & 0xffffu);
        const uint32_t nb0 = __popc(L.m00 >> 16) + __popc(L.m01 >> 16), nb1 = __popc(L.m10 >> 16) + __popc(L.m11 >> 16);
        const uint32_t my0 = sub ? nb0 : na0;
        const uint32_t off0 = (sub ? na0 + na1 : 0u) + __popc((L.m00 >> hshift) & lt) + __popc((L.m01 >> hshift) & lt);
        const uint32_t off1 = (sub ? na0 + na1 : 0u) + my0 + __popc((L.m10 >> hshift) & lt) + __popc((L.m11 >> hshift) & lt);
        if (j == 0 && (L.ex0 || L.ex1)) {                            // rhs: prescribed columns ascending, -(K u), then + f
            uint32_t pend = ((L.um0 | L.um1) >> hshift) & 0xffffu;
            double s0 = 0.0, s1 = 0.0;
            while (pend) {
                const int jj = __ffs(pend) - 1;
                pend &= pend - 1u;
                const uint32_t c2 = L.cols_t[jj];
                if ((L.um0 >> (hshift + jj)) & 1u) {
                    const double u = E.ux[c2];
                    s0 = __dadd_rn(s0, __dmul_rn(__dmul_rn(L.acc_t[4 * jj], u), -1.0));
                    s1 = __dadd_rn(s1, __dmul_rn(__dmul_rn(L.acc_t[4 * jj + 2], u), -1.0));
                }
                if ((L.um1 >> (hshift + jj)) & 1u) {
                    const double u = E.uy[c2];
                    s0 = __dadd_rn(s0, __dmul_rn(__dmul_rn(L.acc_t[4 * jj + 1], u), -1.0));
                    s1 = __dadd_rn(s1, __dmul_rn(__dmul_rn(L.acc_t[4 * jj + 3], u), -1.0));
                }
            }
            if (L.ex0) rhs[L.g0 - E.row_lo] = __dadd_rn(s0, E.fx[L.node]);
            if (L.ex1) rhs[L.g1 - E.row_lo] = __dadd_rn(s1, E.fy[L.node]);
        }
        // CSR position of the pair's first row: the first node of the pair that has rows
        const uint32_t cnt_a = na0 + na1;
        const int has_a = __shfl_sync(0xffffffffu, (int)(L.ex0 || L.ex1), 0);
        const uint32_t base_a = __shfl_sync(0xffffffffu, my_base, 2 * it), base_b = __shfl_sync(0xffffffffu, my_base, 2 * it + 1);
        const uint32_t gbase = has_a ? base_a : base_b;
        (void)cnt_a;
        __syncwarp();                                                // every lane holds its block: the two tables are dead
        double *stage = acc_all + (size_t)(wbase + 2 * it) * kAccStride;
        if (L.keep00) { stage[off0] = L.k00; col[gbase + off0] = (int32_t)L.cid.x; if (L.cid.x == L.g0) diag[L.g0 - E.row_lo] = L.k00; }
        if (L.keep01) { const uint32_t o = off0 + (L.keep00 ? 1u : 0u); stage[o] = L.k01; col[gbase + o] = (int32_t)L.cid.y; if (L.cid.y == L.g0) diag[L.g0 - E.row_lo] = L.k01; }
        if (L.keep10) { stage[off1] = L.k10; col[gbase + off1] = (int32_t)L.cid.x; if (L.cid.x == L.g1) diag[L.g1 - E.row_lo] = L.k10; }
        if (L.keep11) { const uint32_t o = off1 + (L.keep10 ? 1u : 0u); stage[o] = L.k11; col[gbase + o] = (int32_t)L.cid.y; if (L.cid.y == L.g1) diag[L.g1 - E.row_lo] = L.k11; }
        __syncwarp();
        const uint32_t ptot = na0 + na1 + nb0 + nb1;                 // the pair's rows are consecutive in the CSR arrays
        for (uint32_t q = lane; q < ptot; q += 32) val[gbase + q] = stage[q];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) blocks += __shfl_xor_sync(0xffffffffu, blocks, off);
    if (lane == 0 && blocks) atomicAdd(n_blocks_total, (unsigned long long)blocks);         // integer: deterministic
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

static void ensure_fused_attrs(mag_ctx *ctx);

template <int FILL>
static void launch_fused_rows(mag_ctx *ctx, const DevBuf<double2> &xy, const DevBuf<uint32_t> &n0,
                              const DevBuf<uint32_t> &n1, const DevBuf<uint32_t> &n2, const uint32_t *elist,
                              const double *kblk, const Incidence &I, uint32_t n_own, const ElimView &E, uint32_t *row_nnz,
                              const uint32_t *rowptr, int32_t *col, double *val, double *rhs, double *diag,
                              unsigned long long *n_blocks_total) {
    if (!n_own) return;
    ensure_fused_attrs(ctx);
    MAG_LAUNCH(ctx, fused_rows_kernel<FILL>, cdiv(n_own, kFusedThreads), kFusedThreads, kFusedSmem,
               (const double2 *)xy.p, (const uint32_t *)n0.p, (const uint32_t *)n1.p, (const uint32_t *)n2.p, elist,
               kblk, (const uint32_t *)I.pay.p, (const uint32_t *)I.nptr.p, n_own, E, row_nnz, rowptr, col, val, rhs,
               diag, n_blocks_total);
}

static void ensure_fused_attrs(mag_ctx *ctx) {
    if (ctx->fused_attr_set) return;
    MAG_CUDA(cudaFuncSetAttribute(fused_rows_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem));
    MAG_CUDA(cudaFuncSetAttribute(fused_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem));
    MAG_CUDA(cudaFuncSetAttribute(fused_rows_onepass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmem));
    ctx->fused_attr_set = true;
}

}  // namespace mag
