// assembly.cuh — deterministic global stiffness assembly (replaces the dense
// `+=` scatter of reference src/solver.rs:290-331).
//
//   emit   9 COO keys per triangle, one per (row node, col node) pair:
//          key = row_node << bits | col_node, payload = local_element*9 + pair
//   sort   stable LSD radix sort (radix_sort.cuh): equal keys stay in ascending
//          element order
//   reduce warp-shuffle segmented reduction over the sorted stream: every lane
//          gathers the 32-byte 2x2 block its payload points at, segment heads
//          add their followers left to right (shuffles inside the warp, global
//          loads past its end) — the reference's accumulation order, no float
//          atomics, bit-identical run to run
//   out    the full K as 2x2-block CSR (BSR) over nodes: browptr/bcol/bval.
//          Global DOF = 2*node + axis (solver.rs:306-309).
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace mag {

constexpr uint64_t kKeySentinel = ~0ull;

inline int bits_for(uint64_t n) {   // bits needed to represent values in [0, n)
    int b = 1;
    while (b < 63 && (1ull << b) < n) ++b;
    return b;
}

// One thread per (local element, pair).  Rows outside [node_lo, node_hi) belong
// to another rank: they get the sentinel key and sort to the end.
__global__ void emit_keys_kernel(const uint32_t *__restrict__ n0, const uint32_t *__restrict__ n1,
                                 const uint32_t *__restrict__ n2, const uint32_t *__restrict__ elist,
                                 size_t n_local, int bits, uint32_t node_lo, uint32_t node_hi,
                                 uint64_t *__restrict__ keys, uint32_t *__restrict__ payload) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    const size_t e = elist ? elist[i] : i;
    const uint32_t nd[3] = {n0[e], n1[e], n2[e]};
#pragma unroll
    for (int lr = 0; lr < 3; ++lr)
#pragma unroll
        for (int lc = 0; lc < 3; ++lc) {
            const int p = lr * 3 + lc;
            const bool mine = nd[lr] >= node_lo && nd[lr] < node_hi;
            keys[i * 9 + p] = mine ? (((uint64_t)nd[lr] << bits) | nd[lc]) : kKeySentinel;
            payload[i * 9 + p] = (uint32_t)(i * 9 + p);
        }
}

// head[i] = 1 iff sorted entry i starts a new (row,col) segment; sentinel
// entries are never heads.  Also counts blocks per owned node row.
__global__ void mark_heads_kernel(const uint64_t *__restrict__ keys, size_t n, int bits,
                                  uint32_t node_lo, uint32_t *__restrict__ head,
                                  uint32_t *__restrict__ row_blocks) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = keys[i];
    const bool h = (k != kKeySentinel) && (i == 0 || keys[i - 1] != k);
    head[i] = h ? 1u : 0u;
    if (h) atomicAdd(&row_blocks[(uint32_t)(k >> bits) - node_lo], 1u);   // integer: deterministic
}

// Segmented reduction of the 2x2 blocks, one lane per sorted entry.
//   uid[i]   exclusive scan of head[] — the BSR slot of entry i's segment
//   kblk     block-major element matrices (element.cuh), 4 doubles per payload
__global__ void __launch_bounds__(256)
segment_reduce_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ payload,
                      const uint32_t *__restrict__ head, const uint32_t *__restrict__ uid,
                      size_t n_valid, int bits, uint32_t node_lo, const double *__restrict__ kblk,
                      uint32_t *__restrict__ brow, uint32_t *__restrict__ bcol, double *__restrict__ bval) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool ok = i < n_valid;
    const uint64_t key = ok ? keys[i] : kKeySentinel;
    const bool is_head = ok && head[i] != 0;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    if (ok) {
        const double2 *src = reinterpret_cast<const double2 *>(kblk + (size_t)payload[i] * 4);
        const double2 a = __ldg(src), b = __ldg(src + 1);
        v0 = a.x; v1 = a.y; v2 = b.x; v3 = b.y;
    }
    // lanes of this warp that start a segment; the next head after me bounds my segment
    const uint32_t heads = __ballot_sync(0xffffffffu, is_head || !ok || key == kKeySentinel);
    const uint32_t after = (lane == 31) ? 0u : (heads >> (lane + 1));
    const int in_warp_len = after ? __ffs(after) : (32 - lane);   // entries of my segment inside the warp
    const bool open_end = is_head && after == 0;                  // may continue in the next warp
    int steps = is_head ? in_warp_len - 1 : 0;
#pragma unroll 1
    for (int off = 16; off > 0; off >>= 1) steps = max(steps, __shfl_xor_sync(0xffffffffu, steps, off));
    // the reference adds every contribution, the first included, to a zeroed dense entry (solver.rs:295-296,
    // 304-323): 0.0 + (-0.0) = +0.0, so a lone -0.0 contribution is stored as +0.0
    double a0 = __dadd_rn(0.0, v0), a1 = __dadd_rn(0.0, v1), a2 = __dadd_rn(0.0, v2), a3 = __dadd_rn(0.0, v3);
    for (int j = 1; j <= steps; ++j) {
        const double t0 = __shfl_down_sync(0xffffffffu, v0, j);
        const double t1 = __shfl_down_sync(0xffffffffu, v1, j);
        const double t2 = __shfl_down_sync(0xffffffffu, v2, j);
        const double t3 = __shfl_down_sync(0xffffffffu, v3, j);
        if (is_head && j < in_warp_len) {
            a0 = __dadd_rn(a0, t0); a1 = __dadd_rn(a1, t1);
            a2 = __dadd_rn(a2, t2); a3 = __dadd_rn(a3, t3);
        }
    }
    if (open_end) {   // followers that live in the next warp(s): plain global loads
        size_t j = i + in_warp_len;
        while (j < n_valid && keys[j] == key) {
            const double2 *src = reinterpret_cast<const double2 *>(kblk + (size_t)payload[j] * 4);
            const double2 a = __ldg(src), b = __ldg(src + 1);
            a0 = __dadd_rn(a0, a.x); a1 = __dadd_rn(a1, a.y);
            a2 = __dadd_rn(a2, b.x); a3 = __dadd_rn(a3, b.y);
            ++j;
        }
    }
    if (is_head) {
        const uint32_t u = uid[i];
        brow[u] = (uint32_t)(key >> bits) - node_lo;
        bcol[u] = (uint32_t)(key & ((1ull << bits) - 1ull));
        double2 *dst = reinterpret_cast<double2 *>(bval + (size_t)u * 4);
        dst[0] = make_double2(a0, a1);
        dst[1] = make_double2(a2, a3);
    }
}

// The assembled full stiffness matrix of one rank: rows = owned nodes.
struct BsrMatrix {
    uint32_t node_lo = 0, node_hi = 0;   // owned node rows [lo, hi)
    uint32_t n_blocks = 0;
    DevBuf<uint32_t> browptr;            // (hi-lo)+1
    DevBuf<uint32_t> brow;               // n_blocks, LOCAL row node (node - node_lo) of every block
    DevBuf<uint32_t> bcol;               // n_blocks, global node ids, ascending per row
    DevBuf<double> bval;                 // n_blocks*4, row-major 2x2
};

}  // namespace mag
