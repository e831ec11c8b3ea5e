// gather_host_test.cpp — TEST HARNESS, not part of the product: runs the per-node core of the gather
// assembly (magnetite_b200/csrc/gather_core.h, the functions the CUDA kernels of gather.cuh call) on the
// CPU so tests/test_gather_core_host.py can compare it bit for bit with the oracle.  The steps around the
// core mirror assemble_gather() in gather.cuh: element list of the rank, (node, incidence) pairs sorted
// stably by node, per-node offsets, count -> scan -> fill.
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -o libgather_host.so gather_host_test.cpp
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "../../magnetite_b200/csrc/gather_core.h"

namespace {
struct P2 { double x, y; };
}

// Returns the number of 2x2 blocks of the node rows [node_lo, node_hi).  browptr (n_own + 1) is always
// written; bcol / bval only when capacity_blocks is large enough.
extern "C" uint64_t gather_host_assemble(uint64_t n_nodes, uint64_t n_elems, const double *x, const double *y,
                                         const uint32_t *n0, const uint32_t *n1, const uint32_t *n2, const double *D9,
                                         double t, uint32_t node_lo, uint32_t node_hi, int use_elist,
                                         uint32_t *browptr, uint32_t *bcol, double *bval, uint64_t capacity_blocks) {
    using namespace mag::gather;
    std::vector<P2> xy(n_nodes);
    for (uint64_t i = 0; i < n_nodes; ++i) xy[i] = P2{x[i], y[i]};
    auto mine = [&](uint32_t v) { return v >= node_lo && v < node_hi; };
    std::vector<uint32_t> elist;
    if (use_elist)                                            // flag_elements + compact_elements (system.cuh)
        for (uint64_t e = 0; e < n_elems; ++e)
            if (mine(n0[e]) || mine(n1[e]) || mine(n2[e])) elist.push_back((uint32_t)e);
    const Conn conn{n0, n1, n2, use_elist ? elist.data() : nullptr};
    const uint32_t n_local = use_elist ? (uint32_t)elist.size() : (uint32_t)n_elems;
    const uint32_t n_own = node_hi - node_lo;

    std::vector<uint32_t> key, pay, cnt(n_own + 1, 0);        // emit_incidence_kernel
    for (uint32_t i = 0; i < n_local; ++i) {
        uint32_t nd[3];
        corner_nodes(conn, i, nd);
        for (uint32_t k = 0; k < 3; ++k)
            if (mine(nd[k])) { key.push_back(nd[k]); pay.push_back(i * 3 + k); ++cnt[nd[k] - node_lo]; }
    }
    std::vector<uint32_t> order(key.size());                  // radix_sort_pairs: stable, by node
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    std::vector<uint32_t> sorted(pay.size());
    for (size_t i = 0; i < order.size(); ++i) sorted[i] = pay[order[i]];
    std::vector<uint32_t> nptr(n_own + 1, 0);                 // exclusive scan of the counts
    for (uint32_t r = 0; r < n_own; ++r) nptr[r + 1] = nptr[r] + cnt[r];

    browptr[0] = 0;                                           // gather_count_kernel + scan
    for (uint32_t r = 0; r < n_own; ++r)
        browptr[r + 1] = browptr[r] + count_cols(conn, sorted.data(), nptr[r], nptr[r + 1]);
    const uint64_t n_blocks = browptr[n_own];
    if (n_blocks > capacity_blocks || !bcol || !bval) return n_blocks;
    for (uint32_t r = 0; r < n_own; ++r)                      // gather_fill_kernel
        fill_row(conn, xy.data(), D9, t, sorted.data(), nptr[r], nptr[r + 1], browptr[r + 1] - browptr[r],
                 bcol + browptr[r], bval + 4 * (size_t)browptr[r]);
    return n_blocks;
}
