// gather_core.h — per-node core of the gather assembly (mag_options.assembly = 1), the alternative to
// the key sort of assembly.cuh for reference src/solver.rs:290-331.
//
// Plain C++ without CUDA headers: nvcc compiles it into the kernels of gather.cuh, and g++ compiles the
// very same functions into tests/native/gather_host_test.cpp, where they are checked bit for bit against
// the oracle on the CPU.  That harness is test infrastructure; the library never runs this on the host.
//
// Idea.  Row node r of K receives 2x2 blocks from the elements that touch r — six on a plate.  Given
// r's incidence list (element, corner) in ascending order, one thread recomputes the two rows
// 2*corner, 2*corner+1 of each K_e and adds the blocks column by column.  The order of the additions
// into an entry is ascending (element, corner lr, corner lc) — the reference's `+=` order
// (solver.rs:299-323) — so the result equals the sorted-key path bit for bit, without float atomics.
// What is sorted is 3E (node, incidence) pairs by a log2(N)-bit key instead of 9E pairs by 2*log2(N) bits.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MAG_HD __host__ __device__ __forceinline__
#else
#define MAG_HD inline
#endif

namespace mag {
namespace gather {

// Explicitly rounded operations: no FMA contraction on the device; the host pass is compiled with
// -ffp-contract=off.
MAG_HD double gmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
MAG_HD double gadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
MAG_HD double gsub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
MAG_HD double gdiv(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

constexpr int kMaxCols = 16;     // distinct column nodes a row keeps in thread-local storage (plate: 7)

struct Conn {                    // connectivity as the kernels see it
    const uint32_t *n0, *n1, *n2;
    const uint32_t *elist;       // local element -> global element (ascending), or null = identity
};

MAG_HD void corner_nodes(const Conn &m, uint32_t local_elem, uint32_t nd[3]) {
    const size_t e = m.elist ? m.elist[local_elem] : local_elem;
    nd[0] = m.n0[e];
    nd[1] = m.n1[e];
    nd[2] = m.n2[e];
}

// Sorted insert of c into cols[0..n) unless present.  Returns the new count, or -1 when c is new and
// the list already holds `cap` columns.
MAG_HD int insert_col(uint32_t *cols, int n, uint32_t c, int cap = kMaxCols) {
    int pos = 0;
    while (pos < n && cols[pos] < c) ++pos;
    if (pos < n && cols[pos] == c) return n;
    if (n == cap) return -1;
    for (int j = n; j > pos; --j) cols[j] = cols[j - 1];
    cols[pos] = c;
    return n + 1;
}

// Smallest column node > prev (any column when !have_prev) among the corners of the incident elements.
MAG_HD bool next_col(const Conn &m, const uint32_t *pay, uint32_t begin, uint32_t end, bool have_prev,
                     uint32_t prev, uint32_t *out) {
    bool found = false;
    uint32_t best = 0;
    for (uint32_t i = begin; i < end; ++i) {
        uint32_t nd[3];
        corner_nodes(m, pay[i] / 3u, nd);
        for (int k = 0; k < 3; ++k) {
            const uint32_t c = nd[k];
            if (have_prev && c <= prev) continue;
            if (!found || c < best) { best = c; found = true; }
        }
    }
    *out = best;
    return found;
}

// Number of distinct column nodes of the row whose incidences are pay[begin, end).
MAG_HD uint32_t count_cols(const Conn &m, const uint32_t *pay, uint32_t begin, uint32_t end) {
    uint32_t cols[kMaxCols];
    int n = 0;
    for (uint32_t i = begin; i < end && n >= 0; ++i) {
        uint32_t nd[3];
        corner_nodes(m, pay[i] / 3u, nd);
        for (int k = 0; k < 3 && n >= 0; ++k) n = insert_col(cols, n, nd[k]);
    }
    if (n >= 0) return (uint32_t)n;
    uint32_t count = 0, c = 0;          // more than kMaxCols neighbours: walk the columns in ascending order
    bool have = false;
    while (next_col(m, pay, begin, end, have, c, &c)) { have = true; ++count; }
    return count;
}

// Rows 2*lr and 2*lr+1 of K_e = ((B^T D) B) A t for the triangle nd[0..3), operation by operation as
// element_stiffness_kernel (element.cuh) and the reference (solver.rs:187-278) compute them.
template <class Pt>
MAG_HD void ke_rows(const Pt *xy, const uint32_t nd[3], int lr, const double *D, double t, double out[2][6]) {
    const Pt p0 = xy[nd[0]], p1 = xy[nd[1]], p2 = xy[nd[2]];
    const double x0 = p0.x, y0 = p0.y, x1 = p1.x, y1 = p1.y, x2 = p2.x, y2 = p2.y;
    const double area = gmul(0.5, gadd(gadd(gmul(x0, gsub(y1, y2)), gmul(x1, gsub(y2, y0))), gmul(x2, gsub(y0, y1))));
    const double den = gmul(2.0, area);
    const double z = gdiv(0.0, den);
    const double qb[3] = {gdiv(gsub(y1, y2), den), gdiv(gsub(y2, y0), den), gdiv(gsub(y0, y1), den)};
    const double qg[3] = {gdiv(gsub(x2, x1), den), gdiv(gsub(x0, x2), den), gdiv(gsub(x1, x0), den)};
    double B[3][6];
    for (int k = 0; k < 3; ++k) {
        B[0][2 * k] = qb[k]; B[0][2 * k + 1] = z;
        B[1][2 * k] = z;     B[1][2 * k + 1] = qg[k];
        B[2][2 * k] = qg[k]; B[2][2 * k + 1] = qb[k];
    }
    // column 2*lr + a of B without a dynamically indexed array (that would put B in local memory on the device)
    const double qbl = lr == 0 ? qb[0] : (lr == 1 ? qb[1] : qb[2]);
    const double qgl = lr == 0 ? qg[0] : (lr == 1 ? qg[1] : qg[2]);
    for (int a = 0; a < 2; ++a) {
        const double b0 = a ? z : qbl, b1 = a ? qgl : z, b2 = a ? qbl : qgl;      // B[0..3][2*lr + a]
        double btd[3];                                   // row of B^T D: k ascending, product then sum
        for (int c = 0; c < 3; ++c) {
            double s = gmul(b0, D[0 * 3 + c]);
            s = gadd(gmul(b1, D[1 * 3 + c]), s);
            s = gadd(gmul(b2, D[2 * 3 + c]), s);
            btd[c] = s;
        }
        for (int c = 0; c < 6; ++c) {
            double s = gmul(btd[0], B[0][c]);
            s = gadd(gmul(btd[1], B[1][c]), s);
            s = gadd(gmul(btd[2], B[2][c]), s);
            s = gmul(s, area);                           // solver.rs:276
            s = gmul(s, t);                              // solver.rs:277
            out[a][c] = s;
        }
    }
}

// Visits the 2x2 blocks of one node row in ascending column order WITHOUT any table: for every column the
// incidence list is walked again and the K_e rows of the matching elements are recomputed (quadratic in the
// degree, no storage: the path of rows with more columns than a table holds, and of the few rows the
// reaction forces need).  f(col, a0, a1, a2, a3) receives the accumulated block, row-major.
template <class Pt, class F>
MAG_HD void for_each_block_serial(const Conn &m, const Pt *xy, const double *D, double t, const uint32_t *pay,
                                  uint32_t begin, uint32_t end, F &&f) {
    uint32_t c = 0;
    bool have = false;
    while (next_col(m, pay, begin, end, have, c, &c)) {
        have = true;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;     // from +0.0, like the reference's zeroed dense matrix
        for (uint32_t i = begin; i < end; ++i) {
            const uint32_t p = pay[i], le = p / 3u;
            const int lr = (int)(p - 3u * le);
            uint32_t nd[3];
            corner_nodes(m, le, nd);
            if (nd[0] != c && nd[1] != c && nd[2] != c) continue;
            double rows[2][6];
            ke_rows(xy, nd, lr, D, t, rows);
            for (int lc = 0; lc < 3; ++lc) {
                if (nd[lc] != c) continue;
                a0 = gadd(a0, rows[0][2 * lc]); a1 = gadd(a1, rows[0][2 * lc + 1]);
                a2 = gadd(a2, rows[1][2 * lc]); a3 = gadd(a3, rows[1][2 * lc + 1]);
            }
        }
        f(c, a0, a1, a2, a3);
    }
}

// Writes the block row of one node: bcol[0..ncols) ascending column nodes, bval[4*j..4*j+4) the 2x2
// block (row-major) of column bcol[j].  ncols = count_cols() of the same list.
template <class Pt>
MAG_HD void fill_row(const Conn &m, const Pt *xy, const double *D, double t, const uint32_t *pay,
                     uint32_t begin, uint32_t end, uint32_t ncols, uint32_t *bcol, double *bval) {
    if (ncols <= (uint32_t)kMaxCols) {
        uint32_t cols[kMaxCols];
        double acc[kMaxCols * 4];
        int n = 0;
        for (uint32_t i = begin; i < end; ++i) {
            uint32_t nd[3];
            corner_nodes(m, pay[i] / 3u, nd);
            for (int k = 0; k < 3; ++k) n = insert_col(cols, n, nd[k]);
        }
        uint32_t seen = 0;
        for (uint32_t i = begin; i < end; ++i) {
            const uint32_t p = pay[i], le = p / 3u;
            const int lr = (int)(p - 3u * le);
            uint32_t nd[3];
            corner_nodes(m, le, nd);
            double rows[2][6];
            ke_rows(xy, nd, lr, D, t, rows);
            for (int lc = 0; lc < 3; ++lc) {
                int slot = 0;
                while (cols[slot] != nd[lc]) ++slot;
                double *a = acc + 4 * slot;
                const double v0 = rows[0][2 * lc], v1 = rows[0][2 * lc + 1], v2 = rows[1][2 * lc], v3 = rows[1][2 * lc + 1];
                if ((seen >> slot) & 1u) {
                    a[0] = gadd(a[0], v0); a[1] = gadd(a[1], v1); a[2] = gadd(a[2], v2); a[3] = gadd(a[3], v3);
                } else {       // the reference's entry starts at +0.0 (zeroed dense matrix): 0.0 + (-0.0) = +0.0
                    a[0] = gadd(0.0, v0); a[1] = gadd(0.0, v1); a[2] = gadd(0.0, v2); a[3] = gadd(0.0, v3);
                    seen |= 1u << slot;
                }
            }
        }
        for (int j = 0; j < n; ++j) {
            bcol[j] = cols[j];
            for (int q = 0; q < 4; ++q) bval[4 * j + q] = acc[4 * j + q];
        }
        return;
    }
    // a node with more than kMaxCols neighbours: one column at a time, K_e rows recomputed per match
    uint32_t j = 0;
    for_each_block_serial(m, xy, D, t, pay, begin, end, [&](uint32_t c, double a0, double a1, double a2, double a3) {
        if (j >= ncols) return;
        bcol[j] = c;
        bval[4 * j] = a0; bval[4 * j + 1] = a1; bval[4 * j + 2] = a2; bval[4 * j + 3] = a3;
        ++j;
    });
}

}  // namespace gather
}  // namespace mag
