// reorder.cpp — host-only node renumbering for meshes whose node order carries no locality
// (gmsh output, src/mesher.rs:663-671 keeps gmsh's tags as ids).  SURVEY §8(e): the row-block
// partition, the halo extents and the 16-bit SELL column offsets all want a small band, which the
// structured plates have by construction and a Delaunay mesh does not.  Reverse Cuthill-McKee on the
// node graph; the caller permutes the mesh before mag_solve and maps the results back, so DOF
// numbering at the boundary stays 2*node + axis of the ORIGINAL ids.  No CUDA here.
#include <algorithm>
#include <cstdint>
#include <new>
#include <string>
#include <vector>

#include "../../include/magnetite_b200.h"

namespace maghost { void set_error(const std::string &msg); }

namespace {

// max |a - b| over the node pairs of every element = half-bandwidth of K in node blocks
uint64_t node_band(uint64_t n_elems, const uint32_t *n0, const uint32_t *n1, const uint32_t *n2,
                   const uint32_t *map) {
    uint64_t band = 0;
    for (uint64_t e = 0; e < n_elems; ++e) {
        uint32_t a = n0[e], b = n1[e], c = n2[e];
        if (map) { a = map[a]; b = map[b]; c = map[c]; }
        const uint32_t lo = std::min(a, std::min(b, c)), hi = std::max(a, std::max(b, c));
        band = std::max<uint64_t>(band, hi - lo);
    }
    return band;
}

struct Graph {
    std::vector<uint64_t> ptr;      // n + 1
    std::vector<uint32_t> adj;      // sorted, unique, no self loops
    uint32_t degree(uint32_t v) const { return (uint32_t)(ptr[v + 1] - ptr[v]); }
};

Graph build_graph(uint64_t n, uint64_t n_elems, const uint32_t *n0, const uint32_t *n1, const uint32_t *n2) {
    Graph g;
    g.ptr.assign(n + 1, 0);
    const uint32_t *c[3] = {n0, n1, n2};
    for (uint64_t e = 0; e < n_elems; ++e)
        for (int k = 0; k < 3; ++k) g.ptr[c[k][e] + 1] += 2;          // each corner meets the other two
    for (uint64_t v = 0; v < n; ++v) g.ptr[v + 1] += g.ptr[v];
    std::vector<uint32_t> raw(g.ptr[n]);
    std::vector<uint64_t> fill(g.ptr.begin(), g.ptr.end() - 1);
    for (uint64_t e = 0; e < n_elems; ++e)
        for (int k = 0; k < 3; ++k) {
            const uint32_t v = c[k][e];
            raw[fill[v]++] = c[(k + 1) % 3][e];
            raw[fill[v]++] = c[(k + 2) % 3][e];
        }
    g.adj.reserve(raw.size() / 2 + n);
    std::vector<uint64_t> ptr2(n + 1, 0);
    for (uint64_t v = 0; v < n; ++v) {
        uint32_t *b = raw.data() + g.ptr[v], *e = raw.data() + g.ptr[v + 1];
        std::sort(b, e);
        e = std::unique(b, e);
        for (uint32_t *p = b; p < e; ++p) if (*p != v) g.adj.push_back(*p);   // degenerate elements repeat a node
        ptr2[v + 1] = g.adj.size();
    }
    g.ptr.swap(ptr2);
    return g;
}

// Breadth-first level structure of the component of `root`; `order` receives the visit order,
// `level[v]` the distance.  `mark[v] == stamp` flags visited nodes (no clearing between calls).
// Returns the eccentricity of root.
uint32_t bfs_levels(const Graph &g, uint32_t root, std::vector<uint32_t> &order, std::vector<uint32_t> &level,
                    std::vector<uint32_t> &mark, uint32_t stamp) {
    order.clear();
    order.push_back(root);
    mark[root] = stamp;
    level[root] = 0;
    for (size_t head = 0; head < order.size(); ++head) {
        const uint32_t v = order[head];
        for (uint64_t k = g.ptr[v]; k < g.ptr[v + 1]; ++k) {
            const uint32_t w = g.adj[k];
            if (mark[w] != stamp) { mark[w] = stamp; level[w] = level[v] + 1; order.push_back(w); }
        }
    }
    return level[order.back()];
}

// George-Liu: walk to a node of (locally) maximal eccentricity, preferring low degree.
uint32_t pseudo_peripheral(const Graph &g, uint32_t start, std::vector<uint32_t> &order, std::vector<uint32_t> &level,
                           std::vector<uint32_t> &mark, uint32_t &stamp) {
    uint32_t root = start;
    uint32_t ecc = bfs_levels(g, root, order, level, mark, ++stamp);
    for (int guard = 0; guard < 64; ++guard) {
        uint32_t best = root, best_deg = UINT32_MAX;
        for (size_t i = order.size(); i-- > 0 && level[order[i]] == ecc;) {
            const uint32_t v = order[i], d = g.degree(v);
            if (d < best_deg || (d == best_deg && v < best)) { best = v; best_deg = d; }
        }
        if (best == root) break;
        const uint32_t ecc2 = bfs_levels(g, best, order, level, mark, ++stamp);
        if (ecc2 <= ecc) break;          // `order`/`level` now belong to `best`, but only `root` is used below
        root = best;
        ecc = ecc2;
    }
    return root;
}

}  // namespace

extern "C" int mag_reorder_rcm(uint64_t n_nodes, uint64_t n_elems, const uint32_t *n0, const uint32_t *n1,
                               const uint32_t *n2, uint32_t *new_of_old, uint64_t *band_before,
                               uint64_t *band_after) {
    if (!new_of_old && n_nodes) { maghost::set_error("mag_reorder_rcm: new_of_old is NULL"); return MAG_ERR_BAD_ARG; }
    if (n_elems && (!n0 || !n1 || !n2)) { maghost::set_error("mag_reorder_rcm: connectivity is NULL"); return MAG_ERR_BAD_ARG; }
    if (n_nodes >= UINT32_MAX) { maghost::set_error("mag_reorder_rcm: node ids are 32-bit"); return MAG_ERR_BAD_ARG; }
    for (uint64_t e = 0; e < n_elems; ++e)
        if (n0[e] >= n_nodes || n1[e] >= n_nodes || n2[e] >= n_nodes) {
            maghost::set_error("mag_reorder_rcm: element " + std::to_string(e) + " references a node >= n_nodes");
            return MAG_ERR_BAD_INDEX;
        }
    try {
        const uint32_t n = (uint32_t)n_nodes;
        const Graph g = build_graph(n, n_elems, n0, n1, n2);
        std::vector<uint32_t> cm;                       // Cuthill-McKee order: cm[k] = old id of the k-th node
        cm.reserve(n);
        std::vector<uint32_t> order, level(n), mark(n, 0), nbr;
        std::vector<uint8_t> placed(n, 0);
        uint32_t stamp = 0;
        for (uint32_t seed = 0; seed < n; ++seed) {
            if (placed[seed] || g.degree(seed) == 0) continue;
            const uint32_t root = pseudo_peripheral(g, seed, order, level, mark, stamp);
            size_t head = cm.size();
            cm.push_back(root);
            placed[root] = 1;
            for (; head < cm.size(); ++head) {
                const uint32_t v = cm[head];
                nbr.clear();
                for (uint64_t k = g.ptr[v]; k < g.ptr[v + 1]; ++k)
                    if (!placed[g.adj[k]]) { placed[g.adj[k]] = 1; nbr.push_back(g.adj[k]); }
                std::sort(nbr.begin(), nbr.end(), [&](uint32_t a, uint32_t b) {
                    const uint32_t da = g.degree(a), db = g.degree(b);
                    return da != db ? da < db : a < b;
                });
                cm.insert(cm.end(), nbr.begin(), nbr.end());
            }
        }
        const uint32_t connected = (uint32_t)cm.size();
        for (uint32_t k = 0; k < connected; ++k) new_of_old[cm[k]] = connected - 1 - k;     // the "reverse"
        uint32_t next = connected;                      // nodes no element references keep their relative order, last
        for (uint32_t v = 0; v < n; ++v) if (!placed[v]) new_of_old[v] = next++;
    } catch (const std::bad_alloc &) {
        maghost::set_error("mag_reorder_rcm: host allocation failed");
        return MAG_ERR_OOM;
    }
    if (band_before) *band_before = node_band(n_elems, n0, n1, n2, nullptr);
    if (band_after) *band_after = node_band(n_elems, n0, n1, n2, new_of_old);
    return MAG_OK;
}

extern "C" int mag_mesh_band(uint64_t n_nodes, uint64_t n_elems, const uint32_t *n0, const uint32_t *n1,
                             const uint32_t *n2, uint64_t *band) {
    if (!band || (n_elems && (!n0 || !n1 || !n2))) { maghost::set_error("mag_mesh_band: NULL argument"); return MAG_ERR_BAD_ARG; }
    for (uint64_t e = 0; e < n_elems; ++e)
        if (n0[e] >= n_nodes || n1[e] >= n_nodes || n2[e] >= n_nodes) {
            maghost::set_error("mag_mesh_band: element " + std::to_string(e) + " references a node >= n_nodes");
            return MAG_ERR_BAD_INDEX;
        }
    *band = node_band(n_elems, n0, n1, n2, nullptr);
    return MAG_OK;
}
