// magnetite_host.cpp — see magnetite_host.hpp.  Host-only C++17; links libmagnetite_b200.so.
#include "magnetite_host.hpp"

#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "../include/magnetite_b200.h"

namespace magnetite {

static const char *kind_name(MagnetiteError::Kind k) {
    switch (k) {
        case MagnetiteError::Kind::Input: return "Input";
        case MagnetiteError::Kind::Mesher: return "Mesher";
        case MagnetiteError::Kind::Solver: return "Solver";
        default: return "Post Processor";
    }
}

MagnetiteError::MagnetiteError(Kind k, const std::string &msg, int c)
    : std::runtime_error(std::string(kind_name(k)) + " error: " + msg), kind(k), code(c), message(msg) {}

namespace {

struct Ctx {                       // RAII around mag_ctx
    mag_ctx *h = nullptr;
    explicit Ctx(int device) {
        if (mag_abi_version() != MAG_ABI_VERSION)      // struct layouts differ: refuse instead of corrupting the stack
            throw MagnetiteError(MagnetiteError::Kind::Solver, "libmagnetite_b200.so ABI version " +
                                 std::to_string(mag_abi_version()) + " != header version " + std::to_string(MAG_ABI_VERSION));
        const int rc = mag_ctx_create(&h, device);
        if (rc != MAG_OK) throw MagnetiteError(MagnetiteError::Kind::Solver, mag_last_error(), rc);
    }
    ~Ctx() { mag_ctx_destroy(h); }
};

struct Flat {                      // Vec<Node>/Vec<Element> flattened for mag_mesh
    std::vector<double> x, y, ux, uy, fx, fy;
    std::vector<std::uint8_t> known;
    std::vector<std::uint32_t> n0, n1, n2;
    mag_mesh view() const {
        mag_mesh m{};
        m.n_nodes = x.size(); m.n_elems = n0.size();
        m.x = x.data(); m.y = y.data();
        m.n0 = n0.data(); m.n1 = n1.data(); m.n2 = n2.data();
        m.ux = ux.data(); m.uy = uy.data(); m.fx = fx.data(); m.fy = fy.data();
        m.known = known.data(); m.on_device = 0;
        return m;
    }
};

Flat flatten(const std::vector<Node> &nodes, const std::vector<Element> &elements) {
    Flat f;
    const std::size_t n = nodes.size(), e = elements.size();
    f.x.resize(n); f.y.resize(n); f.ux.assign(n, 0.0); f.uy.assign(n, 0.0); f.fx.assign(n, 0.0); f.fy.assign(n, 0.0);
    f.known.assign(n, 0);
    for (std::size_t i = 0; i < n; ++i) {
        const Node &nd = nodes[i];
        f.x[i] = nd.vertex.x; f.y[i] = nd.vertex.y;
        if (nd.ux) { f.ux[i] = *nd.ux; f.known[i] |= MAG_KNOWN_UX; }
        if (nd.uy) { f.uy[i] = *nd.uy; f.known[i] |= MAG_KNOWN_UY; }
        if (nd.fx) { f.fx[i] = *nd.fx; f.known[i] |= MAG_KNOWN_FX; }
        if (nd.fy) { f.fy[i] = *nd.fy; f.known[i] |= MAG_KNOWN_FY; }
    }
    f.n0.resize(e); f.n1.resize(e); f.n2.resize(e);
    for (std::size_t i = 0; i < e; ++i) {
        for (std::size_t v : elements[i].nodes)
            if (v > 0xfffffffeull) throw MagnetiteError(MagnetiteError::Kind::Solver, "node index does not fit 32 bits");
        f.n0[i] = (std::uint32_t)elements[i].nodes[0];
        f.n1[i] = (std::uint32_t)elements[i].nodes[1];
        f.n2[i] = (std::uint32_t)elements[i].nodes[2];
    }
    return f;
}

// The same mesh with node i renamed new_of_old[i] (element order and orientation untouched).
Flat permuted(const Flat &f, const std::vector<std::uint32_t> &new_of_old) {
    Flat g = f;
    for (std::size_t i = 0; i < f.x.size(); ++i) {
        const std::uint32_t j = new_of_old[i];
        g.x[j] = f.x[i]; g.y[j] = f.y[i];
        g.ux[j] = f.ux[i]; g.uy[j] = f.uy[i]; g.fx[j] = f.fx[i]; g.fy[j] = f.fy[i];
        g.known[j] = f.known[i];
    }
    for (std::size_t e = 0; e < f.n0.size(); ++e) {
        g.n0[e] = new_of_old[f.n0[e]]; g.n1[e] = new_of_old[f.n1[e]]; g.n2[e] = new_of_old[f.n2[e]];
    }
    return g;
}

}  // namespace

namespace solver {

void run(std::vector<Node> &nodes, std::vector<Element> &elements, const ModelMetadata &md,
         const SolverOptions &so) {
    auto say = [&](const char *s) { if (!so.quiet) std::printf("%s\n", s); };
    say("info: building element stiffness matrices...");           // solver.rs:551
    say("info: building total stiffness matrix...");               // solver.rs:570
    Flat f = flatten(nodes, elements);
    std::vector<std::uint32_t> new_of_old;                         // empty: solved in the caller's numbering
    if (so.reorder) {                                              // gmsh order has no band
        new_of_old.resize(nodes.size());
        std::uint64_t before = 0, after = 0;
        const int rc = mag_reorder_rcm(f.x.size(), f.n0.size(), f.n0.data(), f.n1.data(), f.n2.data(),
                                       new_of_old.data(), &before, &after);
        if (rc != MAG_OK) throw MagnetiteError(MagnetiteError::Kind::Solver, mag_host_last_error(), rc);
        if (after < before) f = permuted(f, new_of_old);
        else new_of_old.clear();
        if (!so.quiet)
            std::printf("info: node band %llu -> %llu (%s)\n", (unsigned long long)before, (unsigned long long)after,
                        new_of_old.empty() ? "kept the mesher's numbering" : "renumbered for the solve");
    }
    const mag_mesh mesh = f.view();
    const mag_material mat{md.youngs_modulus, md.poisson_ratio, md.part_thickness};
    mag_options opt;
    mag_options_default(&opt);
    opt.compat = so.compat ? 1 : 0;
    opt.rel_tol = so.rel_tol;
    const std::size_t n = nodes.size(), e = elements.size();
    std::vector<double> ux(n), uy(n), fx(n), fy(n), stress(e);
    mag_result out{ux.data(), uy.data(), fx.data(), fy.data(), stress.data(), nullptr, 0};
    mag_stats st{};
    say("info: setting up system...");                             // solver.rs:416
    say("info: solving...");                                       // solver.rs:437
    {
        Ctx ctx(so.device);
        const int rc = mag_solve(ctx.h, &mesh, &mat, &opt, &out, &st);
        if (rc != MAG_OK)                                          // solver.rs:160-164
            throw MagnetiteError(MagnetiteError::Kind::Solver,
                                 std::string("Conjugate Gradient error: ") + mag_last_error(), rc);
    }
    if (!so.quiet) {
        std::printf("info: finished conjugate gradient approximation in %llu iterations\n",
                    (unsigned long long)st.iters);                 // solver.rs:101-104
        std::printf("info: solved system in %.3f seconds\n", st.ms_solve / 1e3);   // solver.rs:441
    }
    for (std::size_t i = 0; i < n; ++i) {                          // solver.rs:476-482
        const std::size_t j = new_of_old.empty() ? i : new_of_old[i];
        nodes[i].ux = ux[j]; nodes[i].uy = uy[j];
        nodes[i].fx = fx[j]; nodes[i].fy = fy[j];
    }
    for (std::size_t i = 0; i < e; ++i) elements[i].stress = stress[i];   // solver.rs:532-533
    say("info: solve complete");                                   // solver.rs:484
}

double compute_element_area(const Element &element, const std::vector<Node> &nodes) {
    std::vector<Node> tri = {nodes.at(element.nodes[0]), nodes.at(element.nodes[1]), nodes.at(element.nodes[2])};
    std::vector<Element> one = {Element{{0, 1, 2}, std::nullopt}};
    const Flat f = flatten(tri, one);
    const mag_mesh mesh = f.view();
    double area = 0.0;
    Ctx ctx(0);
    const int rc = mag_element_area(ctx.h, &mesh, &area);
    if (rc != MAG_OK) throw MagnetiteError(MagnetiteError::Kind::Solver, mag_last_error(), rc);
    return area;
}

}  // namespace solver

namespace post_processor {

// Rust's `{}` for a float of type T: the SHORTEST digits that round-trip in that type, positional, zero padded
// (1.2345678901234567e25 -> "12345678901234567000000000"); to_chars(fixed) would print the exact binary
// expansion instead, so take the shortest digits from the scientific form and place the point by hand.
template <class T>
static std::string format_shortest(T v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[64];
    const auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
    std::string s(buf, r.ptr), out;
    if (s[0] == '-') { out = "-"; s.erase(0, 1); }
    const size_t epos = s.find('e');
    const int exp10 = std::atoi(s.c_str() + epos + 1);
    std::string digits;
    for (size_t i = 0; i < epos; ++i) if (s[i] != '.') digits += s[i];
    const int n = (int)digits.size();
    if (exp10 >= n - 1) {
        out += digits + std::string((size_t)(exp10 - (n - 1)), '0');
    } else if (exp10 >= 0) {
        out += digits.substr(0, (size_t)exp10 + 1) + "." + digits.substr((size_t)exp10 + 1);
    } else {
        out += "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
    }
    return out;
}

std::string format_f64(double v) { return format_shortest<double>(v); }
std::string format_f32(float v) { return format_shortest<float>(v); }

// Rust's Display of std::io::Error, which the reference embeds in its message (post_processor.rs:27-29)
static std::string os_error(int err) { return std::string(std::strerror(err)) + " (os error " + std::to_string(err) + ")"; }

void csv_output(const std::vector<Element> &elements, const std::vector<Node> &nodes,
                const std::string &nodes_output, const std::string &elements_output, bool quiet) {
    std::ofstream nf(nodes_output, std::ios::binary | std::ios::trunc);
    if (!nf) throw MagnetiteError(MagnetiteError::Kind::Solver, "Failed to create nodes.csv: " + os_error(errno));
    std::ofstream ef(elements_output, std::ios::binary | std::ios::trunc);
    if (!ef) throw MagnetiteError(MagnetiteError::Kind::Solver, "Failed to create elements.csv: " + os_error(errno));
    nf << "x,y,ux,uy\n";                                           // post_processor.rs:42
    for (const Node &nd : nodes) {
        if (!nd.ux || !nd.uy)                                      // the reference unwrap()s (:50-51)
            throw MagnetiteError(MagnetiteError::Kind::PostProcessor, "node without a displacement: run the solver first");
        nf << format_f64(nd.vertex.x) << ',' << format_f64(nd.vertex.y) << ',' << format_f64(*nd.ux) << ','
           << format_f64(*nd.uy) << '\n';
    }
    ef << "n0,n1,n2,stress\n";                                     // post_processor.rs:60
    for (const Element &el : elements) {
        if (!el.stress)
            throw MagnetiteError(MagnetiteError::Kind::PostProcessor, "element without a stress: run the solver first");
        ef << el.nodes[0] << ',' << el.nodes[1] << ',' << el.nodes[2] << ',' << format_f64(*el.stress) << '\n';
    }
    if (!quiet) std::printf("info: wrote output to %s and %s\n", nodes_output.c_str(), elements_output.c_str());
}

}  // namespace post_processor
}  // namespace magnetite
