// magnetite_b200 — the reference's main.rs:54-76 with the numerical core on the GPU.
//
//   magnetite_b200 input.json geom.msh [--skip]        mesh from gmsh (or geometry.write_msh) + input.json
//                                                      -> nodes.csv, elements.csv in the working directory
//   magnetite_b200 input.json geom.msh --reorder       the same with the nodes renumbered (reverse Cuthill-McKee)
//                                                      around the solve; the CSVs keep the mesh file's numbering
//   magnetite_b200 input.json outline.svg | outer.csv [holes.csv ...]   the reference's own form (main.rs:21-40):
//                                                      outlines -> geom.geo -> `gmsh geom.geo -2 -o geom.msh` -> the
//                                                      same flow; needs a gmsh binary on PATH
//   magnetite_b200 --geo out.geo input.json outline.svg|csv...            write the gmsh script only (no GPU, no gmsh)
//   magnetite_b200 --mesh out.msh input.json outline.svg|csv...           run gmsh on it and keep the mesh (no GPU)
//   magnetite_b200 --dump-rules input.json             print the parsed metadata and boundary rules (no GPU)
//   magnetite_b200 --band geom.msh                     node band of the mesh as numbered and after RCM (no GPU)
//
// The reference shells out to gmsh for the .msh (mesher.rs:501-506); that step is out of scope, so the
// CLI starts from the mesh file.  `--skip` is accepted for compatibility (the plot step never runs here).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../include/magnetite_b200.h"
#include "magnetite_io.hpp"

using namespace magnetite;

static std::string opt(const std::optional<double> &v) {
    return v ? post_processor::format_f64(*v) : std::string("None");
}

int main(int argc, char **argv) {
    try {
        if (argc == 3 && !std::strcmp(argv[1], "--dump-rules")) {
            const Json j = mesher::load_input_file(argv[2]);
            const ModelMetadata md = mesher::parse_input_metadata(j);
            std::printf("metadata %s %s %s %s %s\n", post_processor::format_f64(md.youngs_modulus).c_str(),
                        post_processor::format_f64(md.poisson_ratio).c_str(), post_processor::format_f64(md.part_thickness).c_str(),
                        post_processor::format_f64(md.characteristic_length_min).c_str(),
                        post_processor::format_f64(md.characteristic_length_max).c_str());
            for (const BoundaryRule &r : mesher::parse_boundary_rules(j))
                std::printf("rule %s region %s %s %s %s targets %s %s %s %s\n", r.name.c_str(),
                            post_processor::format_f64(r.region.x_min).c_str(), post_processor::format_f64(r.region.x_max).c_str(),
                            post_processor::format_f64(r.region.y_min).c_str(), post_processor::format_f64(r.region.y_max).c_str(),
                            opt(r.target.ux).c_str(), opt(r.target.uy).c_str(), opt(r.target.fx).c_str(), opt(r.target.fy).c_str());
            return 0;
        }
        if (argc == 3 && !std::strcmp(argv[1], "--band")) {
            std::vector<Node> nodes;
            std::vector<Element> elements;
            mesher::parse_mesh(argv[2], nodes, elements);
            std::vector<std::uint32_t> c[3], new_of_old(nodes.size());
            for (const Element &e : elements)
                for (int k = 0; k < 3; ++k) c[k].push_back((std::uint32_t)e.nodes[k]);
            std::uint64_t before = 0, after = 0;
            const int rc = mag_reorder_rcm(nodes.size(), elements.size(), c[0].data(), c[1].data(), c[2].data(),
                                           new_of_old.data(), &before, &after);
            if (rc != MAG_OK) throw MagnetiteError(MagnetiteError::Kind::Solver, mag_host_last_error(), rc);
            std::printf("nodes %zu elements %zu band %llu rcm %llu\n", nodes.size(), elements.size(),
                        (unsigned long long)before, (unsigned long long)after);
            return 0;
        }
        auto ends_with = [](const std::string &t, const char *suffix) {
            const size_t n = std::strlen(suffix);
            return t.size() >= n && t.compare(t.size() - n, n, suffix) == 0;
        };
        // mesher.rs:946-959: an .svg replaces whatever was read and ends the list, every .csv adds one container
        // (the first is the outer loop), anything else is an error
        auto read_outlines = [&](int first, int last, float cl_min) {
            std::vector<std::vector<Vertex>> containers;
            for (int i = first; i < last; ++i) {
                const std::string geom = argv[i];
                if (geom.rfind("--", 0) == 0) continue;
                if (ends_with(geom, ".svg")) { containers = mesher::parse_svg(geom, cl_min); break; }
                else if (ends_with(geom, ".csv")) containers.push_back(mesher::parse_csv(geom));
                else throw MagnetiteError(MagnetiteError::Kind::Input, "Unrecognized geometry filetype " + geom);
            }
            return containers;
        };
        if (argc >= 5 && !std::strcmp(argv[1], "--geo")) {
            const ModelMetadata md = mesher::parse_input_metadata(mesher::load_input_file(argv[3]));
            mesher::build_geo(read_outlines(4, argc, md.characteristic_length_min), argv[2], md.characteristic_length_min,
                              md.characteristic_length_max);
            return 0;
        }
        if (argc >= 5 && !std::strcmp(argv[1], "--mesh")) {                      // outlines -> gmsh -> out.msh, nothing else
            const ModelMetadata md = mesher::parse_input_metadata(mesher::load_input_file(argv[3]));
            mesher::compute_mesh(read_outlines(4, argc, md.characteristic_length_min), argv[2], md.characteristic_length_min,
                                 md.characteristic_length_max);
            std::vector<Node> nodes;
            std::vector<Element> elements;
            mesher::parse_mesh(argv[2], nodes, elements);
            std::printf("nodes %zu elements %zu\n", nodes.size(), elements.size());
            return 0;
        }
        if (argc < 3) {
            std::fprintf(stderr, "usage: magnetite_b200 input.json geom.msh|outline.csv... [--skip] [--reorder]\n");
            return 2;
        }
        SolverOptions so;
        for (int i = 3; i < argc; ++i)
            if (!std::strcmp(argv[i], "--reorder")) so.reorder = true;
        const std::string input_file = argv[1];
        std::string mesh_file = argv[2];
        const Json j = mesher::load_input_file(input_file);                     // mesher.rs:943-944
        const ModelMetadata md = mesher::parse_input_metadata(j);
        const bool from_outlines = !ends_with(mesh_file, ".msh");
        if (from_outlines) {                                                     // mesher.rs:946-967
            const std::vector<std::vector<Vertex>> containers = read_outlines(2, argc, md.characteristic_length_min);
            mesh_file = "geom.msh";
            mesher::compute_mesh(containers, mesh_file, md.characteristic_length_min, md.characteristic_length_max);
        }
        std::vector<Node> nodes;
        std::vector<Element> elements;
        mesher::parse_mesh(mesh_file, nodes, elements);                          // mesher.rs:969
        if (from_outlines) std::remove(mesh_file.c_str());                       // mesher.rs:701
        mesher::check_ccw(elements, solver::element_areas(elements, nodes));     // mesher.rs:691-693
        std::printf("info: loaded %zu nodes and %zu elements\n", nodes.size(), elements.size());
        mesher::apply_boundary_conditions(j, nodes, false);                      // mesher.rs:971
        solver::run(nodes, elements, md, so);                                    // main.rs:64
        post_processor::csv_output(elements, nodes, "nodes.csv", "elements.csv");   // main.rs:67-69
    } catch (const MagnetiteError &err) {
        std::fprintf(stderr, "Received error: %s\n", err.what());                // main.rs:46
        return 1;
    }
    return 0;
}
