// magnetite_io.hpp — C++ mirror of the input side of the reference that the solver path depends on
// (SURVEY §8(f) ranks 1-2), so the whole main.rs flow minus gmsh runs in a compiled language:
//
//   mesher::load_input_file / parse_input_metadata   <- src/mesher.rs:713-808
//   mesher::apply_boundary_conditions                <- src/mesher.rs:815-930 (strict > / <, later rules win)
//   mesher::parse_mesh                               <- src/mesher.rs:536-704 (MSH 4.x ASCII)
//   mesher::check_ccw                                <- src/mesher.rs:522-526 (flips when area < 1.0)
//   mesher::parse_csv                                <- src/mesher.rs:253-299 (outline vertices, x / y columns)
//   mesher::geo_text / build_geo / compute_mesh      <- src/mesher.rs:305-519 (.geo script, `gmsh geom.geo -2 -o ...`)
//   mesher::parse_svg                                <- src/mesher.rs:26-244  (polygon / polyline / rect with OUTER / INNER* ids)
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "magnetite_host.hpp"

namespace magnetite {

// Minimal JSON value (objects keep their key order: boundary rules are applied in file order).
struct Json {
    enum class Type { Null, Bool, Number, String, Object, Array } type = Type::Null;
    double number = 0.0;
    bool boolean = false;
    std::string string;
    std::vector<std::pair<std::string, Json>> object;
    std::vector<Json> array;
    bool has_key(const std::string &k) const;
    const Json &operator[](const std::string &k) const;     // Null value when absent
    std::optional<double> as_f64() const;                    // like json::JsonValue::as_f64
    static Json parse(const std::string &text);              // throws MagnetiteError(Input)
};

struct BoundaryRegion { double x_min, x_max, y_min, y_max; };                 // datatypes.rs:32-37
struct BoundaryTarget { std::optional<double> ux, uy, fx, fy; };              // datatypes.rs:40-45
struct BoundaryRule { std::string name; BoundaryRegion region; BoundaryTarget target; };

namespace mesher {
Json load_input_file(const std::string &input_file);
ModelMetadata parse_input_metadata(const Json &input_json);
std::vector<BoundaryRule> parse_boundary_rules(const Json &input_json);
void apply_boundary_conditions(const Json &input_json, std::vector<Node> &nodes, bool quiet = true);
void parse_mesh(const std::string &mesh_file, std::vector<Node> &nodes, std::vector<Element> &elements);
// `areas[e]` = signed area of element e (solver::compute_element_area on the GPU, batched)
void check_ccw(std::vector<Element> &elements, const std::vector<double> &areas);
std::vector<Vertex> parse_csv(const std::string &csv_file);
// containers[0] = the OUTER outline, containers[1..] = INNER outlines in document order (polygons and polylines
// first, then rectangles); y is inverted, repeated vertices and vertices closer than min_element_length to the
// previous one are dropped
std::vector<std::vector<Vertex>> parse_svg(const std::string &svg_file, float min_element_length);
// container 0 is the outer loop, the others are holes
std::string geo_text(const std::vector<std::vector<Vertex>> &vertices_containers, float characteristic_length_min,
                     float characteristic_length_max);
void build_geo(const std::vector<std::vector<Vertex>> &vertices_containers, const std::string &output_file,
               float characteristic_length_min, float characteristic_length_max);
// writes geom.geo, runs `gmsh geom.geo -2 -o <output>`, removes geom.geo; throws Mesher("Gmsh failed: ...") only
// when gmsh cannot be started, like the reference
void compute_mesh(const std::vector<std::vector<Vertex>> &vertices, const std::string &output,
                  float characteristic_length_min, float characteristic_length_max, bool quiet = false);
}  // namespace mesher

namespace solver {
// signed areas of all elements in one launch (solver.rs:187-193)
std::vector<double> element_areas(const std::vector<Element> &elements, const std::vector<Node> &nodes, int device = 0);
}

}  // namespace magnetite
