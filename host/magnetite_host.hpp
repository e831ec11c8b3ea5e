// magnetite_host.hpp — C++ host side above the C ABI, mirroring the reference's Rust modules
// name for name (the reference is compiled code and its toolchain is absent from this image):
//
//   magnetite::Vertex / Node / Element / ModelMetadata   <- src/datatypes.rs:2-29
//   magnetite::MagnetiteError                            <- src/error.rs:4-22
//   magnetite::solver::run(nodes, elements, metadata)    <- src/solver.rs:543-586
//   magnetite::solver::compute_element_area              <- src/solver.rs:187-193
//   magnetite::post_processor::csv_output(...)           <- src/post_processor.rs:18-83
//
// Option<f64> becomes std::optional<double>; Result<(), MagnetiteError> becomes a thrown
// MagnetiteError.  All arithmetic runs on the GPU behind include/magnetite_b200.h.
#pragma once
#include <array>
#include <cstddef>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

namespace magnetite {

struct Vertex { double x, y; };                                   // datatypes.rs:2-5
struct Node {                                                     // datatypes.rs:8-14
    Vertex vertex;
    std::optional<double> ux, uy, fx, fy;
};
struct Element {                                                  // datatypes.rs:17-20
    std::array<std::size_t, 3> nodes;
    std::optional<double> stress;
};
struct ModelMetadata {                                            // datatypes.rs:23-29
    double youngs_modulus, poisson_ratio, part_thickness;
    float characteristic_length_min = 0.f, characteristic_length_max = 0.f;
};

class MagnetiteError : public std::runtime_error {                // error.rs:4-22
public:
    enum class Kind { Input, Mesher, Solver, PostProcessor };
    MagnetiteError(Kind k, const std::string &msg, int code = 0);
    Kind kind;
    int code;                  // MAG_ERR_* when raised by the library, else 0
    std::string message;       // what() is "<Kind> error: <message>" like the reference's Display
};

struct SolverOptions {         // the reference has constants only (solver.rs:18-19): these are its values
    bool compat = true;        // plain CG, x0 = 0, absolute cost <= 1e-4; false = Jacobi-PCG to rel_tol
    double rel_tol = 1e-9;
    bool quiet = false;
    int device = 0;
    bool reorder = false;      // renumber the nodes (reverse Cuthill-McKee, mag_reorder_rcm) around the solve when that
                               // narrows the band — meshes in gmsh order; results stay in the caller's numbering
};

namespace solver {
constexpr std::size_t DOF = 2;                 // solver.rs:17
constexpr unsigned long long MAX_CG_ITER = 10000000ull;   // solver.rs:18
constexpr double TARGET_CG_COST = 1e-4;        // solver.rs:19

void run(std::vector<Node> &nodes, std::vector<Element> &elements, const ModelMetadata &model_metadata,
         const SolverOptions &options = SolverOptions());
double compute_element_area(const Element &element, const std::vector<Node> &nodes);
}  // namespace solver

namespace post_processor {
// Rust's `{}` for f64: shortest round-trip decimal, never scientific, no ".0" on integral values.
std::string format_f64(double v);
std::string format_f32(float v);      // the same for f32 (the characteristic lengths of ModelMetadata)
void csv_output(const std::vector<Element> &elements, const std::vector<Node> &nodes,
                const std::string &nodes_output, const std::string &elements_output, bool quiet = false);
}  // namespace post_processor

}  // namespace magnetite
