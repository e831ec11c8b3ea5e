// plate_demo — the reference's main.rs:54-76 flow on a synthetic plate, through the C++ host layer:
// build Vec<Node>/Vec<Element> (mesher defaults + tensile boundary rules), solver::run,
// post_processor::csv_output.   usage: plate_demo NX NY nodes.csv elements.csv
//                                      plate_demo --format-selftest      (no GPU needed)
//                                      plate_demo --format-stdin         hex bit patterns in, formatted f64 out (no GPU needed)
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "magnetite_host.hpp"

using namespace magnetite;

static int format_selftest() {
    struct { double v; std::string s; } cases[] = {
        {3.0, "3"}, {-0.0, "-0"}, {0.0, "0"}, {69e9, "69000000000"}, {1e-7, "0.0000001"},
        {0.1 + 0.2, "0.30000000000000004"}, {-4.5, "-4.5"}, {1e21, "1000000000000000000000"},
        {1.5e-10, "0.00000000015"}, {1.2345678901234567e25, "12345678901234566000000000"}, {123456.789, "123456.789"},
        {-2.5e-5, "-0.000025"}, {1e15, "1000000000000000"}, {5e-324, "0." + std::string(323, '0') + "5"}};
    int bad = 0;
    for (auto &c : cases) {
        const std::string got = post_processor::format_f64(c.v);
        if (got != c.s) { std::printf("MISMATCH %.17g -> %s (want %s)\n", c.v, got.c_str(), c.s.c_str()); ++bad; }
    }
    MagnetiteError e(MagnetiteError::Kind::PostProcessor, "x");
    if (std::string(e.what()) != "Post Processor error: x") ++bad;
    try {                                        // post_processor.rs:24-31 with Rust's io::Error Display
        post_processor::csv_output({}, {}, "/nonexistent-dir/nodes.csv", "/nonexistent-dir/elements.csv", true);
        ++bad;
    } catch (const MagnetiteError &err) {
        if (std::string(err.what()) != "Solver error: Failed to create nodes.csv: No such file or directory (os error 2)") {
            std::printf("MISMATCH %s\n", err.what());
            ++bad;
        }
    }
    std::printf(bad ? "FORMAT_FAIL\n" : "FORMAT_OK\n");
    return bad;
}

// one hexadecimal 64-bit pattern per input line -> the f64 with those bits, formatted like Rust's `{}`
static int format_stdin() {
    char line[64];
    while (std::fgets(line, sizeof line, stdin)) {
        const unsigned long long bits = std::strtoull(line, nullptr, 16);
        double v;
        static_assert(sizeof v == sizeof bits, "f64 is 64 bits");
        std::memcpy(&v, &bits, sizeof v);
        std::printf("%s\n", post_processor::format_f64(v).c_str());
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc == 2 && !std::strcmp(argv[1], "--format-selftest")) return format_selftest();
    if (argc == 2 && !std::strcmp(argv[1], "--format-stdin")) return format_stdin();
    if (argc != 5) { std::fprintf(stderr, "usage: plate_demo NX NY nodes.csv elements.csv\n"); return 2; }
    const std::size_t nx = std::strtoul(argv[1], nullptr, 10), ny = std::strtoul(argv[2], nullptr, 10);
    const double h = 2.0;
    std::vector<Node> nodes;
    std::vector<Element> elements;
    for (std::size_t j = 0; j <= ny; ++j)
        for (std::size_t i = 0; i <= nx; ++i) {
            Node nd{{i * h, j * h}, std::nullopt, std::nullopt, 0.0, 0.0};          // mesher.rs:615-624
            if (i == 0) nd = Node{{i * h, j * h}, 0.0, 0.0, std::nullopt, std::nullopt};   // restraint rule
            else if (i == nx) nd = Node{{i * h, j * h}, 3.0, std::nullopt, std::nullopt, 0.0};   // load rule
            nodes.push_back(nd);
        }
    for (std::size_t j = 0; j < ny; ++j)
        for (std::size_t i = 0; i < nx; ++i) {
            const std::size_t a = j * (nx + 1) + i, b = a + 1, c = a + nx + 1, d = c + 1;
            elements.push_back(Element{{a, b, d}, std::nullopt});
            elements.push_back(Element{{a, d, c}, std::nullopt});
        }
    try {
        const ModelMetadata md{69e9, 0.33, 0.5, 0.f, 0.f};
        solver::run(nodes, elements, md);                                             // main.rs:64
        post_processor::csv_output(elements, nodes, argv[3], argv[4]);                // main.rs:69
        std::printf("area of element 0: %s\n", post_processor::format_f64(solver::compute_element_area(elements[0], nodes)).c_str());
    } catch (const MagnetiteError &err) {
        std::fprintf(stderr, "Received error: %s\n", err.what());                     // main.rs:46
        return 1;
    }
    return 0;
}
