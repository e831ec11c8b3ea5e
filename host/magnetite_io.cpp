// magnetite_io.cpp — see magnetite_io.hpp.
#include "magnetite_io.hpp"

#include <fcntl.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>

#include "../include/magnetite_b200.h"

namespace magnetite {

using Kind = MagnetiteError::Kind;

// ---------------------------------------------------------------------------
// JSON
// ---------------------------------------------------------------------------
namespace {
struct Parser {
    const std::string &s;
    size_t i = 0;
    explicit Parser(const std::string &t) : s(t) {}
    [[noreturn]] void bad(const std::string &why) const {
        throw MagnetiteError(Kind::Input, "Error in input file json: " + why + " at offset " + std::to_string(i));
    }
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
    bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { ++i; return true; } return false; }
    std::string str() {
        if (!eat('"')) bad("expected a string");
        std::string out;
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\' && i + 1 < s.size()) {
                const char e = s[++i];
                out += (e == 'n') ? '\n' : (e == 't') ? '\t' : e;
            } else {
                out += s[i];
            }
            ++i;
        }
        if (i >= s.size()) bad("unterminated string");
        ++i;
        return out;
    }
    Json value() {
        ws();
        if (i >= s.size()) bad("unexpected end");
        Json v;
        const char c = s[i];
        if (c == '{') {
            ++i; v.type = Json::Type::Object;
            if (eat('}')) return v;
            do {
                std::string k = str();
                if (!eat(':')) bad("expected ':'");
                v.object.emplace_back(std::move(k), value());
            } while (eat(','));
            if (!eat('}')) bad("expected '}'");
        } else if (c == '[') {
            ++i; v.type = Json::Type::Array;
            if (eat(']')) return v;
            do v.array.push_back(value()); while (eat(','));
            if (!eat(']')) bad("expected ']'");
        } else if (c == '"') {
            v.type = Json::Type::String; v.string = str();
        } else if (s.compare(i, 4, "null") == 0) {
            i += 4;
        } else if (s.compare(i, 4, "true") == 0) {
            i += 4; v.type = Json::Type::Bool; v.boolean = true;
        } else if (s.compare(i, 5, "false") == 0) {
            i += 5; v.type = Json::Type::Bool;
        } else {
            char *end = nullptr;
            v.number = std::strtod(s.c_str() + i, &end);
            if (end == s.c_str() + i) bad("unexpected character");
            i = (size_t)(end - s.c_str());
            v.type = Json::Type::Number;
        }
        return v;
    }
};
const Json kNull;
}  // namespace

Json Json::parse(const std::string &text) {
    Parser p(text);
    Json v = p.value();
    p.ws();
    if (p.i != text.size()) p.bad("trailing characters");
    return v;
}
bool Json::has_key(const std::string &k) const {
    for (const auto &kv : object) if (kv.first == k) return true;
    return false;
}
const Json &Json::operator[](const std::string &k) const {
    for (const auto &kv : object) if (kv.first == k) return kv.second;
    return kNull;
}
std::optional<double> Json::as_f64() const {
    if (type == Type::Number) return number;
    return std::nullopt;
}

// ---------------------------------------------------------------------------
// mesher
// ---------------------------------------------------------------------------
namespace mesher {

// Rust's Display of std::io::Error
static std::string os_error_text(int err) { return std::string(std::strerror(err)) + " (os error " + std::to_string(err) + ")"; }

Json load_input_file(const std::string &input_file) {          // mesher.rs:713-760
    std::ifstream f(input_file);
    if (!f) throw MagnetiteError(Kind::Input, "Unable to open input file " + input_file);
    std::stringstream ss;
    ss << f.rdbuf();
    Json j = Json::parse(ss.str());
    if (!j.has_key("metadata")) throw MagnetiteError(Kind::Input, "Input json missing metadata field");
    if (!j.has_key("boundary_conditions"))
        throw MagnetiteError(Kind::Input, "Input json missing boundary_conditions field in metadata section");
    for (const char *k : {"part_thickness", "material_elasticity", "poisson_ratio"})
        if (!j["metadata"].has_key(k))
            throw MagnetiteError(Kind::Input, std::string("Input json missing ") + k + " field in metadata section");
    return j;
}

ModelMetadata parse_input_metadata(const Json &j) {             // mesher.rs:769-808
    const Json &md = j["metadata"];
    const auto e = md["material_elasticity"].as_f64(), t = md["part_thickness"].as_f64();
    const auto nu = md["poisson_ratio"].as_f64();
    const auto cmin = md["characteristic_length_min"].as_f64(), cmax = md["characteristic_length_max"].as_f64();
    if (!e) throw MagnetiteError(Kind::Input, "Input json missing material elasticity");
    if (!nu) throw MagnetiteError(Kind::Input, "Input json missing poisson ratio");
    if (!cmin) throw MagnetiteError(Kind::Input, "Input json missing minimum characteristic length");
    if (!cmax) throw MagnetiteError(Kind::Input, "Input json missing maximum characteristic length");
    if (!t) throw MagnetiteError(Kind::Input, "Input json missing part thickness");     // the reference unwrap()s
    return ModelMetadata{*e, *nu, *t, (float)*cmin, (float)*cmax};
}

std::vector<BoundaryRule> parse_boundary_rules(const Json &j) { // mesher.rs:822-903
    std::vector<BoundaryRule> rules;
    constexpr double lo = std::numeric_limits<double>::lowest(), hi = std::numeric_limits<double>::max();
    for (const auto &kv : j["boundary_conditions"].object) {
        const std::string &name = kv.first;
        const Json &rule = kv.second;
        if (!rule.has_key("region")) throw MagnetiteError(Kind::Input, "Boundary rule " + name + " is missing region field");
        if (!rule.has_key("targets")) throw MagnetiteError(Kind::Input, "Boundary rule " + name + " is missing target field");
        BoundaryRegion reg{lo, hi, lo, hi};
        const Json &r = rule["region"];
        auto bound = [&](const char *key, double &dst) {
            if (!r.has_key(key)) return;
            const auto v = r[key].as_f64();
            if (!v) throw MagnetiteError(Kind::Input, std::string("Bad value for ") + key + " in " + name);
            dst = *v;
        };
        bound("x_target_min", reg.x_min); bound("x_target_max", reg.x_max);
        bound("y_target_min", reg.y_min); bound("y_target_max", reg.y_max);
        const Json &tg = rule["targets"];
        BoundaryTarget t{tg["ux"].as_f64(), tg["uy"].as_f64(), tg["fx"].as_f64(), tg["fy"].as_f64()};
        if (reg.x_min > reg.x_max) throw MagnetiteError(Kind::Input, "Boundary '" + name + "' has x_target_min greater than x_target_max");
        if (reg.y_min > reg.y_max) throw MagnetiteError(Kind::Input, "Boundary '" + name + "' has y_target_min greater than y_target_max");
        if (!t.fx && !t.ux) throw MagnetiteError(Kind::Input, "Boundary '" + name + "' is under-constrained in x-axis");
        if (!t.fy && !t.uy) throw MagnetiteError(Kind::Input, "Boundary '" + name + "' is under-constrained in y-axis");
        if (t.fx && t.ux) throw MagnetiteError(Kind::Input, "Boundary '" + name + "' is over-constrained in x-axis");
        if (t.fy && t.uy) throw MagnetiteError(Kind::Input, "Boundary '" + name + "' is over-constrained in y-axis");
        rules.push_back(BoundaryRule{name, reg, t});
    }
    return rules;
}

void apply_boundary_conditions(const Json &j, std::vector<Node> &nodes, bool quiet) {   // mesher.rs:905-927
    const std::vector<BoundaryRule> rules = parse_boundary_rules(j);
    if (!quiet) std::printf("info: loaded %zu boundary rules from input file\n", rules.size());
    for (Node &n : nodes)
        for (const BoundaryRule &r : rules)
            if (n.vertex.x > r.region.x_min && n.vertex.x < r.region.x_max && n.vertex.y > r.region.y_min &&
                n.vertex.y < r.region.y_max) {
                n.ux = r.target.ux; n.uy = r.target.uy; n.fx = r.target.fx; n.fy = r.target.fy;
            }
}

void parse_mesh(const std::string &mesh_file, std::vector<Node> &nodes, std::vector<Element> &elements) {
    std::ifstream f(mesh_file);                                  // mesher.rs:536-704 (file kept, no check_ccw)
    if (!f) throw MagnetiteError(Kind::Mesher, "Unable to open auto-generated mesh file: " + mesh_file);
    enum { Limbo, Entities, Nodes, Elements } state = Limbo;
    bool seen_meta = false;
    std::vector<Node> unordered;
    std::vector<size_t> indexes;
    elements.clear();
    std::string line;
    auto ints = [](const std::string &l) {
        std::vector<long long> v; std::stringstream ss(l); long long x;
        while (ss >> x) v.push_back(x);
        return v;
    };
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        if (line.rfind("$End", 0) == 0) state = Limbo;
        if (state == Limbo) {
            seen_meta = false;
            if (line.rfind("$Entities", 0) == 0) state = Entities;
            else if (line.rfind("$Node", 0) == 0) state = Nodes;
            else if (line.rfind("$Elements", 0) == 0) state = Elements;
            continue;
        }
        if (state == Nodes) {
            if (!seen_meta) { seen_meta = true; continue; }
            const auto head = ints(line);
            if (head.size() < 4) throw MagnetiteError(Kind::Mesher, "Unexpected non-int in mesh data");
            const size_t n_local = (size_t)head[3];
            std::vector<size_t> tags(n_local);
            for (size_t k = 0; k < n_local; ++k) { std::getline(f, line); tags[k] = (size_t)std::stoull(line); }
            for (size_t k = 0; k < n_local; ++k) {
                std::getline(f, line);
                std::stringstream ss(line);
                double x = 0, y = 0; ss >> x >> y;
                unordered.push_back(Node{{x, y}, std::nullopt, std::nullopt, 0.0, 0.0});   // mesher.rs:615-624
                indexes.push_back(tags[k] - 1);
            }
        } else if (state == Elements) {
            if (!seen_meta) { seen_meta = true; continue; }
            const auto head = ints(line);
            if (head.size() < 4) throw MagnetiteError(Kind::Mesher, "Unexpected non-int in mesh data");
            const long long entity_dim = head[0];
            for (long long k = 0; k < head[3]; ++k) {
                std::getline(f, line);
                const auto meta = ints(line);
                if (entity_dim != 2) continue;
                if (meta.size() < 4) throw MagnetiteError(Kind::Mesher, "Unexpected non-int in mesh data");
                elements.push_back(Element{{(size_t)meta[1] - 1, (size_t)meta[2] - 1, (size_t)meta[3] - 1}, std::nullopt});
            }
        }
    }
    nodes.assign(unordered.size(), Node{{0, 0}, std::nullopt, std::nullopt, std::nullopt, std::nullopt});
    std::vector<char> seen(unordered.size(), 0);
    for (size_t k = 0; k < unordered.size(); ++k) {
        if (indexes[k] >= nodes.size()) throw MagnetiteError(Kind::Mesher, "node tags are not dense 1..N");
        nodes[indexes[k]] = unordered[k];
        seen[indexes[k]] = 1;
    }
    for (char s : seen) if (!s) throw MagnetiteError(Kind::Mesher, "node tags are not dense 1..N");
}

void check_ccw(std::vector<Element> &elements, const std::vector<double> &areas) {     // mesher.rs:522-526
    for (size_t e = 0; e < elements.size(); ++e)
        if (areas[e] < 1.0) std::swap(elements[e].nodes[0], elements[e].nodes[2]);
}

std::vector<Vertex> parse_csv(const std::string &csv_file) {              // mesher.rs:253-299
    std::ifstream in(csv_file, std::ios::binary);
    if (!in) throw MagnetiteError(MagnetiteError::Kind::Input, "Unable to open csv file " + csv_file);
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string contents = ss.str();
    auto trim = [](std::string t) {                     // str::trim: Unicode white space; ASCII is what CSVs hold
        const char *ws = " \t\r\n\v\f";
        const size_t a = t.find_first_not_of(ws);
        if (a == std::string::npos) return std::string();
        return t.substr(a, t.find_last_not_of(ws) - a + 1);
    };
    auto split = [&](const std::string &line) {
        std::vector<std::string> cells;
        size_t start = 0;
        for (;;) {
            const size_t comma = line.find(',', start);
            cells.push_back(trim(line.substr(start, comma == std::string::npos ? std::string::npos : comma - start)));
            if (comma == std::string::npos) break;
            start = comma + 1;
        }
        return cells;
    };
    std::vector<Vertex> vertices;
    bool have_header = false;
    size_t xi = 0, yi = 0, pos = 0;
    while (pos <= contents.size()) {
        const size_t nl = contents.find('\n', pos);
        const std::string line = contents.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
        pos = (nl == std::string::npos) ? contents.size() + 1 : nl + 1;
        if (line.empty()) continue;
        const std::vector<std::string> cells = split(line);
        if (!have_header) {
            const auto fx = std::find(cells.begin(), cells.end(), "x"), fy = std::find(cells.begin(), cells.end(), "y");
            if (fx == cells.end() || fy == cells.end())
                throw MagnetiteError(MagnetiteError::Kind::Input, "Error in csv file: Missing x and/or y field");
            xi = (size_t)(fx - cells.begin()); yi = (size_t)(fy - cells.begin());
            have_header = true;
            continue;
        }
        std::vector<double> vals;
        for (const std::string &c : cells) {
            char *end = nullptr;
            const double v = std::strtod(c.c_str(), &end);
            if (c.empty() || end != c.c_str() + c.size())                      // the reference panics (expect)
                throw MagnetiteError(MagnetiteError::Kind::Input, "Non-float value in csv points");
            vals.push_back(v);
        }
        if (xi >= vals.size() || yi >= vals.size())                            // the reference panics (index)
            throw MagnetiteError(MagnetiteError::Kind::Input, "Non-float value in csv points");
        vertices.push_back(Vertex{vals[xi], vals[yi]});
    }
    return vertices;
}

namespace {

// Just enough XML for SVG outlines: elements in document order with their attributes and parent.  Prolog,
// comments, DOCTYPE, CDATA and text are skipped; attribute values get the XML normalisation (tab / newline ->
// space) and the five predefined entities plus numeric character references.
struct XmlNode {
    std::string name;                                          // local name (prefix stripped)
    std::vector<std::pair<std::string, std::string>> attrs;
    int parent = -1;
    const std::string *attr(const char *key) const {
        for (const auto &kv : attrs) if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

[[noreturn]] void xml_fail(const std::string &what) {
    throw MagnetiteError(MagnetiteError::Kind::Input, "Error in svg file: " + what);      // the reference unwrap()s the parser
}

std::string xml_unescape(const std::string &raw) {
    std::string out;
    for (size_t i = 0; i < raw.size(); ++i) {
        const char c = raw[i];
        if (c == '\t' || c == '\n' || c == '\r') { out += ' '; continue; }
        if (c != '&') { out += c; continue; }
        const size_t semi = raw.find(';', i);
        if (semi == std::string::npos) xml_fail("unterminated entity");
        const std::string ent = raw.substr(i + 1, semi - i - 1);
        if (ent == "amp") out += '&';
        else if (ent == "lt") out += '<';
        else if (ent == "gt") out += '>';
        else if (ent == "quot") out += '"';
        else if (ent == "apos") out += '\'';
        else if (ent.size() > 1 && ent[0] == '#') {
            const long code = (ent[1] == 'x' || ent[1] == 'X') ? std::strtol(ent.c_str() + 2, nullptr, 16)
                                                               : std::strtol(ent.c_str() + 1, nullptr, 10);
            if (code <= 0 || code > 0x7f) out += '?'; else out += (char)code;      // ids and numbers are ASCII
        } else xml_fail("unknown entity &" + ent + ";");
        i = semi;
    }
    return out;
}

std::vector<XmlNode> parse_xml(const std::string &t) {
    std::vector<XmlNode> nodes;
    std::vector<int> open;
    auto is_space = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; };
    size_t i = 0;
    while (i < t.size()) {
        if (t[i] != '<') { ++i; continue; }                                        // text
        if (t.compare(i, 4, "<!--") == 0) {
            const size_t e = t.find("-->", i + 4);
            if (e == std::string::npos) xml_fail("unterminated comment");
            i = e + 3;
        } else if (t.compare(i, 9, "<![CDATA[") == 0) {
            const size_t e = t.find("]]>", i + 9);
            if (e == std::string::npos) xml_fail("unterminated CDATA section");
            i = e + 3;
        } else if (t.compare(i, 2, "<?") == 0) {
            const size_t e = t.find("?>", i + 2);
            if (e == std::string::npos) xml_fail("unterminated processing instruction");
            i = e + 2;
        } else if (t.compare(i, 2, "<!") == 0) {                                   // DOCTYPE, possibly with an internal subset
            int depth = 0;
            size_t e = i + 2;
            for (; e < t.size(); ++e) {
                if (t[e] == '[') ++depth;
                else if (t[e] == ']') --depth;
                else if (t[e] == '>' && depth <= 0) break;
            }
            if (e >= t.size()) xml_fail("unterminated declaration");
            i = e + 1;
        } else if (t.compare(i, 2, "</") == 0) {
            const size_t e = t.find('>', i + 2);
            if (e == std::string::npos || open.empty()) xml_fail("unbalanced end tag");
            open.pop_back();
            i = e + 1;
        } else {
            size_t p = i + 1;
            const size_t name_start = p;
            while (p < t.size() && !is_space(t[p]) && t[p] != '>' && t[p] != '/') ++p;
            if (p == name_start) xml_fail("empty tag name");
            XmlNode node;
            node.name = t.substr(name_start, p - name_start);
            const size_t colon = node.name.rfind(':');
            if (colon != std::string::npos) node.name.erase(0, colon + 1);
            node.parent = open.empty() ? -1 : open.back();
            bool self_closing = false;
            for (;;) {
                while (p < t.size() && is_space(t[p])) ++p;
                if (p >= t.size()) xml_fail("unterminated start tag");
                if (t[p] == '>') { ++p; break; }
                if (t[p] == '/') {
                    if (p + 1 >= t.size() || t[p + 1] != '>') xml_fail("stray '/' in a tag");
                    self_closing = true;
                    p += 2;
                    break;
                }
                const size_t key_start = p;
                while (p < t.size() && !is_space(t[p]) && t[p] != '=' && t[p] != '>' && t[p] != '/') ++p;
                const std::string key = t.substr(key_start, p - key_start);
                while (p < t.size() && is_space(t[p])) ++p;
                if (p >= t.size() || t[p] != '=') xml_fail("attribute " + key + " has no value");
                ++p;
                while (p < t.size() && is_space(t[p])) ++p;
                if (p >= t.size() || (t[p] != '"' && t[p] != '\'')) xml_fail("attribute " + key + " is not quoted");
                const char quote = t[p++];
                const size_t end = t.find(quote, p);
                if (end == std::string::npos) xml_fail("unterminated value of attribute " + key);
                node.attrs.emplace_back(key, xml_unescape(t.substr(p, end - p)));
                p = end + 1;
            }
            nodes.push_back(std::move(node));
            if (!self_closing) open.push_back((int)nodes.size() - 1);
            i = p;
        }
    }
    if (!open.empty()) xml_fail("unclosed element <" + nodes[open.back()].name + ">");
    return nodes;
}

double svg_number(const std::string &text) {
    char *end = nullptr;
    const double v = std::strtod(text.c_str(), &end);
    if (text.empty() || end != text.c_str() + text.size())                        // the reference panics (expect)
        throw MagnetiteError(MagnetiteError::Kind::Input, "Non-float value in svg points");
    return v;
}

}  // namespace

std::vector<std::vector<Vertex>> parse_svg(const std::string &svg_file, float min_element_length) {   // mesher.rs:26-244
    std::ifstream in(svg_file, std::ios::binary);
    if (!in) throw MagnetiteError(MagnetiteError::Kind::Input, "Unable to open svg file " + svg_file);
    std::stringstream ss;
    ss << in.rdbuf();
    const std::vector<XmlNode> nodes = parse_xml(ss.str());
    std::vector<std::vector<Vertex>> containers(1);                               // [0]: placeholder for OUTER
    auto trim_left = [](const std::string &t) {
        const size_t a = t.find_first_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : t.substr(a);
    };
    auto file_under = [&](const XmlNode &el, std::vector<Vertex> pts) {            // mesher.rs:97-128, 213-236
        const std::string *id = el.attr("id");
        if (!id && el.parent >= 0) id = nodes[(size_t)el.parent].attr("id");
        if (!id) throw MagnetiteError(MagnetiteError::Kind::Input, "Error in svg file. Missing id field on polyline");
        const std::string t = trim_left(*id);
        if (t.rfind("INNER", 0) == 0) containers.push_back(std::move(pts));
        else if (t.rfind("OUTER", 0) == 0) {
            if (!containers[0].empty()) throw MagnetiteError(MagnetiteError::Kind::Input, "Multiple OUTER geometries in SVG");
            containers[0] = std::move(pts);
        }                                                                          // other ids: the reference warns and skips
    };
    for (const XmlNode &el : nodes) {
        if (el.name != "polyline" && el.name != "polygon") continue;
        const std::string *raw = el.attr("points");
        if (!raw) throw MagnetiteError(MagnetiteError::Kind::Input, "Error in svg file. No points in polyline element");
        std::vector<double> flat;
        for (size_t a = 0; a <= raw->size();) {
            const size_t b = raw->find(' ', a);
            const std::string piece = raw->substr(a, b == std::string::npos ? std::string::npos : b - a);
            if (!piece.empty()) flat.push_back(svg_number(piece));
            if (b == std::string::npos) break;
            a = b + 1;
        }
        std::vector<Vertex> pts;
        for (size_t i = 0; i + 1 < flat.size(); i += 2) {
            const Vertex v{flat[i], -flat[i + 1]};                                 // mesher.rs:73: y inverted
            bool seen = false;
            for (const Vertex &p : pts) if (p.x == v.x && p.y == v.y) { seen = true; break; }
            if (seen) continue;
            if (!pts.empty() && std::hypot(pts.back().x - v.x, pts.back().y - v.y) < (double)min_element_length) continue;
            pts.push_back(v);
        }
        file_under(el, std::move(pts));
    }
    for (const XmlNode &el : nodes) {
        if (el.name != "rect") continue;
        const double x = el.attr("x") ? svg_number(*el.attr("x")) : 0.0, y = el.attr("y") ? svg_number(*el.attr("y")) : 0.0;
        if (!el.attr("width") || !el.attr("height"))
            throw MagnetiteError(MagnetiteError::Kind::Input, "Error in svg file. No width/height definition in rectangle.");
        const double w = svg_number(*el.attr("width")), h = svg_number(*el.attr("height"));
        file_under(el, {Vertex{x, -y}, Vertex{x + w, -y}, Vertex{x + w, -y - h}, Vertex{x, -y - h}});
    }
    if (containers[0].empty()) throw MagnetiteError(MagnetiteError::Kind::Input, "No OUTER geometry");
    return containers;
}

std::string geo_text(const std::vector<std::vector<Vertex>> &vc, float cl_min, float cl_max) {   // mesher.rs:305-472
    if (vc.empty()) throw MagnetiteError(MagnetiteError::Kind::Input, "no geometry was given");   // the reference panics
    const auto f = post_processor::format_f64;
    auto point = [&](size_t id, const Vertex &v) {
        return "Point(" + std::to_string(id) + ") = { " + f(v.x) + ", " + f(v.y) + ", 0, 1.0 };\n";
    };
    std::string out = "// Define outer points\n";
    for (size_t i = 0; i < vc[0].size(); ++i) out += point(i, vc[0][i]);
    out += "\n// Define inner points\n";
    size_t offset = vc[0].size();
    std::vector<size_t> inner_offsets{0};
    for (size_t c = 1; c < vc.size(); ++c) {
        inner_offsets.push_back(offset);
        for (size_t i = 0; i < vc[c].size(); ++i) out += point(i + offset, vc[c][i]);
        offset += vc[c].size();
    }
    out += "\n// Connect points\n";
    auto line = [](size_t id, size_t a, size_t b) {
        return "Line(" + std::to_string(id) + ") = { " + std::to_string(a) + ", " + std::to_string(b) + " };\n";
    };
    for (size_t c = 0; c < vc.size(); ++c) {
        out += "\n// Point connections for surface " + std::to_string(c) + "\n";
        const size_t off = inner_offsets[c], n = vc[c].size();
        for (size_t k = 1; k < n; ++k) out += line(k + off - 1, k + off - 1, k + off);
        out += line(n + off - 1, n + off - 1, off);
    }
    out += "\n//Register loops\n";
    for (size_t c = 0; c < vc.size(); ++c) {
        out += "Line Loop(" + std::to_string(c + 1) + ") = {";
        for (size_t k = 0; k < vc[c].size(); ++k) out += std::string(k ? "," : "") + " " + std::to_string(k + inner_offsets[c]);
        out += " };\n";
    }
    out += "\n//Define surface\nPlane Surface(1) = {";
    for (size_t k = 0; k < vc.size(); ++k) {                                   // mesher.rs:424-430
        const size_t loop = vc.size() > 2 ? k : vc.size() - 1 - k;
        out += std::string(k ? "," : "") + " " + std::to_string(loop + 1);
    }
    out += " };\n";
    out += "\n// Define Mesh Settings\nMesh.ElementOrder = 1;\nMesh.Algorithm  = 1;\nMesh.CharacteristicLengthMin = " +
           post_processor::format_f32(cl_min) + ";\nMesh.CharacteristicLengthMax = " + post_processor::format_f32(cl_max) +
           ";\nMesh 2;\n";
    return out;
}

void build_geo(const std::vector<std::vector<Vertex>> &vc, const std::string &output_file, float cl_min, float cl_max) {
    const std::string text = geo_text(vc, cl_min, cl_max);
    std::ofstream out(output_file, std::ios::binary | std::ios::trunc);
    if (!out) throw MagnetiteError(MagnetiteError::Kind::Mesher, "Failed to create .geo file");   // the reference panics
    out << text;
}

void compute_mesh(const std::vector<std::vector<Vertex>> &vertices, const std::string &output, float cl_min, float cl_max,
                  bool quiet) {                                                // mesher.rs:481-519
    const std::string geo_filepath = "geom.geo";
    if (!quiet) std::printf("info: building .geo for Gmsh with %.3f< CL < %.3f\n", (double)cl_min, (double)cl_max);
    build_geo(vertices, geo_filepath, cl_min, cl_max);
    if (!quiet) std::printf("info: running gmsh...\n");
    std::fflush(stdout);
    // fork/exec with a close-on-exec pipe: the child reports errno through it only if exec itself fails
    int report[2];
    if (pipe2(report, O_CLOEXEC) != 0) throw MagnetiteError(MagnetiteError::Kind::Mesher, "Gmsh failed: " + os_error_text(errno));
    const pid_t pid = fork();
    if (pid < 0) {
        const int err = errno;
        close(report[0]); close(report[1]);
        throw MagnetiteError(MagnetiteError::Kind::Mesher, "Gmsh failed: " + os_error_text(err));
    }
    if (pid == 0) {                                                            // child: gmsh geom.geo -2 -o <output>, output discarded
        const int devnull = open("/dev/null", O_WRONLY);
        if (devnull >= 0) { dup2(devnull, 1); dup2(devnull, 2); }
        execlp("gmsh", "gmsh", geo_filepath.c_str(), "-2", "-o", output.c_str(), (char *)nullptr);
        const int err = errno;
        if (write(report[1], &err, sizeof err) < 0) {}
        _exit(127);
    }
    close(report[1]);
    int exec_errno = 0;
    ssize_t got;
    while ((got = read(report[0], &exec_errno, sizeof exec_errno)) < 0 && errno == EINTR) {}
    close(report[0]);
    int status = 0;
    while (waitpid(pid, &status, 0) < 0 && errno == EINTR) {}
    // Only a gmsh that could not be STARTED is an error (mesher.rs:503-513); its exit status is ignored and a
    // failed meshing run surfaces in parse_mesh ("Unable to open auto-generated mesh file").
    if (got == (ssize_t)sizeof exec_errno)
        throw MagnetiteError(MagnetiteError::Kind::Mesher, "Gmsh failed: " + os_error_text(exec_errno));
    std::remove(geo_filepath.c_str());
}

}  // namespace mesher

namespace solver {
std::vector<double> element_areas(const std::vector<Element> &elements, const std::vector<Node> &nodes, int device) {
    const size_t n = nodes.size(), e = elements.size();
    std::vector<double> x(n), y(n), area(e);
    std::vector<std::uint8_t> known(n, 0);
    std::vector<std::uint32_t> n0(e), n1(e), n2(e);
    for (size_t i = 0; i < n; ++i) { x[i] = nodes[i].vertex.x; y[i] = nodes[i].vertex.y; }
    for (size_t i = 0; i < e; ++i) {
        n0[i] = (std::uint32_t)elements[i].nodes[0]; n1[i] = (std::uint32_t)elements[i].nodes[1];
        n2[i] = (std::uint32_t)elements[i].nodes[2];
    }
    mag_mesh m{};
    m.n_nodes = n; m.n_elems = e; m.x = x.data(); m.y = y.data();
    m.n0 = n0.data(); m.n1 = n1.data(); m.n2 = n2.data(); m.known = known.data();
    mag_ctx *ctx = nullptr;
    int rc = mag_ctx_create(&ctx, device);
    if (rc != MAG_OK) throw MagnetiteError(Kind::Solver, mag_last_error(), rc);
    rc = mag_element_area(ctx, &m, area.data());
    const std::string msg = rc != MAG_OK ? mag_last_error() : "";
    mag_ctx_destroy(ctx);
    if (rc != MAG_OK) throw MagnetiteError(Kind::Solver, msg, rc);
    return area;
}
}  // namespace solver

}  // namespace magnetite
