// csv.cpp — host-only output stage of the library: the reference's csv_output
// (src/post_processor.rs:18-83) with Rust's `{}` formatting for f64, buffered.  The reference issues
// one unbuffered write per row; at 8 M nodes + 16 M elements that alone would take longer than the
// whole GPU solve, so the writer formats into 1 MiB chunks.  No CUDA here.
#include <charconv>
#include <cmath>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/magnetite_b200.h"

namespace {

// Rust `Display` for f64: shortest round-trip digits, positional (never scientific), zero padded,
// no ".0" on integral values, "NaN" / "inf" / "-inf".  Returns the number of characters written.
size_t format_f64(double v, char *out) {
    if (std::isnan(v)) { std::memcpy(out, "NaN", 3); return 3; }
    if (std::isinf(v)) { const char *s = v > 0 ? "inf" : "-inf"; const size_t n = std::strlen(s); std::memcpy(out, s, n); return n; }
    char sci[40];
    const auto r = std::to_chars(sci, sci + sizeof sci, v, std::chars_format::scientific);
    char *p = out;
    const char *s = sci;
    if (*s == '-') { *p++ = '-'; ++s; }
    const char *e = s;
    while (e < r.ptr && *e != 'e') ++e;
    // to_chars does not NUL-terminate: parse the exponent bounded by r.ptr ("e+05" / "e-324").
    int exp10 = 0;
    {
        const char *q = e + 1;
        const bool neg = q < r.ptr && *q == '-';
        if (q < r.ptr && (*q == '+' || *q == '-')) ++q;
        for (; q < r.ptr; ++q) exp10 = exp10 * 10 + (*q - '0');
        if (neg) exp10 = -exp10;
    }
    char digits[24];
    int n = 0;
    for (const char *c = s; c < e; ++c) if (*c != '.') digits[n++] = *c;
    if (exp10 >= n - 1) {
        std::memcpy(p, digits, (size_t)n); p += n;
        std::memset(p, '0', (size_t)(exp10 - (n - 1))); p += exp10 - (n - 1);
    } else if (exp10 >= 0) {
        std::memcpy(p, digits, (size_t)exp10 + 1); p += exp10 + 1;
        *p++ = '.';
        std::memcpy(p, digits + exp10 + 1, (size_t)(n - exp10 - 1)); p += n - exp10 - 1;
    } else {
        *p++ = '0'; *p++ = '.';
        std::memset(p, '0', (size_t)(-exp10 - 1)); p += -exp10 - 1;
        std::memcpy(p, digits, (size_t)n); p += n;
    }
    return (size_t)(p - out);
}

// Rust's Display of std::io::Error, which the reference embeds in its message (post_processor.rs:27-29)
std::string os_error(int err) { return std::string(std::strerror(err)) + " (os error " + std::to_string(err) + ")"; }

struct Chunked {
    std::FILE *f;
    std::string buf;
    explicit Chunked(std::FILE *file) : f(file) { buf.reserve((1u << 20) + 2048); }
    bool flush() {
        const bool ok = buf.empty() || std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
        buf.clear();
        return ok;
    }
    bool room() { return buf.size() < (1u << 20) || flush(); }
};

}  // namespace

// message of the last failed host-only entry point on this thread (csv.cpp, reorder.cpp)
namespace maghost {
thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
}  // namespace maghost

extern "C" size_t mag_format_f64(double v, char *out /* >= 400 bytes */) {
    const size_t n = format_f64(v, out);
    out[n] = '\0';
    return n;
}

extern "C" const char *mag_host_last_error(void) { return maghost::g_error.c_str(); }

extern "C" int mag_csv_output(const char *nodes_path, const char *elements_path, uint64_t n_nodes, const double *x,
                              const double *y, const double *ux, const double *uy, uint64_t n_elems,
                              const uint32_t *n0, const uint32_t *n1, const uint32_t *n2, const double *stress) {
    if (!nodes_path || !elements_path || (n_nodes && (!x || !y || !ux || !uy)) ||
        (n_elems && (!n0 || !n1 || !n2 || !stress))) {
        maghost::set_error("null argument");
        return MAG_ERR_BAD_ARG;
    }
    std::FILE *nf = std::fopen(nodes_path, "wb");                 // post_processor.rs:24-31
    if (!nf) { maghost::set_error("Failed to create nodes.csv: " + os_error(errno)); return MAG_ERR_BAD_ARG; }
    std::FILE *ef = std::fopen(elements_path, "wb");              // post_processor.rs:32-39
    if (!ef) {
        maghost::set_error("Failed to create elements.csv: " + os_error(errno));
        std::fclose(nf);
        return MAG_ERR_BAD_ARG;
    }
    bool ok = true;
    char tmp[512];
    {
        Chunked w(nf);
        w.buf += "x,y,ux,uy\n";                                   // post_processor.rs:42
        for (uint64_t i = 0; i < n_nodes && ok; ++i) {
            const double v[4] = {x[i], y[i], ux[i], uy[i]};
            for (int k = 0; k < 4; ++k) {
                w.buf.append(tmp, format_f64(v[k], tmp));
                w.buf += (k < 3) ? ',' : '\n';
            }
            ok = w.room();
        }
        ok = w.flush() && ok;
    }
    {
        Chunked w(ef);
        w.buf += "n0,n1,n2,stress\n";                             // post_processor.rs:60
        for (uint64_t i = 0; i < n_elems && ok; ++i) {
            const uint32_t c[3] = {n0[i], n1[i], n2[i]};
            for (int k = 0; k < 3; ++k) {
                const auto r = std::to_chars(tmp, tmp + 16, c[k]);
                w.buf.append(tmp, (size_t)(r.ptr - tmp));
                w.buf += ',';
            }
            w.buf.append(tmp, format_f64(stress[i], tmp));
            w.buf += '\n';
            ok = w.room();
        }
        ok = w.flush() && ok;
    }
    ok = (std::fclose(nf) == 0) && ok;
    ok = (std::fclose(ef) == 0) && ok;
    if (!ok) { maghost::set_error("short write"); return MAG_ERR_BAD_ARG; }
    return MAG_OK;
}
