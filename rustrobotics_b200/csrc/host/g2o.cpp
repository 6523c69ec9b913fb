#include "g2o.hpp"

#include <cerrno>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace robotics { namespace mapping {

namespace {
const int NVAL[3] = {3, 2, 7}, DIM[3] = {3, 2, 6}, NMEAS[3] = {3, 2, 7}, NINFO[3] = {6, 3, 21};
const char *VTAG[3] = {"VERTEX_SE2", "VERTEX_XY", "VERTEX_SE3:QUAT"};
const char *ETAG[3] = {"EDGE_SE2", "EDGE_SE2_XY", "EDGE_SE3:QUAT"};

bool parse_u32(const char *t, uint32_t &v) {           // Rust's str::parse::<u32>: digits only (optional '+')
    if (*t == '+') t++;
    if (!*t) return false;
    uint64_t a = 0;
    for (; *t; t++) { if (*t < '0' || *t > '9') return false; a = a * 10 + (uint64_t)(*t - '0'); if (a > 0xffffffffull) return false; }
    v = (uint32_t)a;
    return true;
}
// Rust's str::parse::<f64>: decimal / exponent notation, "inf" / "nan", optional sign; no hex floats, and independent of the
// process locale (std::from_chars, unlike strtod, never looks at LC_NUMERIC)
bool parse_f64(const char *t, double &v) {
    const char *end = t + std::strlen(t);
    if (*t == '+' && t[1] != '+' && t[1] != '-') t++;
    if (t == end) return false;
    const std::from_chars_result r = std::from_chars(t, end, v, std::chars_format::general);
    return r.ec == std::errc() && r.ptr == end;
}
// shortest-round-trip-safe %.17g, locale-independent
void put_f64(FILE *f, double v) {
    char buf[40];
    const std::to_chars_result r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::general, 17);
    std::fputc(' ', f);
    std::fwrite(buf, 1, (size_t)(r.ptr - buf), f);
}
} // namespace

bool parse_g2o(const std::string &filename, G2oGraph &g, std::string &error) {
    g = G2oGraph();
    FILE *f = std::fopen(filename.c_str(), "rb");
    if (!f) { error = filename + ": " + std::strerror(errno); return false; }
    std::string buf;
    {
        char tmp[1 << 16]; size_t n;
        while ((n = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.append(tmp, n);
        std::fclose(f);
    }
    std::vector<char *> tok;
    int64_t lineno = 0;
    size_t pos = 0;
    auto fail = [&](const std::string &m) { error = filename + ":" + std::to_string(lineno) + ": " + m; return false; };
    while (pos < buf.size()) {                          // str::lines(): split on \n, strip a trailing \r
        size_t e = buf.find('\n', pos);
        if (e == std::string::npos) e = buf.size();
        size_t le = e;
        if (le > pos && buf[le - 1] == '\r') le--;
        lineno++;
        tok.clear();
        char *s = &buf[pos], *end = &buf[0] + le;
        *end = 0;
        while (s < end) {                               // split(' ') and drop empty tokens (g2o.rs:52)
            while (s < end && *s == ' ') s++;
            if (s >= end) break;
            tok.push_back(s);
            while (s < end && *s != ' ') s++;
            if (s < end) *s++ = 0;
        }
        pos = e + 1;
        if (tok.empty()) return fail("blank line");
        int vk = -1, ek = -1;
        for (int k = 0; k < 3; k++) { if (!std::strcmp(tok[0], VTAG[k])) vk = k; if (!std::strcmp(tok[0], ETAG[k])) ek = k; }
        if (vk < 0 && ek < 0) return fail(std::string("not implemented: ") + tok[0]);
        if (vk >= 0) {
            uint32_t id;
            if (tok.size() < 2 || !parse_u32(tok[1], id)) return fail("bad vertex id");
            if ((int)tok.size() - 2 != NVAL[vk]) return fail("wrong number of fields");
            for (int i = 0; i < NVAL[vk]; i++) { double v; if (!parse_f64(tok[2 + i], v)) return fail(std::string("bad number ") + tok[2 + i]); g.vertex_values.push_back(v); }
            g.vertex_id.push_back(id); g.vertex_kind.push_back((uint8_t)vk);
            g.len += DIM[vk];
        } else {
            uint32_t a, b;
            if (tok.size() < 3 || !parse_u32(tok[1], a) || !parse_u32(tok[2], b)) return fail("bad edge endpoint id");
            if ((int)tok.size() - 3 != NMEAS[ek] + NINFO[ek]) return fail("wrong number of fields");
            for (int i = 0; i < NMEAS[ek]; i++) { double v; if (!parse_f64(tok[3 + i], v)) return fail(std::string("bad number ") + tok[3 + i]); g.edge_meas.push_back(v); }
            for (int i = 0; i < NINFO[ek]; i++) { double v; if (!parse_f64(tok[3 + NMEAS[ek] + i], v)) return fail(std::string("bad number ") + tok[3 + NMEAS[ek] + i]); g.edge_info_upper.push_back(v); }
            g.edge_kind.push_back((uint8_t)ek); g.edge_from.push_back(a); g.edge_to.push_back(b);
        }
    }
    return true;
}

bool validate_graph_arrays(size_t n_vertices, const uint8_t *vertex_kind, size_t n_values, size_t n_edges, const uint8_t *edge_kind,
                           size_t n_meas, size_t n_info, std::string &error) {
    size_t nval = 0, nmeas = 0, ninfo = 0;
    for (size_t i = 0; i < n_vertices; i++) {
        if (vertex_kind[i] > 2) { error = "graph arrays: vertex_kind[" + std::to_string(i) + "] = " + std::to_string(vertex_kind[i]) + " (must be 0, 1 or 2)"; return false; }
        nval += (size_t)NVAL[vertex_kind[i]];
    }
    for (size_t i = 0; i < n_edges; i++) {
        if (edge_kind[i] > 2) { error = "graph arrays: edge_kind[" + std::to_string(i) + "] = " + std::to_string(edge_kind[i]) + " (must be 0, 1 or 2)"; return false; }
        nmeas += (size_t)NMEAS[edge_kind[i]]; ninfo += (size_t)NINFO[edge_kind[i]];
    }
    if (n_values != nval) { error = "graph arrays: " + std::to_string(n_values) + " vertex values, the vertex kinds need " + std::to_string(nval); return false; }
    if (n_meas != nmeas) { error = "graph arrays: " + std::to_string(n_meas) + " edge measurement values, the edge kinds need " + std::to_string(nmeas); return false; }
    if (n_info != ninfo) { error = "graph arrays: " + std::to_string(n_info) + " edge information values, the edge kinds need " + std::to_string(ninfo); return false; }
    return true;
}

bool validate_graph(const G2oGraph &g, std::string &error) {
    if (g.vertex_kind.size() != g.vertex_id.size()) { error = "graph arrays: vertex_kind and vertex_id differ in length"; return false; }
    if (g.edge_from.size() != g.edge_kind.size() || g.edge_to.size() != g.edge_kind.size()) { error = "graph arrays: edge_kind / edge_from / edge_to differ in length"; return false; }
    return validate_graph_arrays(g.vertex_id.size(), g.vertex_kind.data(), g.vertex_values.size(), g.edge_kind.size(), g.edge_kind.data(),
                                 g.edge_meas.size(), g.edge_info_upper.size(), error);
}

bool write_g2o(const std::string &filename, const G2oGraph &g, std::string &error) {
    if (!validate_graph(g, error)) return false;
    FILE *f = std::fopen(filename.c_str(), "wb");
    if (!f) { error = filename + ": " + std::strerror(errno); return false; }
    const double *v = g.vertex_values.data();
    for (size_t i = 0; i < g.vertex_id.size(); i++) {
        int k = g.vertex_kind[i];
        std::fprintf(f, "%s %u", VTAG[k], g.vertex_id[i]);
        for (int c = 0; c < NVAL[k]; c++) put_f64(f, *v++);
        std::fputc('\n', f);
    }
    const double *m = g.edge_meas.data(), *w = g.edge_info_upper.data();
    for (size_t i = 0; i < g.edge_kind.size(); i++) {
        int k = g.edge_kind[i];
        std::fprintf(f, "%s %u %u", ETAG[k], g.edge_from[i], g.edge_to[i]);
        for (int c = 0; c < NMEAS[k]; c++) put_f64(f, *m++);
        for (int c = 0; c < NINFO[k]; c++) put_f64(f, *w++);
        std::fputc('\n', f);
    }
    bool ok = std::fclose(f) == 0;
    if (!ok) error = filename + ": write failed";
    return ok;
}

}} // namespace robotics::mapping
