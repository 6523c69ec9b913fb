#include "g2o.hpp"

#include <cerrno>
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
bool parse_f64(const char *t, double &v) {
    char *end = nullptr;
    v = std::strtod(t, &end);
    return end != t && *end == 0;
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

bool write_g2o(const std::string &filename, const G2oGraph &g, std::string &error) {
    FILE *f = std::fopen(filename.c_str(), "wb");
    if (!f) { error = filename + ": " + std::strerror(errno); return false; }
    const double *v = g.vertex_values.data();
    for (size_t i = 0; i < g.vertex_id.size(); i++) {
        int k = g.vertex_kind[i];
        std::fprintf(f, "%s %u", VTAG[k], g.vertex_id[i]);
        for (int c = 0; c < NVAL[k]; c++) std::fprintf(f, " %.17g", *v++);
        std::fputc('\n', f);
    }
    const double *m = g.edge_meas.data(), *w = g.edge_info_upper.data();
    for (size_t i = 0; i < g.edge_kind.size(); i++) {
        int k = g.edge_kind[i];
        std::fprintf(f, "%s %u %u", ETAG[k], g.edge_from[i], g.edge_to[i]);
        for (int c = 0; c < NMEAS[k]; c++) std::fprintf(f, " %.17g", *m++);
        for (int c = 0; c < NINFO[k]; c++) std::fprintf(f, " %.17g", *w++);
        std::fputc('\n', f);
    }
    bool ok = std::fclose(f) == 0;
    if (!ok) error = filename + ": write failed";
    return ok;
}

}} // namespace robotics::mapping
