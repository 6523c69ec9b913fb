#include "pose_graph.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <future>
#include <sys/stat.h>
#include <utility>

namespace robotics { namespace mapping {

void PoseGraph::check(int rc, const char *what) const {
    if (rc == PGO_OK) return;
    throw Error(std::string(what) + ": " + pgo_last_error(h_));
}

void PoseGraph::init(const pgo_options *options) {
    const G2oGraph &g = graph_;
    std::string verr;
    if (!validate_graph(g, verr)) throw Error(verr);
    int rc = pgo_create(&h_, options, (int64_t)g.vertex_id.size(), g.vertex_id.data(), g.vertex_kind.data(), g.vertex_values.data(),
                        (int64_t)g.edge_kind.size(), g.edge_kind.data(), g.edge_from.data(), g.edge_to.data(), g.edge_meas.data(),
                        g.edge_info_upper.data());
    if (rc != PGO_OK) throw Error(std::string("pgo_create: ") + pgo_last_error(nullptr));
}

PoseGraph::PoseGraph(const std::string &file_path, PoseGraphSolver solver, const pgo_options *options) : solver_(solver) {
    std::string err;
    if (!parse_g2o(file_path, graph_, err)) throw Error(err);
    // name = file stem (:217)
    size_t s = file_path.find_last_of('/');
    name_ = file_path.substr(s == std::string::npos ? 0 : s + 1);
    size_t d = name_.find_last_of('.');
    if (d != std::string::npos && d > 0) name_ = name_.substr(0, d);
    init(options);
}

PoseGraph::PoseGraph(const G2oGraph &graph, const std::string &name, PoseGraphSolver solver, const pgo_options *options)
    : graph_(graph), name_(name), solver_(solver) {
    init(options);
}

PoseGraph::PoseGraph(G2oGraph &&graph, const std::string &name, PoseGraphSolver solver, const pgo_options *options)
    : graph_(std::move(graph)), name_(name), solver_(solver) {
    init(options);
}

PoseGraph::PoseGraph(G2oGraph &&graph, const std::string &name, PoseGraphSolver solver, pgo_handle *adopted)
    : graph_(std::move(graph)), name_(name), solver_(solver), h_(adopted) {}

PoseGraph::~PoseGraph() { pgo_destroy(h_); }

double PoseGraph::global_error() const {
    double c = 0;
    check(pgo_chi2(h_, &c), "global_error");
    return c;
}

std::vector<double> PoseGraph::linearize_and_solve() {
    int32_t it = 0;
    check(pgo_linearize_and_solve(h_, &it), "linearize_and_solve");
    std::vector<double> dx(graph_.len);
    check(pgo_get_dx(h_, dx.data(), graph_.len), "get_dx");
    return dx;
}

// optimize (:247-303).  tolerance 1e-4 (:253), lambda 0.01 (:254); LM doubles lambda and undoes the
// step when the error went up, halves it otherwise (:275-282); the error pushed to the history is
// the one measured right after the step even when the step is rejected (:284-286).
std::vector<double> PoseGraph::optimize(size_t num_iterations, bool log, bool do_plot) {
    const double tolerance = 1e-4;
    double lambda = 0.01;
    double last_error = global_error();
    std::vector<double> errors{last_error};
    if (log) {
        std::printf("Loaded graph with %zu nodes and %zu edges\n", num_nodes(), num_edges());
        std::printf("initial error :%.5f\n", errors.back());
    }
    if (do_plot) plot();
    const bool lm = solver_ == PoseGraphSolver::LevenbergMarquardt;
    for (size_t i = 0; i < num_iterations; i++) {
        iteration_ += 1;
        double norm_dx = 0, error = 0;
        int32_t it = 0;
        int rc = pgo_gn_step(h_, lambda, lm ? 1 : 0, &norm_dx, &error, &it);
        if (rc != PGO_OK && rc != PGO_ERR_NOT_CONVERGED) check(rc, "gn_step");
        if (lm) {
            if (last_error < error) { check(pgo_undo_last_step(h_), "undo_last_step"); lambda *= 2.0; }
            else lambda /= 2.0;
        }
        last_error = error;
        norms_.push_back(norm_dx);
        pcg_iters_.push_back(it);
        errors.push_back(error);
        if (log) std::printf("step %3zu : |dx| = %3.5f, error = %3.5f\n", i, norm_dx, errors.back());
        if (do_plot) plot();
        if (norm_dx < tolerance) break;
    }
    return errors;
}

std::vector<Pose> PoseGraph::poses() const {
    std::vector<double> v(graph_.vertex_values.size());
    check(pgo_get_poses(h_, v.data(), (int64_t)v.size()), "get_poses");
    std::vector<Pose> out(graph_.vertex_id.size());
    const double *p = v.data();
    for (size_t i = 0; i < out.size(); i++) {
        const uint8_t k = graph_.vertex_kind[i];
        out[i] = Pose{graph_.vertex_id[i], k, p[0], p[1], k == 0 ? p[2] : 0.0};
        p += (k == 0 ? 3 : k == 1 ? 2 : 7);
    }
    return out;
}

// plot (:375-431): poses as blue dots, the pose sequence (sorted by id) as a red line, landmarks as
// red stars, equal axes; the reference goes through plotpy/matplotlib, here the SVG is written directly.
void PoseGraph::plot() const {
    std::vector<Pose> ps = poses();
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (const Pose &p : ps) { x0 = std::min(x0, p.x); x1 = std::max(x1, p.x); y0 = std::min(y0, p.y); y1 = std::max(y1, p.y); }
    if (ps.empty()) { x0 = y0 = 0; x1 = y1 = 1; }
    const double span = std::max(std::max(x1 - x0, y1 - y0), 1e-9), W = 800, M = 20, sc = (W - 2 * M) / span;
    auto X = [&](double x) { return M + (x - x0) * sc; };
    auto Y = [&](double y) { return W - M - (y - y0) * sc; };
    ::mkdir("img", 0777);
    const char *sname = solver_ == PoseGraphSolver::GaussNewton ? "GaussNewton" : "LevenbergMarquardt";
    const std::string fn = "img/" + name_ + "-" + std::to_string(iteration_) + "-" + sname + ".svg";
    FILE *f = std::fopen(fn.c_str(), "w");
    if (!f) throw Error(fn + ": cannot write");
    std::fprintf(f, "<svg xmlns=\"http://www.w3.org/2000/svg\" width=\"%g\" height=\"%g\" viewBox=\"0 0 %g %g\">\n", W, W, W, W);
    std::vector<const Pose *> seq;
    for (const Pose &p : ps) if (p.kind == 0) seq.push_back(&p);
    std::sort(seq.begin(), seq.end(), [](const Pose *a, const Pose *b) { return a->id < b->id; });
    std::fprintf(f, "<polyline fill=\"none\" stroke=\"red\" stroke-width=\"1\" points=\"");
    for (const Pose *p : seq) std::fprintf(f, "%.2f,%.2f ", X(p->x), Y(p->y));
    std::fprintf(f, "\"/>\n");
    for (const Pose &p : ps) {
        if (p.kind == 0) std::fprintf(f, "<circle cx=\"%.2f\" cy=\"%.2f\" r=\"1.5\" fill=\"blue\"/>\n", X(p.x), Y(p.y));
        else std::fprintf(f, "<text x=\"%.2f\" y=\"%.2f\" fill=\"red\" font-size=\"10\" text-anchor=\"middle\">*</text>\n", X(p.x), Y(p.y) + 4);
    }
    std::fprintf(f, "</svg>\n");
    std::fclose(f);
}

}} // namespace robotics::mapping

// ------------------------------------------------------------------------------------------------
// C wrappers over the C++ mirror so that Python (ctypes) drives exactly the code a C++ user would.
using namespace robotics::mapping;
namespace { thread_local std::string g_err; }

extern "C" {

const char *pg_last_error(void) { return g_err.c_str(); }

void *pg_new(const char *file_path, int solver, const pgo_options *opt) {
    try { return new PoseGraph(file_path, solver ? PoseGraphSolver::LevenbergMarquardt : PoseGraphSolver::GaussNewton, opt); }
    catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

void *pg_from_arrays(const char *name, int solver, const pgo_options *opt,
                     int64_t nv, const uint32_t *vid, const uint8_t *vkind, const double *vval, int64_t n_values,
                     int64_t ne, const uint8_t *ekind, const uint32_t *efrom, const uint32_t *eto,
                     const double *emeas, int64_t n_meas, const double *einfo, int64_t n_info) {
    try {
        if (nv < 0 || ne < 0 || n_values < 0 || n_meas < 0 || n_info < 0) throw Error("graph arrays: negative length");
        std::string verr;
        if (!validate_graph_arrays((size_t)nv, vkind, (size_t)n_values, (size_t)ne, ekind, (size_t)n_meas, (size_t)n_info, verr)) throw Error(verr);
        // the PoseGraph owns a copy of the graph (like the reference's, :157-159); the ~0.4 GB copy at 1M poses runs beside pgo_create
        std::future<G2oGraph> copy = std::async(std::launch::async, [&]() {
            G2oGraph g;
            g.vertex_id.assign(vid, vid + nv); g.vertex_kind.assign(vkind, vkind + nv); g.vertex_values.assign(vval, vval + n_values);
            g.edge_kind.assign(ekind, ekind + ne); g.edge_from.assign(efrom, efrom + ne); g.edge_to.assign(eto, eto + ne);
            g.edge_meas.assign(emeas, emeas + n_meas); g.edge_info_upper.assign(einfo, einfo + n_info);
            for (int64_t i = 0; i < nv; i++) g.len += vkind[i] == 0 ? 3 : vkind[i] == 1 ? 2 : 6;
            return g;
        });
        pgo_handle *h = nullptr;
        const int rc = pgo_create(&h, opt, nv, vid, vkind, vval, ne, ekind, efrom, eto, emeas, einfo);
        G2oGraph g;
        try { g = copy.get(); } catch (...) { pgo_destroy(h); throw; }          // (out of memory in the copy)
        if (rc != PGO_OK) throw Error(std::string("pgo_create: ") + pgo_last_error(nullptr));
        try {
            return new PoseGraph(std::move(g), name ? name : "graph", solver ? PoseGraphSolver::LevenbergMarquardt : PoseGraphSolver::GaussNewton, h);
        } catch (...) { pgo_destroy(h); throw; }
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

void pg_free(void *pg) { delete static_cast<PoseGraph *>(pg); }
pgo_handle *pg_handle(void *pg) { return static_cast<PoseGraph *>(pg)->handle(); }
int64_t pg_num_nodes(void *pg) { return (int64_t)static_cast<PoseGraph *>(pg)->num_nodes(); }
int64_t pg_num_edges(void *pg) { return (int64_t)static_cast<PoseGraph *>(pg)->num_edges(); }
int64_t pg_len(void *pg) { return (int64_t)static_cast<PoseGraph *>(pg)->len(); }

// returns the number of chi2 values written (history length), or -1 on error
int64_t pg_optimize(void *pg, int64_t num_iterations, int log, int plot, double *errors_out, int64_t cap,
                    double *norms_out, int32_t *pcg_iters_out) {
    try {
        PoseGraph *g = static_cast<PoseGraph *>(pg);
        const size_t n0 = g->norms().size();
        std::vector<double> e = g->optimize((size_t)num_iterations, log != 0, plot != 0);
        for (size_t i = 0; i < e.size() && (int64_t)i < cap; i++) errors_out[i] = e[i];
        for (size_t i = n0; i < g->norms().size() && (int64_t)(i - n0) < cap; i++) {
            if (norms_out) norms_out[i - n0] = g->norms()[i];
            if (pcg_iters_out) pcg_iters_out[i - n0] = g->pcg_iterations()[i];
        }
        return (int64_t)e.size();
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

int pg_global_error(void *pg, double *out) {
    try { *out = static_cast<PoseGraph *>(pg)->global_error(); return 0; }
    catch (const std::exception &e) { g_err = e.what(); return -1; }
}

int pg_plot(void *pg) {
    try { static_cast<PoseGraph *>(pg)->plot(); return 0; }
    catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// g2o loader on its own: two-call pattern (sizes, then fill)
void *pg_parse_g2o(const char *path) {
    G2oGraph *g = new G2oGraph();
    std::string err;
    if (!parse_g2o(path, *g, err)) { g_err = err; delete g; return nullptr; }
    return g;
}
void pg_graph_free(void *g) { delete static_cast<G2oGraph *>(g); }
void pg_graph_sizes(void *gp, int64_t *nv, int64_t *ne, int64_t *len, int64_t *nval, int64_t *nmeas, int64_t *ninfo) {
    G2oGraph *g = static_cast<G2oGraph *>(gp);
    *nv = (int64_t)g->vertex_id.size(); *ne = (int64_t)g->edge_kind.size(); *len = g->len;
    *nval = (int64_t)g->vertex_values.size(); *nmeas = (int64_t)g->edge_meas.size(); *ninfo = (int64_t)g->edge_info_upper.size();
}
void pg_graph_fill(void *gp, uint32_t *vid, uint8_t *vkind, double *vval, uint8_t *ekind, uint32_t *efrom, uint32_t *eto,
                   double *emeas, double *einfo) {
    G2oGraph *g = static_cast<G2oGraph *>(gp);
    auto cp = [](auto *dst, const auto &v) { if (!v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    cp(vid, g->vertex_id); cp(vkind, g->vertex_kind); cp(vval, g->vertex_values); cp(ekind, g->edge_kind);
    cp(efrom, g->edge_from); cp(eto, g->edge_to); cp(emeas, g->edge_meas); cp(einfo, g->edge_info_upper);
}
int pg_write_g2o(const char *path, int64_t nv, const uint32_t *vid, const uint8_t *vkind, const double *vval, int64_t n_values,
                 int64_t ne, const uint8_t *ekind, const uint32_t *efrom, const uint32_t *eto,
                 const double *emeas, int64_t n_meas, const double *einfo, int64_t n_info) {
    G2oGraph g;
    g.vertex_id.assign(vid, vid + nv); g.vertex_kind.assign(vkind, vkind + nv); g.vertex_values.assign(vval, vval + n_values);
    g.edge_kind.assign(ekind, ekind + ne); g.edge_from.assign(efrom, efrom + ne); g.edge_to.assign(eto, eto + ne);
    g.edge_meas.assign(emeas, emeas + n_meas); g.edge_info_upper.assign(einfo, einfo + n_info);
    std::string err;
    if (!write_g2o(path, g, err)) { g_err = err; return -1; }
    return 0;
}

} // extern "C"
