// C++ mirror of the reference's public API (src/mapping/mod.rs:6):
//   robotics::mapping::{PoseGraph, PoseGraphSolver}
// same names, argument meaning and error behaviour as pose_graph_optimization.rs:214-432, with
// the Gauss-Newton loop body delegated to the CUDA library through the C ABI (pgo_b200.h).
// Rust's Result<_, Box<dyn Error>> becomes a thrown robotics::mapping::Error.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/pgo_b200.h"
#include "g2o.hpp"

namespace robotics { namespace mapping {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

enum class PoseGraphSolver { GaussNewton = 0, LevenbergMarquardt = 1 };   // :28-32

struct Pose { uint32_t id; uint8_t kind; double x, y, theta; };             // theta = 0 for landmarks

class PoseGraph {
public:
    // PoseGraph::new(file_path, solver) (:215-227)
    PoseGraph(const std::string &file_path, PoseGraphSolver solver, const pgo_options *options = nullptr);
    // same, from an already loaded graph (synthetic benchmarks)
    PoseGraph(const G2oGraph &graph, const std::string &name, PoseGraphSolver solver, const pgo_options *options = nullptr);
    PoseGraph(G2oGraph &&graph, const std::string &name, PoseGraphSolver solver, const pgo_options *options = nullptr);
    // adopts a handle that pgo_create has already built from these very arrays (pg_from_arrays copies the caller's arrays into
    // `graph` while the handle is being created)
    PoseGraph(G2oGraph &&graph, const std::string &name, PoseGraphSolver solver, pgo_handle *adopted);
    ~PoseGraph();
    PoseGraph(const PoseGraph &) = delete;
    PoseGraph &operator=(const PoseGraph &) = delete;

    // optimize(num_iterations, log, plot) -> chi2 history [e0, e1, ...] (:247-303)
    std::vector<double> optimize(size_t num_iterations, bool log, bool plot);
    // plot() (:375-431): writes img/{name}-{iteration}-{solver}.svg
    void plot() const;
    // the reference has no accessor for `nodes` (private, :157); the north star asks for the poses
    std::vector<Pose> poses() const;
    double global_error() const;                       // global_error (:537-574)
    std::vector<double> linearize_and_solve();         // linearize_and_solve (:371-373)

    size_t num_nodes() const { return graph_.vertex_id.size(); }
    size_t num_edges() const { return graph_.edge_kind.size(); }
    size_t len() const { return (size_t)graph_.len; }
    const std::vector<double> &norms() const { return norms_; }          // |dx| per iteration (:255, :285)
    const std::vector<int32_t> &pcg_iterations() const { return pcg_iters_; }
    pgo_handle *handle() const { return h_; }
    const G2oGraph &graph() const { return graph_; }

private:
    void init(const pgo_options *options);
    void check(int rc, const char *what) const;
    G2oGraph graph_;
    std::string name_;
    PoseGraphSolver solver_;
    size_t iteration_ = 0;                             // accumulates across optimize calls (:160, :270)
    pgo_handle *h_ = nullptr;
    std::vector<double> norms_;
    std::vector<int32_t> pcg_iters_;
};

}} // namespace robotics::mapping
