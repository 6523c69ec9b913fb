// Example driver, the counterpart of the reference's examples/mapping/pose_graph_optimization.rs:
//     PoseGraph::new(filename, solver)?.optimize(50, true, plot)?            (:49-50)
// The reference picks the file / solver / plot flag from an interactive dialoguer menu (:10-47); here they are
// command-line arguments so that the example can run unattended on a GPU box:
//     pose_graph_optimization <file.g2o> [GaussNewton|LevenbergMarquardt] [plot]
// Build: make -C examples   (links the in-tree libpgo_b200.so)
#include <cstdio>
#include <cstring>
#include <sys/stat.h>

#include "../rustrobotics_b200/csrc/host/pose_graph.hpp"

using namespace robotics::mapping;

int main(int argc, char **argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <file.g2o> [GaussNewton|LevenbergMarquardt] [plot]\n", argv[0]);
        return 2;
    }
    ::mkdir("./img", 0777);                                                // std::fs::create_dir_all("./img") (:8)
    const PoseGraphSolver solver = (argc > 2 && !std::strcmp(argv[2], "LevenbergMarquardt")) ? PoseGraphSolver::LevenbergMarquardt
                                                                                             : PoseGraphSolver::GaussNewton;
    const bool plot = argc > 3 && !std::strcmp(argv[3], "plot");
    try {
        PoseGraph graph(argv[1], solver);
        std::vector<double> errors = graph.optimize(50, /*log=*/true, plot);
        std::printf("final error %.6f after %zu iteration(s)\n", errors.back(), errors.size() - 1);
    } catch (const Error &e) {                                             // Err(Box<dyn Error>)
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
