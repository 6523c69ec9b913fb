// Example driver, the counterpart of the reference's examples/mapping/pose_graph_optimization.rs:
//     PoseGraph::new(filename, solver)?.optimize(50, true, plot)?            (:49-50)
// The reference picks the file / solver / plot flag from an interactive dialoguer menu (:10-47); here they are
// command-line arguments so that the example can run unattended on a GPU box:
//     pose_graph_optimization <file.g2o> [GaussNewton|LevenbergMarquardt] [plot|noplot] [--devices 0,1,2,3]
// --devices: the ONE PoseGraph below drives a shard on each listed GPU (pgo_options.n_gpus / device_ids); listing a GPU twice
// makes two shards share it.  Build: make -C examples   (links the in-tree libpgo_b200.so)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/stat.h>
#include <vector>

#include "../rustrobotics_b200/csrc/host/pose_graph.hpp"

using namespace robotics::mapping;

int main(int argc, char **argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <file.g2o> [GaussNewton|LevenbergMarquardt] [plot|noplot] [--devices 0,1,...]\n", argv[0]);
        return 2;
    }
    ::mkdir("./img", 0777);                                                // std::fs::create_dir_all("./img") (:8)
    const PoseGraphSolver solver = (argc > 2 && !std::strcmp(argv[2], "LevenbergMarquardt")) ? PoseGraphSolver::LevenbergMarquardt
                                                                                             : PoseGraphSolver::GaussNewton;
    const bool plot = argc > 3 && !std::strcmp(argv[3], "plot");
    pgo_options opt;
    pgo_default_options(&opt);
    std::vector<int32_t> devices;
    for (int a = 2; a + 1 < argc; a++)
        if (!std::strcmp(argv[a], "--devices"))
            for (const char *p = argv[a + 1]; *p; ) { devices.push_back((int32_t)std::strtol(p, const_cast<char **>(&p), 10)); if (*p == ',') p++; }
    if (devices.size() > 1) { opt.n_gpus = (int32_t)devices.size(); opt.device_ids = devices.data(); }
    try {
        PoseGraph graph(argv[1], solver, &opt);
        std::vector<double> errors = graph.optimize(50, /*log=*/true, plot);
        std::printf("final error %.6f after %zu iteration(s)\n", errors.back(), errors.size() - 1);
    } catch (const Error &e) {                                             // Err(Box<dyn Error>)
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
