// Counterpart of the reference's criterion bench benches/graph_slam.rs:6-13 (`graph_slam_intel`): time
//     PoseGraph::new("dataset/g2o/intel.g2o", GaussNewton)?.optimize(10, false, false)
// end to end (parse + symbolic pass + upload + 10 Gauss-Newton iterations), `iters` times after one warm-up.
//     graph_slam <intel.g2o> [iters]
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../rustrobotics_b200/csrc/host/pose_graph.hpp"

using namespace robotics::mapping;

int main(int argc, char **argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s <file.g2o> [iters]\n", argv[0]); return 2; }
    const int iters = argc > 2 ? std::atoi(argv[2]) : 10;
    try {
        double best = 1e300, sum = 0.0, last = 0.0;
        for (int i = -1; i < iters; i++) {
            const auto t0 = std::chrono::steady_clock::now();
            PoseGraph g(argv[1], PoseGraphSolver::GaussNewton);
            last = g.optimize(10, false, false).back();
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (i < 0) continue;                                           // warm-up (CUDA context, first-touch)
            best = ms < best ? ms : best; sum += ms;
        }
        std::printf("{\"bench\": \"graph_slam_intel\", \"iters\": %d, \"mean_ms\": %.3f, \"best_ms\": %.3f, \"final_chi2\": %.6f}\n",
                    iters, sum / iters, best, last);
    } catch (const Error &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
