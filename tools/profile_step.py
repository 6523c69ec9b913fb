"""Short profiling driver (run under ncu on the B200 box): the BASELINE configs[3] graph (or --poses P), one
Gauss-Newton step with the PCG capped at --pcg-iters iterations so that the whole launch list stays short."""
import argparse
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from rustrobotics_b200 import Options, PoseGraph  # noqa: E402
from rustrobotics_b200.synthetic import manhattan_se2, sphere_se3  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--poses", type=int, default=1_000_000)
ap.add_argument("--pcg-iters", type=int, default=16)
ap.add_argument("--preconditioner", type=int, default=1)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--se3", action="store_true")
a = ap.parse_args()
g = sphere_se3(max(2, a.poses // 500), 500) if a.se3 else manhattan_se2(a.poses)
pg = PoseGraph(graph=g, options=Options(pcg_max_iterations=a.pcg_iters, preconditioner=a.preconditioner))
print("chi2", pg.global_error())
for _ in range(a.steps):
    print(pg.gn_step(), pg.timings())
