"""PCG iteration counts per GN step of the single-GPU path on the sharded-parity cases (to compare with shard_worker's)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import graph_of, load_golden  # noqa: E402
from rustrobotics_b200 import Options, PoseGraph  # noqa: E402
from rustrobotics_b200.synthetic import manhattan_se2, sphere_se3  # noqa: E402
for case in ["simulation-pose-pose", "intel", "dlr", "manhattan10000", "manhattan100000", "sphere2500", "sphere40x50", "sphere200x200"]:
    if case.startswith("manhattan"):
        g = manhattan_se2(int(case[len("manhattan"):]))
    elif case.startswith("sphere") and "x" in case:
        g = sphere_se3(*[int(t) for t in case[len("sphere"):].split("x")])
    else:
        g = graph_of(load_golden(case))
    big = len(g["vertex_id"]) > 20000
    pg = PoseGraph(graph=g, options=Options())
    errs = pg.optimize(3 if big else 8)
    print(f"single: {case} levels {pg.level_sizes()[0]} pcg={pg.pcg_iterations} chi2={errs[-1]:.6f}", flush=True)
    pg.close()
