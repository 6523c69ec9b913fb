"""Wall time of PoseGraph construction (pgo_create: host symbolic pass + uploads) at BASELINE configs[3]; PGO_SYM_TIMING=1 prints the phases.
usage: python tools/time_create.py [--poses N] [--host-only] [--repeats R]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from rustrobotics_b200 import Options, PoseGraph  # noqa: E402
from rustrobotics_b200.synthetic import manhattan_se2, sphere_se3  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--poses", type=int, default=1000000)
ap.add_argument("--se3", action="store_true")
ap.add_argument("--host-only", action="store_true", help="structure-only handle (device = -2): the symbolic pass alone, no GPU needed")
ap.add_argument("--repeats", type=int, default=3)
a = ap.parse_args()
g = sphere_se3(a.poses) if a.se3 else manhattan_se2(a.poses)
for i in range(a.repeats):
    t = time.perf_counter()
    pg = PoseGraph(graph=g, options=Options(device=-2) if a.host_only else Options())
    dt = time.perf_counter() - t
    print("create %d: %.3f s" % (i, dt), file=sys.stderr, flush=True)
    pg.close()
