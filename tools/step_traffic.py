"""Run under ncu (see tools/gpu_session.sh): ONE converged Gauss-Newton step on BASELINE configs[3] inside a cudaProfilerStart/Stop
window, after an unprofiled warm-up step -- the launch list with dram__bytes_{read,write}.sum per kernel gives the measured HBM
traffic of a whole step (bench.py: step_roofline.traffic).   python tools/step_traffic.py [--poses P] [--rtol R] [--se3]"""
import argparse
import ctypes
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from rustrobotics_b200 import Options, PoseGraph  # noqa: E402
from rustrobotics_b200.synthetic import manhattan_se2, sphere_se3  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--poses", type=int, default=1_000_000)
ap.add_argument("--rtol", type=float, default=None)
ap.add_argument("--se3", action="store_true")
a = ap.parse_args()
if a.rtol is None:
    import bench
    a.rtol = bench.DEFAULT_PCG_RTOL
g = sphere_se3(max(2, a.poses // 500), 500) if a.se3 else manhattan_se2(a.poses)
pg = PoseGraph(graph=g, options=Options(pcg_rtol=a.rtol))
rt = ctypes.CDLL("libcudart.so.12")
pg.snapshot_poses()
print("warm-up", pg.gn_step(), flush=True)
pg.restore_poses()
rt.cudaProfilerStart()
r = pg.gn_step()
rt.cudaProfilerStop()
print("profiled step", r, pg.timings(), flush=True)
