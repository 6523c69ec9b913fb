"""Accuracy of the first Gauss-Newton step on BASELINE configs[3] (1M poses) against the golden fixture
(tests/golden/manhattan_1m_step1.npz, the refined true solution) as a function of pcg_rtol -- the measurement behind the
benchmark's default tolerance.   python tools/rtol_sweep.py [rtol ...] [--opts k=v,...] [--gpus N]"""
import hashlib
import json
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
from rustrobotics_b200 import Options, PoseGraph
from rustrobotics_b200.synthetic import manhattan_se2

args = [a for a in sys.argv[1:] if not a.startswith("--")]
flags = {a.split("=", 1)[0]: (a.split("=", 1)[1] if "=" in a else "1") for a in sys.argv[1:] if a.startswith("--")}
rtols = [float(a) for a in args] or [1e-8, 1e-9, 1e-10, 1e-11, 1e-12]
extra = {k: (float(v) if "." in v or "e" in v else int(v)) for k, v in (kv.split("=") for kv in flags.get("--opts", "").split(",") if kv)}
ngpu = int(flags.get("--gpus", "1"))
gold = np.load(ROOT / "tests" / "golden" / "manhattan_1m_step1.npz")
g = manhattan_se2(int(gold["n_poses"]))
h = hashlib.sha256()
for k in ("vertex_id", "vertex_kind", "vertex_values", "edge_kind", "edge_from", "edge_to", "edge_meas", "edge_info_upper"):
    h.update(np.ascontiguousarray(g[k]).tobytes())
assert h.hexdigest() == str(gold["graph_sha256"]), "generator produced a different graph"
s = gold["sample"]
full = ROOT / "tests" / "golden" / "_manhattan_1m_step1_dx_full.npy"
dx_full = np.load(full).reshape(-1, 3) if full.exists() else None
print(f"truth: chi2_1 {float(gold['chi2_1']):.6f} |dx| {float(gold['norm_dx']):.6f}  fp64 noise floor xy {float(gold['fp64_noise_xy']):.2e} theta {float(gold['fp64_noise_theta']):.2e}", flush=True)
# the solution of the system THIS build assembles (pgo_options.refine = 1: exact to ~1e-8 m): separates the solver's own error from
# the distance between two fp64 assemblies of the step
own = None
if "--own" in flags and ngpu == 1:
    pg = PoseGraph(graph=g, options=Options(pcg_rtol=1e-9, refine=1, **extra))
    own = pg.linearize_and_solve()[0].reshape(-1, 3)
    pg.close()
    e = np.abs(own[s] - gold["dx_sample"])
    print(f"own exact solution vs golden: xy {e[:, :2].max():.2e} theta {e[:, 2].max():.2e}", flush=True)
for rtol in rtols:
    kw = dict(pcg_rtol=rtol, **extra)
    if ngpu > 1:
        import torch
        nd = torch.cuda.device_count()
        kw.update(n_gpus=ngpu, device_ids=[k % nd for k in range(ngpu)])
    pg = PoseGraph(graph=g, options=Options(**kw))
    pg.snapshot_poses()
    pg.gn_step(allow_not_converged=False)          # warm-up (graph capture, omega)
    pg.restore_poses()
    t = time.perf_counter()
    nd_, c2, it = pg.gn_step(allow_not_converged=False)
    wall = time.perf_counter() - t
    tm = pg.timings()
    dx = pg.dx().reshape(-1, 3)
    e = np.abs(dx[s] - gold["dx_sample"])
    row = dict(rtol=rtol, its=it, step_ms=sum(v[0] for k, v in tm.items() if k != "spmv_fine"), pcg_ms=tm["pcg"][0], wall_ms=wall * 1e3,
               chi2_rel=abs(c2 - float(gold["chi2_1"])) / float(gold["chi2_1"]), norm_dx_abs=abs(nd_ - float(gold["norm_dx"])),
               max_xy=float(e[:, :2].max()), max_theta=float(e[:, 2].max()), rms=float(np.sqrt((e ** 2).mean())))
    if own is not None:
        eo = np.abs(dx - own)
        row.update(solver_err_xy=float(eo[:, :2].max()), solver_err_theta=float(eo[:, 2].max()))
    if dx_full is not None:
        ef = np.abs(dx - dx_full)
        row.update(full_max_xy=float(ef[:, :2].max()), full_max_theta=float(ef[:, 2].max()))
    print(json.dumps(row), flush=True)
    pg.close()
