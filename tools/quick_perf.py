"""Quick perf probe on the B200 box: SpMV roofline + GN step phases on the BASELINE configs[3] graph."""
import argparse, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from rustrobotics_b200 import Options, PoseGraph  # noqa: E402
from rustrobotics_b200.synthetic import manhattan_se2, sphere_se3  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--poses", type=int, default=1_000_000)
ap.add_argument("--se3", action="store_true")      # BASELINE configs[4]: sphere, --poses = levels * 500
ap.add_argument("--opts", default="")     # e.g. "amg_kcycle=1,amg_aggregate_size=8;amg_dense_max=256"  (';' separates configurations)
a = ap.parse_args()
g = sphere_se3(max(2, a.poses // 500), 500) if a.se3 else manhattan_se2(a.poses)
D = 6 if a.se3 else 3
for cfg in (a.opts.split(";") if a.opts else [""]):
    kw = {k: (float(v) if "." in v or "e" in v else int(v)) for k, v in (kv.split("=") for kv in cfg.split(",") if kv)}
    kw.setdefault("pcg_rtol", 1e-8)
    pg = PoseGraph(graph=g, options=Options(**kw))
    st = pg.stats()
    pg.snapshot_poses()
    res = []
    for i in range(3):
        pg.restore_poses()
        r = pg.gn_step()
        t = pg.timings()
        res.append((r, {k: round(v[0], 3) for k, v in t.items()}, sum(v[1] for v in t.values())))
    ms = pg.time_spmv(50)
    nb = st["block_rows"] + st["offdiag_blocks"]
    gbs = ((8 * D * D + 4) * nb + (4 + 16 * D) * st["block_rows"]) / ms / 1e6
    print(f"cfg[{cfg}] levels {pg.level_sizes()[0]} spmv {ms*1e3:.1f} us {gbs:.0f} GB/s | step {res[-1]}", flush=True)
    pg.close()
