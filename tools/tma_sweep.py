"""B200 probe: the TMA-staged sliced SpMV (PGO_SPMV_TMA64 / PGO_SPMV_TMA32 = ring depth, 0 = register-staged kernel)."""
import argparse, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from rustrobotics_b200 import Options, PoseGraph  # noqa: E402
from rustrobotics_b200.synthetic import manhattan_se2, sphere_se3  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--poses", type=int, default=1_000_000)
ap.add_argument("--se3", action="store_true")
ap.add_argument("--cfgs", default="0:0,4:4")     # ns64:ns32 pairs
a = ap.parse_args()
g = sphere_se3(max(2, a.poses // 500), 500) if a.se3 else manhattan_se2(a.poses)
D = 6 if a.se3 else 3
for cfg in a.cfgs.split(","):
    n64, n32 = cfg.split(":")
    os.environ["PGO_SPMV_TMA64"] = n64; os.environ["PGO_SPMV_TMA32"] = n32
    out = []
    for f32 in (0, 1):
        os.environ.pop("PGO_TIME_SPMV_F32", None)
        if f32:
            os.environ["PGO_TIME_SPMV_F32"] = "1"
        pg = PoseGraph(graph=g, options=Options(pcg_rtol=1e-8))
        pg.snapshot_poses()
        for i in range(2):
            pg.restore_poses(); r = pg.gn_step(); t = pg.timings()
        ms = min(pg.time_spmv(50) for _ in range(3))
        st = pg.stats()
        b = ((4 if f32 else 8) * D * D + 4) * st["offdiag_blocks"] + 8 * D * D * st["block_rows"] + (4 + 16 * D) * st["block_rows"]
        out.append(f"{'fp32' if f32 else 'fp64'} {ms*1e3:.1f} us {b/ms/1e6:.0f} GB/s")
        pg.close()
    print(f"TMA64={n64} TMA32={n32}: " + " | ".join(out) + f" | step {r} pcg {t['pcg'][0]:.2f} ms", flush=True)
