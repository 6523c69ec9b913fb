"""B200 probe: fine-level SpMV time with fp64 and fp32 block storage, the GN step phases, and the time of one coarse solve per
AMG level.  (The r01l sweep over entries-per-iteration / register caps that picked the kernel's shape is in
profiles/r01l_spmv_sweep.log; those variants are no longer compiled.)"""
import argparse, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from rustrobotics_b200 import Options, PoseGraph  # noqa: E402
from rustrobotics_b200.synthetic import manhattan_se2, sphere_se3  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--poses", type=int, default=1_000_000)
ap.add_argument("--se3", action="store_true")
ap.add_argument("--opts", default="")     # "k=v,k=v;k=v" option sets for the step timing
a = ap.parse_args()
g = sphere_se3(max(2, a.poses // 500), 500) if a.se3 else manhattan_se2(a.poses)
D = 6 if a.se3 else 3

for f32 in (0, 1):
    os.environ.pop("PGO_TIME_SPMV_F32", None)
    if f32:
        os.environ["PGO_TIME_SPMV_F32"] = "1"
    pg = PoseGraph(graph=g, options=Options(pcg_rtol=1e-8))
    pg.gn_step()      # assembles H, builds the hierarchy (fp32 copies included)
    ms = min(pg.time_spmv(50) for _ in range(3))
    st = pg.stats()
    b = ((4 if f32 else 8) * D * D + 4) * st["offdiag_blocks"] + 8 * D * D * st["block_rows"] + (4 + 16 * D) * st["block_rows"]
    print(f"{'fp32' if f32 else 'fp64'} blocks: {ms*1e3:.1f} us  {b/ms/1e6:.0f} GB/s (bytes {b/1e6:.0f} MB)", flush=True)
    pg.close()
os.environ.pop("PGO_TIME_SPMV_F32", None)
for cfg in (a.opts.split(";") if a.opts else [""]):
    kw = {k: (float(v) if "." in v or "e" in v else int(v)) for k, v in (kv.split("=") for kv in cfg.split(",") if kv)}
    kw.setdefault("pcg_rtol", 1e-8)
    pg = PoseGraph(graph=g, options=Options(**kw))
    pg.snapshot_poses()
    for i in range(3):
        pg.restore_poses(); r = pg.gn_step(); t = pg.timings()
    print(f"step [{cfg}]: {r} " + str({k: round(v[0], 3) for k, v in t.items()}), flush=True)
    nl = len(pg.level_sizes()[0])
    print("   levels", pg.level_sizes()[0], "coarse solve per call (us): " + ", ".join(f"level {l}: {pg.time_coarse(l, 50)*1e3:.1f}" for l in range(1, nl)), flush=True)
    pg.close()
