"""Run under compute-sanitizer (tools/gpu_session.sh): small graphs through every kernel of the hot path, single GPU and the
single-process multi-GPU handle with two shards on one device.   python tools/sanitize.py [n_gpus]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from conftest import graph_of, load_golden
from rustrobotics_b200 import Options, PoseGraph
from rustrobotics_b200.synthetic import manhattan_se2, sphere_se3

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
quick = len(sys.argv) > 2 and sys.argv[2] == "quick"      # the AMG path only, on the two synthetic graphs (dense inverse of 180^2 and 2400^2)
cases = (("intel", graph_of(load_golden("intel"))), ("simulation-pose-landmark", graph_of(load_golden("simulation-pose-landmark"))),
         ("manhattan1000", manhattan_se2(1000)), ("sphere8x50", sphere_se3(8, 50)))
for name, g in (cases[2:] if quick else cases):
    for pre in ((1,) if quick else (1, 0)):
        kw = dict(preconditioner=pre, pcg_max_iterations=400)
        if n > 1:
            import torch
            nd = torch.cuda.device_count()
            kw.update(device_ids=[k % nd for k in range(n)])      # distinct GPUs when the box has them, else shards share GPU 0
        pg = PoseGraph(graph=g, options=Options(**kw))
        errs = pg.optimize(2)
        v = pg.poses()
        print(f"sanitize {name} n_gpus={n} precond={pre}: chi2 {errs} pcg {pg.pcg_iterations} finite {bool(np.isfinite(v).all())}", flush=True)
        pg.close()
