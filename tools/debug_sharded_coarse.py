"""debug: sharded coarse levels (PGO_REPL_MAX_ROWS small) with N shards on the available GPUs"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
from rustrobotics_b200 import Options, PoseGraph, PgoError
from rustrobotics_b200.synthetic import manhattan_se2
n = int(sys.argv[1]); poses = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[3:])}
nd = torch.cuda.device_count()
g = manhattan_se2(poses)
t = time.time()
try:
    pg = PoseGraph(graph=g, options=Options(device_ids=[k % nd for k in range(n)], **kw))
    print("levels", pg.level_sizes()[0], flush=True)
    r = pg.optimize(2)
    print(f"OK n={n} poses={poses} {kw} env REPL={os.environ.get('PGO_REPL_MAX_ROWS')} WHILE={os.environ.get('PGO_WHILE')} PDL={os.environ.get('PGO_PDL')}: {r} pcg {pg.pcg_iterations} {time.time()-t:.1f}s", flush=True)
except PgoError as e:
    print(f"FAIL n={n} poses={poses} {kw} env REPL={os.environ.get('PGO_REPL_MAX_ROWS')} WHILE={os.environ.get('PGO_WHILE')} PDL={os.environ.get('PGO_PDL')}: {e} {time.time()-t:.1f}s", flush=True)
