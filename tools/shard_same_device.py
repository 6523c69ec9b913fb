"""Diagnostic: the sharded path with BOTH ranks on cuda:0 (peer memory over CUDA IPC inside one device, gloo for the handle
exchange), so that sharded numerics can be examined on a 1-GPU box.  Slow (the ranks time-slice one GPU and the stage barriers
spin), only for iteration-count experiments:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_same_device.py [poses] [k=v,...]
"""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
from rustrobotics_b200 import Options, PoseGraph
from rustrobotics_b200.synthetic import manhattan_se2

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(0)
dist.init_process_group("gloo")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
kw = {k: (float(v) if "." in v or "e" in v else int(v)) for k, v in (kv.split("=") for kv in (sys.argv[2] if len(sys.argv) > 2 else "").split(",") if kv)}
g = manhattan_se2(n)
pg = PoseGraph(graph=g, options=Options(device=0, world=world, rank=rank, pcg_rtol=1e-8, **kw))
errs = pg.optimize(2)
if rank == 0:
    print(f"same-device world {world} poses {n} opts {kw} env REPL={os.environ.get('PGO_REPL_MAX_ROWS')}: levels {pg.level_sizes()[0]} pcg {pg.pcg_iterations} chi2 {errs[-1]:.6f}", flush=True)
pg.close()
dist.barrier()
dist.destroy_process_group()
