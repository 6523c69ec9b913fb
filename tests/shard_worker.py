"""Sharded-mode parity worker: run under torchrun on a multi-GPU B200 box (one process per GPU),

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/shard_worker.py [cases...]

Every rank passes the whole graph to pgo_create with (world, rank) and owns a contiguous vertex range;
results must match the CPU oracle exactly as the single-GPU path does (same tolerances as test_gpu_parity.py).
Exits non-zero on the first mismatch; rank 0 prints one line per case."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

CHI2_RTOL, POSE_ATOL = 1e-6, 1e-6


def main():
    import torch
    import torch.distributed as dist
    from bench import TorchComm
    from conftest import graph_of, load_golden
    from oracle.oracle import OraclePoseGraph
    from rustrobotics_b200 import Options, PoseGraph
    from rustrobotics_b200.synthetic import manhattan_se2, sphere_se3
    from test_gpu_parity import _pose_diff
    from test_gpu_se3 import se3_pose_diff

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = sys.argv[1:] or ["simulation-pose-pose", "intel", "dlr", "manhattan10000", "manhattan100000", "sphere2500", "sphere40x50",
                             "sphere200x200"]
    for case in cases:
        precond = 1
        if case.endswith(":bj"):
            case, precond = case[:-3], 0
        if case.startswith("manhattan"):
            g = manhattan_se2(int(case[len("manhattan"):]))
        elif case.startswith("sphere") and "x" in case:                    # SE3 sphere, levels x poses per level
            g = sphere_se3(*[int(t) for t in case[len("sphere"):].split("x")])
        else:
            g = graph_of(load_golden(case))
        se3 = int(g["vertex_kind"][0]) == 2
        big = len(g["vertex_id"]) > 20000
        o = OraclePoseGraph.from_arrays(**g)
        c_o = o.global_error()
        sls = o.build_linear_system() if not big else None
        its = 3 if big else 8
        errs_o = o.optimize(its)
        _, _, _, vo = o.vertices()
        dist.barrier()
        pg = PoseGraph(graph=g, options=Options(device=local, world=world, rank=rank, preconditioner=precond), comm=TorchComm())
        part = pg.partition()
        assert part["world"] == world and part["rank"] == rank
        c_g = pg.global_error()
        assert abs(c_g - c_o) <= 1e-12 * c_o, (case, c_g, c_o)
        if sls is not None:
            cp, ri, vals, b = pg.system()
            assert np.array_equal(cp, sls.col_ptr) and np.array_equal(ri, sls.row_idx), case
            assert np.abs(vals - sls.vals).max() <= 1e-12 * np.abs(sls.vals).max(), case
            assert np.abs(b - sls.b).max() <= 1e-11 * max(np.abs(sls.b).max(), 1.0), case
        errs_g = pg.optimize(its)
        assert len(errs_g) == len(errs_o), (case, errs_g, errs_o)
        np.testing.assert_allclose(errs_g, errs_o, rtol=CHI2_RTOL, err_msg=case)
        dxy, dth = se3_pose_diff(pg.poses(), vo) if se3 else _pose_diff(g, pg.poses(), vo)
        assert dxy < POSE_ATOL and dth < POSE_ATOL, (case, dxy, dth)
        if rank == 0:
            print(f"shard ok: {case} precond={'amg' if precond else 'bj'} world={world} ranges={part['vertex_range']} "
                  f"remote_blocks={part['remote_blocks']} pcg={pg.pcg_iterations} chi2={errs_g[-1]:.6f}", flush=True)
        pg.close()
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
