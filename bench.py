#!/usr/bin/env python
"""bench.py -- headline benchmark of the pose-graph-optimization hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--poses P]

A "step" is ONE Gauss-Newton iteration (linearise + assemble + PCG solve + retract + chi2,
reference pose_graph_optimization.rs:271-274) on BASELINE.json configs[3]: the synthetic
Manhattan-world SE(2) graph with 1M poses / 4M edges (seed 42).  Every timed step starts from
the same initial guess (device-side pose snapshot restored before the step), so all K steps
do identical work.  The graph's working set (~2.4 GB) is far larger than the 126 MB L2, so no
L2 flush is needed between steps.

Prints ONE JSON line (rank 0).  value = edges processed per second by the whole job
(|E| x GN iterations / s) with the graph resident in HBM, timed with CUDA events on the
library's stream; e2e = the same metric through the public PoseGraph API with HOST buffers
(pinned), i.e. set_poses (H2D) + gn_step + poses() (D2H) inside the timed region.

--impl reference times the CPU restatement of the reference's own path (oracle/: sequential COO
assembly in the reference's put order + COO->CSC + sparse LU + retract + chi2) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "gn_edges_per_sec"
UNIT = "edges/s"
DEFAULT_PCG_RTOL = 1e-9             # the tolerance the golden parity test of the headline config runs at (tests/test_gpu_parity.py); see profiles/r03a_rtol_sweep.log
# CPU arm (oracle port: reference-order assembly + SciPy SuperLU).  SuperLU (32-bit indices) cannot factorise the 1M-pose system of
# configs[3] in this image ("Not enough memory to perform factorization", tests/golden/make_golden_1m.py), so the CPU arm runs the
# largest sample that fits its time budget and says so; the direct solve grows faster than linearly with the graph (measured in the
# authoring container, 1 core: 2.4 s / 16.7 s / 33 s per GN iteration at 100k / 300k / 500k poses), so a smaller sample FLATTERS the CPU.
CPU_SAMPLE_POSES = 300_000          # cpu_baseline leg of the default run: one GN iteration, ~20 s
CPU_REF_SAMPLE_POSES = 500_000      # --impl reference: ~35 s per GN iteration, at most 1 warm-up + 2 timed steps
CPU_SAMPLE_POSES_SE3 = 25_000       # sphere SE3 sample (50 levels x 500)
CPU_EXTRAPOLATION = ("same_config false: the oracle's SuperLU cannot factorise the 1M-pose system here; measured 2.4 / 16.7 / 33 s per GN "
                     "iteration at 100k / 300k / 500k poses (1 core, authoring container; the GPU box's host is ~1.7x faster) => about 100 s at 1M poses (SURVEY App. C probe: 105 s), i.e. ~4e4 edges/s")


BUNDLED = {   # the reference's bundled datasets (dataset/g2o/*.g2o), shipped as parsed arrays in tests/golden/*.npz (make_golden.py)
    "pose-pose": "simulation-pose-pose", "pose-landmark": "simulation-pose-landmark", "intel": "intel", "dlr": "dlr",
    "m3500": "input_M3500_g2o", "sphere2500": "sphere2500", "garage": "parking-garage",
}
GRAPH_KEYS = ("vertex_id", "vertex_kind", "vertex_values", "edge_kind", "edge_from", "edge_to", "edge_meas", "edge_info_upper")


def make_graph(workload: str, n_poses: int):
    """(graph arrays, block dimension, description) of a BASELINE.json workload"""
    from rustrobotics_b200.synthetic import manhattan_se2, sphere_se3
    if workload in BUNDLED:
        import numpy as np
        z = np.load(ROOT / "tests" / "golden" / f"{BUNDLED[workload]}.npz")
        g = {k: z[k] for k in GRAPH_KEYS}
        D = 6 if int(g["vertex_kind"][0]) == 2 else 3
        return g, D, f"bundled dataset/g2o/{BUNDLED[workload]}.g2o of the reference (BASELINE configs[0..1])" + ", {} vertices / {} edges"
    if workload == "sphere":
        g = sphere_se3(max(2, n_poses // 500), 500)
        return g, 6, "synthetic sphere SE3 pose graph (6x6 blocks; repo-defined SE3 semantics, parity unpinned), {} poses / {} edges, seed 42 (BASELINE configs[4])"
    g = manhattan_se2(n_poses)
    return g, 3, "synthetic Manhattan-world SE2 pose graph, {} poses / {} edges, seed 42 (BASELINE configs[3])"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class TorchComm:
    """The launcher-side communicator of the process-per-GPU mode (torchrun): torch.distributed is only plumbing here -- one
    all-gather of an 80-byte blob (CUDA IPC handle + GPU UUID) per rank at start-up, and the merge of the ranks' pose spans for the getters.
    The product package imports no torch; its single-process multi-GPU mode (Options(n_gpus=N)) needs none of this."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)

    def all_gather_bytes(self, mine: bytes) -> bytes:
        parts = [None] * self.world
        self.dist.all_gather_object(parts, bytes(mine), group=self.group)
        return b"".join(parts)

    def all_reduce_sum(self, a):
        import torch
        t = torch.from_numpy(a)
        if self.dist.get_backend(self.group) == "nccl":
            t = t.cuda()
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def barrier(self):
        self.dist.barrier(self.group)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line).  The sampler is started before the warm-up
    (nvidia-smi needs a few hundred ms to come up), writes line-buffered time-stamped rows every 25 ms, and only the rows whose
    time stamp falls inside [begin(), end()] -- the timed region -- are reported."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = self.t1 = None

    def start(self):
        import shutil
        cmd = ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(self.device)]
        if shutil.which("stdbuf"):
            cmd = ["stdbuf", "-oL"] + cmd
        try:
            self.p = subprocess.Popen(cmd, stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def stop(self) -> dict:
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        parsed = []
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                parsed.append((ts, float(r[1]), float(r[2]), float(r[3]), [nm for nm, v in zip(names, r[5:9]) if v.strip().lower().startswith("active")]))
            except ValueError:
                continue
        if rows and not parsed:
            log(f"[clocks] could not parse nvidia-smi rows, first row: {rows[0]}")
        inside = [q for q in parsed if self.t0 is not None and self.t1 is not None and self.t0 <= q[0] <= self.t1]
        window = "timed region"
        if not inside and parsed:            # a timed region shorter than the sampling period: the nearest rows under load
            inside = [q for q in parsed if self.t0 is not None and q[0] >= self.t0 - 0.5] or parsed
            window = "timed region shorter than the 25 ms sampling period: rows from the warm-up + timed region"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(q[1] for q in inside)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(q[2] for q in inside), "power_w_max": max(q[3] for q in inside),
                "samples": len(sm), "reasons": sorted({nm for q in inside for nm in q[4]}), "window": window}


def measured_peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(name="spmv_traffic.json"):
    """DRAM bytes measured by ncu (dram__bytes_read.sum + dram__bytes_write.sum), from the capture committed under profiles/ this
    round (tools/gpu_session.sh writes them; a number measured under a profiler is never a bench value, only the traffic is used)"""
    p = ROOT / "profiles" / name
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            return None
    return None


def golden_check(pg_chi2_after, pg_norm_dx, n_poses, workload):
    """the step's result against the committed golden of the headline config (tests/golden/manhattan_1m_step1.npz)"""
    import numpy as np
    p = ROOT / "tests" / "golden" / "manhattan_1m_step1.npz"
    if workload != "manhattan" or n_poses != 1_000_000 or not p.exists():
        return None
    z = np.load(p)
    c, nd = float(z["chi2_1"]), float(z["norm_dx"])
    return {"chi2_after_step_golden": c, "chi2_rel_err": abs(pg_chi2_after - c) / c, "norm_dx_golden": nd, "norm_dx_abs_err": abs(pg_norm_dx - nd),
            "golden": "tests/golden/manhattan_1m_step1.npz (true solution of the reference's first GN system; tests/test_gpu_parity.py checks sampled poses)"}


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(n_poses: int, steps: int, warmup: int, workload: str = "manhattan"):
    """The reference's CPU path (oracle restatement) on a bounded sample: a graph of `n_poses` poses of the workload.
    Every step is one GN iteration from the same initial guess.  -> (edges/s, seconds per step, description, cores)"""
    import numpy as np
    from oracle.oracle import OraclePoseGraph
    g, _, _ = make_graph(workload, n_poses)
    ne = len(g["edge_from"])
    o = OraclePoseGraph.from_arrays(**g)
    s0 = o.state().copy()
    try:
        from threadpoolctl import threadpool_info
        blas_threads = max([t.get("num_threads", 1) for t in threadpool_info()] or [1])
    except Exception:
        blas_threads = 1
    if len(g["vertex_id"]) <= 20000:
        warmup = max(warmup, 2)          # small graphs: keep Python / SciPy first-call costs out of the timed steps
    ts = []
    for i in range(warmup + steps):
        o.set_state(s0)
        t = time.perf_counter()
        dx = o.build_linear_system(0.0).solve()
        o.update_nodes(dx)
        float(np.linalg.norm(dx))
        o.global_error()
        dt = time.perf_counter() - t
        if i >= warmup:
            ts.append(dt)
        log(f"[cpu] GN iteration {i}: {dt:.2f} s")
    sec = sum(ts) / len(ts)
    kind = f"bundled {BUNDLED[workload]}.g2o" if workload in BUNDLED else ("sphere SE3 (seed 42)" if workload == "sphere" else "Manhattan SE2 (seed 42)")
    sample = (f"{kind} {len(g['vertex_id'])} vertices / {ne} edges, {steps} GN iteration(s) from the initial guess; "
              f"oracle/ restatement: sequential COO assembly + COO->CSC + SciPy SuperLU (stand-in for UMFPACK) + retract + chi2; "
              f"assembly single-threaded like the reference, BLAS threads available to SuperLU: {blas_threads}")
    return ne / sec, sec, sample, 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = min(args.poses, CPU_SAMPLE_POSES_SE3 if args.workload == "sphere" else CPU_REF_SAMPLE_POSES)      # bundled graphs: the whole graph
    big = args.workload not in BUNDLED and n > 100_000
    if big:                              # ~35 s per GN iteration: the whole arm must end within a few minutes
        args.steps, args.warmup = min(args.steps, 2), min(args.warmup, 1)
    val, sec, sample, cores = cpu_reference_run(n, args.steps, args.warmup, args.workload)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "bundled dataset" if args.workload in BUNDLED else "synthetic", "gn_iterations_per_sec": 1.0 / sec,
        "config": {"workload": (f"bundled {BUNDLED[args.workload]}.g2o (whole graph); " if args.workload in BUNDLED else
                                f"synthetic sphere SE3, {args.poses} poses (BASELINE configs[4]); " if args.workload == "sphere" else
                                f"synthetic Manhattan SE2, {args.poses} poses / {4 * args.poses} edges (BASELINE configs[3]); ") +
                               "CPU arm runs the bounded sample below", "sample_poses": n,
                   "same_config": args.workload in BUNDLED or n == args.poses,
                   "extrapolation": None if (args.workload in BUNDLED or n == args.poses) else CPU_EXTRAPOLATION},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_cpus": os.cpu_count(),
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # `python bench.py --gpus N` WITHOUT torchrun: the single-process multi-GPU handle (Options(n_gpus=N)): one process, one PoseGraph,
    # N shards.  Under torchrun (what the driver does for N > 1) the same shards are one process each (Options(world, rank)).
    n_shards = args.gpus if (world == 1 and args.gpus > 1) else 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from rustrobotics_b200 import Options, PoseGraph, _build
    _build.build()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t0 = time.perf_counter()
    g, D, wl_desc = make_graph(args.workload, args.poses)
    n_poses, n_edges = len(g["vertex_id"]), len(g["edge_from"])
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    # world > 1: ONE graph sharded by contiguous vertex ranges, one process per GPU; halo rows are read from
    # peer HBM over NVLink inside the kernels, dot products reduced by a device-side peer-memory all-reduce
    extra = {k: (float(v) if "." in v or "e" in v else int(v)) for k, v in (kv.split("=") for kv in args.opts.split(",") if kv)}
    if n_shards > 1:
        nd = torch.cuda.device_count()
        extra = dict(extra, device_ids=[k % nd for k in range(n_shards)])       # fewer GPUs than shards: shards share devices (functional test only)
    pg = PoseGraph(graph=g, options=Options(device=local, world=world, rank=rank, pcg_rtol=args.pcg_rtol,
                                            preconditioner=args.preconditioner, **extra), comm=TorchComm() if world > 1 else None)
    t_create = time.perf_counter() - t0
    log(f"[rank {rank}] graph {n_poses} poses / {n_edges} edges generated in {t_gen:.1f}s, created in {t_create:.1f}s, {pg.stats()}")
    chi2_0 = pg.global_error()
    pg.snapshot_poses()

    def one_step():
        pg.restore_poses()
        r = pg.gn_step(allow_not_converged=False)
        t = pg.timings()
        return r, t

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler.begin()
    w0 = time.perf_counter()
    dev_ms, launches, pcg_its, phases = 0.0, 0, [], {}
    last = None
    for _ in range(args.steps):
        (nd, c2, it), t = one_step()
        last = (nd, c2)
        pcg_its.append(it)
        for k, (ms, ln) in t.items():
            if k == "spmv_fine":
                continue
            dev_ms += ms; launches += ln
            phases[k] = phases.get(k, 0.0) + ms / args.steps
    barrier()
    wall_s = time.perf_counter() - w0
    sampler.end()
    clocks = sampler.stop()

    # ---- dominant kernel: fine-level BSR SpMV, timed live with CUDA events on the library's stream
    barrier()
    spmv_ms = pg.time_spmv(50)
    barrier()
    st = pg.stats()
    nb = st["block_rows"] + st["offdiag_blocks"]          # blocks of H incl. diagonal
    # 8 D^2 B (values) + 4B (col) + 4N (row ptr) + 8 D N (x) + 8 D N (y)  [D = 3: 76B + 52N]
    spmv_bytes = (8 * D * D + 4) * nb + (4 + 16 * D) * st["block_rows"]
    if n_shards > 1:                                      # one handle, N shards: the handle reports totals, the kernel time is the slowest shard's
        spmv_bytes /= n_shards                            # (the partition balances stored blocks, so this is the mean = about the largest shard)
    peak, peak_src = measured_peak_hbm()
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    tr = ncu_traffic()

    # ---- e2e through the public API with host buffers (pinned)
    init_host = torch.from_numpy(np.ascontiguousarray(g["vertex_values"])).pin_memory()
    out_host = torch.empty_like(init_host).pin_memory()
    init_np, out_np = init_host.numpy(), out_host.numpy()
    from rustrobotics_b200.mapping._lib import lib, ptr
    L = lib()
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        pg.set_poses(init_np)                                               # H2D: this step's input poses
        r = pg.gn_step(allow_not_converged=False)                           # D2H: |dx|, chi2, PCG iterations
        rc = L.pgo_get_poses(pg._h, ptr(out_np), len(out_np))               # D2H: the updated poses
        assert rc == 0
        return r
    e2e_step()
    barrier()
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - e0) / e2e_steps

    # ---- reduce over ranks: max time
    step_ms = dev_ms / args.steps
    tt = torch.tensor([step_ms, wall_s / args.steps * 1e3, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_ms, wall_ms, e2e_ms = tt.tolist()
    total_edges = n_edges                                                  # one graph, sharded over the ranks (strong scaling)
    value = total_edges / (step_ms * 1e-3)

    # ---- whole-step figure by SURVEY 8(d)'s definition: the compulsory bytes of linearise+assemble, block-Jacobi setup, k PCG
    # iterations of the block-Jacobi form (SpMV + preconditioner apply + vector updates), retraction and chi2, with k = the PCG
    # iterations actually run.  For the AMG path this UNDER-counts what an iteration really moves (three fine SpMVs, the cycle's
    # vector kernels, the coarse levels), so the fraction is a lower bound there; for --preconditioner 0 it is the exact model.
    N_, E_ = n_poses, n_edges
    k_its = sum(pcg_its) / len(pcg_its)
    if D == 3:
        step_bytes = (240 * E_ + 120 * N_) + 144 * N_ + k_its * (152 * E_ + 464 * N_) + 72 * N_ + (80 * E_ + 24 * N_)
    else:     # 6x6 blocks: 72 -> 288 per block, 24 -> 48 per vector record, 56-byte poses, 21-value information triangle
        step_bytes = ((8 + 56 + 168) * E_ + 16 * E_ + 288 * (N_ + 2 * E_) + (56 + 48) * N_) + 2 * 288 * N_ + \
                     k_its * ((288 + 4) * (N_ + 2 * E_) + (4 + 96 + 288 + 96 + 6 * 48) * N_) + (56 * 2 + 48) * N_ + ((8 + 56 + 168) * E_ + 56 * N_)
    part = pg.partition() if world * n_shards > 1 else None
    step_tr = (ncu_traffic("step_traffic.json") or {}) if (D == 3 and world == 1 and n_poses == 1_000_000 and args.preconditioner == 1) else {}
    # ---- the reference's benches/graph_slam.rs shape (:7-11): PoseGraph::new + optimize(5) per sample, host set-up included
    new_opt5 = None
    if world == 1 and n_shards == 1 and not args.no_secondary:
        pg.close()
        t0 = time.perf_counter()
        pg2 = PoseGraph(graph=g, options=Options(device=local, pcg_rtol=args.pcg_rtol, preconditioner=args.preconditioner, **extra))
        t_new = time.perf_counter() - t0
        errs = pg2.optimize(5)
        t_all = time.perf_counter() - t0
        new_opt5 = {"new_s": t_new, "new_plus_optimize5_s": t_all, "gn_iterations_run": len(errs) - 1, "final_chi2": errs[-1],
                    "what": "PoseGraph(graph) + optimize(5) through the public API, host symbolic pass and uploads included (reference benches/graph_slam.rs:7-11 shape)"}
        pg2.close()
    # ---- the same step with pgo_options.refine = 1 (one refinement round on a double-double residual): the mode that solves the assembled
    # system exactly (1.4e-8 m; tests/test_gpu_parity.py::test_config4_solver_error_is_separated_from_the_conditioning_of_the_step)
    refined = None
    if world == 1 and n_shards == 1 and args.workload == "manhattan" and not args.no_secondary:
        pg3 = PoseGraph(graph=g, options=Options(device=local, pcg_rtol=args.pcg_rtol, preconditioner=args.preconditioner, refine=1, **extra))
        pg3.snapshot_poses()
        msr, itr = [], []
        for i in range(2 + 3):
            pg3.restore_poses()
            ndr, c2r, kr = pg3.gn_step(allow_not_converged=False)
            if i >= 2:
                msr.append(sum(v[0] for kk, v in pg3.timings().items() if kk != "spmv_fine")); itr.append(kr)
        refined = {"ms_per_step": sum(msr) / len(msr), "pcg_iterations_per_step": sum(itr) / len(itr), "steps": 3, "warmup": 2,
                   "parity": golden_check(c2r, ndr, n_poses, args.workload),
                   "what": "same step with pgo_options.refine = 1: + one round of iterative refinement, residual in double-double arithmetic"}
        pg3.close()
    # ---- BASELINE configs[4] (SE3 sphere, 250k poses / 1M edges; repo-defined SE3 semantics, parity unpinned) as a secondary line
    secondary = None
    if world == 1 and n_shards == 1 and args.workload == "manhattan" and n_poses == 1_000_000 and not args.no_secondary:
        g3, _, d3 = make_graph("sphere", 250_000)
        p3 = PoseGraph(graph=g3, options=Options(device=local, pcg_rtol=args.pcg_rtol))
        p3.snapshot_poses()
        ms3, it3 = [], []
        for i in range(3 + 3):
            p3.restore_poses()
            _, c3, k3 = p3.gn_step(allow_not_converged=False)
            if i >= 3:
                ms3.append(sum(v[0] for kk, v in p3.timings().items() if kk != "spmv_fine")); it3.append(k3)
        secondary = {"config": {"workload": d3.format(len(g3["vertex_id"]), len(g3["edge_from"])), "pcg_rtol": args.pcg_rtol}, "metric": METRIC, "unit": UNIT,
                     "value": len(g3["edge_from"]) / (sum(ms3) / len(ms3) * 1e-3), "ms_per_step": sum(ms3) / len(ms3), "steps": 3, "warmup": 3,
                     "pcg_iterations_per_step": sum(it3) / len(it3), "chi2_after_step": c3, "parity": "unpinned (the reference's SE3 optimise is todo!())"}
        p3.close()
    line = None
    if rank == 0:
        cpu = None
        if world == 1 and n_shards == 1 and not args.no_cpu_baseline:
            v, sec, sample, cores = cpu_reference_run(CPU_SAMPLE_POSES_SE3 if D == 6 else CPU_SAMPLE_POSES, 1 if args.workload == "manhattan" else 2, 0,
                                                      args.workload)   # bundled graphs ignore the size
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "s_per_gn_iteration": sec,
                   "extrapolation": CPU_EXTRAPOLATION if args.workload == "manhattan" and n_poses > CPU_SAMPLE_POSES else None}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world * n_shards, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "bundled dataset" if args.workload in BUNDLED else "synthetic",
            "config": {"workload": wl_desc.format(n_poses, n_edges) + "; 1 step = 1 Gauss-Newton iteration from the initial guess",
                       "poses": n_poses, "edges": n_edges, "pcg_rtol": args.pcg_rtol, "option_overrides": args.opts or None,
                       "preconditioner": "aggregation-AMG K-cycle (flexible PCG); the cycle's SpMVs read fp32 copies of the stored blocks and accumulate in fp64, the PCG operator / residual / dot products are fp64" if args.preconditioner == 1 else "block-Jacobi",
                       "parallelism": "single GPU" if world * n_shards == 1 else
                       f"1 graph sharded over {n_shards} GPUs by contiguous vertex ranges, ONE process / one PoseGraph handle (pgo_options.n_gpus); halo rows read "
                       f"from peer HBM (NVLink), all-reduces fused into the producing kernels' last blocks" if n_shards > 1 else
                       f"1 graph sharded over {world} GPUs by contiguous vertex ranges; halo rows read from peer HBM (NVLink), "
                       f"device-side peer-memory all-reduce for the dot products",
                       "l2": (f"working set {st['device_bytes'] * world / 1e9:.1f} GB >> 126 MB L2, no flush needed" if st['device_bytes'] * world > 1e9 else
                              f"working set {st['device_bytes'] * world / 1e6:.1f} MB fits the 126 MB L2 (small bundled graph: L2-resident by nature, not flushed)")},
            "gn_iterations_per_sec": 1e3 / step_ms, "pcg_iterations_per_step": sum(pcg_its) / len(pcg_its),
            "wall_ms_per_step": wall_ms, "phase_ms": phases, "create_s": t_create, "e2e_new_plus_optimize5": new_opt5,
            "partition": part,
            "chi2": {"initial": chi2_0, "after_step": last[1], "norm_dx": last[0]},
            "roofline": {"bound": "hbm", "kernel": f"k_spmv<{D},0> (fine-level BSR SpMV)" + ("" if world * n_shards == 1 else ", the largest shard"), "achieved": achieved, "peak": peak,
                         "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak, "ms_per_launch": spmv_ms,
                         "algorithmic_bytes_per_launch": spmv_bytes,
                         "traffic": (tr or {}).get("dram_bytes_per_launch") if (D == 3 and world == 1 and n_poses == 1_000_000) else None,
                         "traffic_source": (tr or {}).get("source") if (D == 3 and world == 1 and n_poses == 1_000_000) else None},
            "step_roofline": {"definition": "SURVEY 8(d): compulsory bytes of assemble + block-Jacobi setup + k block-Jacobi-form PCG iterations + retract + chi2, k = PCG iterations run; a lower bound for the AMG path",
                              "bytes_per_step": step_bytes, "achieved": step_bytes / (step_ms * 1e-3) / 1e9 / (world * n_shards), "unit": "GB/s per GPU",
                              "frac": step_bytes / (step_ms * 1e-3) / 1e9 / (world * n_shards) / peak,
                              # what the step REALLY moves: ncu dram__bytes over every kernel of one converged GN step (profiles/)
                              "traffic": step_tr.get("dram_bytes_per_step"), "traffic_source": step_tr.get("source"),
                              "traffic_pcg_iterations": step_tr.get("pcg_iterations"),
                              "hbm_utilisation": (step_tr["dram_bytes_per_step"] * (k_its / step_tr["pcg_iterations"] if step_tr.get("pcg_iterations") else 1.0)
                                                  / (step_ms * 1e-3) / 1e9 / peak) if step_tr.get("dram_bytes_per_step") else None},
            "parity": golden_check(last[1], last[0], n_poses, args.workload),
            "secondary": secondary, "refined": refined,
            "cpu_baseline": cpu,
            "e2e": {"value": total_edges / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(init_np.nbytes), "d2h_bytes_per_step": int(out_np.nbytes) + 20 * world},
            "gpu_launches": int(launches), "clocks": clocks, "host_cpus": os.cpu_count(),
        }
        emit(line)
    pg.close()                           # idempotent
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line: dict):
    """the ONE JSON line goes to the process's real stdout; everything else a library prints (e.g. NCCL's version banner)
    was redirected to stderr by main()"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                       # fd 1 -> stderr for the rest of the run (C libraries included)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="manhattan", choices=["manhattan", "sphere"] + list(BUNDLED),
                    help="manhattan = BASELINE configs[3] (the headline, default); sphere = configs[4] (SE3, 6x6 blocks); "
                         "pose-pose / pose-landmark / intel / dlr / m3500 / sphere2500 / garage = the reference's bundled g2o graphs")
    ap.add_argument("--poses", type=int, default=None, help="default: 1M (manhattan) / 250k (sphere)")
    ap.add_argument("--pcg-rtol", type=float, default=DEFAULT_PCG_RTOL)
    ap.add_argument("--preconditioner", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs[4] secondary line and the new+optimize(5) figure")
    ap.add_argument("--opts", default="", help="extra pgo_options overrides, k=v,k=v (e.g. amg_kcycle3=0,amg_fp64_storage=1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.poses is None:
        args.poses = 250_000 if args.workload == "sphere" else 1_000_000
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
