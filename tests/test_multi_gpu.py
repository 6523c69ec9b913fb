"""Single-process multi-GPU handle (pgo_options.n_gpus, include/pgo_b200.h): ONE PoseGraph in ONE process drives N shards,
one per GPU -- `PoseGraph::new(path, solver)?.optimize(n, ..)` (pose_graph_optimization.rs:215, :247) unchanged for the caller.

GPU tests: every case runs (a) with the shards on N distinct GPUs when the box has them, else (b) with all N shards sharing
cuda:0 (device_ids = [0] * N) -- the identical code path (vertex-range partition, halo pulls from the peers' arenas, cross-shard
reductions and barriers in peer memory), so the sharded numerics are covered on the one-GPU lease of the test driver too.
Results must match the CPU oracle exactly as the single-GPU path does: chi2 per iteration 1e-6 relative, poses 1e-6 m / rad."""
import numpy as np
import pytest

from conftest import graph_of, load_golden

CHI2_RTOL, POSE_ATOL = 1e-6, 1e-6


def _graph(case):
    from rustrobotics_b200.synthetic import manhattan_se2, sphere_se3
    if case.startswith("manhattan"):
        return manhattan_se2(int(case[len("manhattan"):]))
    if case.startswith("sphere") and "x" in case:
        return sphere_se3(*[int(t) for t in case[len("sphere"):].split("x")])
    return graph_of(load_golden(case))


def _device_ids(n):
    import torch
    nd = torch.cuda.device_count()
    return list(range(n)) if nd >= n else [k % max(nd, 1) for k in range(n)]


_ORACLE = {}


def _oracle_run(case, its):
    if case not in _ORACLE:
        from oracle.oracle import OraclePoseGraph
        g = _graph(case)
        o = OraclePoseGraph.from_arrays(**g)
        c0 = o.global_error()
        errs = o.optimize(its)
        _ORACLE[case] = (g, c0, errs, o.vertices()[3])
    return _ORACLE[case]


def test_options_struct_carries_the_device_list(built):
    from rustrobotics_b200 import Options
    o = Options(device_ids=[0, 0, 0])
    assert o.n_gpus == 3 and [o.device_ids[k] for k in range(3)] == [0, 0, 0]
    assert Options().n_gpus == 0 and not Options().device_ids


def test_multi_handle_needs_a_device_and_says_so(built):
    """no GPU in the authoring container: creating a multi-GPU handle fails loudly (no CPU fallback); on a GPU box a bad
    device list is rejected"""
    import torch
    from rustrobotics_b200 import Options, PgoError, PoseGraph
    g = graph_of(load_golden("simulation-pose-pose"))
    if not torch.cuda.is_available():
        with pytest.raises(PgoError, match="no CUDA device"):
            PoseGraph(graph=g, options=Options(n_gpus=2))
    else:
        with pytest.raises(PgoError, match="device_ids"):
            PoseGraph(graph=g, options=Options(device_ids=[0, 99]))
    with pytest.raises(PgoError, match="exclusive"):
        PoseGraph(graph=g, options=Options(n_gpus=2, world=2, rank=1))
    with pytest.raises(PgoError, match="n_gpus > 8"):
        PoseGraph(graph=g, options=Options(n_gpus=9))


@pytest.mark.gpu
@pytest.mark.parametrize("n_gpus", [2, 4])
@pytest.mark.parametrize("case", ["simulation-pose-pose", "intel", "dlr", "manhattan100000", "sphere40x50"])
def test_multi_gpu_handle_matches_the_oracle(built, case, n_gpus):
    from rustrobotics_b200 import Options, PoseGraph
    from test_gpu_parity import _pose_diff
    from test_gpu_se3 import se3_pose_diff
    big = case == "manhattan100000"
    its = 3 if big else 8
    g, c0, errs_o, vo = _oracle_run(case, its)
    pg = PoseGraph(graph=g, options=Options(device_ids=_device_ids(n_gpus)))
    part = pg.partition()
    assert part["world"] == n_gpus and part["vertex_range"][0] == 0 and part["vertex_range"][-1] == len(g["vertex_id"])
    assert abs(pg.global_error() - c0) <= 1e-12 * c0
    errs_g = pg.optimize(its)
    assert len(errs_g) == len(errs_o)
    np.testing.assert_allclose(errs_g, errs_o, rtol=CHI2_RTOL)
    se3 = int(g["vertex_kind"][0]) == 2
    dxy, dth = se3_pose_diff(pg.poses(), vo) if se3 else _pose_diff(g, pg.poses(), vo)
    assert dxy < POSE_ATOL and dth < POSE_ATOL
    pg.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["manhattan100000", "sphere40x50"])
def test_multi_gpu_handle_with_sharded_coarse_levels(built, monkeypatch, case):
    """what a graph too large to replicate level 1 looks like (8M poses on 8 GPUs), forced on a small one: PGO_REPL_MAX_ROWS makes the
    coarse levels above 700 rows SHARDED too, so the K-cycle's inner products, halo pulls and barriers cross the GPUs on every level"""
    from rustrobotics_b200 import Options, PoseGraph
    from test_gpu_parity import _pose_diff
    from test_gpu_se3 import se3_pose_diff
    monkeypatch.setenv("PGO_REPL_MAX_ROWS", "700")
    g, c0, errs_o, vo = _oracle_run(case, 3 if case == "manhattan100000" else 8)
    pg = PoseGraph(graph=g, options=Options(device_ids=_device_ids(3)))
    errs_g = pg.optimize(len(errs_o) - 1)
    np.testing.assert_allclose(errs_g, errs_o[:len(errs_g)], rtol=CHI2_RTOL)
    se3 = int(g["vertex_kind"][0]) == 2
    dxy, dth = se3_pose_diff(pg.poses(), vo) if se3 else _pose_diff(g, pg.poses(), vo)
    assert dxy < POSE_ATOL and dth < POSE_ATOL
    pg.close()


@pytest.mark.gpu
def test_multi_gpu_handle_assembles_the_same_system_and_same_api(built):
    """pattern / H / b bit-identical in structure and 1e-12 in value to the oracle, set/get/snapshot/undo behave like 1 GPU"""
    from oracle.oracle import OraclePoseGraph
    from rustrobotics_b200 import Options, PgoError, PoseGraph
    g = graph_of(load_golden("intel"))
    o = OraclePoseGraph.from_arrays(**g)
    sls = o.build_linear_system()
    pg = PoseGraph(graph=g, options=Options(device_ids=_device_ids(3)))
    cp, ri, vals, b = pg.system()
    assert np.array_equal(cp, sls.col_ptr) and np.array_equal(ri, sls.row_idx)
    assert np.abs(vals - sls.vals).max() <= 1e-12 * np.abs(sls.vals).max()
    assert np.abs(b - sls.b).max() <= 1e-11 * max(np.abs(sls.b).max(), 1.0)
    dx, _ = pg.linearize_and_solve()
    assert np.abs(dx - sls.solve()).max() <= 1e-8 * np.abs(dx).max()
    v0 = pg.poses()
    np.testing.assert_allclose(v0, g["vertex_values"], atol=1e-15)
    pg.snapshot_poses()
    c0 = pg.global_error()
    pg.gn_step()
    pg.undo_last_step()
    with pytest.raises(PgoError):
        pg.undo_last_step()                          # a step can be undone once
    np.testing.assert_allclose(pg.poses(), v0, atol=1e-9)
    pg.gn_step()
    pg.restore_poses()
    assert pg.global_error() == c0
    pg.set_poses(v0 + 0.01)
    from test_gpu_parity import _pose_diff
    dxy, dth = _pose_diff(g, pg.poses(), v0 + 0.01)              # theta comes back wrapped into (-pi, pi]
    assert dxy <= 1e-12 and dth <= 1e-12
    assert pg.stats()["block_rows"] == len(g["vertex_id"]) and pg.time_spmv(3) > 0
    pg.close()
    # the refinement round (double-double residual over the halo-extended x) on a sharded handle
    pr = PoseGraph(graph=g, options=Options(device_ids=_device_ids(3), refine=1))
    dxr, _ = pr.linearize_and_solve()
    assert np.abs(dxr - sls.solve()).max() <= 1e-8 * np.abs(dxr).max()
    pr.close()
