"""Host-side logic on CPU: C++ g2o loader/writer, C ABI surface, symbolic pass (bit-exact structure vs the
oracle), synthetic generator.  No compute entry point is called (there is no GPU here)."""
import ctypes as C
import hashlib
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import KEYS, ROOT, SE2_GRAPHS, graph_of, load_golden
import reference_kat as KAT
from oracle.oracle import OraclePoseGraph


def test_library_exports_every_header_symbol(built):
    from rustrobotics_b200.mapping._lib import ABI_SYMBOLS
    header = (ROOT / "include" / "pgo_b200.h").read_text()
    declared = set(re.findall(r"\b(pgo_[a-z0-9_]+)\s*\(", header))
    assert declared == set(ABI_SYMBOLS)
    for s in declared:
        assert hasattr(built, s), f"libpgo_b200.so does not export {s}"
    assert b"sm_100a" in built.pgo_version()


@pytest.mark.parametrize("name", list(KAT.FROM_G2O))
def test_parse_g2o_counts_and_roundtrip(g2o_files, name):      # g2o.rs:149-175
    from rustrobotics_b200 import parse_g2o
    ln, g = parse_g2o(g2o_files[name])
    assert (len(g["vertex_id"]), len(g["edge_kind"]), ln) == KAT.FROM_G2O[name]
    gold = load_golden(name)
    for k in KEYS:                                              # %.17g text round trip is bit exact
        assert np.array_equal(g[k], gold[k]), k
    o = OraclePoseGraph.from_g2o(g2o_files[name])               # oracle's own parser agrees
    oa = o.arrays()
    for k in ("vertex_id", "vertex_kind", "edge_kind", "edge_from", "edge_to", "edge_info_upper"):
        assert np.array_equal(g[k], oa[k]), k


def test_parse_g2o_tolerates_double_and_trailing_spaces(tmp_path):   # g2o.rs:52
    from rustrobotics_b200 import parse_g2o
    p = tmp_path / "a.g2o"
    p.write_text("VERTEX_SE2 0  0 0 0 \nVERTEX_SE2 1 1 0 0.5\r\nVERTEX_XY 7 2 2\n"
                 "EDGE_SE2 0 1 1 0 0.5  1 0 0 1 0 1 \nEDGE_SE2_XY 1 7 1 2 1 0 1\n")
    ln, g = parse_g2o(p)
    assert ln == 8 and g["vertex_id"].tolist() == [0, 1, 7] and g["edge_kind"].tolist() == [0, 1]
    assert g["edge_info_upper"].tolist() == [1, 0, 0, 1, 0, 1, 1, 0, 1]


@pytest.mark.parametrize("text,what", [
    ("VERTEX_SE2 0 0 0 0\n\nVERTEX_SE2 1 0 0 0\n", "blank line"),              # index panic, g2o.rs:53
    ("VERTEX_SE2 0 0 0 0\nFIX 0\n", "not implemented"),                        # unimplemented!, :138
    ("VERTEX_SE2 0 0 0\n", "wrong number"),                                    # todo!() arm, :56-58
    ("VERTEX_SE2 a 0 0 0\n", "bad vertex id"),                                 # parse()? -> Err, :55
    ("VERTEX_SE2 0 0 zero 0\n", "bad number"),                                 # unwrap panic, :30
    ("VERTEX_SE2 0 0 0 0\nEDGE_SE2 0 1 1 0 0 1 0 0 1 0\n", "wrong number"),
])
def test_parse_g2o_errors(tmp_path, text, what):
    from rustrobotics_b200 import parse_g2o
    p = tmp_path / "bad.g2o"
    p.write_text(text)
    with pytest.raises(ValueError, match=what):
        parse_g2o(p)
    with pytest.raises(ValueError):
        OraclePoseGraph.from_g2o(p)


def test_parse_g2o_rejects_hex_floats_and_ignores_the_locale(tmp_path, built):
    """Rust's str::parse::<f64> (g2o.rs:24-32) knows no hex floats and no locale; neither does the C++ loader"""
    import locale
    from rustrobotics_b200 import parse_g2o, write_g2o
    p = tmp_path / "hex.g2o"
    p.write_text("VERTEX_SE2 0 0x10 1 0\n")
    with pytest.raises(ValueError, match="bad number"):
        parse_g2o(p)
    p.write_text("VERTEX_SE2 0 +1.5 -2.5e-1 1E2\nVERTEX_XY 1 inf .5\n")
    _, g = parse_g2o(p)
    assert g["vertex_values"].tolist() == [1.5, -0.25, 100.0, float("inf"), 0.5]
    old = locale.setlocale(locale.LC_NUMERIC)
    try:
        for loc in ("de_DE.UTF-8", "fr_FR.UTF-8", "de_DE", "C.UTF-8"):       # whichever decimal-comma locale the image has
            try:
                locale.setlocale(locale.LC_NUMERIC, loc)
                break
            except locale.Error:
                continue
        p.write_text("VERTEX_SE2 0 1.5 2.25 0.125\n")
        _, g = parse_g2o(p)
        assert g["vertex_values"].tolist() == [1.5, 2.25, 0.125]
        q = tmp_path / "out.g2o"
        write_g2o(q, g)
        assert q.read_text() == "VERTEX_SE2 0 1.5 2.25 0.125\n"
    finally:
        locale.setlocale(locale.LC_NUMERIC, old)


def test_raw_reference_lines_parse_like_the_oracle(tmp_path, built):
    """the reference's own bytes (committed excerpts of dataset/g2o/sphere2500.g2o -- every line ends in a space, fields separated
    by two spaces before the information block -- and dlr.g2o), not text re-written by write_g2o: g2o.rs:52"""
    from rustrobotics_b200 import parse_g2o
    for name in ("sphere2500_head", "dlr_head"):
        p = ROOT / "tests" / "golden" / "raw" / f"{name}.g2o"
        assert name != "sphere2500_head" or p.read_bytes().split(b"\n")[0].endswith(b" ")
        ln, g = parse_g2o(p)
        oa = OraclePoseGraph.from_g2o(p).arrays()
        for k in KEYS:
            if k in ("vertex_values", "edge_meas"):    # the oracle hands back its stored state (unit complex / normalised quaternion)
                np.testing.assert_allclose(g[k], oa[k], atol=2e-6 if name == "sphere2500_head" else 1e-12)
            else:
                assert np.array_equal(g[k], oa[k]), (name, k)
        # every number exactly as Python's (correctly rounded) float() reads the same token
        toks = [ln_.split() for ln_ in p.read_text().splitlines()]
        want_v = [float(t) for tk in toks if tk[0].startswith("VERTEX") for t in tk[2:]]
        want_e = [float(t) for tk in toks if tk[0].startswith("EDGE") for t in tk[3:]]
        assert g["vertex_values"].tolist() == want_v
        nm = {0: 3, 1: 2, 2: 7}
        got_e, im, ii = [], 0, 0
        for kind in g["edge_kind"]:
            m, w = nm[int(kind)], {0: 6, 1: 3, 2: 21}[int(kind)]
            got_e += g["edge_meas"][im:im + m].tolist() + g["edge_info_upper"][ii:ii + w].tolist()
            im += m; ii += w
        assert got_e == want_e
        assert ln == int(np.sum(np.array([3, 2, 6])[g["vertex_kind"]]))


@pytest.mark.parametrize("name", SE2_GRAPHS + ["sphere2500", "parking-garage"])
def test_raw_reference_files_parse_to_the_committed_fixtures(built, name):
    """the whole raw files, when the reference checkout is present (authoring container; skipped on the GPU box)"""
    from rustrobotics_b200 import parse_g2o
    p = Path("/root/reference/dataset/g2o") / f"{name}.g2o"
    if not p.exists():
        pytest.skip("reference checkout not present")
    ln, g = parse_g2o(p)
    gold = load_golden(name)
    assert ln == int(gold["len"])
    se3 = name in ("sphere2500", "parking-garage")
    for k in KEYS:
        if k in ("vertex_values", "edge_meas"):   # the fixtures hold the oracle's stored state (theta through atan2(sin, cos); normalised quaternions)
            np.testing.assert_allclose(g[k], gold[k], atol=5e-6 if se3 else 1e-12)
        else:
            assert np.array_equal(g[k], gold[k]), k


def test_graph_arrays_are_validated(tmp_path, built):
    """kinds outside 0..2 and value arrays that do not match the per-kind counts are rejected before anything reads them"""
    from rustrobotics_b200 import Options, PgoError, PoseGraph, write_g2o
    g = graph_of(load_golden("simulation-pose-landmark"))
    for key, val, what in (("vertex_kind", 200, "vertex_kind"), ("edge_kind", 7, "edge_kind")):
        bad = dict(g); bad[key] = g[key].copy(); bad[key][0] = val
        with pytest.raises(PgoError, match=what):
            PoseGraph(graph=bad, options=Options(device=-2))
        with pytest.raises(ValueError, match=what):
            write_g2o(tmp_path / "bad.g2o", bad)
    for key, what in (("vertex_values", "vertex values"), ("edge_meas", "measurement values"), ("edge_info_upper", "information values")):
        bad = dict(g); bad[key] = g[key][:-1]
        with pytest.raises(PgoError, match=what):
            PoseGraph(graph=bad, options=Options(device=-2))
        with pytest.raises(ValueError, match=what):
            write_g2o(tmp_path / "bad.g2o", bad)
    # the C ABI itself rejects bad kinds (it cannot know the array lengths: documented caller contract)
    from rustrobotics_b200.mapping._lib import lib, ptr
    L = lib()
    a = {k: np.ascontiguousarray(g[k]) for k in KEYS}
    a["vertex_kind"] = a["vertex_kind"].copy(); a["vertex_kind"][3] = 9
    h = C.c_void_p()
    o = Options(device=-2)
    rc = L.pgo_create(C.byref(h), C.byref(o), len(a["vertex_id"]), ptr(a["vertex_id"]), ptr(a["vertex_kind"]), ptr(a["vertex_values"]),
                      len(a["edge_kind"]), ptr(a["edge_kind"]), ptr(a["edge_from"]), ptr(a["edge_to"]), ptr(a["edge_meas"]), ptr(a["edge_info_upper"]))
    assert rc == 1 and b"vertex_kind[3]" in L.pgo_last_error(None)


def test_missing_file_is_an_error(tmp_path, built):                # fs::read_to_string(...)? -> Err, g2o.rs:51
    from rustrobotics_b200 import PgoError, PoseGraph, parse_g2o
    with pytest.raises(ValueError):
        parse_g2o(tmp_path / "nope.g2o")
    with pytest.raises(PgoError):
        PoseGraph(tmp_path / "nope.g2o")


def _structure_handle(graph):
    from rustrobotics_b200 import Options, PoseGraph
    return PoseGraph(graph=graph, options=Options(device=-2))


@pytest.mark.parametrize("name", SE2_GRAPHS + ["sphere2500", "parking-garage"])
def test_pattern_bit_exact(built, name):
    """the symbolic pass's scalar CSC pattern == the pattern the reference's COO puts turn into (oracle)"""
    gold = load_golden(name)
    pg = _structure_handle(graph_of(gold))
    cp, ri = pg.pattern()
    sls = OraclePoseGraph.from_arrays(**graph_of(gold)).build_linear_system()
    assert np.array_equal(cp, sls.col_ptr) and np.array_equal(ri, sls.row_idx)
    assert len(ri) == int(gold["nnz"])
    h = hashlib.sha256(cp.tobytes() + ri.tobytes()).hexdigest()
    assert h == str(gold["pattern_sha256"])


@pytest.mark.parametrize("name", SE2_GRAPHS + ["sphere2500", "parking-garage"])
def test_block_structure_and_slot_map(built, name):
    """block CSR + edge->slot map against a direct restatement of update_linear_system's four set_matrix calls"""
    gold = load_golden(name)
    pg = _structure_handle(graph_of(gold))
    rp, bc, es = pg.block_structure()
    o = OraclePoseGraph.from_arrays(**graph_of(gold))
    _, fi, ti = o.edge_endpoints()
    nv = o.n_vertices
    pairs = set(zip(fi.tolist(), ti.tolist())) | set(zip(ti.tolist(), fi.tolist())) | {(v, v) for v in range(nv)}
    want = sorted(pairs)
    got = [(r, int(c)) for r in range(nv) for c in bc[rp[r]:rp[r + 1]]]
    assert got == want
    assert len(bc) == nv + 2 * len(set(zip(np.minimum(fi, ti).tolist(), np.maximum(fi, ti).tolist())))
    index = {rc: k for k, rc in enumerate(want)}
    exp = np.array([[index[(a, a)], index[(a, b)], index[(b, a)], index[(b, b)]] for a, b in zip(fi.tolist(), ti.tolist())])
    assert np.array_equal(es, exp)


def test_anchor_is_first_pose_pose_edges_from(built):               # :330-336, SURVEY fact 6
    for name, want_id in (("intel", 2), ("simulation-pose-landmark", 100), ("dlr", 0)):
        gold = load_golden(name)
        pg = _structure_handle(graph_of(gold))
        assert int(gold["vertex_id"][pg.anchor()]) == want_id


def test_create_rejects_malformed_graphs(built):
    from rustrobotics_b200 import Options, PgoError, PoseGraph
    g = graph_of(load_golden("simulation-pose-landmark"))
    bad = dict(g); bad["edge_to"] = g["edge_to"].copy(); bad["edge_to"][0] = 99999
    with pytest.raises(PgoError, match="unknown vertex id"):
        PoseGraph(graph=bad, options=Options(device=-2))
    bad = dict(g); bad["edge_kind"] = g["edge_kind"].copy(); bad["edge_kind"][1] = 0   # EDGE_SE2 onto a landmark
    bad["edge_meas"] = np.concatenate([g["edge_meas"], [0.0]]); bad["edge_info_upper"] = np.concatenate([g["edge_info_upper"], [0.0] * 3])
    with pytest.raises(PgoError, match="kinds"):
        PoseGraph(graph=bad, options=Options(device=-2))
    bad = dict(g); bad["vertex_id"] = g["vertex_id"].copy(); bad["vertex_id"][1] = bad["vertex_id"][0]
    with pytest.raises(PgoError, match="duplicate"):
        PoseGraph(graph=bad, options=Options(device=-2))


def test_landmark_graphs_aggregate_poses_and_attach_landmarks(built):
    """level-0 aggregation of a graph with VERTEX_XY: poses are aggregated over pose-pose edges (pairs of chain aggregates), every
    landmark joins the aggregate of most of its observers -- never a landmark-rooted star (dlr: 250 -> 41 PCG iterations)"""
    g = graph_of(load_golden("dlr"))
    pg = _structure_handle(g)
    rows, _ = pg.level_sizes()
    assert rows[0] == 3873 and 400 <= rows[1] <= 640            # ~6 poses per aggregate, and the dense coarsest level right below
    agg = pg.aggregates(0)
    kind = g["vertex_kind"]
    lut = {int(v): i for i, v in enumerate(g["vertex_id"])}
    # every landmark sits in the aggregate that holds the plurality of its observing poses
    obs = {}
    for a, b, k in zip(g["edge_from"], g["edge_to"], g["edge_kind"]):
        if k == 1:
            obs.setdefault(lut[int(b)], []).append(agg[lut[int(a)]])
    for lm, aggs in obs.items():
        vals, cnt = np.unique(aggs, return_counts=True)
        assert cnt[vals == agg[lm]].sum() == cnt.max(), lm
    # no aggregate consists of landmarks only (unless the landmark has no observer at all)
    pose_aggs = set(agg[kind == 0].tolist())
    assert all(a in pose_aggs for a in agg[kind == 1].tolist())
    sizes = np.bincount(agg[kind == 0])
    assert sizes.max() <= 2 * 16


def test_structure_only_handle_refuses_to_compute(built):
    from rustrobotics_b200 import PgoError
    pg = _structure_handle(graph_of(load_golden("simulation-pose-landmark")))
    with pytest.raises(PgoError, match="no CPU fallback"):
        pg.global_error()
    with pytest.raises(PgoError):
        pg.gn_step()


def test_no_device_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rustrobotics_b200 import PgoError, PoseGraph
    with pytest.raises(PgoError, match="no CPU fallback"):
        PoseGraph(graph=graph_of(load_golden("simulation-pose-landmark")))


def test_synthetic_manhattan_is_deterministic_and_well_formed():
    from rustrobotics_b200.synthetic import manhattan_se2
    a, b = manhattan_se2(5000), manhattan_se2(5000)
    for k in KEYS:
        assert np.array_equal(a[k], b[k])
    c = manhattan_se2(5000, seed=7)
    assert not np.array_equal(a["vertex_values"], c["vertex_values"])
    f, t = a["edge_from"].astype(np.int64), a["edge_to"].astype(np.int64)
    assert np.all(f < t)                                              # from < to always
    assert len(set(zip(f.tolist(), t.tolist()))) == len(f)            # no duplicate pairs
    odo = (t - f) == 1
    assert odo.sum() == 5000 - 1 and np.all((t - f)[~odo] > 10)       # N-1 odometry edges, closures span > 10
    assert 3.5 * 5000 <= len(f) <= 4 * 5000
    g = manhattan_se2(20000)
    assert len(g["edge_from"]) == 4 * 20000
    # converges under the oracle in a handful of Gauss-Newton iterations
    o = OraclePoseGraph.from_arrays(**manhattan_se2(2000))
    errs = o.optimize(10)
    assert errs[-1] < 0.2 * errs[0] and len(errs) <= 8


def test_synthetic_sphere_is_well_formed():
    from rustrobotics_b200.synthetic import sphere_se3
    g = sphere_se3(20, 25)
    n = 500
    assert len(g["vertex_id"]) == n and len(g["vertex_values"]) == 7 * n
    q = g["vertex_values"].reshape(n, 7)[:, 3:]
    np.testing.assert_allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-12)
    assert 3.8 * n <= len(g["edge_from"]) <= 4 * n


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the CPU arm: the oracle port on the host cores) runs without a GPU and prints ONE JSON line
    on stdout with the contract's keys; everything else goes to stderr"""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "pose-landmark", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gn_edges_per_sec" and d["unit"] == "edges/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_symbolic_pass_is_independent_of_the_host_thread_count(built, monkeypatch):
    """the symbolic pass runs its independent loops on the host's cores: same structure for 1 and for many threads"""
    import numpy as np
    from rustrobotics_b200 import Options, PoseGraph
    from rustrobotics_b200.synthetic import manhattan_se2
    g = manhattan_se2(20000)
    out = []
    for nt in ("1", "7"):
        monkeypatch.setenv("PGO_HOST_THREADS", nt)
        pg = PoseGraph(graph=g, options=Options(device=-2))
        rows, blocks = pg.level_sizes()
        out.append((rows, blocks, pg.aggregates(0).copy(), [a.copy() for a in pg.block_structure()]))
        pg.close()
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1]
    assert np.array_equal(out[0][2], out[1][2])
    for a, b in zip(out[0][3], out[1][3]):
        assert np.array_equal(a, b)
