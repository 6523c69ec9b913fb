"""Generates tests/golden/*.npz from the reference's bundled g2o datasets (run in the authoring container,
where /root/reference exists; the GPU box only sees the committed .npz files).

Each fixture holds (a) the graph as flat arrays, parsed by the ORACLE's restatement of parse_g2o, so that tests can
rebuild the exact g2o text with write_g2o; (b) the oracle's results on it (chi2 history, |dx| history, final poses,
first Gauss-Newton dx, sha256 of the CSC pattern).  The reference's own known-answer values live in
tests/reference_kat.py and are checked against these in tests/test_oracle_kat.py.

    python tests/golden/make_golden.py
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.oracle import OraclePoseGraph  # noqa: E402

DATASETS = Path("/root/reference/dataset/g2o")
SE2_FILES = ["simulation-pose-pose", "simulation-pose-landmark", "intel", "dlr", "input_M3500_g2o"]
# SE(3): the reference only PARSES these (optimize is todo!() for SE3, pose_graph_optimization.rs:241,357,570); the results stored
# for them are the oracle's own (repo-defined semantics) -- parity unpinned.
SE3_FILES = ["sphere2500", "parking-garage"]


def pattern_hash(col_ptr, row_idx):
    return hashlib.sha256(np.ascontiguousarray(col_ptr, np.int32).tobytes() + np.ascontiguousarray(row_idx, np.int32).tobytes()).hexdigest()


def main():
    out = Path(__file__).resolve().parent
    for name in SE2_FILES + SE3_FILES:
        g = OraclePoseGraph.from_g2o(DATASETS / f"{name}.g2o")
        arrays = g.arrays()
        sls = g.build_linear_system()
        dx0 = sls.solve()
        errs, norms = g.optimize(100 if name in SE2_FILES else 12, return_norms=True)
        _, _, _, final = g.vertices()
        np.savez_compressed(out / f"{name}.npz", **arrays, len=np.int64(g.len), chi2_history=np.array(errs),
                            norm_history=np.array(norms), final_values=final, dx0=dx0, puts=np.int64(sls.puts),
                            nnz=np.int64(len(sls.row_idx)), pattern_sha256=np.array(pattern_hash(sls.col_ptr, sls.row_idx)))
        print(name, g.n_vertices, g.n_edges, g.len, len(errs) - 1, errs[0], errs[-1])


if __name__ == "__main__":
    main()
