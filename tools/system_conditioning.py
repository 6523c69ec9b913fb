"""Where does the ~3e-5 m between the GPU's first Gauss-Newton step on BASELINE configs[3] and the golden come from?  

Run on the GPU box.  (1) the GPU assembles H, b (pgo_get_system) -- they differ from the oracle's in the last bits only (different
summation order); (2) the GPU's REFINED solve (pgo_options.refine = 1) is compared with the true solution of the GPU-ASSEMBLED system,
computed on the CPU by the golden generator's solver (SciPy AMG-PCG + long-double refinement); (3) both are compared with the golden =
the true solution of the ORACLE-assembled system.  If (2) agrees to ~1e-8 m while (3) is ~3e-5 m, the gap is the conditioning of the
problem (cond(H) * eps * |dx|), not the solver.    python tools/system_conditioning.py [n_poses]"""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests" / "golden"))
import numpy as np
import scipy.sparse as sp
from make_golden_1m import true_solution
from oracle.oracle import OraclePoseGraph
from rustrobotics_b200 import Options, PoseGraph
from rustrobotics_b200.synthetic import manhattan_se2

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
g = manhattan_se2(n)
pos = g["vertex_values"].reshape(n, 3)[:, :2].copy()
t = time.time()
pg = PoseGraph(graph=g, options=Options(pcg_rtol=1e-9, refine=1))
cp, ri, vals, b = pg.system()
dx_ref, its = pg.linearize_and_solve()
print(f"GPU system + refined solve: {time.time() - t:.1f}s, {its} PCG iterations", flush=True)
pg2 = PoseGraph(graph=g, options=Options(pcg_rtol=1e-9))
dx_plain, its2 = pg2.linearize_and_solve()
Hg = sp.csc_matrix((vals, ri, cp), shape=(3 * n, 3 * n))
o = OraclePoseGraph.from_arrays(**g)
sls = o.build_linear_system()
print("assembled systems, GPU vs oracle: max |dH| / max |H| = %.2e, max |db| / max |b| = %.2e" %
      (np.abs(vals - sls.vals).max() / np.abs(sls.vals).max(), np.abs(b - sls.b).max() / np.abs(sls.b).max()), flush=True)
x_gpu_sys, last, _ = true_solution(Hg, b, pos)
def err(a, c):
    e = np.abs((a - c).reshape(n, 3))
    return "max xy %.2e m, theta %.2e rad" % (e[:, :2].max(), e[:, 2].max())
print("GPU refined dx  vs TRUE solution of the GPU-assembled system   :", err(dx_ref, x_gpu_sys), f"(truth pinned to {last:.1e})", flush=True)
print("GPU plain dx    vs TRUE solution of the GPU-assembled system   :", err(dx_plain, x_gpu_sys), flush=True)
gold = ROOT / "tests" / "golden" / "manhattan_1m_step1.npz"
if n == 1_000_000 and gold.exists():
    z = np.load(gold); s = z["sample"]
    e = np.abs(x_gpu_sys.reshape(n, 3)[s] - z["dx_sample"])
    print("TRUE solution of the GPU-assembled system vs golden (oracle-assembled), sampled: max xy %.2e m, theta %.2e rad" % (e[:, :2].max(), e[:, 2].max()), flush=True)
else:
    x_or, _, _ = true_solution(sls.csc(), sls.b, pos)
    print("TRUE solution of the GPU-assembled system vs TRUE solution of the oracle-assembled system:", err(x_gpu_sys, x_or), flush=True)
