"""Stage-by-stage GPU-vs-oracle diagnostics (prints, never asserts): run on the B200 box via gpurun."""
import sys, time, traceback
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from conftest import SE2_GRAPHS, graph_of, load_golden
from oracle.oracle import OraclePoseGraph
from rustrobotics_b200 import Options, PoseGraph
from rustrobotics_b200.synthetic import manhattan_se2


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def check_graph(name, graph, precond, its=30, rtol=1e-10, full=True):
    print(f"=== {name} precond={'amg' if precond else 'bj'}", flush=True)
    o = OraclePoseGraph.from_arrays(**graph)
    t = time.time()
    pg = PoseGraph(graph=graph, options=Options(preconditioner=precond, pcg_rtol=rtol))
    print(f"  create {time.time()-t:.2f}s stats {pg.stats()} levels {pg.level_sizes()}")
    c_g, c_o = pg.global_error(), o.global_error()
    print(f"  chi2 gpu {c_g:.10g} oracle {c_o:.10g} rel {abs(c_g-c_o)/c_o:.2e}")
    if full:
        sls = o.build_linear_system()
        cp, ri, vals, b = pg.system()
        print(f"  pattern eq {np.array_equal(cp, sls.col_ptr) and np.array_equal(ri, sls.row_idx)}  H rel {rel(vals, sls.vals):.2e} max|dH| {np.abs(vals-sls.vals).max():.2e}  b rel {rel(b, sls.b):.2e}")
        dx_o = sls.solve()
        t = time.time(); dx_g, it = pg.linearize_and_solve(); dt = time.time() - t
        print(f"  first dx: pcg its {it} time {dt*1e3:.1f}ms rel {rel(dx_g, dx_o):.2e} maxabs {np.abs(dx_g-dx_o).max():.2e}")
    t = time.time(); errs_o, norms_o = o.optimize(its, return_norms=True); t_o = time.time() - t
    t = time.time(); errs_g = pg.optimize(its); t_g = time.time() - t
    n = min(len(errs_g), len(errs_o))
    print(f"  optimize: gpu {len(errs_g)-1} its {t_g:.3f}s | oracle {len(errs_o)-1} its {t_o:.3f}s | pcg its {pg.pcg_iterations}")
    print(f"  chi2 hist rel err {[float(f'{abs(a-b)/b:.1e}') for a, b in zip(errs_g[:n], errs_o[:n])]}")
    print(f"  final chi2 gpu {errs_g[-1]:.10g} oracle {errs_o[-1]:.10g}")
    _, _, _, vo = o.vertices()
    vg = pg.poses()
    print(f"  final poses max abs diff {np.abs(vg - vo).max():.3e}")
    print(f"  timings(last step) {pg.timings()}")
    pg.close()


def main():
    which = sys.argv[1:] or ["small", "synth"]
    if "small" in which:
        for name in SE2_GRAPHS:
            for pc in (0, 1):
                try:
                    check_graph(name, graph_of(load_golden(name)), pc)
                except Exception:
                    traceback.print_exc()
    if "synth" in which:
        for n, pcs in ((10000, (0, 1)), (100000, (1,))):
            g = manhattan_se2(n)
            for pc in pcs:
                try:
                    check_graph(f"manhattan{n}", g, pc, its=8)
                except Exception:
                    traceback.print_exc()
    if "big" in which:
        g = manhattan_se2(1000000)
        try:
            t = time.time()
            pg = PoseGraph(graph=g, options=Options(preconditioner=1, pcg_rtol=1e-8))
            print(f"=== manhattan 1M create {time.time()-t:.1f}s {pg.stats()} levels {pg.level_sizes()}")
            print("  chi2", pg.global_error())
            for i in range(4):
                t = time.time(); r = pg.gn_step(); print(f"  step {i}: {r} {time.time()-t:.3f}s {pg.timings()}", flush=True)
            print("  spmv ms", pg.time_spmv(20))
        except Exception:
            traceback.print_exc()


if __name__ == "__main__":
    main()
