"""CPU research (NOT product code, not used by tests): does a SMOOTHED prolongator on level 0 only make the K-cycle size-robust?
Backs "next" item 1 of DESIGN.md section 8.   python tools/research/smoothed_level0.py [poses]

Level 0 uses the library's own aggregates (structure-only handle); P0 = (I - w Dinv H) P_tentative with w = 0.66 / rho(Dinv H);
the coarser levels are plain root+neighbours aggregation of the (denser) Galerkin operator (variant 2: level 1 aggregated over the
unsmoothed coarse graph instead, so that its coarsening ratio stays 16:1), rigid-motion coarse spaces as in the
library, K-cycle steps as the library runs them ((3, 2) up to four levels, 3 on every level beyond)."""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools" / "research"))
from amg_proto import *            # noqa: E402,F401,F403
from amg_kcycle import fcg         # noqa: E402
import hierarchy_study as hs       # noqa: E402


def mk_level(H):
    L = Level(); L.H = H; L.Dinv = block_diag_inv(H)
    v = np.random.default_rng(0).standard_normal(H.shape[0])
    for _ in range(15):
        v = L.Dinv @ (H @ v); rho = np.linalg.norm(v); v /= rho
    L.rho = rho; L.omega = 1.5 / rho
    return L


hs.mk_level = mk_level


def hierarchy(g, H0, pos0, smooth, dense_max=640):
    from rustrobotics_b200 import Options, PoseGraph
    pg = PoseGraph(graph=g, options=Options(device=-2))
    agg = pg.aggregates(0).astype(np.int64)
    pg.close()
    uniq, comp = np.unique(agg, return_inverse=True)
    L0 = mk_level(H0)
    P, cen = tentative_P(comp, len(uniq), pos0)
    if smooth:
        P = (P - (0.66 / L0.rho) * (L0.Dinv @ (H0 @ P))).tocsr()
    L0.P = P
    Pt = tentative_P(comp, len(uniq), pos0)[0]
    A1 = adjacency((Pt.T @ H0 @ Pt).tocsr())          # smooth == 2: level 1 is aggregated over the UNSMOOTHED coarse graph
    levels, H, pos = [L0], (P.T @ H0 @ P).tocsr(), cen
    while True:
        L = mk_level(H); levels.append(L)
        if H.shape[0] // 3 <= dense_max or len(levels) > 8:
            L.lu = spla.splu(H.tocsc()); break
        a, nc = aggregate_graph(A1 if (smooth == 2 and len(levels) == 2) else adjacency(H), 16)
        L.P, cen = tentative_P(a, nc, pos)
        H = (L.P.T @ H @ L.P).tocsr(); pos = cen
    return levels


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
    g, H0, b, pos0 = build(n)
    for smooth in (0, 1, 2):
        t = time.time()
        lv = hierarchy(g, H0, pos0, smooth)
        ms = (3, 2) if len(lv) <= 4 else (3, 3, 3, 3, 3)
        its = fcg(H0, b, lambda r: hs.cyc(lv, 0, r, ms), rtol=1e-9)[1]
        print(f"{n} poses, smoothed P0 = {smooth}: levels {[L.H.shape[0] // 3 for L in lv]}, blocks per row of level 1: "
              f"{lv[1].H.nnz / 9 / (lv[1].H.shape[0] // 3):.1f}, level-1 / level-0 nnz {lv[1].H.nnz / H0.nnz:.2f}, "
              f"nnz(P0) / rows {lv[0].P.nnz / 3 / (H0.shape[0] // 3):.1f}; K-cycle{ms[:2]} PCG its {its}  ({time.time() - t:.0f} s)", flush=True)
