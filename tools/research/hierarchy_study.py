"""CPU research (NOT product code, not used by tests): how the multilevel K-cycle depends on the shape of the hierarchy.
Backs the statements of DESIGN.md sections 4b and 6.  Uses the LIBRARY's own aggregates (pgo_get_aggregates on a structure-only
handle, no GPU needed) inside the scipy prototype of tools/research/amg_proto.py.

    python tools/research/hierarchy_study.py worlds   [poses]   # library hierarchies of world 1 / 2 / 8: two-grid and K-cycle PCG counts
    python tools/research/hierarchy_study.py ordering [poses]   # same level 1, deeper levels aggregated in different root orders
    python tools/research/hierarchy_study.py variants [poses]   # aggregation algorithm variants on every level

Results at 1M poses (seed 42): worlds -> K-cycle(3,2) 42 vs 51 iterations for world 1 vs 2 (two-grid: equal); ordering -> natural 42,
Morton 56, RCM 57, degree 50, random 93; variants -> root+neighbours 16: 38, pass-2 'smallest': 60, max 12: 53, max 24: 37,
pairwise x3: 195, pairwise x4: 412."""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools" / "research"))
from amg_proto import *            # noqa: E402,F401,F403  (build, Level, block_diag_inv, adjacency, aggregate_graph, tentative_P, np, sp, spla)
from amg_kcycle import fcg         # noqa: E402


def mk_level(H):
    L = Level(); L.H = H; L.Dinv = block_diag_inv(H)
    v = np.random.default_rng(0).standard_normal(H.shape[0])
    for _ in range(15):
        v = L.Dinv @ (H @ v); rho = np.linalg.norm(v); v /= rho
    L.omega = min(1.0, 4.0 / (3.0 * 1.1 * rho))
    return L


def inner_fcg(Hc, B, rhs, m):
    """m steps of flexible CG with full orthogonalisation of the search directions, from zero"""
    x = np.zeros_like(rhs); r = rhs.copy(); ps, qs, pqs = [], [], []
    for _ in range(m):
        z = B(r); p = z.copy()
        for pj, qj, pqj in zip(ps, qs, pqs):
            p -= (z @ qj) / pqj * pj
        q = Hc @ p; pq = p @ q
        a = (p @ r) / pq
        x += a * p; r -= a * q
        ps.append(p); qs.append(q); pqs.append(pq)
    return x


def cyc(levels, l, r, msteps):
    """one cycle at level l; msteps[l] inner FCG steps solve level l+1 (the last level is solved exactly)"""
    L = levels[l]
    if l == len(levels) - 1:
        return L.lu.solve(r)
    x = L.omega * (L.Dinv @ r)
    rc = L.P.T @ (r - L.H @ x)
    B = lambda v: cyc(levels, l + 1, v, msteps)
    m = msteps[l] if l < len(msteps) else msteps[-1]
    ec = B(rc) if (l + 1 == len(levels) - 1 or m <= 1) else inner_fcg(levels[l + 1].H, B, rc, m)
    x = x + L.P @ ec
    return x + L.omega * (L.Dinv @ (r - L.H @ x))


def library_hierarchy(g, H0, pos0, world):
    from rustrobotics_b200 import Options, PoseGraph
    pg = PoseGraph(graph=g, options=Options(device=-2, world=world, rank=0))
    sizes = pg.level_sizes()[0]
    levels, H, pos, l = [], H0, pos0, 0
    agg = pg.aggregates(0).astype(np.int64)
    while True:
        L = mk_level(H); levels.append(L)
        if l == len(sizes) - 1:
            L.lu = spla.splu(H.tocsc()); break
        uniq, comp = np.unique(agg, return_inverse=True)
        assert len(uniq) == sizes[l + 1]
        L.P, cen = tentative_P(comp, len(uniq), pos)
        H = (L.P.T @ H @ L.P).tocsr(); pos = cen; l += 1
        if l < len(sizes) - 1:
            a = None
            for cand in range(int(np.ceil((uniq.max() + 1) / 32.0) * 32), int(uniq.max()) + 32 * 12, 32):   # padded rows of the level
                try:
                    a = pg.aggregates(l, cand); break
                except Exception:
                    continue
            agg = a[uniq].astype(np.int64)
    pg.close()
    return levels, sizes


def generic_hierarchy(H0, pos0, level1, fn, dense_max=640):
    """levels below `level1` (a (P, H1, pos1) triple or None) built with the aggregation function fn(A, pos) -> (agg, nc)"""
    levels, H, pos = [], H0, pos0
    while True:
        L = mk_level(H); levels.append(L)
        if H.shape[0] // 3 <= dense_max or len(levels) > 7:
            L.lu = spla.splu(H.tocsc()); break
        if len(levels) == 1 and level1 is not None:
            L.P, H, pos = level1
            continue
        a, nc = fn(adjacency(H), pos)
        L.P, cen = tentative_P(a, nc, pos)
        H = (L.P.T @ H @ L.P).tocsr(); pos = cen
    return levels


def morton(pos):
    p = pos - pos.min(0); s = (p / max(p.max(), 1e-9) * 65535).astype(np.uint64)

    def spread(x):
        x = (x | (x << 16)) & 0x0000FFFF0000FFFF; x = (x | (x << 8)) & 0x00FF00FF00FF00FF
        x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0F; x = (x | (x << 2)) & 0x3333333333333333
        return (x | (x << 1)) & 0x5555555555555555
    return np.argsort(spread(s[:, 0]) | (spread(s[:, 1]) << np.uint64(1)), kind="stable")


def ordered(kind):
    from scipy.sparse.csgraph import reverse_cuthill_mckee

    def fn(A, pos):
        n = A.shape[0]
        perm = {"natural": lambda: np.arange(n), "random": lambda: np.random.default_rng(1).permutation(n), "morton": lambda: morton(pos),
                "rcm": lambda: np.asarray(reverse_cuthill_mckee(A.tocsr(), symmetric_mode=True)),
                "degree": lambda: np.argsort(-np.diff(A.tocsr().indptr), kind="stable")}[kind]()
        a, nc = aggregate_graph(A.tocsr()[perm][:, perm], 16)
        out = np.empty_like(a); out[perm] = a
        return out, nc
    return fn


def pairwise(passes):
    def fn(A, pos):
        n = A.shape[0]; total = np.arange(n)
        for _ in range(passes):
            A = A.tocsr(); ptr, nbr = A.indptr, A.indices
            m = A.shape[0]; agg = -np.ones(m, np.int64); nc = 0
            for i in range(m):
                if agg[i] >= 0: continue
                agg[i] = nc
                for j in nbr[ptr[i]:ptr[i + 1]]:
                    if agg[j] < 0 and j != i: agg[j] = nc; break
                nc += 1
            total = agg[total]
            Pm = sp.csr_matrix((np.ones(m), (np.arange(m), agg)), shape=(m, nc))
            A = (Pm.T @ A @ Pm).tocsr(); A.setdiag(0); A.eliminate_zeros()
        return total, nc
    return fn


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "worlds"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 300000
    g, H0, b, pos0 = build(n)
    report = lambda name, lv: print(f"{name:36s} levels {[L.H.shape[0] // 3 for L in lv]} K-cycle(3,2) PCG its "
                                    f"{fcg(H0, b, lambda r: cyc(lv, 0, r, (3, 2)))[1]}", flush=True)
    if what == "worlds":
        for w in (1, 2, 8):
            lv, sizes = library_hierarchy(g, H0, pos0, w)
            two = [lv[0], mk_level(lv[1].H)]; two[1].lu = spla.splu(lv[1].H.tocsc())
            print(f"world {w}: two-grid PCG its {fcg(H0, b, lambda r: cyc(two, 0, r, (1,)))[1]}", flush=True)
            report(f"world {w} (library hierarchy)", lv)
    elif what == "ordering":
        lv, _ = library_hierarchy(g, H0, pos0, 1)
        P0 = lv[0].P; H1 = lv[1].H
        cen = np.zeros((H1.shape[0] // 3, 2)); cnt = np.zeros(H1.shape[0] // 3)
        comp = np.asarray(P0.tocsr()[0::3][:, 0::3].argmax(axis=1)).ravel()
        np.add.at(cen, comp, pos0); np.add.at(cnt, comp, 1); cen /= cnt[:, None]
        for kind in ("natural", "morton", "rcm", "degree", "random"):
            report(f"order {kind}", generic_hierarchy(H0, pos0, (P0, H1, cen), ordered(kind)))
    else:
        base = lambda A, pos: aggregate_graph(A, 16)
        report("root+neighbours 16", generic_hierarchy(H0, pos0, None, base))
        report("root+neighbours 24", generic_hierarchy(H0, pos0, None, lambda A, pos: aggregate_graph(A, 24)))
        report("root+neighbours 12", generic_hierarchy(H0, pos0, None, lambda A, pos: aggregate_graph(A, 12)))
        report("pairwise x3", generic_hierarchy(H0, pos0, None, pairwise(3)))
