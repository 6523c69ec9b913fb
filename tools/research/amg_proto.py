"""CPU research prototype (NOT product code, not used by tests): explores AMG variants for the pose-graph
Gauss-Newton system in scipy, to decide what the CUDA V-cycle should implement."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from oracle.oracle import OraclePoseGraph
from rustrobotics_b200.synthetic import manhattan_se2


def build(n):
    g = manhattan_se2(n)
    o = OraclePoseGraph.from_arrays(**g)
    sls = o.build_linear_system()
    H = sls.csc().tocsr()
    pos = g["vertex_values"].reshape(-1, 3)[:, :2].copy()
    return g, H, sls.b, pos


def block_diag_inv(H, bs=3):
    n = H.shape[0] // bs
    B = sp.bsr_matrix(H, blocksize=(bs, bs))
    B.sort_indices()
    D = np.zeros((n, bs, bs))
    for i in range(n):
        s, e = B.indptr[i], B.indptr[i + 1]
        j = np.searchsorted(B.indices[s:e], i)
        D[i] = B.data[s + j]
    Dinv = np.linalg.inv(D)
    return sp.bsr_matrix((Dinv, np.arange(n), np.arange(n + 1)), shape=H.shape).tocsr()


def adjacency(H, bs=3):
    n = H.shape[0] // bs
    B = sp.bsr_matrix(H, blocksize=(bs, bs))
    A = sp.csr_matrix((np.ones(len(B.indices)), B.indices, B.indptr), shape=(n, n))
    A.setdiag(0); A.eliminate_zeros()
    return A


def aggregate_chain(A, run=4):
    n = A.shape[0]
    agg = np.empty(n, np.int64)
    nc, cur = 0, 0
    A = A.tocsr()
    for v in range(n):
        join = v > 0 and 0 < cur < run and A[v, v - 1] != 0
        if join:
            agg[v] = nc - 1; cur += 1
        else:
            agg[v] = nc; nc += 1; cur = 1
    return agg, nc


def aggregate_graph(A, max_size=16):
    n = A.shape[0]
    A = A.tocsr()
    ptr, nbr = A.indptr, A.indices
    agg = -np.ones(n, np.int64)
    nc = 0
    for i in range(n):
        if agg[i] >= 0: continue
        nb = nbr[ptr[i]:ptr[i + 1]]
        if np.any(agg[nb] >= 0): continue
        agg[i] = nc
        agg[nb[:max_size - 1]] = nc
        nc += 1
    snap = agg.copy()
    for i in range(n):
        if agg[i] >= 0: continue
        nb = nbr[ptr[i]:ptr[i + 1]]
        c = snap[nb]; c = c[c >= 0]
        if len(c) == 0: continue
        vals, cnt = np.unique(c, return_counts=True)
        agg[i] = vals[np.argmax(cnt)]
    for i in range(n):
        if agg[i] >= 0: continue
        agg[i] = nc
        nb = nbr[ptr[i]:ptr[i + 1]]
        for j in nb[:max_size - 1]:
            if agg[j] < 0: agg[j] = nc
        nc += 1
    return agg, nc


def tentative_P(agg, nc, pos):
    n = len(agg)
    cen = np.zeros((nc, 2)); cnt = np.bincount(agg, minlength=nc)
    np.add.at(cen, agg, pos)
    cen /= cnt[:, None]
    d = pos - cen[agg]
    blocks = np.zeros((n, 3, 3))
    blocks[:, 0, 0] = 1; blocks[:, 1, 1] = 1; blocks[:, 2, 2] = 1
    blocks[:, 0, 2] = -d[:, 1]; blocks[:, 1, 2] = d[:, 0]
    P = sp.bsr_matrix((blocks, agg, np.arange(n + 1)), shape=(3 * n, 3 * nc)).tocsr()
    return P, cen


class Level: pass


def setup(H, pos, smooth_P=False, first="chain", run=4, max_size=16, coarsest=200, omega_scale=1.0, max_levels=12):
    levels = []
    while True:
        L = Level(); L.H = H; L.Dinv = block_diag_inv(H)
        # rho(Dinv H)
        v = np.random.default_rng(0).standard_normal(H.shape[0])
        for _ in range(15):
            v = L.Dinv @ (H @ v); rho = np.linalg.norm(v); v /= rho
        L.rho = rho; L.omega = min(1.0, omega_scale * 4.0 / (3.0 * 1.1 * rho))
        levels.append(L)
        n = H.shape[0] // 3
        if n <= coarsest or len(levels) >= max_levels:
            L.dense = np.linalg.inv(H.toarray()) if n <= 2000 else None
            L.lu = None if L.dense is not None else spla.splu(H.tocsc())
            break
        A = adjacency(H)
        if len(levels) == 1 and first == "chain":
            agg, nc = aggregate_chain(A, run)
            if nc > 0.6 * n: agg, nc = aggregate_graph(A, max_size)
        else:
            agg, nc = aggregate_graph(A, max_size)
        if nc >= 0.9 * n:
            L.dense = None; L.lu = spla.splu(H.tocsc()); break
        P, cen = tentative_P(agg, nc, pos)
        if smooth_P:
            P = (P - (4.0 / (3.0 * rho)) * (L.Dinv @ (H @ P))).tocsr()
        L.P = P
        H = (P.T @ H @ P).tocsr()
        pos = cen
    return levels


def vcycle(levels, l, r, nu=1, gamma=1, cheb=0):
    L = levels[l]
    if l == len(levels) - 1:
        return L.dense @ r if L.dense is not None else L.lu.solve(r)
    x = L.omega * (L.Dinv @ r)
    for _ in range(nu - 1):
        x = x + L.omega * (L.Dinv @ (r - L.H @ x))
    res = r - L.H @ x
    rc = L.P.T @ res
    ec = vcycle(levels, l + 1, rc, nu, gamma)
    if gamma == 2 and l + 1 < len(levels) - 1:   # W-cycle: second coarse visit on the remaining residual
        Lc = levels[l + 1]
        ec = ec + vcycle(levels, l + 1, rc - Lc.H @ ec, nu, gamma)
    x = x + L.P @ ec
    for _ in range(nu):
        x = x + L.omega * (L.Dinv @ (r - L.H @ x))
    return x


def pcg(H, b, M, rtol=1e-8, maxit=5000):
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); rz = r @ z; rz0 = rz
    for it in range(1, maxit + 1):
        q = H @ p; a = rz / (p @ q); x += a * p; r -= a * q
        z = M(r); rz1 = r @ z
        if rz1 <= rtol * rtol * rz0: return x, it
        p = z + (rz1 / rz) * p; rz = rz1
    return x, maxit


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    g, H, b, pos = build(n)
    print("n", n, "nnz", H.nnz)
    xd = spla.splu(H.tocsc()).solve(b)
    for name, kw, cyc in [
        ("current: chain4+graph16, V(1,1)", dict(), dict()),
        ("V(2,2)", dict(), dict(nu=2)),
        ("W(1,1)", dict(), dict(gamma=2)),
        ("smoothed P, V(1,1)", dict(smooth_P=True), dict()),
        ("smoothed P, graph only, V(1,1)", dict(smooth_P=True, first="graph"), dict()),
        ("graph16 only V(1,1)", dict(first="graph"), dict()),
        ("chain8 V(1,1)", dict(run=8), dict()),
    ]:
        t = time.time(); lv = setup(H, pos, **kw); ts = time.time() - t
        sizes = [L.H.shape[0] // 3 for L in lv]; nnzs = [L.H.nnz for L in lv]
        t = time.time(); x, it = pcg(H, b, lambda r: vcycle(lv, 0, r, **cyc)); tp = time.time() - t
        print(f"{name:40s} its {it:5d} err {np.linalg.norm(x-xd)/np.linalg.norm(xd):.1e} levels {sizes} opcx {sum(nnzs)/nnzs[0]:.2f} setup {ts:.1f}s pcg {tp:.1f}s rho {[round(L.rho,2) for L in lv]}", flush=True)
