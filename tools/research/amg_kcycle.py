import sys, time
sys.path.insert(0, "tools/research")
from amg_proto import *

def smooth_pre(L, r, nu):
    x = L.omega * (L.Dinv @ r)
    for _ in range(nu - 1):
        x = x + L.omega * (L.Dinv @ (r - L.H @ x))
    return x

def kcycle(levels, l, r, klevels, nu=1, count=None):
    """preconditioner application at level l: smoothing + coarse correction; the coarse system is solved by
    2 FCG steps preconditioned by kcycle(l+1) if (l+1) in klevels else by one kcycle(l+1) (V)."""
    L = levels[l]
    if count is not None: count[l] = count.get(l, 0) + 1
    if l == len(levels) - 1:
        return L.dense @ r if L.dense is not None else L.lu.solve(r)
    x = smooth_pre(L, r, nu)
    res = r - L.H @ x
    rc = L.P.T @ res
    Lc = levels[l + 1]
    B = lambda v: kcycle(levels, l + 1, v, klevels, nu, count)
    if (l + 1) in klevels and l + 1 < len(levels) - 1:
        c1 = B(rc); v1 = Lc.H @ c1; rho1 = c1 @ v1; a1 = c1 @ rc
        r1 = rc - (a1 / rho1) * v1
        c2 = B(r1); v2 = Lc.H @ c2; gam = c2 @ v1; beta = c2 @ v2; a2 = c2 @ r1
        rho2 = beta - gam * gam / rho1
        ec = (a1 / rho1 - gam * a2 / (rho1 * rho2)) * c1 + (a2 / rho2) * c2
    else:
        ec = B(rc)
    x = x + L.P @ ec
    for _ in range(nu):
        x = x + L.omega * (L.Dinv @ (r - L.H @ x))
    return x

def fcg(H, b, M, rtol=1e-8, maxit=3000):
    """flexible CG (Notay FCG(1)): beta from (z_new . q_old) -- works with a variable preconditioner"""
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); rz0 = r @ z
    for it in range(1, maxit + 1):
        q = H @ p; pq = p @ q; a = (p @ r) / pq; x += a * p; r -= a * q
        z = M(r); rz1 = r @ z
        if rz1 <= rtol * rtol * rz0: return x, it
        beta = -(z @ q) / pq
        p = z + beta * p
    return x, maxit

if __name__ == "__main__":
    n = int(sys.argv[1])
    g, H, b, pos = build(n)
    xd = spla.splu(H.tocsc()).solve(b)
    for name, kw, kl in [
        ("chain4 V (fcg)", dict(), ()),
        ("chain4 K@1", dict(), (1,)),
        ("chain4 K@1,2", dict(), (1, 2)),
        ("chain4 K@all", dict(), (1, 2, 3, 4, 5)),
        ("graph16 K@1", dict(first="graph"), (1,)),
        ("graph16 K@all", dict(first="graph"), (1, 2, 3, 4)),
        ("graph8  K@all", dict(first="graph", max_size=8), (1, 2, 3, 4, 5)),
    ]:
        t = time.time(); lv = setup(H, pos, **kw); ts = time.time() - t
        sizes = [L.H.shape[0] // 3 for L in lv]; nnzs = [L.H.nnz for L in lv]
        cnt = {}
        t = time.time(); x, it = fcg(H, b, lambda r: kcycle(lv, 0, r, kl, 1, cnt)); tp = time.time() - t
        work = sum(cnt.get(l, 0) * nnzs[l] for l in range(len(lv))) / nnzs[0] / it
        print(f"{name:20s} its {it:5d} err {np.linalg.norm(x-xd)/np.linalg.norm(xd):.1e} levels {sizes} visits/it {[cnt.get(l,0)//it for l in range(len(lv))]} work/it {work:.2f} setup {ts:.1f}s pcg {tp:.1f}s", flush=True)
