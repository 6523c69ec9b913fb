"""CPU research prototype (NOT product code): does solving the gauge-free system H0 x = b (H0 = H without the 1e7 anchor term,
singular but consistent because b is orthogonal to the global rigid motions) and fixing the gauge afterwards in closed form
give the reference's dx more accurately / in fewer iterations than PCG on the anchored system at the same rtol?"""
import sys, time
sys.path.insert(0, "tools/research")
from amg_proto import *
from amg_kcycle import kcycle, fcg

def gauge_modes(pos):
    n = len(pos)
    V = np.zeros((3 * n, 3))
    V[0::3, 0] = 1; V[1::3, 1] = 1
    V[0::3, 2] = -pos[:, 1]; V[1::3, 2] = pos[:, 0]; V[2::3, 2] = 1
    return V

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    g, H, b, pos = build(n)
    a = int(g["edge_from"][0])
    w = 1e7
    E = sp.csr_matrix((np.full(3, w), (np.arange(3) + 3 * a, np.arange(3) + 3 * a)), shape=H.shape)
    H0 = (H - E).tocsr()
    V = gauge_modes(pos)
    print("n", n, "|H0 V|", np.abs(H0 @ V).max(), "|V^T b|", np.abs(V.T @ b), flush=True)
    xd = spla.splu(H.tocsc()).solve(b)
    print("direct: x_a", xd[3 * a:3 * a + 3], "|x|", np.linalg.norm(xd))
    lv = setup(H, pos, first="graph", max_size=16, coarsest=640)
    print("levels", [L.H.shape[0] // 3 for L in lv])
    kl = (1, 2, 3, 4)
    M = lambda r: kcycle(lv, 0, r, kl, 1)
    def report(tag, x, it):
        e = (x - xd).reshape(-1, 3)
        print(f"{tag:28s} its {it:4d} max|e_xy| {np.abs(e[:, :2]).max():.2e} max|e_th| {np.abs(e[:, 2]).max():.2e} rms {np.sqrt((e**2).mean()):.2e} rel {np.linalg.norm(x-xd)/np.linalg.norm(xd):.1e}", flush=True)
    for rtol in (1e-6, 1e-8, 1e-10):
        x, it = fcg(H, b, M, rtol=rtol)
        report(f"anchored rtol {rtol:g}", x, it)
        x, it = fcg(H0, b, M, rtol=rtol)
        Va = V[3 * a:3 * a + 3]
        c = np.linalg.solve(Va, x[3 * a:3 * a + 3])
        report(f"gauge-free rtol {rtol:g}", x - V @ c, it)
