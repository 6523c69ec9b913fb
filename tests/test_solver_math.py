"""CPU checks of the algebra two CUDA kernels implement (numpy mirrors of the device code; the kernels themselves are checked on
the GPU through the parity suites): the scalar recurrences of the three-step K-cycle (kernels.cuh: finalize FIN_K1 / FIN_K2 /
FIN_K3) and the blocked ping-pong Gauss-Jordan inverse of the dense coarsest level (k_dense_invert)."""
import numpy as np


def _spd(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, n))
    return a @ a.T + n * np.eye(n)


def test_three_step_kcycle_coefficients_reproduce_flexible_cg():
    """x = coef1 c1 + coef2 c2 + coef3 c3 with the FIN_K1/K2/K3 recurrences equals three steps of flexible CG with full
    orthogonalisation of the search directions (any preconditioner: here a fixed SPD matrix B)"""
    n = 40
    A, B = _spd(n, 0), np.linalg.inv(_spd(n, 1))
    rhs = np.random.default_rng(2).standard_normal(n)
    # reference: textbook FCG with explicit directions
    x = np.zeros(n); r = rhs.copy(); ds, ws, xs = [], [], []
    for _ in range(3):
        c = B @ r
        d = c.copy()
        for dj, wj in zip(ds, ws):                      # A-orthogonalise against every previous direction
            d -= (c @ wj) / (dj @ wj) * dj
        w = A @ d
        al = (d @ r) / (d @ w)
        x = x + al * d; r = r - al * w
        ds.append(d); ws.append(w); xs.append(x.copy())
    # device recurrences (kernels.cuh finalize): dots are taken with the UN-orthogonalised c_i and v_i = A c_i
    c1 = B @ rhs; v1 = A @ c1
    rho1, a1 = c1 @ v1, c1 @ rhs                      # FIN_K1: {c1.v1, c1.rhs}
    alpha = a1 / rho1
    r1 = rhs - alpha * v1                             # k_kresid_dinv<1>
    c2 = B @ r1; v2 = A @ c2
    gam, beta, a2 = c2 @ v1, c2 @ v2, c2 @ r1         # FIN_K2: {c2.v1, c2.v2, c2.r1}
    rho2 = beta - gam * gam / rho1
    coef1 = a1 / rho1 - gam * a2 / (rho1 * rho2); coef2 = a2 / rho2
    alpha2 = a2 / rho2; e2 = alpha2; e1 = alpha2 * gam / rho1
    np.testing.assert_allclose(coef1 * c1 + coef2 * c2, xs[1], rtol=1e-10, atol=1e-13)      # Notay's two-step K-cycle
    r2 = r1 - e2 * v2 + e1 * v1                       # k_kresid_dinv<2>
    c3 = B @ r2; v3 = A @ c3
    t = [c3 @ v1, c3 @ v2, c3 @ v3, c3 @ r2]          # FIN_K3
    g31 = t[0]; g32 = t[1] - (gam / rho1) * t[0]
    rho3 = t[2] - g31 * g31 / rho1 - g32 * g32 / rho2
    a3 = t[3] / rho3; b32, b31, b21 = g32 / rho2, g31 / rho1, gam / rho1
    coef1 += a3 * (b32 * b21 - b31); coef2 -= a3 * b32; coef3 = a3
    np.testing.assert_allclose(coef1 * c1 + coef2 * c2 + coef3 * c3, x, rtol=1e-9, atol=1e-12)


def test_blocked_pingpong_gauss_jordan_inverts():
    """k_dense_invert: panels of 32, every panel step reads `src` and writes the whole matrix to `dst`
    (A[pp] <- P^-1, A[p,r] <- P^-1 A[p,r], A[r,p] <- -A[r,p] P^-1, A[r,r] <- A[r,r] - A[r,p] P^-1 A[p,r]); the matrix is treated
    as padded with an identity block up to a multiple of the panel width"""
    W = 32
    for m in (1, 31, 32, 33, 100, 257):
        A = _spd(m, m)
        src = A.copy()
        for p0 in range(0, m, W):
            w = min(W, m - p0)
            p = slice(p0, p0 + w)
            rest = np.r_[0:p0, p0 + w:m]
            Pinv = np.linalg.inv(src[p, p])
            dst = np.empty_like(src)
            R = Pinv @ src[p][:, rest]
            dst[np.ix_(range(p0, p0 + w), range(p0, p0 + w))] = Pinv
            dst[np.ix_(range(p0, p0 + w), rest)] = R
            dst[np.ix_(rest, range(p0, p0 + w))] = -src[rest][:, p] @ Pinv
            dst[np.ix_(rest, rest)] = src[np.ix_(rest, rest)] - src[rest][:, p] @ R
            src = dst
        np.testing.assert_allclose(src @ A, np.eye(m), atol=1e-9)
