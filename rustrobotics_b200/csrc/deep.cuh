// Cluster-resident "deep levels" kernel: the K-cycle below the first small coarse level in ONE launch of ONE thread-block
// cluster.
//
// The deep levels of the hierarchy (3.7k / 0.4k block rows at BASELINE configs[3]) are a chain of ~16 dependent stages
// per coarse solve, each a 3-9 us kernel of 30-120 CTAs whose cost is launch ramp + one dependent-load chain + drain;
// three such solves per PCG iteration are ~18 % of a Gauss-Newton step.  Here the host records the stage program of
// one coarse solve once (the recursion coarse_solve / cycle of pgo_b200.cu unrolled) and ONE cluster of 16 (or 8) CTAs
// x 512 threads executes it with a hardware cluster barrier (barrier.cluster, ~0.2 us, also invalidates L1) between
// stages instead of a kernel boundary (~3 us).  Every stage is a cluster-stride loop over the level's rows; matrices
// and vectors stay in L2 (a level is < 2 MB); the K-cycle dot products are reduced through distributed shared memory:
// every CTA leaves its partial sums in its own shared memory, and after the stage barrier every CTA adds all of them
// in rank order (identical bits everywhere) and runs the scalar recurrence itself -- no global atomics, no second
// barrier.  A grid-wide version of this idea (tail.cuh, cooperative launch, grid.sync) LOST 8 % because grid.sync costs
// what a kernel boundary costs; the cluster barrier is the difference.
//
// The dense coarsest solve is a bandwidth problem (12 MB of fp64 inverse per apply at configs[3]) that 16 SMs read slower
// than the full-grid kernel does, so the program is cut at the dense stages: segments between them run here, the dense
// applies stay separate launches (PGO_DEEP_DENSE=1 runs them inline).
//
// STATUS: correct (parity-tested with PGO_DEEP=1) but NOT the default.  Measured on B200 at configs[3]
// (profiles/r01x_deep_experiment.log): one level-2 coarse solve 86 us here vs 53 us as 16 graph-launched kernels; the SE3
// sphere 137 vs 86 us.  The barrier is indeed cheap, but a stage is a chain of dependent L2 accesses (~0.7 us each)
// either way, and 8192 threads give a row 2 lanes where the full-grid kernels give it 8: the chains get longer.  What
// would win is holding the level (0.7 MB) and its vectors in the cluster's distributed shared memory; not built.
#pragma once
#include <type_traits>

#include "kernels.cuh"

namespace pgo {

enum { DOP_DINV = 0, DOP_SPMV = 1, DOP_RESTRICT = 2, DOP_PROLONG = 3, DOP_PROLONGK = 4, DOP_KRESID = 5, DOP_KCOMBINE = 6, DOP_DENSE = 7 };

struct DeepOp {
    int type, lvl, mode, fin;          // SPMV: mode 0/1/2, fin FIN_NONE / FIN_K1..3 ; KRESID: mode = step (1, 2) ; KCOMBINE: mode = which (1, 2)
    const double *a, *b, *c, *d;       // DINV: a = rhs ; SPMV: a = x, b = r, c = u1, d = u2 ; RESTRICT: a = res ; PROLONG: a = ec ;
                                       // PROLONGK: a, b, c = c1, c2, c3 (c3 may be null) ; KRESID: a = rhs, b = v1, c = w, d = xa (written) ;
                                       // KCOMBINE: a, b, c = c1, c2, c3 ; DENSE: a = rhs
    double *out;
    double omega;
};

struct DeepCtx {
    const DeepOp *prog; int op0, op1;  // this launch runs prog[op0 .. op1)
    const LevelDev *lv;
    int dense_m; const double *Ainv;
    Scalars *S;
};

constexpr int DEEP_NT = 512;

template <int VS> __device__ __forceinline__ void ldcg_vec(const double *p, double *o) {
#pragma unroll
    for (int i = 0; i < VS; i += 2) { const double2 t = __ldcg(reinterpret_cast<const double2 *>(p + i)); o[i] = t.x; o[i + 1] = t.y; }
}

// the K-cycle parts of finalize() on a KScal held in shared memory
__device__ __forceinline__ void deep_finalize(int fin, KScal &K, const double *t) {
    if (fin == FIN_K1) {
        K.rho1 = t[0]; K.a1 = t[1];
        K.alpha = (t[0] > 0.0) ? t[1] / t[0] : 0.0;
    } else if (fin == FIN_K2) {
        const double rho1 = K.rho1, a1 = K.a1, gam = t[0], beta = t[1], a2 = t[2];
        const double rho2 = beta - gam * gam / rho1;
        const bool ok = rho1 > 0.0 && rho2 > 1e-12 * beta && rho2 == rho2;
        if (ok) { K.coef1 = a1 / rho1 - gam * a2 / (rho1 * rho2); K.coef2 = a2 / rho2; }
        else { K.coef1 = K.alpha; K.coef2 = 0.0; }
        K.coef3 = 0.0;
        K.rho2 = ok ? rho2 : 0.0; K.gam21 = ok ? gam : 0.0;
        K.alpha2 = ok ? a2 / rho2 : 0.0;
        K.e2 = K.alpha2; K.e1 = ok ? K.alpha2 * gam / rho1 : 0.0;
    } else if (fin == FIN_K3) {
        const double rho1 = K.rho1, rho2 = K.rho2, g21 = K.gam21;
        if (rho1 > 0.0 && rho2 > 0.0) {
            const double g31 = t[0], g32 = t[1] - (g21 / rho1) * t[0];
            const double rho3 = t[2] - g31 * g31 / rho1 - g32 * g32 / rho2;
            if (rho3 > 1e-12 * t[2] && rho3 == rho3) {
                const double a3 = t[3] / rho3, b32 = g32 / rho2, b31 = g31 / rho1, b21 = g21 / rho1;
                K.coef1 += a3 * (b32 * b21 - b31);
                K.coef2 -= a3 * b32;
                K.coef3 = a3;
            }
        }
    }
}

// y = (MODE) of H x on a block-CSR level, LPR lanes per row, cluster-stride over the rows; dots accumulated per thread
template <int D, typename VT, int LPR>
__device__ __forceinline__ void deep_spmv(const LevelDev &L, const DeepOp &op, unsigned gtid, unsigned G, double *dots) {
    constexpr int DD = D * D, VS = VecStride<D>::value;
    const int sub = gtid & (LPR - 1);
    const VT *__restrict__ vals = level_val<VT>(L);
    for (int64_t base = 0; base < L.n; base += G / LPR) {        // same trip count for every thread: the shuffles below are full-warp
        const int64_t row = base + gtid / LPR;
        const bool live = row < L.n;
        double acc[D];
#pragma unroll
        for (int a = 0; a < D; a++) acc[a] = 0.0;
        if (live) {
            const int64_t b = __ldg(L.slice_ptr + row), e = __ldg(L.slice_ptr + row + 1);
            for (int64_t s = b + sub; s < e; s += LPR) {
                const uint32_t c = __ldg(L.col + s);
                double xj[VS];
                ldcg_vec<VS>(op.a + (int64_t)(c & COL_LOCAL_MASK) * VS, xj);
                const VT *v = vals + s * DD;
#pragma unroll
                for (int a = 0; a < D; a++)
#pragma unroll
                    for (int q = 0; q < D; q++) acc[a] = fma((double)__ldg(v + a * D + q), xj[q], acc[a]);
            }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
            for (int a = 0; a < D; a++) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
        }
        if (live && sub == 0) {
            double xi[VS], out[VS];
#pragma unroll
            for (int a = 0; a < VS; a++) out[a] = 0.0;
            ldcg_vec<VS>(op.a + row * VS, xi);
            const double *dg = L.diag + row;
#pragma unroll
            for (int a = 0; a < D; a++)
#pragma unroll
                for (int q = 0; q < D; q++) acc[a] = fma(__ldcg(dg + (int64_t)(a * D + q) * L.n_pad), xi[q], acc[a]);
            if (op.mode == 0) {
#pragma unroll
                for (int a = 0; a < D; a++) out[a] = acc[a];
                if (op.fin == FIN_K1) {
                    double ui[VS];
                    ldcg_vec<VS>(op.c + row * VS, ui);
#pragma unroll
                    for (int a = 0; a < D; a++) { dots[0] = fma(xi[a], acc[a], dots[0]); dots[1] = fma(xi[a], ui[a], dots[1]); }
                } else if (op.fin == FIN_K2) {
                    double ui[VS], wi[VS];
                    ldcg_vec<VS>(op.c + row * VS, ui);
                    ldcg_vec<VS>(op.d + row * VS, wi);
#pragma unroll
                    for (int a = 0; a < D; a++) { dots[0] = fma(xi[a], ui[a], dots[0]); dots[1] = fma(xi[a], acc[a], dots[1]); dots[2] = fma(xi[a], wi[a], dots[2]); }
                } else if (op.fin == FIN_K3) {
                    double ui[VS], wi[VS], zi[VS];
                    ldcg_vec<VS>(op.c + row * VS, ui);
                    ldcg_vec<VS>(op.d + row * VS, wi);
                    ldcg_vec<VS>(op.b + row * VS, zi);
#pragma unroll
                    for (int a = 0; a < D; a++) {
                        dots[0] = fma(xi[a], ui[a], dots[0]); dots[1] = fma(xi[a], wi[a], dots[1]);
                        dots[2] = fma(xi[a], acc[a], dots[2]); dots[3] = fma(xi[a], zi[a], dots[3]);
                    }
                }
            } else {
                double ri[VS];
                ldcg_vec<VS>(op.b + row * VS, ri);
                if (op.mode == 1) {
#pragma unroll
                    for (int a = 0; a < D; a++) out[a] = ri[a] - acc[a];
                } else {
                    const double *di = L.dinv + row;
                    double t[D];
#pragma unroll
                    for (int a = 0; a < D; a++) t[a] = ri[a] - acc[a];
#pragma unroll
                    for (int a = 0; a < D; a++) {
                        double s = 0.0;
#pragma unroll
                        for (int q = 0; q < D; q++) s = fma(__ldcg(di + (int64_t)(a * D + q) * L.n_pad), t[q], s);
                        out[a] = fma(op.omega, s, xi[a]);
                    }
                }
            }
            st_vec<VS>(op.out + row * VS, out);
        }
    }
}

template <int D>
__device__ __forceinline__ void deep_dinv_row(const LevelDev &L, int64_t row, const double *rv, double omega, double *out) {
    const double *di = L.dinv + row;
#pragma unroll
    for (int c = 0; c < D; c++) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < D; q++) s = fma(__ldcg(di + (int64_t)(c * D + q) * L.n_pad), rv[q], s);
        out[c] = omega * s;
    }
}

template <int D, bool LOWP>
__global__ void __launch_bounds__(DEEP_NT, 1) k_deep(const __grid_constant__ DeepCtx T) {
    PDL_ENTER();
    if (ld_done(T.S)) return;                       // set only by fine-level kernels: every CTA of the cluster sees the same value
    using VT = typename std::conditional<LOWP, float, double>::type;
    constexpr int VS = VecStride<D>::value;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned NC = cluster.num_blocks(), crank = cluster.block_rank();
    const unsigned G = NC * DEEP_NT, gtid = crank * DEEP_NT + threadIdx.x;
    const unsigned lane = threadIdx.x & 31, gw = gtid >> 5, GW = G >> 5;
    __shared__ KScal ks[MAX_LEVELS];
    __shared__ double red[2][4];                     // this CTA's partial sums, double-buffered by reduction parity
    __shared__ double wsum[DEEP_NT / 32][4];
    extern __shared__ double deep_sr[];              // dense stage: the right-hand side (dense_m doubles)
    for (int i = threadIdx.x; i < MAX_LEVELS; i += DEEP_NT) ks[i] = T.S->k[i];
    __syncthreads();
    int rpar = 0;
    for (int i = T.op0; i < T.op1; i++) {
        const DeepOp op = T.prog[i];
        const LevelDev &L = T.lv[op.lvl];
        double dots[4] = {0.0, 0.0, 0.0, 0.0};
        switch (op.type) {
        case DOP_DINV:
            for (int64_t row = gtid; row < L.n_pad; row += G) {
                double rv[VS], out[VS];
                ldcg_vec<VS>(op.a + row * VS, rv);
#pragma unroll
                for (int c = 0; c < VS; c++) out[c] = 0.0;
                deep_dinv_row<D>(L, row, rv, op.omega, out);
                st_vec<VS>(op.out + row * VS, out);
            }
            break;
        case DOP_SPMV:
            if (L.n * 4 <= (int64_t)G) deep_spmv<D, VT, 4>(L, op, gtid, G, dots);
            else deep_spmv<D, VT, 2>(L, op, gtid, G, dots);
            break;
        case DOP_RESTRICT: {                         // one warp per coarse row
            const LevelDev &C = T.lv[op.lvl + 1];
            for (int64_t I0 = 0; I0 < C.n_pad; I0 += GW) {
                const int64_t I = I0 + gw;
                if (I >= C.n_pad) continue;          // warp-uniform
                double s[VS];
#pragma unroll
                for (int a = 0; a < VS; a++) s[a] = 0.0;
                if (I < C.n) {
                    for (int64_t m = C.mem_ptr[I] + lane; m < C.mem_ptr[I + 1]; m += 32) {
                        const int64_t r = C.mem_idx[m];
                        double rv[VS];
                        ldcg_vec<VS>(op.a + r * VS, rv);
                        xfer_restrict(xfer_own<D>(L, r), rv, s);
                    }
#pragma unroll
                    for (int a = 0; a < D; a++) s[a] = warp_sum(s[a]);
                }
                if (lane == 0) st_vec<VS>(op.out + I * VS, s);
            }
            break;
        }
        case DOP_PROLONG:
        case DOP_PROLONGK: {
            const KScal &K = ks[op.lvl + 1];
            for (int64_t r = gtid; r < L.n; r += G) {
                const int64_t I = L.agg[r];
                double e[VS], e2[VS], xi[VS];
                ldcg_vec<VS>(op.a + I * VS, e);
                if (op.type == DOP_PROLONGK) {
                    ldcg_vec<VS>(op.b + I * VS, e2);
#pragma unroll
                    for (int c = 0; c < VS; c++) e[c] = fma(K.coef1, e[c], K.coef2 * e2[c]);
                    if (op.c) {
                        ldcg_vec<VS>(op.c + I * VS, e2);
#pragma unroll
                        for (int c = 0; c < VS; c++) e[c] = fma(K.coef3, e2[c], e[c]);
                    }
                }
                ldcg_vec<VS>(op.out + r * VS, xi);
                xfer_prolong(xfer_own<D>(L, r), e, xi);
                st_vec<VS>(op.out + r * VS, xi);
            }
            break;
        }
        case DOP_KRESID: {
            const KScal &K = ks[op.lvl];
            double *xa = const_cast<double *>(op.d);
            for (int64_t row = gtid; row < L.n_pad; row += G) {
                double a[VS], b[VS], out[VS];
                ldcg_vec<VS>(op.a + row * VS, a);
                ldcg_vec<VS>(op.b + row * VS, b);
                if (op.mode == 1) {
#pragma unroll
                    for (int c = 0; c < VS; c++) a[c] = fma(-K.alpha, b[c], a[c]);
                } else {
                    double wv[VS];
                    ldcg_vec<VS>(op.c + row * VS, wv);
#pragma unroll
                    for (int c = 0; c < VS; c++) a[c] = fma(K.e1, b[c], fma(-K.e2, wv[c], a[c]));
                }
#pragma unroll
                for (int c = 0; c < VS; c++) out[c] = 0.0;
                st_vec<VS>(op.out + row * VS, a);
                deep_dinv_row<D>(L, row, a, op.omega, out);
                st_vec<VS>(xa + row * VS, out);
            }
            break;
        }
        case DOP_KCOMBINE: {
            const KScal &K = ks[op.lvl];
            const int64_t nd = L.n_pad * VS;
            for (int64_t j = (int64_t)gtid * 2; j < nd; j += (int64_t)G * 2) {
                const double2 av = __ldcg(reinterpret_cast<const double2 *>(op.a + j)), bv = __ldcg(reinterpret_cast<const double2 *>(op.b + j));
                double2 o;
                o.x = fma(K.coef1, av.x, K.coef2 * bv.x); o.y = fma(K.coef1, av.y, K.coef2 * bv.y);
                if (op.mode == 2) {
                    const double2 cv = __ldcg(reinterpret_cast<const double2 *>(op.c + j));
                    o.x = fma(K.coef3, cv.x, o.x); o.y = fma(K.coef3, cv.y, o.y);
                }
                *reinterpret_cast<double2 *>(op.out + j) = o;
            }
            break;
        }
        case DOP_DENSE: {                            // out = Ainv rhs, one warp per scalar row, rhs staged in shared memory
            const int m = T.dense_m;
            for (int t = threadIdx.x; t < m; t += DEEP_NT) deep_sr[t] = __ldcg(op.a + (int64_t)(t / D) * VS + (t % D));
            __syncthreads();
            for (int srow = gw; srow < m; srow += GW) {
                double s = 0.0;
                const double *a = T.Ainv + (int64_t)srow * m;       // fp64: an fp32 copy of the explicit inverse of this ill-conditioned
                for (int j = lane; j < m; j += 32) s = fma(__ldg(a + j), deep_sr[j], s);   // matrix (1e7 anchor) tripled the PCG count
                s = warp_sum(s);
                if (lane == 0) op.out[(int64_t)(srow / D) * VS + (srow % D)] = s;
            }
            break;
        }
        }
        const bool has_dots = op.type == DOP_SPMV && op.fin != FIN_NONE;
        if (has_dots) {                              // this CTA's partial sums -> red[rpar]
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double w = warp_sum(dots[k]);
                if (lane == 0) wsum[threadIdx.x >> 5][k] = w;
            }
            __syncthreads();
            if (threadIdx.x < 4) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < DEEP_NT / 32; w++) s += wsum[w][threadIdx.x];
                red[rpar][threadIdx.x] = s;
            }
        }
        cluster.sync();                              // stage barrier (release / acquire at cluster scope, L1 invalidated)
        if (has_dots) {
            if (threadIdx.x == 0) {
                double t[4] = {0.0, 0.0, 0.0, 0.0};
                for (unsigned r = 0; r < NC; r++) {
                    const double *rr = cluster.map_shared_rank(&red[rpar][0], r);
#pragma unroll
                    for (int k = 0; k < 4; k++) t[k] += rr[k];
                }
                deep_finalize(op.fin, ks[op.lvl], t);
            }
            __syncthreads();
            rpar ^= 1;
        }
    }
    // the K-cycle scalars travel through S->k: the caller's prolongation (and a later segment, when the program is cut at the
    // dense stages) reads them there
    if (crank == 0 && threadIdx.x < MAX_LEVELS) T.S->k[threadIdx.x] = ks[threadIdx.x];
    cluster.sync();                                  // nobody leaves while a peer may still read its shared memory
}

} // namespace pgo
