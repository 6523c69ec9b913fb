// Persistent coarse-tail kernel: the whole K-cycle below the fine level(s) in ONE cooperative launch.
//
// The coarse levels of the hierarchy (62k / 3.7k / 0.4k block rows at BASELINE configs[3]) are L2-resident and
// latency-bound: as separate kernels one PCG iteration needs ~60 launches for them, 5-6 us each, i.e. ~40 % of the
// Gauss-Newton step at 1 GPU and most of it at 8.  Here the host records, once, the linear "stage program" of one
// coarse solve (the recursion cycle / coarse_solve of pgo_b200.cu unrolled: ~100 stages of the six kinds below);
// k_tail runs it with one grid-wide barrier (cooperative groups) between stages.  Every stage executes the same
// __device__ body as the stand-alone kernel it replaces, over "virtual blocks" handed out round-robin to the resident
// CTAs, so results are bit-identical to the unfused path (same per-block partial sums, same last-block finalisation).
//
// STATUS: correct (the whole GPU parity suite passes with it) but NOT the default -- measured on B200 at configs[3] it
// loses to the stage-per-kernel CUDA graph: 1030 vs 954 us per PCG iteration (profiles/r01j_tail_experiment.log).  A
// stage is a ~3 us chain of dependent L2 loads in both forms; grid.sync() costs about what a graph kernel boundary
// costs; and the persistent CTAs (80 registers, 3 per SM) keep fewer warps in flight for the 62k-row level than the
// stand-alone launches (8 CTAs per SM).  Enable with PGO_TAIL=1; kept as the starting point for a cluster-scoped variant.
#pragma once
#include "kernels.cuh"

namespace pgo {

enum { TOP_DINV = 0, TOP_SPMV = 1, TOP_RESTRICT = 2, TOP_PROLONG = 3, TOP_KCOMBINE = 4, TOP_DENSE = 5 };

struct TailOp {
    int type, lvl, mode, fin;       // SPMV: mode 0/1/2, fin FIN_NONE/FIN_K1/FIN_K2 ; KCOMBINE: mode = which
    int lpr, nvb;                   // SPMV: lanes per row (8 / 32) ; number of virtual blocks (of 256 threads) of the stage
    const double *a, *b, *c, *d;    // DINV: a = rhs ; SPMV: a = x, b = r, c = u1, d = u2 ; RESTRICT: a = res ; PROLONG: a = ec ;
                                    // KCOMBINE: a, b ; DENSE: a = rhs
    double *out;
    double omega;
};

struct TailCtx {
    const TailOp *prog; int n_ops;
    const LevelDev *lv;             // [n_levels] device copies of the level descriptors
    DenseMap dmap; int dense_m; const double *Ainv;
    Scalars *S; double *partials;
};

template <int D, int MODE, int LPR>
__device__ __forceinline__ void tail_spmv(const TailOp &op, const LevelDev &L, const XRef &xr, Scalars *S, double *partials, unsigned vb) {
    if (op.fin == FIN_K1) spmv_csr_body<D, MODE, FIN_K1, false, LPR>(L, xr, op.a, op.b, op.out, op.omega, op.c, op.d, S, partials, op.lvl, vb, op.nvb);
    else if (op.fin == FIN_K2) spmv_csr_body<D, MODE, FIN_K2, false, LPR>(L, xr, op.a, op.b, op.out, op.omega, op.c, op.d, S, partials, op.lvl, vb, op.nvb);
    else spmv_csr_body<D, MODE, FIN_NONE, false, LPR>(L, xr, op.a, op.b, op.out, op.omega, op.c, op.d, S, partials, op.lvl, vb, op.nvb);
}

template <int D>
__global__ void __launch_bounds__(256, 3) k_tail(const __grid_constant__ TailCtx T) {
    PDL_ENTER();
    if (ld_done(T.S)) return;                       // set only by fine-level kernels, so every CTA sees the same value
    cg::grid_group grid = cg::this_grid();
    XRef none{};
    for (int i = 0; i < T.n_ops; i++) {
        const TailOp op = T.prog[i];
        const LevelDev &L = T.lv[op.lvl];
        for (unsigned vb = blockIdx.x; vb < (unsigned)op.nvb; vb += gridDim.x) {
            __syncthreads();                        // the static shared scratch of the previous virtual block is free again
            switch (op.type) {
            case TOP_DINV: dinv_apply_body<D, FIN_NONE, 256>(L, op.a, op.out, op.omega, nullptr, T.S, T.partials, vb, op.nvb); break;
            case TOP_SPMV:
                if (op.lpr == 8) {
                    if (op.mode == 0) tail_spmv<D, 0, 8>(op, L, none, T.S, T.partials, vb);
                    else if (op.mode == 1) tail_spmv<D, 1, 8>(op, L, none, T.S, T.partials, vb);
                    else tail_spmv<D, 2, 8>(op, L, none, T.S, T.partials, vb);
                } else {
                    if (op.mode == 0) tail_spmv<D, 0, 32>(op, L, none, T.S, T.partials, vb);
                    else if (op.mode == 1) tail_spmv<D, 1, 32>(op, L, none, T.S, T.partials, vb);
                    else tail_spmv<D, 2, 32>(op, L, none, T.S, T.partials, vb);
                }
                break;
            case TOP_RESTRICT: restrict_body<D>(L, T.lv[op.lvl + 1], op.a, op.out, vb); break;
            case TOP_PROLONG: prolong_body<D, 256>(L, op.a, op.out, vb); break;
            case TOP_KCOMBINE:
                if (op.mode == 0) kcombine_body<0>(L.n_pad * VecStride<D>::value, op.a, op.b, op.out, T.S, op.lvl, vb);
                else kcombine_body<1>(L.n_pad * VecStride<D>::value, op.a, op.b, op.out, T.S, op.lvl, vb);
                break;
            case TOP_DENSE: {
                XRef rr{};
                rr.p[0] = op.a;
                dense_apply_body<D>(L.n, T.dmap, 0, 1, T.dense_m, T.Ainv, rr, op.out, vb);
                break;
            }
            }
        }
        if (i + 1 < T.n_ops) grid.sync();
    }
}

} // namespace pgo
