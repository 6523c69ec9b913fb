// CUDA kernels (sm_100a, fp64) of the pose-graph-optimization hot path.
//
// All of this is HBM-bound 3x3 / 6x6 block work: no tensor cores (a 3x3 block product is not a
// dense contraction).  The design rules that matter: one thread per block row, one warp per
// 32-row slice streaming ONE contiguous blob of matrix data front to back with fully coalesced
// loads (sliced jagged storage, pgo_internal.h), 32-byte pose / vector records so a gather is
// exactly one sector, no atomics on the Gauss-Newton system (every block has a single writer),
// deterministic two-stage reductions ("last block finalises"), and no host in the PCG loop.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pgo_internal.h"

namespace pgo {

// ------------------------------------------------------------------------------------------------
struct Scalars {
    double rz, rz0, pq, alpha, beta, tol2;
    double norm2_dx, chi2;
    int iters, max_iters, done, status;
    unsigned counter[8];
};
enum { ST_OK = 0, ST_BREAKDOWN = 1, ST_MAXIT = 2 };
enum { FIN_NONE = 0, FIN_PQ = 1, FIN_RZ = 2, FIN_RZ_INIT = 3, FIN_NORM = 4, FIN_CHI2 = 5 };

struct LevelDev {
    int64_t n, n_pad, n_slices, n_slots;
    const int64_t *slice_ptr; const int32_t *deg; const uint32_t *col;
    double *val, *diag, *dinv;
    double *pos;                 // [2][n_pad] planes: position of each row (centroid on coarse levels)
    const uint8_t *vkind;        // level 0 only (nullptr on coarse levels)
    const int32_t *agg; const int64_t *ctgt; const int32_t *cstr;   // towards the coarser level
    const int64_t *mem_ptr; const int32_t *mem_idx;                 // members in the finer level
};

template <int D> struct VecStride { static constexpr int value = (D == 3) ? 4 : D; };

__device__ __forceinline__ int ld_done(const Scalars *S) { return *(const volatile int *)&S->done; }

template <int VS> __device__ __forceinline__ void ld_vec(const double *p, double *o) {
#pragma unroll
    for (int i = 0; i < VS; i += 2) { double2 t = *reinterpret_cast<const double2 *>(p + i); o[i] = t.x; o[i + 1] = t.y; }
}
template <int VS> __device__ __forceinline__ void st_vec(double *p, const double *v) {
#pragma unroll
    for (int i = 0; i < VS; i += 2) *reinterpret_cast<double2 *>(p + i) = make_double2(v[i], v[i + 1]);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-level sum of v; the block writes its partial, the LAST block to arrive sums all partials in a
// fixed order (deterministic) and returns true in thread 0 with `total` set.  counter wraps to 0.
template <int NT> __device__ bool block_sum_last(double v, double *partials, unsigned *counter, double &total) {
    __shared__ double sm[NT / 32];
    __shared__ bool is_last;
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) s += sm[w];
        partials[blockIdx.x] = s;
        __threadfence();
        unsigned t = atomicInc(counter, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double s = 0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += NT) s += __ldcg(partials + i);
    s = warp_sum(s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
#pragma unroll
        for (int w = 0; w < NT / 32; w++) t += sm[w];
        total = t;
        return true;
    }
    return false;
}

__device__ __forceinline__ void finalize(int FIN, Scalars *S, double total) {
    if (FIN == FIN_PQ) {
        S->pq = total;
        if (!(total > 0.0)) { S->status = ST_BREAKDOWN; S->done = 1; S->alpha = 0.0; }
        else S->alpha = S->rz / total;
    } else if (FIN == FIN_RZ) {
        double rz = S->rz;
        S->beta = total / rz;
        S->rz = total;
        int it = S->iters + 1;
        S->iters = it;
        if (!(total >= 0.0)) { S->status = ST_BREAKDOWN; S->done = 1; }
        else if (total <= S->tol2 * S->rz0) S->done = 1;
        else if (it >= S->max_iters) { S->status = ST_MAXIT; S->done = 1; }
    } else if (FIN == FIN_RZ_INIT) {
        S->rz = total; S->rz0 = total; S->beta = 0.0;
        if (!(total >= 0.0)) { S->status = ST_BREAKDOWN; S->done = 1; }
        else if (total == 0.0) S->done = 1;
    } else if (FIN == FIN_NORM) {
        S->norm2_dx = total;
    } else if (FIN == FIN_CHI2) {
        S->chi2 = total;
    }
    __threadfence();
}

// ------------------------------------------------------------------------------------------------
// BSR SpMV over the sliced jagged storage, one thread per block row.
//   MODE 0: y = H x                      (+ partial x.y  -> FIN_PQ)
//   MODE 1: y = r - H x                  (residual)
//   MODE 2: y = x + omega Dinv (r - H x) (damped block-Jacobi sweep; + partial r.y -> FIN_RZ*)
template <int D, int MODE, int FIN>
__global__ void __launch_bounds__(128) k_spmv(LevelDev L, const double *__restrict__ x, const double *__restrict__ r,
                                               double *__restrict__ y, double omega, Scalars *S, double *partials, int check_done) {
    if (check_done && ld_done(S)) return;
    constexpr int DD = D * D, VS = VecStride<D>::value;
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int64_t slice = row >> 5;
    double acc[D], xi[VS];
#pragma unroll
    for (int a = 0; a < D; a++) acc[a] = 0.0;
#pragma unroll
    for (int a = 0; a < VS; a++) xi[a] = 0.0;
    const bool live = slice < L.n_slices;
    if (live) {
        const int mydeg = L.deg[row];
        const int maxdeg = __shfl_sync(0xffffffffu, mydeg, 0);
        ld_vec<VS>(x + row * VS, xi);
        const double *dg = L.diag + row;
#pragma unroll
        for (int a = 0; a < D; a++)
#pragma unroll
            for (int b = 0; b < D; b++) acc[a] = fma(__ldg(dg + (int64_t)(a * D + b) * L.n_pad), xi[b], acc[a]);
        const int64_t base = L.slice_ptr[slice];
        int64_t off = 0;
        for (int k = 0; k < maxdeg; k++) {
            const bool active = k < mydeg;
            const int cnt = __popc(__ballot_sync(0xffffffffu, active));
            if (active) {
                const uint32_t c = __ldg(L.col + base + off + lane) & COL_MASK;
                double xj[VS];
                ld_vec<VS>(x + (int64_t)c * VS, xj);
                const double *v = L.val + (base + off) * DD + lane;
                double h[DD];
#pragma unroll
                for (int q = 0; q < DD; q++) h[q] = __ldg(v + (int64_t)q * cnt);
#pragma unroll
                for (int a = 0; a < D; a++)
#pragma unroll
                    for (int b = 0; b < D; b++) acc[a] = fma(h[a * D + b], xj[b], acc[a]);
            }
            off += cnt;
        }
    }
    double dot = 0.0;
    if (live) {
        double out[VS];
#pragma unroll
        for (int a = 0; a < VS; a++) out[a] = 0.0;
        if (MODE == 0) {
#pragma unroll
            for (int a = 0; a < D; a++) { out[a] = acc[a]; dot = fma(xi[a], acc[a], dot); }
        } else {
            double ri[VS];
            ld_vec<VS>(r + row * VS, ri);
            if (MODE == 1) {
#pragma unroll
                for (int a = 0; a < D; a++) out[a] = ri[a] - acc[a];
            } else {
                double t[D];
#pragma unroll
                for (int a = 0; a < D; a++) t[a] = ri[a] - acc[a];
                const double *di = L.dinv + row;
#pragma unroll
                for (int a = 0; a < D; a++) {
                    double s = 0.0;
#pragma unroll
                    for (int b = 0; b < D; b++) s = fma(__ldg(di + (int64_t)(a * D + b) * L.n_pad), t[b], s);
                    out[a] = fma(omega, s, xi[a]);
                    dot = fma(ri[a], out[a], dot);
                }
            }
        }
        st_vec<VS>(y + row * VS, out);
    }
    if (FIN != FIN_NONE) {
        double total;
        if (block_sum_last<128>(dot, partials, &S->counter[FIN], total)) finalize(FIN, S, total);
    }
}

// x = omega Dinv r  (first half of the V-cycle pre-smoothing; with FIN: block-Jacobi z = Dinv r and r.z)
template <int D, int FIN>
__global__ void __launch_bounds__(128) k_dinv_apply(LevelDev L, const double *__restrict__ r, double *__restrict__ x, double omega,
                                                     Scalars *S, double *partials, int check_done) {
    if (check_done && ld_done(S)) return;
    constexpr int VS = VecStride<D>::value;
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    double dot = 0.0;
    if (row < L.n_pad) {
        double ri[VS], out[VS];
        ld_vec<VS>(r + row * VS, ri);
#pragma unroll
        for (int a = 0; a < VS; a++) out[a] = 0.0;
        const double *di = L.dinv + row;
#pragma unroll
        for (int a = 0; a < D; a++) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < D; b++) s = fma(__ldg(di + (int64_t)(a * D + b) * L.n_pad), ri[b], s);
            out[a] = omega * s;
            dot = fma(ri[a], out[a], dot);
        }
        st_vec<VS>(x + row * VS, out);
    }
    if (FIN != FIN_NONE) {
        double total;
        if (block_sum_last<128>(dot, partials, &S->counter[FIN], total)) finalize(FIN, S, total);
    }
}

// x += alpha p ; r -= alpha q
template <int D>
__global__ void __launch_bounds__(256) k_update_xr(int64_t n_pad, double *__restrict__ x, double *__restrict__ r,
                                                    const double *__restrict__ p, const double *__restrict__ q, const Scalars *S) {
    if (ld_done(S)) return;
    constexpr int VS = VecStride<D>::value;
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 2;       // in doubles
    if (i >= n_pad * VS) return;
    const double a = S->alpha;
    double2 xv = *reinterpret_cast<double2 *>(x + i), rv = *reinterpret_cast<double2 *>(r + i);
    const double2 pv = *reinterpret_cast<const double2 *>(p + i), qv = *reinterpret_cast<const double2 *>(q + i);
    xv.x = fma(a, pv.x, xv.x); xv.y = fma(a, pv.y, xv.y);
    rv.x = fma(-a, qv.x, rv.x); rv.y = fma(-a, qv.y, rv.y);
    *reinterpret_cast<double2 *>(x + i) = xv;
    *reinterpret_cast<double2 *>(r + i) = rv;
}

// p = z + beta p
template <int D>
__global__ void __launch_bounds__(256) k_update_p(int64_t n_pad, double *__restrict__ p, const double *__restrict__ z, const Scalars *S) {
    if (ld_done(S)) return;
    constexpr int VS = VecStride<D>::value;
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 2;
    if (i >= n_pad * VS) return;
    const double b = S->beta;
    double2 pv = *reinterpret_cast<double2 *>(p + i);
    const double2 zv = *reinterpret_cast<const double2 *>(z + i);
    pv.x = fma(b, pv.x, zv.x); pv.y = fma(b, pv.y, zv.y);
    *reinterpret_cast<double2 *>(p + i) = pv;
}

__global__ void k_fill(double *p, int64_t n, double v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// Aggregation AMG transfer operators (D = 3).  The coarse unknown of an aggregate is a rigid motion
// (tx, ty, theta) of its members about the aggregate centroid c:  P_i = [[1,0,-dy],[0,1,dx],[0,0,pz]],
// d = pos_i - c ; pz = 0 for landmark rows (their third, padding, unknown stays decoupled).
// Global rigid motions -- the near-null space of H that the 1e7 anchor barely pins -- are
// represented exactly on every level.
__device__ __forceinline__ double row_pz(const LevelDev &L, int64_t row) { return (L.vkind && L.vkind[row] == 1) ? 0.0 : 1.0; }

// rc_I = sum_{i in I} P_i^T res_i   (one thread per coarse row; deterministic)
__global__ void __launch_bounds__(128) k_restrict3(LevelDev F, LevelDev C, const double *__restrict__ res, double *__restrict__ rc,
                                                    const Scalars *S) {
    if (ld_done(S)) return;
    const int64_t I = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (I >= C.n_pad) return;
    double s0 = 0, s1 = 0, s2 = 0;
    if (I < C.n) {
        const double cx = C.pos[I], cy = C.pos[C.n_pad + I];
        for (int64_t m = C.mem_ptr[I]; m < C.mem_ptr[I + 1]; m++) {
            const int64_t i = C.mem_idx[m];
            double r[4];
            ld_vec<4>(res + i * 4, r);
            const double dx = F.pos[i] - cx, dy = F.pos[F.n_pad + i] - cy;
            s0 += r[0]; s1 += r[1];
            s2 += fma(-dy, r[0], fma(dx, r[1], row_pz(F, i) * r[2]));
        }
    }
    double out[4] = {s0, s1, s2, 0.0};
    st_vec<4>(rc + I * 4, out);
}

// x_i += P_i e_{agg(i)}
__global__ void __launch_bounds__(128) k_prolong3(LevelDev F, LevelDev C, const double *__restrict__ ec, double *__restrict__ x,
                                                   const Scalars *S) {
    if (ld_done(S)) return;
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= F.n) return;
    const int64_t I = F.agg[i];
    double e[4], xi[4];
    ld_vec<4>(ec + I * 4, e);
    ld_vec<4>(x + i * 4, xi);
    const double dx = F.pos[i] - C.pos[I], dy = F.pos[F.n_pad + i] - C.pos[C.n_pad + I];
    xi[0] += fma(-dy, e[2], e[0]);
    xi[1] += fma(dx, e[2], e[1]);
    xi[2] += row_pz(F, i) * e[2];
    st_vec<4>(x + i * 4, xi);
}

// centroid of the members
__global__ void __launch_bounds__(128) k_coarse_pos(LevelDev F, LevelDev C) {
    const int64_t I = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (I >= C.n_pad) return;
    double sx = 0, sy = 0;
    if (I < C.n) {
        const int64_t b = C.mem_ptr[I], e = C.mem_ptr[I + 1];
        for (int64_t m = b; m < e; m++) { const int64_t i = C.mem_idx[m]; sx += F.pos[i]; sy += F.pos[F.n_pad + i]; }
        sx /= (double)(e - b); sy /= (double)(e - b);
    }
    C.pos[I] = sx; C.pos[C.n_pad + I] = sy;
}

// G = P_i^T H P_j for 3x3 blocks
__device__ __forceinline__ void ptap3(const double *h, double dxi, double dyi, double pzi, double dxj, double dyj, double pzj, double *g) {
    double m[9];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        m[3 * a + 0] = h[3 * a + 0];
        m[3 * a + 1] = h[3 * a + 1];
        m[3 * a + 2] = fma(-dyj, h[3 * a + 0], fma(dxj, h[3 * a + 1], pzj * h[3 * a + 2]));
    }
#pragma unroll
    for (int b = 0; b < 3; b++) {
        g[b] = m[b];
        g[3 + b] = m[3 + b];
        g[6 + b] = fma(-dyi, m[b], fma(dxi, m[3 + b], pzi * m[6 + b]));
    }
}

// Galerkin product Hc = P^T H P: every fine block adds P_i^T H_ij P_j into the coarse block of
// (agg i, agg j).  Several fine blocks share a coarse block, hence atomics (coarse levels only;
// the Gauss-Newton system itself is assembled without atomics).
__global__ void __launch_bounds__(128) k_galerkin3(LevelDev F, LevelDev C) {
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int64_t slice = row >> 5;
    if (slice >= F.n_slices) return;
    const int mydeg = F.deg[row];
    const int maxdeg = __shfl_sync(0xffffffffu, mydeg, 0);
    const bool real = row < F.n;
    int64_t I = 0; double dxi = 0, dyi = 0, pzi = 1;
    if (real) {
        I = F.agg[row];
        dxi = F.pos[row] - C.pos[I]; dyi = F.pos[F.n_pad + row] - C.pos[C.n_pad + I];
        pzi = row_pz(F, row);
        double h[9], g[9];
#pragma unroll
        for (int q = 0; q < 9; q++) h[q] = F.diag[(int64_t)q * F.n_pad + row];
        ptap3(h, dxi, dyi, pzi, dxi, dyi, pzi, g);
#pragma unroll
        for (int q = 0; q < 9; q++) atomicAdd(C.diag + (int64_t)q * C.n_pad + I, g[q]);
    }
    const int64_t base = F.slice_ptr[slice];
    int64_t off = 0;
    for (int k = 0; k < maxdeg; k++) {
        const bool active = k < mydeg;
        const int cnt = __popc(__ballot_sync(0xffffffffu, active));
        if (active) {
            const int64_t slot = base + off + lane;
            const int64_t j = F.col[slot] & COL_MASK;
            const int64_t J = F.agg[j];
            const double dxj = F.pos[j] - C.pos[J], dyj = F.pos[F.n_pad + j] - C.pos[C.n_pad + J];
            const double *v = F.val + (base + off) * 9 + lane;
            double h[9], g[9];
#pragma unroll
            for (int q = 0; q < 9; q++) h[q] = v[(int64_t)q * cnt];
            ptap3(h, dxi, dyi, pzi, dxj, dyj, row_pz(F, j), g);
            const int64_t t = F.ctgt[slot];
            double *dst; int64_t str;
            if (t & CTGT_DIAG) { dst = C.diag + (t & ~CTGT_DIAG); str = C.n_pad; }
            else { dst = C.val + t; str = F.cstr[slot]; }
#pragma unroll
            for (int q = 0; q < 9; q++) atomicAdd(dst + (int64_t)q * str, g[q]);
        }
        off += cnt;
    }
}

// 3x3 inverse (cofactors)
__device__ __forceinline__ void inv3(const double *a, double *o) {
    const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    const double id = 1.0 / det;
    o[0] = c00 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = c01 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = c02 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

__global__ void __launch_bounds__(128) k_invert_diag3(LevelDev L) {
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (row >= L.n_pad) return;
    double a[9], o[9];
#pragma unroll
    for (int q = 0; q < 9; q++) a[q] = L.diag[(int64_t)q * L.n_pad + row];
    if (row < L.n) {
        // an aggregate made of landmarks that all sit on its centroid (e.g. a single landmark) has no
        // rotational unknown: P^T H P is exactly singular in theta.  Decouple that unknown (its restricted
        // residual is always 0) so that the level stays SPD.
        if (!(a[8] > 1e-14 * (a[0] + a[4]))) {
            a[2] = a[5] = a[6] = a[7] = 0.0; a[8] = 1.0;
            L.diag[(int64_t)2 * L.n_pad + row] = 0.0; L.diag[(int64_t)5 * L.n_pad + row] = 0.0;
            L.diag[(int64_t)6 * L.n_pad + row] = 0.0; L.diag[(int64_t)7 * L.n_pad + row] = 0.0;
            L.diag[(int64_t)8 * L.n_pad + row] = 1.0;
        }
        inv3(a, o);
    } else {
#pragma unroll
        for (int q = 0; q < 9; q++) o[q] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < 9; q++) L.dinv[(int64_t)q * L.n_pad + row] = o[q];
}

// Coarsest level: explicit dense inverse (m = D n <= 150 unknowns) by Gauss-Jordan in shared memory.
// H_c is SPD, so no pivoting.  One CTA.
template <int D>
__global__ void __launch_bounds__(256) k_dense_invert(LevelDev L, double *__restrict__ Ainv) {
    extern __shared__ double sA[];
    constexpr int DD = D * D;
    const int m = (int)L.n * D;
    for (int i = threadIdx.x; i < m * m; i += 256) sA[i] = 0.0;
    __syncthreads();
    for (int row = threadIdx.x; row < (int)L.n; row += 256) {
        for (int a = 0; a < D; a++)
            for (int b = 0; b < D; b++) sA[(row * D + a) * m + row * D + b] = L.diag[(int64_t)(a * D + b) * L.n_pad + row];
        // walk this row's slots
        const int slice = row >> 5, lane = row & 31;
        const int64_t base = L.slice_ptr[slice];
        int64_t off = 0;
        const int mydeg = L.deg[row];
        for (int k = 0; k < mydeg; k++) {
            int cnt = 0;
            while (cnt < 32 && L.deg[slice * 32 + cnt] > k) cnt++;
            const int j = (int)(L.col[base + off + lane] & COL_MASK);
            const double *v = L.val + (base + off) * DD + lane;
            for (int a = 0; a < D; a++)
                for (int b = 0; b < D; b++) sA[(row * D + a) * m + j * D + b] += v[(int64_t)(a * D + b) * cnt];
            off += cnt;
        }
    }
    __syncthreads();
    // in-place Gauss-Jordan: for pivot p, A[i][j] -= A[i][p] A[p][j] / A[p][p] (i,j != p),
    // A[i][p] = -A[i][p] / A[p][p], A[p][j] /= A[p][p], A[p][p] = 1 / A[p][p]
    __shared__ double colp[256];
    for (int p = 0; p < m; p++) {
        if ((int)threadIdx.x < m) colp[threadIdx.x] = sA[threadIdx.x * m + p];
        __syncthreads();
        const double ip = 1.0 / colp[p];
        for (int idx = threadIdx.x; idx < m * m; idx += 256) {
            const int i = idx / m, j = idx - i * m;
            if (i != p && j != p) sA[idx] = fma(-colp[i] * ip, sA[p * m + j], sA[idx]);
        }
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += 256) {
            if (j == p) continue;
            sA[p * m + j] *= ip;
            sA[j * m + p] = -colp[j] * ip;
        }
        if (threadIdx.x == 0) sA[p * m + p] = ip;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < m * m; i += 256) Ainv[i] = sA[i];
}

// x = Ainv r on the coarsest level (Ainv symmetric: column reads are coalesced)
template <int D>
__global__ void __launch_bounds__(256) k_dense_apply(LevelDev L, const double *__restrict__ Ainv, const double *__restrict__ r,
                                                      double *__restrict__ x, const Scalars *S) {
    if (ld_done(S)) return;
    constexpr int VS = VecStride<D>::value;
    __shared__ double sr[256];
    const int m = (int)L.n * D;
    const int t = threadIdx.x;
    if (t < m) sr[t] = r[(t / D) * VS + (t % D)];
    __syncthreads();
    if (t < m) {
        double s = 0.0;
        for (int j = 0; j < m; j++) s = fma(__ldg(Ainv + (int64_t)j * m + t), sr[j], s);
        x[(t / D) * VS + (t % D)] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// SE(2) linearisation (restated from pose_graph_optimization.rs:434-486, 516-535; closed forms in
// SURVEY.md Appendix A).  Poses are (x, y, cos, sin): the reference stores the heading as a unit
// complex (g2o.rs:14-16), so no sincos per edge, one atan2.
struct PP { double e[3]; double m11, m12, a0, a1; };

__device__ __forceinline__ void pose_pose(const double *x1, const double *x2, const double *z, PP &o) {
    const double c1 = x1[2], s1 = x1[3], cz = z[2], sz = z[3];
    const double dx = x2[0] - x1[0], dy = x2[1] - x1[1];
    // u = R1^T d ; e_t = Rz^T (u - tz)
    const double u0 = c1 * dx + s1 * dy, u1 = -s1 * dx + c1 * dy;
    const double w0 = u0 - z[0], w1 = u1 - z[1];
    o.e[0] = cz * w0 + sz * w1;
    o.e[1] = -sz * w0 + cz * w1;
    // rotation part: conj(rz) conj(r1) r2
    const double re12 = c1 * x2[2] + s1 * x2[3], im12 = c1 * x2[3] - s1 * x2[2];
    o.e[2] = atan2(cz * im12 - sz * re12, cz * re12 + sz * im12);
    // M = Rz^T R1^T = [[m11, m12], [-m12, m11]]
    o.m11 = cz * c1 - sz * s1; o.m12 = cz * s1 + sz * c1;
    // a = Rz^T (D R1)^T d ,  (D R1)^T = [[-s1, c1], [-c1, -s1]]
    const double v0 = -s1 * dx + c1 * dy, v1 = -c1 * dx - s1 * dy;
    o.a0 = cz * v0 + sz * v1; o.a1 = -sz * v0 + cz * v1;
}

// C = X^T W Y for 3x3 row-major X, Y and symmetric W given as upper triangle (w00 w01 w02 w11 w12 w22)
__device__ __forceinline__ void xtwy3(const double *X, const double *w, const double *Y, double *C) {
    double WY[9];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        WY[c] = w[0] * Y[c] + w[1] * Y[3 + c] + w[2] * Y[6 + c];
        WY[3 + c] = w[1] * Y[c] + w[3] * Y[3 + c] + w[4] * Y[6 + c];
        WY[6 + c] = w[2] * Y[c] + w[4] * Y[3 + c] + w[5] * Y[6 + c];
    }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) C[3 * r + c] = X[r] * WY[c] + X[3 + r] * WY[3 + c] + X[6 + r] * WY[6 + c];
}

// Fused linearise + assemble + block-Jacobi setup.  One thread per block row walks the row's half
// edges (sorted by destination block row = segmented accumulation with a single writer per block):
// recomputes e, A, B of each incident edge from the two 32-byte pose records, accumulates the
// diagonal block and gradient in registers, streams the off-diagonal block into the slice blob, and
// finally writes diag, its inverse (the block-Jacobi preconditioner), r = b = -g and the row position.
// hz: measurement stream laid out like val with 10 components (z: x y cos sin ; Omega upper 6).
__global__ void __launch_bounds__(128) k_assemble_se2(LevelDev L, const double *__restrict__ poses, const double *__restrict__ hz,
                                                       double *__restrict__ rvec, int64_t anchor_row, double anchor_w, double lambda) {
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int64_t slice = row >> 5;
    if (slice >= L.n_slices) return;
    const int mydeg = L.deg[row];
    const int maxdeg = __shfl_sync(0xffffffffu, mydeg, 0);
    double xi[4];
    ld_vec<4>(poses + row * 4, xi);
    double Hd[9], g[3];
#pragma unroll
    for (int q = 0; q < 9; q++) Hd[q] = 0.0;
    g[0] = g[1] = g[2] = 0.0;
    const int64_t base = L.slice_ptr[slice];
    int64_t off = 0;
    for (int k = 0; k < maxdeg; k++) {
        const bool active = k < mydeg;
        const int cnt = __popc(__ballot_sync(0xffffffffu, active));
        if (active) {
            const uint32_t cw = __ldg(L.col + base + off + lane);
            const bool to_side = (cw & COL_ROLE_TO) != 0;
            double xj[4], z[4], w[6];
            ld_vec<4>(poses + (int64_t)(cw & COL_MASK) * 4, xj);
            const double *m = hz + (base + off) * 10 + lane;
#pragma unroll
            for (int q = 0; q < 4; q++) z[q] = __ldg(m + (int64_t)q * cnt);
#pragma unroll
            for (int q = 0; q < 6; q++) w[q] = __ldg(m + (int64_t)(4 + q) * cnt);
            const double *x1 = to_side ? xj : xi, *x2 = to_side ? xi : xj;
            double A[9], B[9], e[3], T[9], Own[9], gi[3];
            if (!(cw & COL_EDGE_XY)) {
                PP p;
                pose_pose(x1, x2, z, p);
                e[0] = p.e[0]; e[1] = p.e[1]; e[2] = p.e[2];
                A[0] = -p.m11; A[1] = -p.m12; A[2] = p.a0;
                A[3] = p.m12;  A[4] = -p.m11; A[5] = p.a1;
                A[6] = 0.0;    A[7] = 0.0;    A[8] = -1.0;
                B[0] = p.m11;  B[1] = p.m12;  B[2] = 0.0;
                B[3] = -p.m12; B[4] = p.m11;  B[5] = 0.0;
                B[6] = 0.0;    B[7] = 0.0;    B[8] = 1.0;
            } else {
                // pose-landmark (:449-455, 516-535), x1 = pose, x2 = landmark; third error row is padding.
                const double c = x1[2], s = x1[3], dx = x2[0] - x1[0], dy = x2[1] - x1[1];
                e[0] = (c * dx + s * dy) - z[0]; e[1] = (-s * dx + c * dy) - z[1]; e[2] = 0.0;
                A[0] = -c; A[1] = -s; A[2] = -s * dx + c * dy;
                A[3] = s;  A[4] = -c; A[5] = -c * dx - s * dy;
                A[6] = A[7] = A[8] = 0.0;
                B[0] = c;  B[1] = s;  B[2] = 0.0;
                B[3] = -s; B[4] = c;  B[5] = 0.0;
                B[6] = B[7] = B[8] = 0.0;
                // Omega is 2x2: stored as (w11 w12 w22 . . .) -> expand to the 3x3 upper triangle
                const double w11 = w[0], w12 = w[1], w22 = w[2];
                w[0] = w11; w[1] = w12; w[2] = 0.0; w[3] = w22; w[4] = 0.0; w[5] = 0.0;
            }
            xtwy3(A, w, B, T);                       // H_ij = A^T W B ; H_ji is its transpose (:176-177)
            const double *J = to_side ? B : A;
            xtwy3(J, w, J, Own);
            const double we0 = w[0] * e[0] + w[1] * e[1] + w[2] * e[2];
            const double we1 = w[1] * e[0] + w[3] * e[1] + w[4] * e[2];
            const double we2 = w[2] * e[0] + w[4] * e[1] + w[5] * e[2];
#pragma unroll
            for (int r = 0; r < 3; r++) gi[r] = J[r] * we0 + J[3 + r] * we1 + J[6 + r] * we2;
#pragma unroll
            for (int q = 0; q < 9; q++) Hd[q] += Own[q];
            g[0] += gi[0]; g[1] += gi[1]; g[2] += gi[2];
            double *v = L.val + (base + off) * 9 + lane;
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) v[(int64_t)(3 * r + c) * cnt] = to_side ? T[3 * c + r] : T[3 * r + c];
        }
        off += cnt;
    }
    const bool real = row < L.n;
    if (real) {
        if (row == anchor_row) { Hd[0] += anchor_w; Hd[4] += anchor_w; Hd[8] += anchor_w; }   // :330-336
        if (L.vkind[row] == 1) Hd[8] = 1.0;                                                     // landmark padding unknown
        Hd[0] += lambda; Hd[4] += lambda; Hd[8] += lambda;                                      // LM, :362-366
    }
    double Di[9];
    if (real) inv3(Hd, Di);
    else {
#pragma unroll
        for (int q = 0; q < 9; q++) Di[q] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < 9; q++) {
        L.diag[(int64_t)q * L.n_pad + row] = Hd[q];
        L.dinv[(int64_t)q * L.n_pad + row] = Di[q];
    }
    double out[4] = {-g[0], -g[1], -g[2], 0.0};                                                 // b = -g (:361)
    st_vec<4>(rvec + row * 4, out);
    L.pos[row] = xi[0]; L.pos[L.n_pad + row] = xi[1];
}

// global_error (:537-574): one thread per edge, edge-ordered SoA copy of the measurements.
// ed: [10][n_edges] planes (z: x y cos sin ; Omega upper 6 -- for XY edges w11 w12 w22 in the first three)
__global__ void __launch_bounds__(256) k_chi2_se2(int64_t n_edges, const int2 *__restrict__ ends, const double *__restrict__ ed,
                                                   const double *__restrict__ poses, Scalars *S, double *partials) {
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double c = 0.0;
    if (k < n_edges) {
        const int2 en = ends[k];
        double x1[4], x2[4], z[4], w[6];
        const bool xy = en.y < 0;
        ld_vec<4>(poses + (int64_t)en.x * 4, x1);
        ld_vec<4>(poses + (int64_t)(xy ? ~en.y : en.y) * 4, x2);
#pragma unroll
        for (int q = 0; q < 4; q++) z[q] = __ldg(ed + (int64_t)q * n_edges + k);
#pragma unroll
        for (int q = 0; q < 6; q++) w[q] = __ldg(ed + (int64_t)(4 + q) * n_edges + k);
        if (!xy) {
            PP p;
            pose_pose(x1, x2, z, p);
            const double e0 = p.e[0], e1 = p.e[1], e2 = p.e[2];
            c = e0 * (w[0] * e0 + w[1] * e1 + w[2] * e2) + e1 * (w[1] * e0 + w[3] * e1 + w[4] * e2) + e2 * (w[2] * e0 + w[4] * e1 + w[5] * e2);
        } else {
            const double cs = x1[2], sn = x1[3], dx = x2[0] - x1[0], dy = x2[1] - x1[1];
            const double e0 = (cs * dx + sn * dy) - z[0], e1 = (-sn * dx + cs * dy) - z[1];
            c = e0 * (w[0] * e0 + w[1] * e1) + e1 * (w[1] * e0 + w[2] * e1);
        }
    }
    double total;
    if (block_sum_last<256>(c, partials, &S->counter[FIN_CHI2], total)) finalize(FIN_CHI2, S, total);
}

// update_nodes (:229-245): t += dx.xy (global frame), r <- r * (cos dth, sin dth) without
// renormalisation; landmarks l += dx.  Also ||dx||^2 (:273).  sign = -1 undoes a step (:277).
__global__ void __launch_bounds__(256) k_retract_se2(LevelDev L, double *__restrict__ poses, const double *__restrict__ dx, double sign,
                                                      Scalars *S, double *partials) {
    const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double n2 = 0.0;
    if (row < L.n) {
        double p[4], d[4];
        ld_vec<4>(poses + row * 4, p);
        ld_vec<4>(dx + row * 4, d);
        p[0] = fma(sign, d[0], p[0]); p[1] = fma(sign, d[1], p[1]);
        n2 = d[0] * d[0] + d[1] * d[1];
        if (L.vkind[row] == 0) {
            double sn, cs;
            sincos(sign * d[2], &sn, &cs);
            const double re = p[2] * cs - p[3] * sn, im = p[2] * sn + p[3] * cs;
            p[2] = re; p[3] = im;
            n2 = fma(d[2], d[2], n2);
        }
        st_vec<4>(poses + row * 4, p);
    }
    double total;
    if (block_sum_last<256>(n2, partials, &S->counter[FIN_NORM], total)) finalize(FIN_NORM, S, total);
}

// pgo_set_poses / pgo_get_poses: g2o-layout vertex values (x y theta | x y, lut order, packed) <-> the
// 32-byte device pose records (x, y, cos, sin) in storage order.  iso2 (g2o.rs:14-16) on the way in,
// atan2(im, re) on the way out.
__global__ void __launch_bounds__(256) k_import_poses(int64_t n, const int64_t *__restrict__ row_valofs, const uint8_t *__restrict__ vkind,
                                                       const double *__restrict__ values, double *__restrict__ poses) {
    const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= n) return;
    const double *v = values + row_valofs[row];
    double p[4] = {v[0], v[1], 1.0, 0.0};
    if (vkind[row] == 0) sincos(v[2], &p[3], &p[2]);
    st_vec<4>(poses + row * 4, p);
}
__global__ void __launch_bounds__(256) k_export_poses(int64_t n, const int64_t *__restrict__ row_valofs, const uint8_t *__restrict__ vkind,
                                                       const double *__restrict__ poses, double *__restrict__ values) {
    const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= n) return;
    double p[4];
    ld_vec<4>(poses + row * 4, p);
    double *v = values + row_valofs[row];
    v[0] = p[0]; v[1] = p[1];
    if (vkind[row] == 0) v[2] = atan2(p[3], p[2]);
}

} // namespace pgo
