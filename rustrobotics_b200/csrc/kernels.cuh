// CUDA kernels (sm_100a, fp64) of the pose-graph-optimization hot path.
//
// All of this is HBM-bound 3x3 block work: no tensor cores (a 3x3 block product is not a dense contraction).
// Design rules: level 0 (the Gauss-Newton system, far larger than L2) is streamed by one thread per block row, one
// warp per 32-row slice reading ONE contiguous blob front to back with coalesced loads (sliced jagged storage,
// pgo_internal.h); coarse AMG levels (L2-resident, latency-bound) use one warp per block row over block CSR; 32-byte
// pose / vector records so a gather is exactly one sector; neighbour rows are addressed as (owner rank, local row)
// through a table of peer pointers (XRef), so that the same kernels run sharded over NVLink peer memory; no atomics on
// the Gauss-Newton system (every block has a single writer); deterministic two-stage reductions ("last block
// finalises") with all PCG / K-cycle scalars resident on the device; no host in the PCG loop.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pgo_internal.h"

namespace pgo {

// ------------------------------------------------------------------------------------------------
constexpr int MAX_LEVELS = 12;
struct KScal { double alpha, coef1, coef2, coef3, rho1, a1, rho2, gam21, alpha2, e1, e2; int steps, pad_; };
// Cross-rank synchronisation block (sharded handles): one per rank at the start of its peer-visible arena, see xrank_exchange
struct Comm {
    unsigned long long flags[MAX_RANKS];
    double vals[2][MAX_RANKS][4];
};
struct CommRef { Comm *p[MAX_RANKS]; };

struct Scalars {
    double rz, rz0, pq, alpha, beta, tol2;
    double norm2_dx, chi2;
    int iters, max_iters, done, status;
    unsigned counter[12];           // everything from here on survives the per-solve reset
    int world, repl_from;           // sharded handles: levels >= repl_from are replicated (their sums are already global)
    unsigned long long epoch;       // cross-rank barrier epoch (peer.cuh)
    KScal k[MAX_LEVELS];
    double loc[4];                  // (unused since the cross-rank reduction is fused into the producing kernel)
    Comm *comm_mine; Comm *comm_peer[MAX_RANKS]; int rank, pad2_;   // sharded handles: where the last block of a reducing kernel meets its peers
    long long spin_limit;           // clock cycles a rank waits for its peers before it gives up with ST_COMM (~20 s; PGO_COMM_TIMEOUT_S)
};
enum { ST_OK = 0, ST_BREAKDOWN = 1, ST_MAXIT = 2, ST_COMM = 3 };
enum { FIN_NONE = 0, FIN_PQ = 1, FIN_RZ = 2, FIN_RZ_INIT = 3, FIN_NORM = 4, FIN_CHI2 = 5, FIN_K1 = 6, FIN_K2 = 7, FIN_K3 = 8 };
__host__ __device__ constexpr int fin_ndot(int FIN) { return FIN == FIN_K3 ? 4 : FIN == FIN_K2 ? 3 : (FIN == FIN_RZ || FIN == FIN_K1) ? 2 : FIN == FIN_NONE ? 0 : 1; }

// one vector (or pose / lever-arm array) as seen from this rank: p[k] = base of rank k's segment
struct XRef { const double *p[MAX_RANKS]; };
template <int STRIDE> __device__ __forceinline__ const double *xgather(const XRef &x, uint32_t colword) {
    return x.p[(colword >> COL_OWNER_SHIFT) & (MAX_RANKS - 1)] + (int64_t)(colword & COL_LOCAL_MASK) * STRIDE;
}

// element ranges (rows or stored blocks) produced by each rank on the first replicated level
struct SegMap { int64_t off[MAX_RANKS + 1]; };

struct LevelDev {
    int64_t n, n_pad, n_slices, n_slots;      // local to this rank
    const int64_t *slice_ptr;                 // JDS: [n_slices + 1] ; CSR: row_ptr [n_pad + 1]
    const int32_t *deg; const uint32_t *col;
    double *val, *diag, *dinv;
    float *diagf, *dinvf;        // fp32 copies of diag / dinv (sliced levels only), same use as valf
    float *valf;                 // fp32 copy of val (same layout) read by the SpMVs INSIDE the multigrid cycle when the
                                 // preconditioner is stored in reduced precision (the PCG operator itself stays fp64)
    double *pos;                 // [2][n_pad] planes: position of each row (centroid on coarse levels)
    double *lev;                 // [n_pad][2]: lever arm of each row about its aggregate's centroid (peer-visible)
    const uint8_t *vkind;        // level 0 only (nullptr on coarse levels)
    const double *quat;          // level 0 of an SE3 graph: the (halo-extended) pose records, 8 doubles per row, unit
                                 // quaternion (w,x,y,z) at +4 (the rotation unknown is body-frame); nullptr otherwise
    const int32_t *agg; const int32_t *ctgt; const int32_t *cstr;   // towards the coarser level (local indices)
    const int32_t *gptr; const int32_t *gsrc; int64_t n_gblk;       // deterministic Galerkin product: contributors of every coarse block
    const int64_t *mem_ptr; const int32_t *mem_idx;                 // members in the finer level (local rows)
};

// per block dimension (3: SE2 / XY graphs, 6: SE3 graphs): doubles per vector record, per pose record, geometry
// dimension (positions / lever arms of the AMG transfer operators), doubles per lever-arm record, components of a
// measurement record (z + upper triangle of Omega)
template <int D> struct VecStride { static constexpr int value = (D == 3) ? 4 : D; };
template <int D> struct Dim;
template <> struct Dim<3> { static constexpr int VS = 4, PS = 4, NG = 2, LS = 2, NZ = 4, NW = 6, NM = 10; };
template <> struct Dim<6> { static constexpr int VS = 6, PS = 8, NG = 3, LS = 4, NZ = 7, NW = 21, NM = 28; };

// Programmatic dependent launch: the FIRST statement of every kernel.  A grid launched with the programmatic-serialisation
// attribute (launch_k in pgo_b200.cu) is scheduled while its predecessor drains; griddepcontrol.wait blocks until the
// predecessor grid has completed and its writes are visible, so nothing may be read before it -- not even the `done`
// flag.  Every kernel executes the wait (transitively every earlier kernel of the stream has completed by then) and
// only then lets its own successor start launching.  A no-op for grids launched without the attribute.
#define PDL_ENTER()                                                 \
    do {                                                            \
        asm volatile("griddepcontrol.wait;" ::: "memory");          \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
    } while (0)

template <typename VT> __device__ __forceinline__ const VT *level_val(const LevelDev &L);
template <> __device__ __forceinline__ const double *level_val<double>(const LevelDev &L) { return L.val; }
template <> __device__ __forceinline__ const float *level_val<float>(const LevelDev &L) { return L.valf; }
template <typename VT> __device__ __forceinline__ const VT *level_diag(const LevelDev &L);
template <> __device__ __forceinline__ const double *level_diag<double>(const LevelDev &L) { return L.diag; }
template <> __device__ __forceinline__ const float *level_diag<float>(const LevelDev &L) { return L.diagf; }
template <typename VT> __device__ __forceinline__ const VT *level_dinv(const LevelDev &L);
template <> __device__ __forceinline__ const double *level_dinv<double>(const LevelDev &L) { return L.dinv; }
template <> __device__ __forceinline__ const float *level_dinv<float>(const LevelDev &L) { return L.dinvf; }

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

// Block-level sum of NV values; the block writes its partials, the LAST block to arrive sums all partials in a
// fixed order (deterministic) and returns true in thread 0 with `total` set.  counter wraps to 0.
// (vb, nvb) = this block's index and the block count of the launch: blockIdx.x / gridDim.x for a plain kernel, the virtual
// block of a stage when the stage runs inside the persistent coarse-tail kernel (k_tail).
template <int NT, int NV> __device__ bool block_sum_last(const double *v, double *partials, unsigned *counter, double *total, unsigned vb, unsigned nvb) {
    __shared__ double sm[NV][NT / 32];
    __shared__ bool is_last;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double w = warp_sum(v[k]);
        if ((threadIdx.x & 31) == 0) sm[k][threadIdx.x >> 5] = w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double s = 0;
#pragma unroll
            for (int w = 0; w < NT / 32; w++) s += sm[k][w];
            partials[(size_t)vb * NV + k] = s;
        }
        __threadfence();
        unsigned t = atomicInc(counter, nvb - 1);
        is_last = (t == nvb - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double s[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) s[k] = 0;
    for (unsigned i = threadIdx.x; i < nvb; i += NT)
#pragma unroll
        for (int k = 0; k < NV; k++) s[k] += __ldcg(partials + (size_t)i * NV + k);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double w = warp_sum(s[k]);
        if ((threadIdx.x & 31) == 0) sm[k][threadIdx.x >> 5] = w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double t = 0;
#pragma unroll
            for (int w = 0; w < NT / 32; w++) t += sm[k][w];
            total[k] = t;
        }
    }
    return true;                                   // every thread of the last block; `total` is valid in thread 0
}

// Cross-rank all-reduce / barrier over NVLink peer memory, executed by ONE CTA per rank (the last block of a reducing kernel, or
// the single block of k_xbarrier): thread t < world stores this rank's NV partial sums and then the new epoch into rank t's Comm
// block (st.release.sys), and spins on its own Comm until rank t's epoch arrives (ld.acquire.sys); thread 0 then adds the
// partials in rank order (identical bits on every rank).  Kernels of one stream run in order, so every later kernel sees the
// peers' earlier writes; no peer can be more than one epoch ahead (values are double-buffered by epoch parity).  A spin of
// more than ~20 s sets ST_COMM and ends the solve instead of hanging.  Called by all threads of the block; returns false on time-out.
template <int NV> __device__ bool xrank_exchange(Scalars *S, double *total) {
    __shared__ unsigned long long ep;
    __shared__ int bad;
    __shared__ double mine_v[NV > 0 ? NV : 1];
    const int t = threadIdx.x, world = S->world, rank = S->rank;
    if (t == 0) {
        ep = S->epoch + 1; S->epoch = ep; bad = 0;
#pragma unroll
        for (int k = 0; k < NV; k++) mine_v[k] = total[k];
    }
    __syncthreads();
    const unsigned long long epoch = ep;
    const int slot = (int)(epoch & 1);
    Comm *mine = S->comm_mine;
    if (t < world) {
        Comm *p = S->comm_peer[t];
#pragma unroll
        for (int k = 0; k < NV; k++) *((volatile double *)&p->vals[slot][rank][k]) = mine_v[k];
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&p->flags[rank]), "l"(epoch) : "memory");
        const long long t0 = clock64();
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(&mine->flags[t]) : "memory");
            if (v < epoch && clock64() - t0 > S->spin_limit) { bad = 1; break; }
        } while (v < epoch);
    }
    __syncthreads();
    if (bad) {
        if (t == 0) { S->status = ST_COMM; S->done = 1; __threadfence(); }
        return false;
    }
    if (t == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double s = 0.0;
            for (int r = 0; r < world; r++) s += *((volatile double *)&mine->vals[slot][r][k]);
            total[k] = s;
        }
    }
    return true;
}

// what the owner of a global sum does with it (PCG / flexible-CG / K-cycle scalar recurrences)
__device__ __forceinline__ void finalize(int FIN, Scalars *S, const double *t, int lvl) {
    if (FIN == FIN_PQ) {
        S->pq = t[0];
        if (!(t[0] > 0.0)) { S->status = ST_BREAKDOWN; S->done = 1; S->alpha = 0.0; }
        else S->alpha = S->rz / t[0];
    } else if (FIN == FIN_RZ) {
        // flexible CG (the K-cycle preconditioner is a variable operator): beta = -(z.q)/(p.q); t = {r.z, z.q}
        S->beta = -t[1] / S->pq;
        S->rz = t[0];
        int it = S->iters + 1;
        S->iters = it;
        if (!(t[0] >= 0.0)) { S->status = ST_BREAKDOWN; S->done = 1; }
        else if (t[0] <= S->tol2 * S->rz0) S->done = 1;
        else if (it >= S->max_iters) { S->status = ST_MAXIT; S->done = 1; }
    } else if (FIN == FIN_RZ_INIT) {
        S->rz = t[0]; S->rz0 = t[0]; S->beta = 0.0;
        if (!(t[0] >= 0.0)) { S->status = ST_BREAKDOWN; S->done = 1; }
        else if (t[0] == 0.0) S->done = 1;
    } else if (FIN == FIN_NORM) {
        S->norm2_dx = t[0];
    } else if (FIN == FIN_CHI2) {
        S->chi2 = t[0];
    } else if (FIN == FIN_K1) {
        // first inner FCG step of the K-cycle at level lvl: t = {c1.v1, c1.rhs}
        KScal &K = S->k[lvl];
        K.rho1 = t[0]; K.a1 = t[1];
        K.alpha = (t[0] > 0.0) ? t[1] / t[0] : 0.0;
    } else if (FIN == FIN_K2) {
        // second step: t = {c2.v1, c2.v2, c2.r1}.  Two-step K-cycle (Notay): x = coef1 c1 + coef2 c2.  Three-step K-cycle
        // (flexible CG with full orthogonalisation of the search directions d1 = c1, d2 = c2 - (gam21/rho1) c1, ...):
        // the residual after the second step is r2 = r1 - alpha2 (v2 - (gam21/rho1) v1) = r1 - e2 v2 + e1 v1.
        KScal &K = S->k[lvl];
        const double rho1 = K.rho1, a1 = K.a1, gam = t[0], beta = t[1], a2 = t[2];
        const double rho2 = beta - gam * gam / rho1;
        const bool ok = rho1 > 0.0 && rho2 > 1e-12 * beta && rho2 == rho2;
        if (ok) {
            K.coef1 = a1 / rho1 - gam * a2 / (rho1 * rho2);
            K.coef2 = a2 / rho2;
        } else {
            K.coef1 = K.alpha; K.coef2 = 0.0;
        }
        K.coef3 = 0.0;
        K.rho2 = ok ? rho2 : 0.0; K.gam21 = ok ? gam : 0.0;
        K.alpha2 = ok ? a2 / rho2 : 0.0;
        K.e2 = K.alpha2; K.e1 = ok ? K.alpha2 * gam / rho1 : 0.0;
    } else if (FIN == FIN_K3) {
        // third step: t = {c3.v1, c3.v2, c3.v3, c3.r2};  d3 = c3 - (g31/rho1) c1 - (g32/rho2) d2 with g32 = c3.(A d2)
        KScal &K = S->k[lvl];
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
    __threadfence();
}

template <int NT, int FIN> __device__ __forceinline__ void reduce_and_finalize(const double *dots, Scalars *S, double *partials, int lvl,
                                                                              unsigned vb, unsigned nvb) {
    if (FIN == FIN_NONE) return;
    constexpr int NV = fin_ndot(FIN) > 0 ? fin_ndot(FIN) : 1;
    double total[NV];
    if (!block_sum_last<NT, NV>(dots, partials, &S->counter[FIN], total, vb, nvb)) return;
    // the last block of this rank: on a sharded level it meets the other ranks' last blocks right here (no separate kernel)
    if (S->world > 1 && lvl < S->repl_from && !xrank_exchange<NV>(S, total)) return;
    if (threadIdx.x == 0) finalize(FIN, S, total, lvl);
}

// Tail of the sliced SpMV of one block row (shared by k_spmv and k_spmv_tma): adds the diagonal block, applies the MODE,
// stores y and returns this row's contributions to the fused dot products.
template <int D, int MODE, int FIN, typename VT>
__device__ __forceinline__ void spmv_row_finish(const LevelDev &L, int64_t row, double *acc, const double *xi, const double *__restrict__ r,
                                                double *__restrict__ y, double omega, const double *__restrict__ u1,
                                                const double *__restrict__ u2, double *dots) {
    constexpr int DD = D * D, VS = VecStride<D>::value;
    {   // diagonal block last: its D^2 loads are not held in registers across the loop
        const VT *dg = level_diag<VT>(L) + row;
        VT dgv[DD];
#pragma unroll
        for (int q = 0; q < DD; q++) dgv[q] = __ldg(dg + (int64_t)q * L.n_pad);
#pragma unroll
        for (int a = 0; a < D; a++)
#pragma unroll
            for (int b = 0; b < D; b++) acc[a] = fma((double)dgv[a * D + b], xi[b], acc[a]);
    }
    double out[VS];
#pragma unroll
    for (int a = 0; a < VS; a++) out[a] = 0.0;
    if (MODE == 0) {
#pragma unroll
        for (int a = 0; a < D; a++) out[a] = acc[a];
        if (FIN == FIN_K1) {
            double ui[VS];
            ld_vec<VS>(u1 + row * VS, ui);
#pragma unroll
            for (int a = 0; a < D; a++) { dots[0] = fma(xi[a], acc[a], dots[0]); dots[1] = fma(xi[a], ui[a], dots[1]); }
        } else if (FIN == FIN_K2) {
            double ui[VS], wi[VS];
            ld_vec<VS>(u1 + row * VS, ui);
            ld_vec<VS>(u2 + row * VS, wi);
#pragma unroll
            for (int a = 0; a < D; a++) { dots[0] = fma(xi[a], ui[a], dots[0]); dots[1] = fma(xi[a], acc[a], dots[1]); dots[2] = fma(xi[a], wi[a], dots[2]); }
        } else if (FIN == FIN_K3) {      // {x.u1, x.u2, x.y, x.r}: r carries the third vector (MODE 0 does not use it otherwise)
            double ui[VS], wi[VS], zi[VS];
            ld_vec<VS>(u1 + row * VS, ui);
            ld_vec<VS>(u2 + row * VS, wi);
            ld_vec<VS>(r + row * VS, zi);
#pragma unroll
            for (int a = 0; a < D; a++) {
                dots[0] = fma(xi[a], ui[a], dots[0]); dots[1] = fma(xi[a], wi[a], dots[1]);
                dots[2] = fma(xi[a], acc[a], dots[2]); dots[3] = fma(xi[a], zi[a], dots[3]);
            }
        } else {
#pragma unroll
            for (int a = 0; a < D; a++) dots[0] = fma(xi[a], acc[a], dots[0]);
        }
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
            const VT *di = level_dinv<VT>(L) + row;
#pragma unroll
            for (int a = 0; a < D; a++) {
                double s = 0.0;
#pragma unroll
                for (int b = 0; b < D; b++) s = fma((double)__ldg(di + (int64_t)(a * D + b) * L.n_pad), t[b], s);
                out[a] = fma(omega, s, xi[a]);
                dots[0] = fma(ri[a], out[a], dots[0]);
            }
            if (FIN == FIN_RZ) {
                double ui[VS];
                ld_vec<VS>(u1 + row * VS, ui);
#pragma unroll
                for (int a = 0; a < D; a++) dots[1] = fma(ui[a], out[a], dots[1]);
            }
        }
    }
    st_vec<VS>(y + row * VS, out);
}

// ------------------------------------------------------------------------------------------------
// Level 0: BSR SpMV over the sliced jagged storage, one thread per block row.
//   MODE 0: y = H x                      (+ x.y  -> FIN_PQ)
//   MODE 1: y = r - H x                  (residual)
//   MODE 2: y = x + omega Dinv (r - H x) (damped block-Jacobi sweep; + r.y -> FIN_RZ_INIT, + {r.y, y.u1} -> FIN_RZ)
// Large coarse levels use the same kernel (K-cycle dots FIN_K1 / FIN_K2 as in k_spmv_csr).
// (Loading the block values -- a pure stream -- with the evict-first policy, ld.global.cs, so that they do not push the gathered x
// records and the L2-resident coarse levels out of L2, was measured and does not pay: 135 vs 130 us per launch, the same PCG time;
// profiles/r03b_stream_cs.log.)
template <int D, int MODE, int FIN, bool PEER, typename VT = double, int U = 1>
__global__ void __launch_bounds__(128) k_spmv(LevelDev L, const __grid_constant__ XRef xr, const double *__restrict__ x, const double *__restrict__ r,
                                               double *__restrict__ y, double omega, const double *__restrict__ u1, const double *__restrict__ u2,
                                               Scalars *S, double *partials, int lvl, int check_done) {
    PDL_ENTER();
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
    double dots[4] = {0.0, 0.0, 0.0, 0.0};
    const bool live = slice < L.n_slices;
    if (live) {
        const int mydeg = L.deg[row];
        const int maxdeg = __shfl_sync(0xffffffffu, mydeg, 0);
        const int64_t base = L.slice_ptr[slice];
        const VT *__restrict__ vals = level_val<VT>(L) + base * DD + lane;
        const uint32_t *__restrict__ cols = L.col + base + lane;
        // The kernel is bound by the bytes it keeps in flight (one warp streams one slice; every load is a dependence-free
        // coalesced 8- or 4-byte-per-lane access), so an iteration requests ALL the loads of its U entries ("columns" of the
        // slice) before the first multiply -- U x D^2 block values + U gathered x records -- plus the column words of the
        // NEXT iteration, so the chain col -> gather never sits inside an iteration.  Measured on B200 (1M-pose SE2 graph,
        // profiles/r01l_spmv_sweep.log): U = 1 at 52 registers / 9 CTAs per SM is the optimum, 130 us (fp64 blocks, 86 % of
        // the measured HBM peak; the previous gather-ahead loop: 147 us) and 100 us (fp32 blocks); U = 2 (78 registers) 137 /
        // 104 us, U = 4 (128 registers) 142 / 136 us; capping registers for 12 / 14 CTAs per SM: 139 / 170 us (spills).
        uint32_t cw[U];
        {
            int64_t o = 0;
#pragma unroll
            for (int j = 0; j < U; j++) {
                cw[j] = 0;
                if (j < mydeg) cw[j] = __ldg(cols + o);
                o += __popc(__ballot_sync(0xffffffffu, j < mydeg));
            }
        }
        ld_vec<VS>(x + row * VS, xi);
        int64_t off = 0;
        for (int k = 0; k < maxdeg; k += U) {
            int cn[U];
            int64_t of[U];
#pragma unroll
            for (int j = 0; j < U; j++) {
                cn[j] = __popc(__ballot_sync(0xffffffffu, k + j < mydeg));
                of[j] = off;
                off += cn[j];
            }
            double xj[U][VS];
            VT hv[U][DD];
#pragma unroll
            for (int j = 0; j < U; j++) {
#pragma unroll
                for (int a = 0; a < VS; a++) xj[j][a] = 0.0;
#pragma unroll
                for (int q = 0; q < DD; q++) hv[j][q] = (VT)0;
                if (k + j < mydeg) {
                    ld_vec<VS>(PEER ? xgather<VS>(xr, cw[j]) : x + (int64_t)(cw[j] & COL_LOCAL_MASK) * VS, xj[j]);
                    const VT *v = vals + of[j] * DD;
#pragma unroll
                    for (int q = 0; q < DD; q++) hv[j][q] = __ldg(v + (int64_t)q * cn[j]);
                }
            }
            {
                int64_t o = off;
#pragma unroll
                for (int j = 0; j < U; j++) {
                    const bool on = k + U + j < mydeg;
                    cw[j] = on ? __ldg(cols + o) : 0u;
                    o += __popc(__ballot_sync(0xffffffffu, on));
                }
            }
#pragma unroll
            for (int j = 0; j < U; j++)
#pragma unroll
                for (int a = 0; a < D; a++)
#pragma unroll
                    for (int b = 0; b < D; b++) acc[a] = fma((double)hv[j][a * D + b], xj[j][b], acc[a]);
        }
        spmv_row_finish<D, MODE, FIN, VT>(L, row, acc, xi, r, y, omega, u1, u2, dots);
    }
    reduce_and_finalize<128, FIN>(dots, S, partials, lvl, blockIdx.x, gridDim.x);
}

// ---- residual in extended precision (pgo_options.refine) -----------------------------------------------------------------------
// r = b - H x with every product and sum carried in double-double (error-free two_prod via FMA, two_sum), rounded to fp64 once at the
// end.  At 1M poses the first Gauss-Newton step is ~80 m per pose and cond(H) ~ 1e9: a plain fp64 residual is itself wrong at the level
// of the 1e-6 m pose tolerance, which is why NO fp64 solver gets there; one refinement round x += H^-1 r with THIS
// residual does.  Same streaming pass over the sliced storage as k_spmv (one thread per block row); ~10x the flops, still memory-bound.
struct dd { double hi, lo; };
// (__dadd_rn / __dmul_rn: never contracted into an FMA by the compiler -- the error-free transformations rely on the individual roundings)
__device__ __forceinline__ dd dd_add(dd a, double bh, double bl) {        // a + (bh + bl)
    const double s = __dadd_rn(a.hi, bh), v = __dadd_rn(s, -a.hi);
    const double e = __dadd_rn(__dadd_rn(__dadd_rn(a.hi, -__dadd_rn(s, -v)), __dadd_rn(bh, -v)), __dadd_rn(a.lo, bl));   // TwoSum + the low parts
    dd o; o.hi = __dadd_rn(s, e); o.lo = __dadd_rn(e, -__dadd_rn(o.hi, -s));
    return o;
}
__device__ __forceinline__ dd dd_sub_prod(dd a, double h, double x) {     // a - h * x, the product exact (TwoProduct by FMA)
    const double p = __dmul_rn(h, x), pe = __fma_rn(h, x, -p);
    return dd_add(a, -p, -pe);
}
template <int D>
__global__ void __launch_bounds__(128) k_residual_dd(LevelDev L, const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ r) {
    PDL_ENTER();
    constexpr int DD = D * D, VS = VecStride<D>::value;
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int64_t slice = row >> 5;
    if (slice >= L.n_slices) return;
    const int mydeg = L.deg[row];
    const int maxdeg = __shfl_sync(0xffffffffu, mydeg, 0);
    double xi[VS], bi[VS];
    ld_vec<VS>(x + row * VS, xi);
    ld_vec<VS>(b + row * VS, bi);
    dd acc[D];
#pragma unroll
    for (int a = 0; a < D; a++) { acc[a].hi = bi[a]; acc[a].lo = 0.0; }
    const int64_t base = L.slice_ptr[slice];
    int64_t off = 0;
    for (int k = 0; k < maxdeg; k++) {
        const bool active = k < mydeg;
        const int cnt = __popc(__ballot_sync(0xffffffffu, active));
        if (active) {
            const uint32_t cw = __ldg(L.col + base + off + lane);
            double xj[VS];
            ld_vec<VS>(x + (int64_t)(cw & COL_LOCAL_MASK) * VS, xj);
            const double *v = L.val + (base + off) * DD + lane;
#pragma unroll
            for (int a = 0; a < D; a++)
#pragma unroll
                for (int c = 0; c < D; c++) acc[a] = dd_sub_prod(acc[a], __ldg(v + (int64_t)(a * D + c) * cnt), xj[c]);
        }
        off += cnt;
    }
    double out[VS];
#pragma unroll
    for (int a = 0; a < VS; a++) out[a] = 0.0;
    if (row < L.n) {
#pragma unroll
        for (int a = 0; a < D; a++) {
#pragma unroll
            for (int c = 0; c < D; c++) acc[a] = dd_sub_prod(acc[a], L.diag[(int64_t)(a * D + c) * L.n_pad + row], xi[c]);
            out[a] = acc[a].hi + acc[a].lo;
        }
    }
    st_vec<VS>(r + row * VS, out);
}

// x += y (vector records)
__global__ void __launch_bounds__(256) k_add_to(int64_t n_doubles, double *__restrict__ x, const double *__restrict__ y) {
    PDL_ENTER();
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 2;
    if (i >= n_doubles) return;
    double2 a = *reinterpret_cast<double2 *>(x + i);
    const double2 c = *reinterpret_cast<const double2 *>(y + i);
    a.x += c.x; a.y += c.y;
    *reinterpret_cast<double2 *>(x + i) = a;
}

// ---- TMA-staged variant of the sliced SpMV -----------------------------------------------------------------------
// Same mapping (one thread per block row, one warp per 32-row slice), but the block values -- >= 90 % of the bytes -- do
// not pass through registers while in flight: lane 0 of every warp streams the slice's columns with 1-D bulk copies
// (cp.async.bulk global -> shared, completion on an mbarrier) into a private ring of NS slots, NS columns ahead of the
// multiply, so the bytes in flight per SM are bounded by shared memory (NS x 2.3 KB per warp at D = 3 / fp64) instead
// of by registers x occupancy.  A column starts on an 8-byte (fp64) or 4-byte (fp32) boundary, so the copy is widened to
// the enclosing 16-byte-aligned window (the few extra bytes belong to the neighbouring columns / the allocation's pad).
// The gathered x records stay in registers, requested PD = 2 columns ahead, the column words PD + 1 ahead.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <int D, typename VT> struct TmaSlot { static constexpr int BYTES = (32 * D * D * (int)sizeof(VT) + 16 + 15) / 16 * 16; };
template <int D, typename VT> inline size_t spmv_tma_smem(int ns) { return (size_t)4 * ns * TmaSlot<D, VT>::BYTES + (size_t)4 * ns * 8; }

template <int D, int MODE, int FIN, typename VT>
__global__ void __launch_bounds__(128) k_spmv_tma(LevelDev L, const double *__restrict__ x, const double *__restrict__ r, double *__restrict__ y,
                                                   double omega, const double *__restrict__ u1, const double *__restrict__ u2, Scalars *S,
                                                   double *partials, int lvl, int check_done, int NS) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    constexpr int DD = D * D, VS = VecStride<D>::value, PD = 2, SLOT = TmaSlot<D, VT>::BYTES;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char tma_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int64_t slice = row >> 5;
    double acc[D], xi[VS];
#pragma unroll
    for (int a = 0; a < D; a++) acc[a] = 0.0;
#pragma unroll
    for (int a = 0; a < VS; a++) xi[a] = 0.0;
    double dots[4] = {0.0, 0.0, 0.0, 0.0};
    const bool live = slice < L.n_slices;
    if (live) {
        unsigned char *ring = tma_smem + (size_t)warp * NS * SLOT;
        const uint32_t ring_s = smem_u32(ring);
        const uint32_t bar_s = smem_u32(tma_smem + (size_t)4 * NS * SLOT) + (uint32_t)warp * NS * 8;
        if (lane == 0) {
            for (int q = 0; q < NS; q++) mbar_init(bar_s + 8 * q, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        const int mydeg = L.deg[row];
        const int maxdeg = __shfl_sync(FULL, mydeg, 0);
        const int64_t base = L.slice_ptr[slice];
        const unsigned char *vbytes = reinterpret_cast<const unsigned char *>(level_val<VT>(L) + base * DD);
        const uint32_t *__restrict__ cols = L.col + base + lane;
        // producer cursor: next column to copy, its offset (in stored blocks) inside the slice, its ring slot
        int ki = 0, si = 0;
        int64_t off_i = 0;
        auto issue = [&]() {
            const int c = __popc(__ballot_sync(FULL, ki < mydeg));
            if (ki < maxdeg) {
                const unsigned char *src = vbytes + off_i * (int64_t)(DD * sizeof(VT));
                const uintptr_t a = reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)15;
                const uint32_t bytes = ((uint32_t)(reinterpret_cast<uintptr_t>(src) - a) + (uint32_t)c * DD * (uint32_t)sizeof(VT) + 15u) & ~15u;
                if (lane == 0) {
                    mbar_expect_tx(bar_s + 8 * si, bytes);
                    bulk_g2s(ring_s + (uint32_t)si * SLOT, reinterpret_cast<const void *>(a), bytes, bar_s + 8 * si);
                }
                off_i += c;
            }
            ki++;
            si = (si + 1 == NS) ? 0 : si + 1;
        };
        for (int q = 0; q < NS; q++) issue();
        // column-word cursor
        int kc = 0;
        int64_t off_c = 0;
        auto next_cw = [&]() -> uint32_t {
            const bool on = kc < mydeg;
            const uint32_t w = on ? __ldg(cols + off_c) : 0u;
            off_c += __popc(__ballot_sync(FULL, on));
            kc++;
            return w;
        };
        double xq[PD][VS];
#pragma unroll
        for (int j = 0; j < PD; j++) {
            const uint32_t w = next_cw();
#pragma unroll
            for (int a = 0; a < VS; a++) xq[j][a] = 0.0;
            if (j < mydeg) ld_vec<VS>(x + (int64_t)(w & COL_LOCAL_MASK) * VS, xq[j]);
        }
        uint32_t cwn = next_cw();            // column word of entry PD
        ld_vec<VS>(x + row * VS, xi);
        int64_t off = 0;
        int sc = 0;
        uint32_t parity = 0;
        for (int k = 0; k < maxdeg; k++) {
            const int cnt = __popc(__ballot_sync(FULL, k < mydeg));
            double xj[VS];
#pragma unroll
            for (int a = 0; a < VS; a++) xj[a] = xq[0][a];
#pragma unroll
            for (int j = 0; j + 1 < PD; j++)
#pragma unroll
                for (int a = 0; a < VS; a++) xq[j][a] = xq[j + 1][a];
#pragma unroll
            for (int a = 0; a < VS; a++) xq[PD - 1][a] = 0.0;
            if (k + PD < mydeg) ld_vec<VS>(x + (int64_t)(cwn & COL_LOCAL_MASK) * VS, xq[PD - 1]);
            cwn = next_cw();
            mbar_wait(bar_s + 8 * sc, parity);
            if (k < mydeg) {
                const unsigned char *src = vbytes + off * (int64_t)(DD * sizeof(VT));
                const uint32_t delta = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);
                const VT *v = reinterpret_cast<const VT *>(ring + (size_t)sc * SLOT + delta) + lane;
                VT hv[DD];
#pragma unroll
                for (int q = 0; q < DD; q++) hv[q] = v[q * cnt];
#pragma unroll
                for (int a = 0; a < D; a++)
#pragma unroll
                    for (int b = 0; b < D; b++) acc[a] = fma((double)hv[a * D + b], xj[b], acc[a]);
            }
            off += cnt;
            sc++;
            if (sc == NS) { sc = 0; parity ^= 1u; }
            __syncwarp();                    // every lane has read its values: the slot may be overwritten
            issue();
        }
        spmv_row_finish<D, MODE, FIN, VT>(L, row, acc, xi, r, y, omega, u1, u2, dots);
    }
    reduce_and_finalize<128, FIN>(dots, S, partials, lvl, blockIdx.x, gridDim.x);
}

// Coarse levels (L2-resident, latency-bound): block CSR, LPR lanes per block row (8 when rows are short, else a full
// warp), 256 / LPR rows per CTA.  Same modes; K-cycle dots:
//   FIN_K1: {x.y, x.u1}    FIN_K2: {x.u1, x.y, x.u2}      (x = c, y = H c)
template <int D, int MODE, int FIN, bool PEER, int LPR, typename VT = double>
__device__ __forceinline__ void spmv_csr_body(const LevelDev &L, const XRef &xr, const double *__restrict__ x, const double *__restrict__ r,
                                              double *__restrict__ y, double omega, const double *__restrict__ u1,
                                              const double *__restrict__ u2, Scalars *S, double *partials, int lvl, unsigned vb, unsigned nvb) {
    constexpr int DD = D * D, VS = VecStride<D>::value;
    const int sub = threadIdx.x & (LPR - 1);
    const int64_t row = (int64_t)vb * (256 / LPR) + threadIdx.x / LPR;
    double dots[4] = {0.0, 0.0, 0.0, 0.0};
    // whole warps take the branch together (LPR divides 32 and rows beyond n only occur at the tail)
    const bool live = row < L.n;
    double acc[D];
#pragma unroll
    for (int a = 0; a < D; a++) acc[a] = 0.0;
    // the lane that finishes the row requests everything the epilogue needs (own x record, diagonal block, rhs, inverse
    // diagonal) BEFORE walking the row: these kernels are chains of dependent L2 accesses (row pointer -> column -> x), and
    // the epilogue's loads would otherwise add one more link after the shuffle reduction
    // (6x6 blocks: 177 registers, one 256-thread CTA per SM; capping at 128 registers or keeping the inverse diagonal a late load
    // were both measured slower, profiles/r02a_csr_variants.log)
    constexpr bool PRE_DINV = MODE == 2;
    double xi[VS], dgv[DD], ri[VS], div[PRE_DINV ? DD : 1];
#pragma unroll
    for (int a = 0; a < VS; a++) { xi[a] = 0.0; ri[a] = 0.0; }
    if (live && sub == 0) {
        ld_vec<VS>(x + row * VS, xi);
        const double *dg = L.diag + row;
#pragma unroll
        for (int q = 0; q < DD; q++) dgv[q] = dg[(int64_t)q * L.n_pad];
        if (MODE != 0 || FIN == FIN_K3) ld_vec<VS>(r + row * VS, ri);
        if (PRE_DINV) {
            const double *di = L.dinv + row;
#pragma unroll
            for (int q = 0; q < DD; q++) div[q] = di[(int64_t)q * L.n_pad];
        }
    }
    // (a fixed-stride copy of every row's first column words, so that the x gather would not wait for the row pointer, was
    // measured and did not help: level-1 solve 319 -> 335 us, profiles/r01z_col0_experiment.log)
    if (live) {
        const int64_t b = L.slice_ptr[row], e = L.slice_ptr[row + 1];
        if constexpr (D == 3) {
            // two entries per trip, all of their loads requested before the first multiply: a row longer than LPR entries
            // costs one chain col -> x, not one per LPR entries (6x6 blocks: 174 registers, one CTA per SM -- not worth it)
            for (int64_t s = b + sub; s < e; s += 2 * LPR) {
                const bool two = s + LPR < e;
                const uint32_t ca = __ldg(L.col + s), cb = two ? __ldg(L.col + s + LPR) : 0u;
                double xa[VS], xb[VS];
#pragma unroll
                for (int a = 0; a < VS; a++) xb[a] = 0.0;
                ld_vec<VS>(PEER ? xgather<VS>(xr, ca) : x + (int64_t)(ca & COL_LOCAL_MASK) * VS, xa);
                if (two) ld_vec<VS>(PEER ? xgather<VS>(xr, cb) : x + (int64_t)(cb & COL_LOCAL_MASK) * VS, xb);
                const VT *va = level_val<VT>(L) + s * DD;
                VT ha[DD], hb[DD];
#pragma unroll
                for (int q = 0; q < DD; q++) { ha[q] = va[q]; hb[q] = two ? va[(int64_t)LPR * DD + q] : (VT)0; }
#pragma unroll
                for (int a = 0; a < D; a++)
#pragma unroll
                    for (int q = 0; q < D; q++) acc[a] = fma((double)ha[a * D + q], xa[q], fma((double)hb[a * D + q], xb[q], acc[a]));
            }
        } else {
            for (int64_t s = b + sub; s < e; s += LPR) {
                const uint32_t c = __ldg(L.col + s);
                double xj[VS];
                ld_vec<VS>(PEER ? xgather<VS>(xr, c) : x + (int64_t)(c & COL_LOCAL_MASK) * VS, xj);
                const VT *v = level_val<VT>(L) + s * DD;
#pragma unroll
                for (int a = 0; a < D; a++)
#pragma unroll
                    for (int q = 0; q < D; q++) acc[a] = fma((double)v[a * D + q], xj[q], acc[a]);
            }
        }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < D; a++) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
    }
    if (live && sub == 0) {
        double out[VS];
#pragma unroll
        for (int a = 0; a < VS; a++) out[a] = 0.0;
#pragma unroll
        for (int a = 0; a < D; a++)
#pragma unroll
            for (int q = 0; q < D; q++) acc[a] = fma(dgv[a * D + q], xi[q], acc[a]);
        if (MODE == 0) {
#pragma unroll
            for (int a = 0; a < D; a++) out[a] = acc[a];
            if (FIN == FIN_K1) {
                double ui[VS];
                ld_vec<VS>(u1 + row * VS, ui);
#pragma unroll
                for (int a = 0; a < D; a++) { dots[0] = fma(xi[a], acc[a], dots[0]); dots[1] = fma(xi[a], ui[a], dots[1]); }
            } else if (FIN == FIN_K2) {
                double ui[VS], wi[VS];
                ld_vec<VS>(u1 + row * VS, ui);
                ld_vec<VS>(u2 + row * VS, wi);
#pragma unroll
                for (int a = 0; a < D; a++) { dots[0] = fma(xi[a], ui[a], dots[0]); dots[1] = fma(xi[a], acc[a], dots[1]); dots[2] = fma(xi[a], wi[a], dots[2]); }
            } else if (FIN == FIN_K3) {
                double ui[VS], wi[VS];
                ld_vec<VS>(u1 + row * VS, ui);
                ld_vec<VS>(u2 + row * VS, wi);
#pragma unroll
                for (int a = 0; a < D; a++) {
                    dots[0] = fma(xi[a], ui[a], dots[0]); dots[1] = fma(xi[a], wi[a], dots[1]);
                    dots[2] = fma(xi[a], acc[a], dots[2]); dots[3] = fma(xi[a], ri[a], dots[3]);
                }
            }
        } else {
            if (MODE == 1) {
#pragma unroll
                for (int a = 0; a < D; a++) out[a] = ri[a] - acc[a];
            } else {
                double t[D];
#pragma unroll
                for (int a = 0; a < D; a++) t[a] = ri[a] - acc[a];
#pragma unroll
                for (int a = 0; a < D; a++) {
                    double s = 0.0;
#pragma unroll
                    for (int q = 0; q < D; q++) s = fma(PRE_DINV ? div[PRE_DINV ? a * D + q : 0] : L.dinv[row + (int64_t)(a * D + q) * L.n_pad], t[q], s);
                    out[a] = fma(omega, s, xi[a]);
                }
            }
        }
        st_vec<VS>(y + row * VS, out);
    }
    reduce_and_finalize<256, FIN>(dots, S, partials, lvl, vb, nvb);
}
template <int D, int MODE, int FIN, bool PEER, int LPR, typename VT = double>
__global__ void __launch_bounds__(256) k_spmv_csr(LevelDev L, const __grid_constant__ XRef xr, const double *__restrict__ x, const double *__restrict__ r,
                                                   double *__restrict__ y, double omega, const double *__restrict__ u1,
                                                   const double *__restrict__ u2, Scalars *S, double *partials, int lvl, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    spmv_csr_body<D, MODE, FIN, PEER, LPR, VT>(L, xr, x, r, y, omega, u1, u2, S, partials, lvl, blockIdx.x, gridDim.x);
}

// x = omega Dinv r  (pre-smoothing from a zero guess; with FIN: block-Jacobi z = Dinv r and r.z (, z.u1))
template <int D, int FIN, int NT>
__device__ __forceinline__ void dinv_apply_body(const LevelDev &L, const double *__restrict__ r, double *__restrict__ x, double omega,
                                                const double *__restrict__ u1, Scalars *S, double *partials, unsigned vb, unsigned nvb) {
    constexpr int VS = VecStride<D>::value;
    const int64_t row = (int64_t)vb * NT + threadIdx.x;
    double dots[2] = {0.0, 0.0};
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
            dots[0] = fma(ri[a], out[a], dots[0]);
        }
        if (FIN == FIN_RZ) {
            double ui[VS];
            ld_vec<VS>(u1 + row * VS, ui);
#pragma unroll
            for (int a = 0; a < D; a++) dots[1] = fma(ui[a], out[a], dots[1]);
        }
        st_vec<VS>(x + row * VS, out);
    }
    reduce_and_finalize<NT, FIN>(dots, S, partials, 0, vb, nvb);
}
template <int D, int FIN>
__global__ void __launch_bounds__(128) k_dinv_apply(LevelDev L, const double *__restrict__ r, double *__restrict__ x, double omega,
                                                     const double *__restrict__ u1, Scalars *S, double *partials, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    dinv_apply_body<D, FIN, 128>(L, r, x, omega, u1, S, partials, blockIdx.x, gridDim.x);
}

// fp32 copy of a level's stored blocks (same layout), for the SpMVs inside the multigrid cycle
__global__ void __launch_bounds__(256) k_to_float(int64_t n, const double *__restrict__ src, float *__restrict__ dst) {
    PDL_ENTER();
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 2;
    if (i + 1 < n) {
        const double2 v = *reinterpret_cast<const double2 *>(src + i);
        *reinterpret_cast<float2 *>(dst + i) = make_float2((float)v.x, (float)v.y);
    } else if (i < n) dst[i] = (float)src[i];
}

// x += alpha p ; r -= alpha q
template <int D>
__global__ void __launch_bounds__(256) k_update_xr(int64_t n_pad, double *__restrict__ x, double *__restrict__ r,
                                                    const double *__restrict__ p, const double *__restrict__ q, const Scalars *S, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
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

// x += alpha p ; r -= alpha q ; xa = omega Dinv r   (the PCG update fused with the pre-smoothing step of the cycle that follows:
// the new residual is used while it is still in registers)
template <int D>
__global__ void __launch_bounds__(128) k_update_xr_dinv(LevelDev L, double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
                                                         const double *__restrict__ q, double *__restrict__ xa, double omega, const Scalars *S, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    constexpr int VS = VecStride<D>::value;
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (row >= L.n_pad) return;
    const double a = S->alpha;
    double xv[VS], rv[VS], pv[VS], qv[VS], out[VS];
    ld_vec<VS>(x + row * VS, xv); ld_vec<VS>(p + row * VS, pv);
    ld_vec<VS>(r + row * VS, rv); ld_vec<VS>(q + row * VS, qv);
#pragma unroll
    for (int c = 0; c < VS; c++) { xv[c] = fma(a, pv[c], xv[c]); rv[c] = fma(-a, qv[c], rv[c]); out[c] = 0.0; }
    st_vec<VS>(x + row * VS, xv);
    st_vec<VS>(r + row * VS, rv);
    // xa = omega Dinv r is the first step of the PRECONDITIONER: with reduced-precision preconditioner storage it reads the fp32
    // copy of the inverse diagonal blocks like the cycle's other kernels (half the bytes of this term)
    if (L.dinvf) {
        const float *di = L.dinvf + row;
#pragma unroll
        for (int c = 0; c < D; c++) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < D; b++) s = fma((double)__ldg(di + (int64_t)(c * D + b) * L.n_pad), rv[b], s);
            out[c] = omega * s;
        }
    } else {
        const double *di = L.dinv + row;
#pragma unroll
        for (int c = 0; c < D; c++) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < D; b++) s = fma(__ldg(di + (int64_t)(c * D + b) * L.n_pad), rv[b], s);
            out[c] = omega * s;
        }
    }
    st_vec<VS>(xa + row * VS, out);
}

// p = z + beta p
template <int D>
__global__ void __launch_bounds__(256) k_update_p(int64_t n_pad, double *__restrict__ p, const double *__restrict__ z, const Scalars *S, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    constexpr int VS = VecStride<D>::value, U = 4;          // four 16-byte pairs per thread, all loads in flight before the scalar arrives
    const int64_t n2 = n_pad * VS / 2, i0 = (int64_t)blockIdx.x * 256 * U + threadIdx.x;
    double2 *p2 = reinterpret_cast<double2 *>(p);
    const double2 *z2 = reinterpret_cast<const double2 *>(z);
    double2 pv[U], zv[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int64_t i = i0 + u * 256;
        if (i < n2) { pv[u] = p2[i]; zv[u] = z2[i]; }
    }
    const double b = S->beta;
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int64_t i = i0 + u * 256;
        if (i < n2) p2[i] = make_double2(fma(b, pv[u].x, zv[u].x), fma(b, pv[u].y, zv[u].y));
    }
}

// K-cycle vector updates at level lvl:  WHICH 0: out = a - alpha_l b      WHICH 1: out = coef1_l a + coef2_l b
template <int WHICH>
__device__ __forceinline__ void kcombine_body(int64_t n_doubles, const double *__restrict__ a, const double *__restrict__ b,
                                              double *__restrict__ out, const Scalars *S, int lvl, unsigned vb, const double *__restrict__ third = nullptr) {
    const int64_t i = ((int64_t)vb * 256 + threadIdx.x) * 2;
    if (i >= n_doubles) return;
    const double2 av = *reinterpret_cast<const double2 *>(a + i), bv = *reinterpret_cast<const double2 *>(b + i);
    double2 o;
    if (WHICH == 0) { const double al = S->k[lvl].alpha; o.x = fma(-al, bv.x, av.x); o.y = fma(-al, bv.y, av.y); }
    else { const double c1 = S->k[lvl].coef1, c2 = S->k[lvl].coef2; o.x = fma(c1, av.x, c2 * bv.x); o.y = fma(c1, av.y, c2 * bv.y); }
    if (WHICH == 2) {                                // + coef3 c   (three-step K-cycle; c arrives in `out2`)
        const double c3 = S->k[lvl].coef3;
        const double2 cv = *reinterpret_cast<const double2 *>(third + i);
        o.x = fma(c3, cv.x, o.x); o.y = fma(c3, cv.y, o.y);
    }
    *reinterpret_cast<double2 *>(out + i) = o;
}
template <int WHICH>
__global__ void __launch_bounds__(256) k_kcombine(int64_t n_doubles, const double *__restrict__ a, const double *__restrict__ b,
                                                   double *__restrict__ out, const Scalars *S, int lvl, const double *__restrict__ third, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    kcombine_body<WHICH>(n_doubles, a, b, out, S, lvl, blockIdx.x, third);
}

// K-cycle, fused: r1 = rhs - alpha_l v1 (the residual after the first inner step) and the pre-smoothing step of the cycle
// that follows, xa = omega Dinv r1, while r1 is still in registers
// STEP 2 (three-step K-cycle): r2 = r1 - e2 v2 + e1 v1, in place (rhs == r1), with w = v2
template <int D, int STEP>
__global__ void __launch_bounds__(128) k_kresid_dinv(LevelDev L, const double *rhs, const double *__restrict__ v1, const double *__restrict__ w, double *r1,
                                                      double *__restrict__ xa, double omega, const Scalars *S, int lvl, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    constexpr int VS = VecStride<D>::value;
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (row >= L.n_pad) return;
    double a[VS], b[VS], out[VS];
    ld_vec<VS>(rhs + row * VS, a); ld_vec<VS>(v1 + row * VS, b);
    if (STEP == 1) {
        const double al = S->k[lvl].alpha;
#pragma unroll
        for (int c = 0; c < VS; c++) a[c] = fma(-al, b[c], a[c]);
    } else {
        const double e1 = S->k[lvl].e1, e2 = S->k[lvl].e2;
        double wv[VS];
        ld_vec<VS>(w + row * VS, wv);
#pragma unroll
        for (int c = 0; c < VS; c++) a[c] = fma(e1, b[c], fma(-e2, wv[c], a[c]));
    }
#pragma unroll
    for (int c = 0; c < VS; c++) out[c] = 0.0;
    st_vec<VS>(r1 + row * VS, a);
    const double *di = L.dinv + row;
#pragma unroll
    for (int c = 0; c < D; c++) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < D; q++) s = fma(__ldg(di + (int64_t)(c * D + q) * L.n_pad), a[q], s);
        out[c] = omega * s;
    }
    st_vec<VS>(xa + row * VS, out);
}

// ------------------------------------------------------------------------------------------------
// Aggregation AMG transfer operators.  The coarse unknown of an aggregate is a rigid motion of its members about
// the aggregate centroid c, expressed in the GLOBAL frame: (t, theta) for D = 3, (t, omega) for D = 6.  With the
// lever arm l = pos_i - c of a fine row:
//   D = 3:  P_i = [[1,0,-ly],[0,1,lx],[0,0,pz]]      pz = 0 for landmark rows (their third, padding, unknown stays decoupled)
//   D = 6:  P_i = [[I, X],[0, Q]]   X w = w x l      Q = R_i^T on level 0 (the SE3 rotation unknown is body-frame,
//                                                    R <- R Exp(dw): a global rotation w is dw_i = R_i^T w), Q = I on coarse levels
// Global rigid motions -- the near-null space of H that the 1e7 anchor barely pins -- are represented exactly on
// every level.
__device__ __forceinline__ double row_pz(const LevelDev &L, int64_t row) { return (L.vkind && L.vkind[row] == 1) ? 0.0 : 1.0; }

__device__ __forceinline__ void quat_to_R(const double *q, double *R) {      // q = (w, x, y, z), row-major R
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
}

template <int D> struct Xfer;
template <> struct Xfer<3> { double lx, ly, pz; };
template <> struct Xfer<6> { double l[3]; double Q[9]; };

__device__ __forceinline__ void xfer_rot(const double *quat_rec, double *Q) {    // Q = R^T, or I when there is no rotation record
    if (quat_rec) {
        double q[4], R[9];
        ld_vec<4>(quat_rec, q);
        quat_to_R(q, R);
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) Q[3 * a + b] = R[3 * b + a];
    } else {
#pragma unroll
        for (int a = 0; a < 9; a++) Q[a] = (a % 4 == 0) ? 1.0 : 0.0;
    }
}
// P of one of this rank's rows
template <int D> __device__ __forceinline__ Xfer<D> xfer_own(const LevelDev &F, int64_t i);
template <> __device__ __forceinline__ Xfer<3> xfer_own<3>(const LevelDev &F, int64_t i) {
    const double2 l = *reinterpret_cast<const double2 *>(F.lev + i * 2);
    return Xfer<3>{l.x, l.y, row_pz(F, i)};
}
template <> __device__ __forceinline__ Xfer<6> xfer_own<6>(const LevelDev &F, int64_t i) {
    Xfer<6> P;
    double l[4];
    ld_vec<4>(F.lev + i * 4, l);
    P.l[0] = l[0]; P.l[1] = l[1]; P.l[2] = l[2];
    xfer_rot(F.quat ? F.quat + i * 8 + 4 : nullptr, P.Q);
    return P;
}
// P of the neighbour row a stored block points to (column word cw)
template <int D> __device__ __forceinline__ Xfer<D> xfer_nbr(const LevelDev &F, const XRef &levr, uint32_t cw);
template <> __device__ __forceinline__ Xfer<3> xfer_nbr<3>(const LevelDev &F, const XRef &levr, uint32_t cw) {
    const double2 l = *reinterpret_cast<const double2 *>(xgather<2>(levr, cw));
    // the neighbour is a landmark iff this is a pose-landmark edge seen from its pose (`from`) side (level 0 only)
    const double pz = (F.vkind && (cw & COL_EDGE_XY) && !(cw & COL_ROLE_TO)) ? 0.0 : 1.0;
    return Xfer<3>{l.x, l.y, pz};
}
template <> __device__ __forceinline__ Xfer<6> xfer_nbr<6>(const LevelDev &F, const XRef &levr, uint32_t cw) {
    Xfer<6> P;
    double l[4];
    ld_vec<4>(xgather<4>(levr, cw), l);
    P.l[0] = l[0]; P.l[1] = l[1]; P.l[2] = l[2];
    xfer_rot(F.quat ? F.quat + (int64_t)(cw & COL_LOCAL_MASK) * 8 + 4 : nullptr, P.Q);     // halo rows sit behind the own rows
    return P;
}

// s += P^T r
__device__ __forceinline__ void xfer_restrict(const Xfer<3> &P, const double *r, double *s) {
    s[0] += r[0]; s[1] += r[1];
    s[2] += fma(-P.ly, r[0], fma(P.lx, r[1], P.pz * r[2]));
}
__device__ __forceinline__ void xfer_restrict(const Xfer<6> &P, const double *r, double *s) {
    s[0] += r[0]; s[1] += r[1]; s[2] += r[2];
    // X^T r_t = l x r_t ; Q^T r_w
    s[3] += (P.l[1] * r[2] - P.l[2] * r[1]) + (P.Q[0] * r[3] + P.Q[3] * r[4] + P.Q[6] * r[5]);
    s[4] += (P.l[2] * r[0] - P.l[0] * r[2]) + (P.Q[1] * r[3] + P.Q[4] * r[4] + P.Q[7] * r[5]);
    s[5] += (P.l[0] * r[1] - P.l[1] * r[0]) + (P.Q[2] * r[3] + P.Q[5] * r[4] + P.Q[8] * r[5]);
}
// x += P e
__device__ __forceinline__ void xfer_prolong(const Xfer<3> &P, const double *e, double *x) {
    x[0] += fma(-P.ly, e[2], e[0]);
    x[1] += fma(P.lx, e[2], e[1]);
    x[2] += P.pz * e[2];
}
__device__ __forceinline__ void xfer_prolong(const Xfer<6> &P, const double *e, double *x) {
    // X w = w x l
    x[0] += e[0] + (e[4] * P.l[2] - e[5] * P.l[1]);
    x[1] += e[1] + (e[5] * P.l[0] - e[3] * P.l[2]);
    x[2] += e[2] + (e[3] * P.l[1] - e[4] * P.l[0]);
    x[3] += P.Q[0] * e[3] + P.Q[1] * e[4] + P.Q[2] * e[5];
    x[4] += P.Q[3] * e[3] + P.Q[4] * e[4] + P.Q[5] * e[5];
    x[5] += P.Q[6] * e[3] + P.Q[7] * e[4] + P.Q[8] * e[5];
}
// g = P_i^T h P_j
__device__ __forceinline__ void xfer_ptap(const double *h, const Xfer<3> &Pi, const Xfer<3> &Pj, double *g) {
    double m[9];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        m[3 * a + 0] = h[3 * a + 0];
        m[3 * a + 1] = h[3 * a + 1];
        m[3 * a + 2] = fma(-Pj.ly, h[3 * a + 0], fma(Pj.lx, h[3 * a + 1], Pj.pz * h[3 * a + 2]));
    }
#pragma unroll
    for (int b = 0; b < 3; b++) {
        g[b] = m[b];
        g[3 + b] = m[3 + b];
        g[6 + b] = fma(-Pi.ly, m[b], fma(Pi.lx, m[3 + b], Pi.pz * m[6 + b]));
    }
}
__device__ __forceinline__ void xfer_ptap(const double *h, const Xfer<6> &Pi, const Xfer<6> &Pj, double *g) {
    double m[36];
    const double *l = Pj.l, *Q = Pj.Q;
#pragma unroll
    for (int a = 0; a < 6; a++) {
        const double *ha = h + 6 * a;
        m[6 * a + 0] = ha[0]; m[6 * a + 1] = ha[1]; m[6 * a + 2] = ha[2];
        // row * X = l x row_t ; row_w * Q
        m[6 * a + 3] = (l[1] * ha[2] - l[2] * ha[1]) + (ha[3] * Q[0] + ha[4] * Q[3] + ha[5] * Q[6]);
        m[6 * a + 4] = (l[2] * ha[0] - l[0] * ha[2]) + (ha[3] * Q[1] + ha[4] * Q[4] + ha[5] * Q[7]);
        m[6 * a + 5] = (l[0] * ha[1] - l[1] * ha[0]) + (ha[3] * Q[2] + ha[4] * Q[5] + ha[5] * Q[8]);
    }
    l = Pi.l; Q = Pi.Q;
#pragma unroll
    for (int b = 0; b < 6; b++) {
        const double m0 = m[b], m1 = m[6 + b], m2 = m[12 + b], m3 = m[18 + b], m4 = m[24 + b], m5 = m[30 + b];
        g[b] = m0; g[6 + b] = m1; g[12 + b] = m2;
        // X^T col_t = l x col_t ; Q^T col_w
        g[18 + b] = (l[1] * m2 - l[2] * m1) + (Q[0] * m3 + Q[3] * m4 + Q[6] * m5);
        g[24 + b] = (l[2] * m0 - l[0] * m2) + (Q[1] * m3 + Q[4] * m4 + Q[7] * m5);
        g[30 + b] = (l[0] * m1 - l[1] * m0) + (Q[2] * m3 + Q[5] * m4 + Q[8] * m5);
    }
}

// rc_I = sum_{i in I} P_i^T res_i   (16 lanes per coarse row -- an aggregate has ~16 members --, two rows per warp; fixed summation
// order => deterministic)
template <int D>
__device__ __forceinline__ void restrict_body(const LevelDev &F, const LevelDev &C, const double *__restrict__ res, double *__restrict__ rc, unsigned vb) {
    constexpr int VS = VecStride<D>::value;
    const int sub = threadIdx.x & 15;
    const int64_t I = (int64_t)vb * 16 + (threadIdx.x >> 4);
    double s[VS];
#pragma unroll
    for (int a = 0; a < VS; a++) s[a] = 0.0;
    if (I < C.n) {
        for (int64_t m = C.mem_ptr[I] + sub; m < C.mem_ptr[I + 1]; m += 16) {
            const int64_t i = C.mem_idx[m];
            double r[VS];
            ld_vec<VS>(res + i * VS, r);
            xfer_restrict(xfer_own<D>(F, i), r, s);
        }
    }
#pragma unroll
    for (int a = 0; a < D; a++)
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s[a] += __shfl_xor_sync(0xffffffffu, s[a], o);
    if (sub == 0 && I < C.n_pad) st_vec<VS>(rc + I * VS, s);
}
template <int D>
__global__ void __launch_bounds__(256) k_restrict(LevelDev F, LevelDev C, const double *__restrict__ res, double *__restrict__ rc,
                                                   const Scalars *S, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    restrict_body<D>(F, C, res, rc, blockIdx.x);
}

// x_i += P_i e_{agg(i)}
template <int D, int NT>
__device__ __forceinline__ void prolong_body(const LevelDev &F, const double *__restrict__ ec, double *__restrict__ x, unsigned vb) {
    constexpr int VS = VecStride<D>::value;
    const int64_t i = (int64_t)vb * NT + threadIdx.x;
    if (i >= F.n) return;
    const int64_t I = F.agg[i];
    double e[VS], xi[VS];
    ld_vec<VS>(ec + I * VS, e);
    ld_vec<VS>(x + i * VS, xi);
    xfer_prolong(xfer_own<D>(F, i), e, xi);
    st_vec<VS>(x + i * VS, xi);
}
// x_i += P_i (coef1 c1 + coef2 c2)_{agg(i)}: prolongation fused with the final combination of the coarse level's K-cycle
template <int D>
__global__ void __launch_bounds__(128) k_prolong_k(LevelDev F, const double *__restrict__ c1, const double *__restrict__ c2, const double *__restrict__ c3,
                                                    double *__restrict__ x, const Scalars *S, int clvl, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    constexpr int VS = VecStride<D>::value;
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= F.n) return;
    const double k1 = S->k[clvl].coef1, k2 = S->k[clvl].coef2;
    const int64_t I = F.agg[i];
    double e[VS], e2[VS], xi[VS];
    ld_vec<VS>(c1 + I * VS, e);
    ld_vec<VS>(c2 + I * VS, e2);
    ld_vec<VS>(x + i * VS, xi);
#pragma unroll
    for (int c = 0; c < VS; c++) e[c] = fma(k1, e[c], k2 * e2[c]);
    if (c3) {                                        // three-step K-cycle
        const double k3 = S->k[clvl].coef3;
        ld_vec<VS>(c3 + I * VS, e2);
#pragma unroll
        for (int c = 0; c < VS; c++) e[c] = fma(k3, e2[c], e[c]);
    }
    xfer_prolong(xfer_own<D>(F, i), e, xi);
    st_vec<VS>(x + i * VS, xi);
}
template <int D>
__global__ void __launch_bounds__(128) k_prolong(LevelDev F, const double *__restrict__ ec, double *__restrict__ x, const Scalars *S, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    prolong_body<D, 128>(F, ec, x, blockIdx.x);
}

// centroid of the members (one warp per coarse row); NG position planes
template <int NG>
__global__ void __launch_bounds__(256) k_coarse_pos(LevelDev F, LevelDev C) {
    PDL_ENTER();
    const int lane = threadIdx.x & 31;
    const int64_t I = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (I >= C.n_pad) return;
    double sp[NG];
#pragma unroll
    for (int g = 0; g < NG; g++) sp[g] = 0.0;
    if (I < C.n) {
        const int64_t b = C.mem_ptr[I], e = C.mem_ptr[I + 1];
        for (int64_t m = b + lane; m < e; m += 32) {
            const int64_t i = C.mem_idx[m];
#pragma unroll
            for (int g = 0; g < NG; g++) sp[g] += F.pos[(int64_t)g * F.n_pad + i];
        }
        const double inv = e > b ? 1.0 / (double)(e - b) : 0.0;     // rows built by another rank have no members here
#pragma unroll
        for (int g = 0; g < NG; g++) sp[g] = warp_sum(sp[g]) * inv;
    }
    if (lane == 0) {
#pragma unroll
        for (int g = 0; g < NG; g++) C.pos[(int64_t)g * C.n_pad + I] = sp[g];
    }
}

// lever arm of every fine row about its aggregate's centroid: records of LS = 2 (NG = 2) or 4 (NG = 3) doubles
template <int NG>
__global__ void __launch_bounds__(128) k_lever(LevelDev F, LevelDev C) {
    PDL_ENTER();
    constexpr int LS = NG == 2 ? 2 : 4;
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= F.n_pad) return;
    double l[LS];
#pragma unroll
    for (int g = 0; g < LS; g++) l[g] = 0.0;
    if (i < F.n) {
        const int64_t I = F.agg[i];
#pragma unroll
        for (int g = 0; g < NG; g++) l[g] = F.pos[(int64_t)g * F.n_pad + i] - C.pos[(int64_t)g * C.n_pad + I];
    }
    st_vec<LS>(F.lev + i * LS, l);
}

template <int DD>
__device__ __forceinline__ void galerkin_scatter(const LevelDev &C, int32_t tgt, int32_t stride, const double *g) {
    double *dst = tgt < 0 ? C.diag + (int64_t)(-1 - tgt) : C.val + (int64_t)tgt;
    const int64_t st = tgt < 0 ? C.n_pad : (int64_t)stride;
#pragma unroll
    for (int q = 0; q < DD; q++) atomicAdd(dst + (int64_t)q * st, g[q]);
}

// Galerkin product Hc = P^T H P: every fine block adds P_i^T H_ij P_j into the coarse block of (agg i, agg j), which
// lives in a row this rank owns.  Several fine blocks share a coarse block, hence atomics (coarse levels only; the
// Gauss-Newton system itself is assembled without atomics).  Level-0 source (JDS): one thread per row.
template <int D>
__global__ void __launch_bounds__(128) k_galerkin_jds(LevelDev F, LevelDev C, const __grid_constant__ XRef levr) {
    PDL_ENTER();
    constexpr int DD = D * D;
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int64_t slice = row >> 5;
    if (slice >= F.n_slices) return;
    const int mydeg = F.deg[row];
    const int maxdeg = __shfl_sync(0xffffffffu, mydeg, 0);
    const bool real = row < F.n;
    Xfer<D> Pi = xfer_own<D>(F, real ? row : 0);
    if (real) {
        double h[DD], g[DD];
#pragma unroll
        for (int q = 0; q < DD; q++) h[q] = F.diag[(int64_t)q * F.n_pad + row];
        xfer_ptap(h, Pi, Pi, g);
        galerkin_scatter<DD>(C, -1 - F.agg[row], 0, g);
    }
    const int64_t base = F.slice_ptr[slice];
    int64_t off = 0;
    for (int k = 0; k < maxdeg; k++) {
        const bool active = k < mydeg;
        const int cnt = __popc(__ballot_sync(0xffffffffu, active));
        if (active) {
            const int64_t slot = base + off + lane;
            const uint32_t cw = F.col[slot];
            const Xfer<D> Pj = xfer_nbr<D>(F, levr, cw);
            const double *v = F.val + (base + off) * DD + lane;
            double h[DD], g[DD];
#pragma unroll
            for (int q = 0; q < DD; q++) h[q] = v[(int64_t)q * cnt];
            xfer_ptap(h, Pi, Pj, g);
            galerkin_scatter<DD>(C, F.ctgt[slot], F.cstr[slot], g);
        }
        off += cnt;
    }
}

// coarse source (block CSR): one warp per row
template <int D>
__global__ void __launch_bounds__(256) k_galerkin_csr(LevelDev F, LevelDev C, const __grid_constant__ XRef levr) {
    PDL_ENTER();
    constexpr int DD = D * D;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= F.n) return;
    const Xfer<D> Pi = xfer_own<D>(F, row);
    if (lane == 0) {
        double h[DD], g[DD];
#pragma unroll
        for (int q = 0; q < DD; q++) h[q] = F.diag[(int64_t)q * F.n_pad + row];
        xfer_ptap(h, Pi, Pi, g);
        galerkin_scatter<DD>(C, -1 - F.agg[row], 0, g);
    }
    for (int64_t s = F.slice_ptr[row] + lane; s < F.slice_ptr[row + 1]; s += 32) {
        const Xfer<D> Pj = xfer_nbr<D>(F, levr, F.col[s]);
        const double *v = F.val + s * DD;
        double h[DD], g[DD];
#pragma unroll
        for (int q = 0; q < DD; q++) h[q] = v[q];
        xfer_ptap(h, Pi, Pj, g);
        galerkin_scatter<DD>(C, F.ctgt[s], F.cstr[s], g);
    }
}

// ---- deterministic Galerkin product (default; the atomic kernels above remain for coarse levels in sliced storage) --------------
// Step 1: every fine block is projected, g = P_i^T H_ij P_j, into a staging buffer in STORAGE order, block-major (DD doubles per
// block: stored block s at s, diagonal block of row r at n_slots + r) -- the same streaming pass as above, all stores coalesced.
// Step 2: one group of lanes per coarse block sums its contributors in the fixed order of the symbolic pass' lists.  No atomics,
// a single writer per coarse block: two runs give bit-identical hierarchies (and hence bit-identical Gauss-Newton steps).
template <int D>
__global__ void __launch_bounds__(128) k_galerkin_stage_jds(LevelDev F, const __grid_constant__ XRef levr, double *__restrict__ stage) {
    PDL_ENTER();
    constexpr int DD = D * D;
    __shared__ double sm[4][32 * DD];
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int64_t slice = row >> 5;
    if (slice >= F.n_slices) return;
    double *my = sm[threadIdx.x >> 5];
    const int mydeg = F.deg[row];
    const int maxdeg = __shfl_sync(0xffffffffu, mydeg, 0);
    const bool real = row < F.n;
    Xfer<D> Pi = xfer_own<D>(F, real ? row : 0);
    {
        double h[DD], g[DD];
#pragma unroll
        for (int q = 0; q < DD; q++) { h[q] = real ? F.diag[(int64_t)q * F.n_pad + row] : 0.0; g[q] = 0.0; }
        if (real) xfer_ptap(h, Pi, Pi, g);
#pragma unroll
        for (int q = 0; q < DD; q++) my[lane * DD + q] = g[q];
        __syncwarp();
        double *dst = stage + (F.n_slots + slice * 32) * DD;
        for (int t = lane; t < 32 * DD; t += 32) dst[t] = my[t];
        __syncwarp();
    }
    const int64_t base = F.slice_ptr[slice];
    int64_t off = 0;
    for (int k = 0; k < maxdeg; k++) {
        const bool active = k < mydeg;
        const int cnt = __popc(__ballot_sync(0xffffffffu, active));
        if (active) {
            const uint32_t cw = F.col[base + off + lane];
            const Xfer<D> Pj = xfer_nbr<D>(F, levr, cw);
            const double *v = F.val + (base + off) * DD + lane;
            double h[DD], g[DD];
#pragma unroll
            for (int q = 0; q < DD; q++) h[q] = v[(int64_t)q * cnt];
            xfer_ptap(h, Pi, Pj, g);
#pragma unroll
            for (int q = 0; q < DD; q++) my[lane * DD + q] = g[q];
        }
        __syncwarp();
        double *dst = stage + (base + off) * DD;
        for (int t = lane; t < cnt * DD; t += 32) dst[t] = my[t];
        __syncwarp();
        off += cnt;
    }
}

template <int D>
__global__ void __launch_bounds__(256) k_galerkin_stage_csr(LevelDev F, const __grid_constant__ XRef levr, double *__restrict__ stage) {
    PDL_ENTER();
    constexpr int DD = D * D;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= F.n) return;
    const Xfer<D> Pi = xfer_own<D>(F, row);
    if (lane == 0) {
        double h[DD], g[DD];
#pragma unroll
        for (int q = 0; q < DD; q++) h[q] = F.diag[(int64_t)q * F.n_pad + row];
        xfer_ptap(h, Pi, Pi, g);
        double *dst = stage + (F.n_slots + row) * DD;
#pragma unroll
        for (int q = 0; q < DD; q++) dst[q] = g[q];
    }
    for (int64_t s = F.slice_ptr[row] + lane; s < F.slice_ptr[row + 1]; s += 32) {
        const Xfer<D> Pj = xfer_nbr<D>(F, levr, F.col[s]);
        const double *v = F.val + s * DD;
        double h[DD], g[DD];
#pragma unroll
        for (int q = 0; q < DD; q++) h[q] = v[q];
        xfer_ptap(h, Pi, Pj, g);
        double *dst = stage + s * DD;
#pragma unroll
        for (int q = 0; q < DD; q++) dst[q] = g[q];
    }
}

template <int D> struct GalerkinLanes { static constexpr int value = D * D <= 16 ? 16 : 32; };   // lanes per coarse block

template <int D>
__global__ void __launch_bounds__(256) k_galerkin_reduce(LevelDev F, LevelDev C, const double *__restrict__ stage) {
    PDL_ENTER();
    constexpr int DD = D * D, LPB = GalerkinLanes<D>::value;
    const int64_t b = ((int64_t)blockIdx.x * 256 + threadIdx.x) / LPB;
    const int q0 = threadIdx.x % LPB;
    if (b >= F.n_gblk) return;
    const int32_t p0 = F.gptr[b], p1 = F.gptr[b + 1];
    if (p0 == p1) return;                           // a block of the first replicated level built by another rank: gathered later
    for (int q = q0; q < DD; q += LPB) {
        double s = 0.0;
        int32_t p = p0;
        for (; p + 4 <= p1; p += 4) {               // four independent loads in flight, summed in list order
            const double a0 = stage[(int64_t)F.gsrc[p] * DD + q], a1 = stage[(int64_t)F.gsrc[p + 1] * DD + q];
            const double a2 = stage[(int64_t)F.gsrc[p + 2] * DD + q], a3 = stage[(int64_t)F.gsrc[p + 3] * DD + q];
            s += a0; s += a1; s += a2; s += a3;
        }
        for (; p < p1; p++) s += stage[(int64_t)F.gsrc[p] * DD + q];
        if (b < C.n_pad) C.diag[(int64_t)q * C.n_pad + b] = s;
        else C.val[(b - C.n_pad) * DD + q] = s;
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
// inverse of a small SPD block: cofactors for 3x3, in-register Gauss-Jordan without pivoting for 6x6
template <int D> __device__ __forceinline__ void inv_block(const double *a, double *o) {
    if constexpr (D == 3) { inv3(a, o); } else {
#pragma unroll
    for (int q = 0; q < D * D; q++) o[q] = a[q];
#pragma unroll
    for (int k = 0; k < D; k++) {
        const double piv = 1.0 / o[k * D + k];
#pragma unroll
        for (int j = 0; j < D; j++) if (j != k) o[k * D + j] *= piv;
#pragma unroll
        for (int i = 0; i < D; i++) {
            if (i == k) continue;
            const double f = o[i * D + k];
#pragma unroll
            for (int j = 0; j < D; j++) if (j != k) o[i * D + j] = fma(-f, o[k * D + j], o[i * D + j]);
            o[i * D + k] = -f * piv;
        }
        o[k * D + k] = piv;
    }
    }
}

template <int D>
__global__ void __launch_bounds__(128) k_invert_diag(LevelDev L) {
    PDL_ENTER();
    constexpr int DD = D * D;
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (row >= L.n_pad) return;
    double a[DD], o[DD];
#pragma unroll
    for (int q = 0; q < DD; q++) a[q] = L.diag[(int64_t)q * L.n_pad + row];
    if (row < L.n) {
        // D = 3: an aggregate made of landmarks that all sit on its centroid (e.g. a single landmark) has no
        // rotational unknown: P^T H P is exactly singular in theta.  Decouple that unknown (its restricted
        // residual is always 0) so that the level stays SPD.
        if (D == 3 && !(a[8] > 1e-14 * (a[0] + a[4]))) {
            a[2] = a[5] = a[6] = a[7] = 0.0; a[8] = 1.0;
            L.diag[(int64_t)2 * L.n_pad + row] = 0.0; L.diag[(int64_t)5 * L.n_pad + row] = 0.0;
            L.diag[(int64_t)6 * L.n_pad + row] = 0.0; L.diag[(int64_t)7 * L.n_pad + row] = 0.0;
            L.diag[(int64_t)8 * L.n_pad + row] = 1.0;
        }
        inv_block<D>(a, o);
    } else {
#pragma unroll
        for (int q = 0; q < DD; q++) o[q] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < DD; q++) L.dinv[(int64_t)q * L.n_pad + row] = o[q];
}

// ------------------------------------------------------------------------------------------------
// Coarsest level: explicit dense inverse.  The level's rows are numbered densely across ranks:
// dense index of (rank k, local row i) = dense_off[k] + i ; m = D * sum of real rows.
struct DenseMap { int32_t off[MAX_RANKS + 1]; };

__device__ __forceinline__ int dense_col(const DenseMap &dm, uint32_t colword) {
    return dm.off[(colword >> COL_OWNER_SHIFT) & (MAX_RANKS - 1)] + (int)(colword & COL_LOCAL_MASK);
}

// scatter this rank's block rows into rows [D*dense_off[rank], ...) of the (pre-zeroed) dense matrix A (m x m, row-major)
template <int D, bool JDS>
__global__ void __launch_bounds__(128) k_dense_assemble(LevelDev L, DenseMap dm, int rank, int m, double *__restrict__ A) {
    PDL_ENTER();
    constexpr int DD = D * D;
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (row >= L.n) return;
    const int gi = dm.off[rank] + (int)row;
    for (int a = 0; a < D; a++)
        for (int b = 0; b < D; b++) A[(int64_t)(gi * D + a) * m + gi * D + b] = L.diag[(int64_t)(a * D + b) * L.n_pad + row];
    if (JDS) {
        const int64_t slice = row >> 5; const int lane = (int)(row & 31);
        const int64_t base = L.slice_ptr[slice];
        const int mydeg = L.deg[row];
        int64_t off = 0;
        for (int k = 0; k < mydeg; k++) {
            int cnt = 0;
            while (cnt < 32 && L.deg[slice * 32 + cnt] > k) cnt++;
            const int gj = dense_col(dm, L.col[base + off + lane]);
            const double *v = L.val + (base + off) * DD + lane;
            for (int a = 0; a < D; a++)
                for (int b = 0; b < D; b++) A[(int64_t)(gi * D + a) * m + gj * D + b] += v[(int64_t)(a * D + b) * cnt];   // duplicate edges sum
            off += cnt;
        }
    } else {
        for (int64_t s = L.slice_ptr[row]; s < L.slice_ptr[row + 1]; s++) {
            const int gj = dense_col(dm, L.col[s]);
            const double *v = L.val + s * DD;
            for (int a = 0; a < D; a++)
                for (int b = 0; b < D; b++) A[(int64_t)(gi * D + a) * m + gj * D + b] = v[a * D + b];
        }
    }
}

// Inverse of the dense coarsest matrix (symmetric positive definite, m <= 6 * 1024): blocked SWEEP operator (the symmetric form of
// Gauss-Jordan, no pivoting), panel width 32, as a persistent cooperative kernel with ONE grid barrier per panel.  With pivot block
// P = S[pp], R = S[p,r] and X = P^-1 R the panel step is
//     S[pp] <- -P^-1        S[p,r] <- X        S[r,p] <- X^T        S[r,r] <- S[r,r] - R^T X
// which keeps S symmetric; after the last panel S = -A^-1.  Only the 64x64 tiles on or above the diagonal are stored and updated --
// half the work of plain Gauss-Jordan -- and a tile reads the panel rows / columns it needs from the stored triangle (transposed when
// they lie below it).  Every CTA recomputes the slice of X its tiles need, so nothing has to be exchanged inside a panel step; the
// step reads `src` and writes every stored tile to `dst` (ping-pong, both L2-resident), which removes every read-after-write hazard
// between tiles.  The last panel writes -S, mirrored, so the caller's Ainv is the full matrix.  The matrix is treated as padded with
// an identity block up to a multiple of 32.  The host passes the buffers such that the result of the last panel lands in Ainv.
//   * the 32x32 pivot inverse of panel p+1 is produced DURING panel p (look-ahead) by one CTA that does nothing else: it updates
//     just that block, inverts it (gj_invert32_cols) and publishes it, so after the barrier every CTA only loads 8 KB;
//   * the grid barrier is one atomic arrive + an acquire spin on a counter (the cooperative launch only guarantees co-residency);
//   * the panel product is register-blocked (6 shared loads per 8 FMAs).
// History (1221^2 matrix of config 4): 2.13 ms (redundant elimination per CTA, grid.sync()) -> 1.20 ms (look-ahead, own barrier;
// every tile of the full matrix: 400 tiles on 296 CTAs = two rounds per panel) -> 1.04 ms (symmetric: 210 tiles, one round; timers
// inside the kernel then showed 16.6 us of each 26 us panel in the 32x32 look-ahead inverse on the CTA that also owned a tile,
// profiles/r04a_gj_phases.log) -> 8.5 us for that inverse, on its own CTA, beside the tile updates (~10 us).  A two-phase variant
// (X computed once per column block, a second barrier per panel) measured the same as the second one (profiles/r03n_gj3.log).
constexpr int GJ_W = 32, GJ_T = 64;
constexpr size_t GJ_SMEM = sizeof(double) * (4 * GJ_W * (GJ_W + 1) + GJ_W * (GJ_T + 1) + GJ_W * GJ_T + GJ_T * (GJ_W + 1));

__device__ __forceinline__ void gj_grid_barrier(unsigned *bar, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int)(v - target) < 0);
    }
    __syncthreads();
}

// P (in Pb[0]) <- P^-1 by 32 unblocked Gauss-Jordan steps ping-ponging Pb[0] <-> Pb[1]; all 256 threads; result in Pb[0]
__device__ __forceinline__ void gj_invert32(double (*Pb)[GJ_W][GJ_W + 1]) {
    const int tid = threadIdx.x;
    const int i = tid >> 3, j0 = (tid & 7) * 4;
    for (int k = 0; k < GJ_W; k++) {
        const double (*Pi)[GJ_W + 1] = Pb[k & 1];
        double (*Po)[GJ_W + 1] = Pb[(k & 1) ^ 1];
        const double ikk = 1.0 / Pi[k][k];
        const double pik = Pi[i][k];
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const int j = j0 + jj;
            double v;
            if (i == k) v = (j == k) ? ikk : Pi[k][j] * ikk;
            else if (j == k) v = -pik * ikk;
            else v = Pi[i][j] - pik * Pi[k][j] * ikk;
            Po[i][j] = v;
        }
        __syncthreads();
    }
}

// The same inverse for the look-ahead block, which sits on the critical path of every panel (16.6 us with the version above: a
// block barrier, dynamic indexing and an fp64 division chain per step; 8.5 us with this one).  Thread (warp w, lane i) keeps
// P[i][4w .. 4w+3] in registers through the 32 unrolled steps; the pivot row reaches the other lanes by shuffles inside each warp, the
// multiplier column and the pivot's reciprocal -- computed one step ahead by the thread that owns the next pivot -- go through a
// double-buffered shared-memory line (Pb[1]).
__device__ __forceinline__ void gj_invert32_cols(double (*Pb)[GJ_W][GJ_W + 1]) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int BS = GJ_W + 2;
    double *bc = &Pb[1][0][0];
    double r[4];
#pragma unroll
    for (int c = 0; c < 4; c++) r[c] = Pb[0][lane][4 * w + c];
    double ipn = 1.0 / r[0];                      // warp 0, lane 0: reciprocal of the first pivot
#pragma unroll
    for (int k = 0; k < GJ_W; k++) {
        const int wk = k >> 2, ck = k & 3;
        double *b = bc + (k & 1) * BS;
        if (w == wk) {
            b[lane] = r[ck];
            if (lane == k) b[GJ_W] = ipn;
        }
        __syncthreads();
        const double ip = b[GJ_W];
        const double f = b[lane] * ip;
        double pk[4];
#pragma unroll
        for (int c = 0; c < 4; c++) pk[c] = __shfl_sync(0xffffffffu, r[c], k);
        if (lane == k) {
#pragma unroll
            for (int c = 0; c < 4; c++) r[c] = (w == wk && c == ck) ? ip : pk[c] * ip;
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) r[c] = (w == wk && c == ck) ? -f : fma(-f, pk[c], r[c]);
        }
        if (k + 1 < GJ_W && w == ((k + 1) >> 2)) ipn = 1.0 / r[(k + 1) & 3];
    }
#pragma unroll
    for (int c = 0; c < 4; c++) Pb[0][lane][4 * w + c] = r[c];
    __syncthreads();
}

// number of stored tiles (on or above the diagonal) and the position of stored tile t, rows first
__host__ __device__ inline int gj_num_tiles(int nt) { return nt * (nt + 1) / 2; }
__device__ __forceinline__ void gj_tile_of(int t, int nt, int &I, int &J) {
    I = 0;
    while (t >= nt - I) { t -= nt - I; I++; }
    J = I + t;
}

__global__ void __launch_bounds__(256) k_dense_invert_sym(int m, double *__restrict__ bufA, double *__restrict__ bufB, double *__restrict__ pnext,
                                                           unsigned *bar, unsigned bar_base) {
    PDL_ENTER();
    extern __shared__ double gj_smem[];
    double *sp = gj_smem;
    double (*Pb)[GJ_W][GJ_W + 1] = reinterpret_cast<double (*)[GJ_W][GJ_W + 1]>(sp); sp += 2 * GJ_W * (GJ_W + 1);     // [2][32][33] current pivot inverse
    double (*Pn)[GJ_W][GJ_W + 1] = reinterpret_cast<double (*)[GJ_W][GJ_W + 1]>(sp); sp += 2 * GJ_W * (GJ_W + 1);     // [2][32][33] look-ahead block
    double (*Xr)[GJ_T + 1] = reinterpret_cast<double (*)[GJ_T + 1]>(sp); sp += GJ_W * (GJ_T + 1);                     // [32][65] raw panel rows
    double (*Xb)[GJ_T] = reinterpret_cast<double (*)[GJ_T]>(sp); sp += GJ_W * GJ_T;                                   // [32][64] X (P^-1 in the pivot columns)
    double (*Cb)[GJ_W + 1] = reinterpret_cast<double (*)[GJ_W + 1]>(sp);                                              // [64][33] panel columns
    const int tid = threadIdx.x;
    const int nt = (m + GJ_T - 1) / GJ_T, n_panels = (m + GJ_W - 1) / GJ_W, n_tiles = gj_num_tiles(nt);
    const int G = (int)gridDim.x - 1;                    // CTAs 0 .. G-1 update tiles, CTA G produces the pivot inverses
    const bool pivot_cta = (int)blockIdx.x == G;
    const double *src = bufA;
    double *dst = bufB;
    for (int pi = 0; pi < n_panels; pi++) {
        const int p0 = pi * GJ_W, Ip = p0 / GJ_T;
        const bool last = pi + 1 == n_panels;
        if (pi == 0) {                                   // the first pivot: every CTA inverts it for itself
            for (int t = tid; t < GJ_W * GJ_W; t += 256) {
                const int i = t / GJ_W, j = t % GJ_W;
                Pb[0][i][j] = (i < m && j < m) ? __ldcg(src + (int64_t)i * m + j) : (i == j ? 1.0 : 0.0);
            }
            __syncthreads();
            gj_invert32(Pb);
        } else {
            const double *pn = pnext + (size_t)(pi & 1) * GJ_W * GJ_W;
            for (int t = tid; t < GJ_W * GJ_W; t += 256) Pb[0][t / GJ_W][t % GJ_W] = __ldcg(pn + t);
            __syncthreads();
        }
        const double (*Pinv)[GJ_W + 1] = Pb[0];
        if (pivot_cta) {
            // Look-ahead: the NEXT pivot block as this panel leaves it, S[nn] - S[p,n]^T P^-1 S[p,n], inverted and published while the
            // other CTAs update their tiles (they recompute this block as part of its tile)
            if (!last) {
                const int np0 = p0 + GJ_W;
                const int i = tid >> 3, jq = (tid & 7) * 4;
                for (int t = tid; t < GJ_W * GJ_W; t += 256) {
                    const int l = t / GJ_W, jj = t % GJ_W;
                    const int gi = p0 + l, gj = np0 + jj;
                    Xr[l][jj] = (gi < m && gj < m) ? __ldcg(src + (int64_t)gi * m + gj) : 0.0;      // S[p, n]: on or above the diagonal
                }
                double d[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int gi = np0 + i, gj = np0 + jq + c;
                    d[c] = (gi < m && gj < m) ? __ldcg(src + (int64_t)gi * m + gj) : (gi == gj ? 1.0 : 0.0);
                }
                __syncthreads();
                double x[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
                for (int q = 0; q < GJ_W; q++) {
                    const double pv = Pinv[i][q];
#pragma unroll
                    for (int c = 0; c < 4; c++) x[c] = fma(pv, Xr[q][jq + c], x[c]);
                }
#pragma unroll
                for (int c = 0; c < 4; c++) Xb[i][jq + c] = x[c];
                __syncthreads();
#pragma unroll 8
                for (int l = 0; l < GJ_W; l++) {
                    const double rv = Xr[l][i];
#pragma unroll
                    for (int c = 0; c < 4; c++) d[c] = fma(-rv, Xb[l][jq + c], d[c]);
                }
#pragma unroll
                for (int c = 0; c < 4; c++) Pn[0][i][jq + c] = d[c];
                __syncthreads();
                gj_invert32_cols(Pn);
                double *pn = pnext + (size_t)((pi + 1) & 1) * GJ_W * GJ_W;
                for (int t = tid; t < GJ_W * GJ_W; t += 256) pn[t] = Pn[0][t / GJ_W][t % GJ_W];
            }
        } else
        for (int tile = (int)blockIdx.x; tile < n_tiles; tile += G) {
            int I, J;
            gj_tile_of(tile, nt, I, J);
            const int i0 = I * GJ_T, j0 = J * GJ_T;
            __syncthreads();                              // the previous tile's shared arrays are free
            if (Ip <= J) {                                // panel rows of this column block: stored as they are ...
                for (int t = tid; t < GJ_W * GJ_T; t += 256) {
                    const int l = t / GJ_T, jj = t % GJ_T;
                    const int gi = p0 + l, gj = j0 + jj;
                    Xr[l][jj] = (gi < m && gj < m) ? __ldcg(src + (int64_t)gi * m + gj) : 0.0;
                }
            } else {                                      // ... or below the diagonal: read S[j, p] from the stored triangle
                for (int t = tid; t < GJ_W * GJ_T; t += 256) {
                    const int jj = t / GJ_W, l = t % GJ_W;
                    const int gi = p0 + l, gj = j0 + jj;
                    Xr[l][jj] = (gi < m && gj < m) ? __ldcg(src + (int64_t)gj * m + gi) : 0.0;
                }
            }
            if (I <= Ip) {                                // panel columns of this row block
                for (int t = tid; t < GJ_T * GJ_W; t += 256) {
                    const int ii = t / GJ_W, l = t % GJ_W;
                    const int gi = i0 + ii, gj = p0 + l;
                    Cb[ii][l] = (gi < m && gj < m) ? __ldcg(src + (int64_t)gi * m + gj) : 0.0;
                }
            } else {
                for (int t = tid; t < GJ_T * GJ_W; t += 256) {
                    const int l = t / GJ_T, ii = t % GJ_T;
                    const int gi = i0 + ii, gj = p0 + l;
                    Cb[ii][l] = (gi < m && gj < m) ? __ldcg(src + (int64_t)gj * m + gi) : 0.0;
                }
            }
            const int ty = tid >> 4, tx = tid & 15;
            double aij[4][4];                             // the tile's own entries: requested before the products need them
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int gi = i0 + ty * 4 + a, gj = j0 + tx + 16 * c;
                    aij[a][c] = (gi < m && gj < m) ? __ldcg(src + (int64_t)gi * m + gj) : 0.0;
                }
            __syncthreads();
            {   // X[l][j] = (P^-1 S[p, j])[l] outside the panel columns, P^-1[l][j - p0] inside.  Register-blocked: a warp owns four rows l
                // (its P^-1 operands are warp-uniform shared-memory broadcasts), a lane two columns: 6 shared loads per 8 FMAs
                const int l0 = (tid >> 5) * 4, jj0 = tid & 31;
                double xa[4][2];
#pragma unroll
                for (int a = 0; a < 4; a++) xa[a][0] = xa[a][1] = 0.0;
#pragma unroll 8
                for (int q = 0; q < GJ_W; q++) {
                    const double x0 = Xr[q][jj0], x1 = Xr[q][jj0 + 32];
#pragma unroll
                    for (int a = 0; a < 4; a++) {
                        const double pv = Pinv[l0 + a][q];
                        xa[a][0] = fma(pv, x0, xa[a][0]);
                        xa[a][1] = fma(pv, x1, xa[a][1]);
                    }
                }
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const int jj = jj0 + 32 * c, gj = j0 + jj;
                    const bool jin = gj >= p0 && gj < p0 + GJ_W;
#pragma unroll
                    for (int a = 0; a < 4; a++) Xb[l0 + a][jj] = jin ? Pinv[l0 + a][gj - p0] : xa[a][c];
                }
            }
            __syncthreads();
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[a][c] = 0.0;
#pragma unroll 4
            for (int l = 0; l < GJ_W; l++) {
                double cv[4], xv[4];
#pragma unroll
                for (int a = 0; a < 4; a++) cv[a] = Cb[ty * 4 + a][l];
#pragma unroll
                for (int c = 0; c < 4; c++) xv[c] = Xb[l][tx + 16 * c];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[a][c] = fma(cv[a], xv[c], acc[a][c]);
            }
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const int gi = i0 + ty * 4 + a;
                const bool iin = gi >= p0 && gi < p0 + GJ_W;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int gj = j0 + tx + 16 * c;
                    const bool jin = gj >= p0 && gj < p0 + GJ_W;
                    double v;
                    if (iin) v = jin ? -Xb[gi - p0][tx + 16 * c] : Xb[gi - p0][tx + 16 * c];     // -P^-1 or X
                    else if (jin) v = acc[a][c];                                                   // X^T = S[r,p] P^-1
                    else v = aij[a][c] - acc[a][c];
                    if (gi < m && gj < m) {
                        if (!last) dst[(int64_t)gi * m + gj] = v;
                        else {                            // A^-1 = -S, both triangles (a diagonal tile holds both of its own)
                            dst[(int64_t)gi * m + gj] = -v;
                            if (I != J) dst[(int64_t)gj * m + gi] = -v;
                        }
                    }
                }
            }
        }
        gj_grid_barrier(bar, bar_base + (unsigned)(pi + 1) * gridDim.x);
        const double *t = src; src = dst; dst = const_cast<double *>(t);
    }
}

// (Splitting a row of the inverse over four warps, so that every lane has all its loads of the row in flight at once, was measured
// and LOSES: 8.8 vs 6.9 us per application -- four times as many CTAs each stage the whole right-hand side before they can start;
// profiles/r03d_setup_launches.md.)
template <int D>
__device__ __forceinline__ void dense_apply_body(int64_t n_local, const DenseMap &dm, int rank, int world, int m, const double *__restrict__ Ainv,
                                                 const XRef &rr, double *__restrict__ x, unsigned vb) {
    constexpr int VS = VecStride<D>::value;
    extern __shared__ double sr[];
    const int lane = threadIdx.x & 31;
    const int64_t srow = (int64_t)vb * 8 + (threadIdx.x >> 5);      // local scalar row
    const bool live = srow < n_local * D;
    const double *a = Ainv + ((int64_t)dm.off[rank] * D + (live ? srow : 0)) * m;
    // a row of the inverse is read in batches of DB independent loads per lane (the first one before the right-hand side
    // is staged): the kernel is a chain of L2 round trips, one per batch, not a bandwidth problem
    constexpr int DB = 16;
    double av[DB];
#pragma unroll
    for (int u = 0; u < DB; u++) { const int j = u * 32 + lane; av[u] = (live && j < m) ? __ldg(a + j) : 0.0; }
    for (int t = threadIdx.x; t < m; t += 256) {
        const int g = t / D, c = t - D * g;
        int k = 0;
        while (k + 1 < world && g >= dm.off[k + 1]) k++;
        sr[t] = rr.p[k][(int64_t)(g - dm.off[k]) * VS + c];
    }
    __syncthreads();
    if (!live) return;
    double s = 0.0;
    for (int j0 = 0; j0 < m; j0 += 32 * DB) {
        double nv[DB];
#pragma unroll
        for (int u = 0; u < DB; u++) { const int j = j0 + 32 * DB + u * 32 + lane; nv[u] = j < m ? __ldg(a + j) : 0.0; }    // next batch
#pragma unroll
        for (int u = 0; u < DB; u++) { const int j = j0 + u * 32 + lane; s = fma(av[u], j < m ? sr[j] : 0.0, s); }
#pragma unroll
        for (int u = 0; u < DB; u++) av[u] = nv[u];
    }
    s = warp_sum(s);
    if (lane == 0) x[(srow / D) * VS + (srow % D)] = s;
}
template <int D>
__global__ void __launch_bounds__(256) k_dense_apply(int64_t n_local, DenseMap dm, int rank, int world, int m, const double *__restrict__ Ainv,
                                                      const __grid_constant__ XRef rr, double *__restrict__ x, const Scalars *S, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    dense_apply_body<D>(n_local, dm, rank, world, m, Ainv, rr, x, blockIdx.x);
}

// Halo exchange: v[(n_pad + i) * stride ...] <- the record of row halo_src[i] in its owner's copy of v (peer HBM, NVLink).
// One thread per halo row: independent remote reads, so the NVLink latency is paid once per exchange, not per gather.
// stride (doubles per record) is even.
__global__ void __launch_bounds__(128) k_halo_pull(double *__restrict__ v, const __grid_constant__ XRef peers, const uint32_t *__restrict__ halo_src,
                                                    int64_t n_halo, int64_t n_pad, int stride, const Scalars *S, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n_halo) return;
    const uint32_t cw = halo_src[i];
    const double *src = peers.p[(cw >> COL_OWNER_SHIFT) & (MAX_RANKS - 1)] + (int64_t)(cw & COL_LOCAL_MASK) * stride;
    double *dst = v + (n_pad + i) * stride;
    for (int c = 0; c < stride; c += 2) *reinterpret_cast<double2 *>(dst + c) = *reinterpret_cast<const double2 *>(src + c);
}

// First replicated level of a sharded handle: copy the element ranges the OTHER ranks produced from their (identically
// laid out) arrays in peer HBM.  v has n_planes planes of plane_stride doubles; inside a plane the elements of rank k
// are [seg.off[k] * comps, seg.off[k+1] * comps).
__global__ void __launch_bounds__(256) k_gather_peer(double *__restrict__ v, const __grid_constant__ XRef src, const __grid_constant__ SegMap seg,
                                                      int rank, int world, int comps, int n_planes, int64_t plane_stride,
                                                      const Scalars *S, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= seg.off[world] * comps) return;
    int k = 0;
    while (k + 1 < world && i >= seg.off[k + 1] * comps) k++;
    if (k == rank) return;
    for (int q = 0; q < n_planes; q++) v[q * plane_stride + i] = src.p[k][q * plane_stride + i];
}

// ------------------------------------------------------------------------------------------------
// SE(2) linearisation (restated from pose_graph_optimization.rs:434-486, 516-535; closed forms in
// SURVEY.md Appendix A).  Poses are (x, y, cos, sin): the reference stores the heading as a unit
// complex (g2o.rs:14-16), so no sincos per edge, one atan2.
// One-time (pgo_create): the half-edge measurement stream hz -- laid out exactly like val, NM components per stored block -- is
// gathered on the device from the edge-ordered records `ed` ([NM][ed_stride] planes) through the slot -> edge map, instead of being
// packed on the host and uploaded (640 MB at 1M poses).  One thread per block row, walking its slice like the assembly kernel.
template <int NM>
__global__ void __launch_bounds__(128) k_build_hz(LevelDev L, const int32_t *__restrict__ slot_edge, const double *__restrict__ ed, int64_t ed_stride,
                                                   double *__restrict__ hz) {
    PDL_ENTER();
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int64_t slice = row >> 5;
    if (slice >= L.n_slices) return;
    const int mydeg = L.deg[row];
    const int maxdeg = __shfl_sync(0xffffffffu, mydeg, 0);
    const int64_t base = L.slice_ptr[slice];
    int64_t off = 0;
    for (int k = 0; k < maxdeg; k++) {
        const bool active = k < mydeg;
        const int cnt = __popc(__ballot_sync(0xffffffffu, active));
        if (active) {
            const int64_t e = slot_edge[base + off + lane];
            double *dst = hz + (base + off) * NM + lane;
#pragma unroll
            for (int c = 0; c < NM; c++) dst[(int64_t)c * cnt] = __ldg(ed + (int64_t)c * ed_stride + e);
        }
        off += cnt;
    }
}

// One-time (pgo_create): the edge-ordered measurement records `ed` ([NM][stride] planes; SE2 / XY: z = x y cos sin, Omega upper (6 | 3);
// SE3: z = t(3) q(w,x,y,z) normalised, Omega upper (21)) from the caller's packed g2o-layout arrays (measurement x y theta | x y |
// t(3) q(x,y,z,w); information upper triangle).  Record i is edge inc[i] (all edges in order when inc == nullptr); mofs / iofs are the
// packed offsets of every edge, nullptr when all edges have the same kind.
template <int NM>
__global__ void __launch_bounds__(256) k_build_ed(int64_t n_rec, const int32_t *__restrict__ inc, const uint8_t *__restrict__ ekind,
                                                   const int64_t *__restrict__ mofs, const int64_t *__restrict__ iofs, const double *__restrict__ emeas,
                                                   const double *__restrict__ einfo, double *__restrict__ ed, int64_t stride) {
    PDL_ENTER();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_rec) return;
    const int64_t k = inc ? inc[i] : i;
    const int kind = ekind[k];
    const int nmeas = kind == 2 ? 7 : (kind == 1 ? 2 : 3), ninfo = kind == 2 ? 21 : (kind == 1 ? 3 : 6);
    const double *m = emeas + (mofs ? mofs[k] : k * nmeas), *w = einfo + (iofs ? iofs[k] : k * ninfo);
    double *o = ed + i;
    if (NM == 28) {
        const double nq = sqrt(m[3] * m[3] + m[4] * m[4] + m[5] * m[5] + m[6] * m[6]);
        o[0] = m[0]; o[stride] = m[1]; o[2 * stride] = m[2];
        o[3 * stride] = m[6] / nq; o[4 * stride] = m[3] / nq; o[5 * stride] = m[4] / nq; o[6 * stride] = m[5] / nq;
#pragma unroll
        for (int c = 0; c < 21; c++) o[(int64_t)(7 + c) * stride] = w[c];
    } else {
        double cz = 0.0, sz = 0.0;
        if (kind == 0) sincos(m[2], &sz, &cz);
        o[0] = m[0]; o[stride] = m[1]; o[2 * stride] = cz; o[3 * stride] = sz;
#pragma unroll
        for (int c = 0; c < 6; c++) o[(int64_t)(4 + c) * stride] = c < ninfo ? w[c] : 0.0;
    }
}

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
__global__ void __launch_bounds__(128) k_assemble_se2(LevelDev L, const __grid_constant__ XRef posr, const double *__restrict__ poses, const double *__restrict__ hz,
                                                       double *__restrict__ rvec, int64_t anchor_row, double anchor_w, double lambda) {
    PDL_ENTER();
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
            ld_vec<4>(xgather<4>(posr, cw), xj);
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
            float *vf = L.valf ? L.valf + (base + off) * 9 + lane : nullptr;      // fp32 copy for the cycle's products, written in the same pass
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const double t = to_side ? T[3 * c + r] : T[3 * r + c];
                    v[(int64_t)(3 * r + c) * cnt] = t;
                    if (vf) vf[(int64_t)(3 * r + c) * cnt] = (float)t;
                }
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
        if (L.diagf) { L.diagf[(int64_t)q * L.n_pad + row] = (float)Hd[q]; L.dinvf[(int64_t)q * L.n_pad + row] = (float)Di[q]; }
    }
    double out[4] = {-g[0], -g[1], -g[2], 0.0};                                                 // b = -g (:361)
    st_vec<4>(rvec + row * 4, out);
    L.pos[row] = xi[0]; L.pos[L.n_pad + row] = xi[1];
}

// global_error (:537-574): one thread per edge this rank owns (edges whose `from` vertex it owns),
// edge-ordered SoA copy of the measurements.
// ed: [10][n_edges] planes (z: x y cos sin ; Omega upper 6 -- for XY edges w11 w12 w22 in the first three)
// ends: .x = local row of `from`, .y = column word of `to` (COL_EDGE_XY marks a pose-landmark edge)
__global__ void __launch_bounds__(256) k_chi2_se2(int64_t n_edges, int64_t ed_stride, const uint2 *__restrict__ ends, const double *__restrict__ ed,
                                                   const double *__restrict__ poses, const __grid_constant__ XRef posr, Scalars *S, double *partials) {
    PDL_ENTER();
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double c = 0.0;
    if (k < n_edges) {
        const uint2 en = ends[k];
        double x1[4], x2[4], z[4], w[6];
        const bool xy = (en.y & COL_EDGE_XY) != 0;
        ld_vec<4>(poses + (int64_t)en.x * 4, x1);
        ld_vec<4>(xgather<4>(posr, en.y), x2);
#pragma unroll
        for (int q = 0; q < 4; q++) z[q] = __ldg(ed + (int64_t)q * ed_stride + k);
#pragma unroll
        for (int q = 0; q < 6; q++) w[q] = __ldg(ed + (int64_t)(4 + q) * ed_stride + k);
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
    reduce_and_finalize<256, FIN_CHI2>(&c, S, partials, 0, blockIdx.x, gridDim.x);
}

// update_nodes (:229-245): t += dx.xy (global frame), r <- r * (cos dth, sin dth) without
// renormalisation; landmarks l += dx.  Also ||dx||^2 (:273).  sign = -1 undoes a step (:277).
__global__ void __launch_bounds__(256) k_retract_se2(LevelDev L, double *__restrict__ poses, const double *__restrict__ dx, double sign,
                                                      Scalars *S, double *partials) {
    PDL_ENTER();
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
    reduce_and_finalize<256, FIN_NORM>(&n2, S, partials, 0, blockIdx.x, gridDim.x);
}

// pgo_set_poses / pgo_get_poses: g2o-layout vertex values (x y theta | x y, lut order, packed) <-> the
// 32-byte device pose records (x, y, cos, sin) in storage order.  iso2 (g2o.rs:14-16) on the way in,
// atan2(im, re) on the way out.  row_valofs is relative to the first value this rank owns.
__global__ void __launch_bounds__(256) k_import_poses(int64_t n, const int64_t *__restrict__ row_valofs, const uint8_t *__restrict__ vkind,
                                                       const double *__restrict__ values, double *__restrict__ poses) {
    PDL_ENTER();
    const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= n) return;
    const double *v = values + row_valofs[row];
    double p[4] = {v[0], v[1], 1.0, 0.0};
    if (vkind[row] == 0) sincos(v[2], &p[3], &p[2]);
    st_vec<4>(poses + row * 4, p);
}
__global__ void __launch_bounds__(256) k_export_poses(int64_t n, const int64_t *__restrict__ row_valofs, const uint8_t *__restrict__ vkind,
                                                       const double *__restrict__ poses, double *__restrict__ values) {
    PDL_ENTER();
    const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= n) return;
    double p[4];
    ld_vec<4>(poses + row * 4, p);
    double *v = values + row_valofs[row];
    v[0] = p[0]; v[1] = p[1];
    if (vkind[row] == 0) v[2] = atan2(p[3], p[2]);
}

// ================================================================================================
// SE(3) -- repo-defined semantics, parity unpinned: the reference parses SE3 graphs but optimize is todo!() for them
// (pose_graph_optimization.rs:241, 357, 570); SURVEY 8(c) fixes the natural extension of the SE2 convention, restated
// on the CPU in oracle/pgo_oracle.c (se3_error_jac):
//   E = Z^-1 X1^-1 X2 ;  e = [ Rz^T (R1^T (t2 - t1) - tz) ; Log(qz^-1 q1^-1 q2) ]  in R^6
//   retraction  t += dt (global frame),  q <- q Exp(dw) (body frame), renormalised
//   A = [[-Rz^T R1^T, Rz^T [R1^T d]x], [0, -Jr^-1(e_w) R2^T R1]]    B = [[Rz^T R1^T, 0], [0, Jr^-1(e_w)]]
// Pose record: 8 doubles (x, y, z, 0, qw, qx, qy, qz) = two 32-byte sectors; measurement: t(3), q(w,x,y,z), upper
// triangle of Omega (21), streamed component-major like the SE2 one.
__device__ __forceinline__ void quat_mul(const double *a, const double *b, double *o) {
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
__device__ __forceinline__ void mat3_mul(const double *a, const double *b, double *o) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
__device__ __forceinline__ void mat3_tmul(const double *a, const double *b, double *o) {    // a^T b
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) o[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}

// x1, x2: pose records (8 doubles) ; z: t(3), q(4).  e[6]; with JAC also the 3x3 pieces of A and B:
//   M = Rz^T R1^T, T = Rz^T [u]x, JR = Jri R2^T R1, Jri = Jr^-1(e_w)
template <bool JAC>
__device__ __forceinline__ void se3_edge(const double *x1, const double *x2, const double *z, double *e, double *M, double *T, double *JR, double *Jri) {
    double R1[9], Rz[9];
    quat_to_R(x1 + 4, R1); quat_to_R(z + 3, Rz);
    const double d0 = x2[0] - x1[0], d1 = x2[1] - x1[1], d2 = x2[2] - x1[2];
    double u[3], w[3];
#pragma unroll
    for (int i = 0; i < 3; i++) { u[i] = R1[i] * d0 + R1[3 + i] * d1 + R1[6 + i] * d2; w[i] = u[i] - z[i]; }
#pragma unroll
    for (int i = 0; i < 3; i++) e[i] = Rz[i] * w[0] + Rz[3 + i] * w[1] + Rz[6 + i] * w[2];
    const double qzi[4] = {z[3], -z[4], -z[5], -z[6]}, q1i[4] = {x1[4], -x1[5], -x1[6], -x1[7]};
    double t[4], qe[4];
    quat_mul(qzi, q1i, t); quat_mul(t, x2 + 4, qe);
    {   // Log
        double qw = qe[0], qx = qe[1], qy = qe[2], qz = qe[3];
        if (qw < 0) { qw = -qw; qx = -qx; qy = -qy; qz = -qz; }
        const double n = sqrt(qx * qx + qy * qy + qz * qz);
        const double k = (n < 1e-12) ? 2.0 / qw : 2.0 * atan2(n, qw) / n;
        e[3] = k * qx; e[4] = k * qy; e[5] = k * qz;
    }
    if (!JAC) return;
    double R2[9], R1Rz[9];
    quat_to_R(x2 + 4, R2);
    mat3_mul(R1, Rz, R1Rz);                  // (R1 Rz)^T = Rz^T R1^T
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) M[3 * i + j] = R1Rz[3 * j + i];
    const double Su[9] = {0.0, -u[2], u[1], u[2], 0.0, -u[0], -u[1], u[0], 0.0};
    mat3_tmul(Rz, Su, T);
    {   // inverse right Jacobian of SO(3): I + 1/2 [p]x + k [p]x^2
        const double p0 = e[3], p1 = e[4], p2 = e[5];
        const double th2 = p0 * p0 + p1 * p1 + p2 * p2, th = sqrt(th2);
        const double k = (th < 1e-5) ? (1.0 / 12.0 + th2 / 720.0) : (1.0 / th2 - (1.0 + cos(th)) / (2.0 * th * sin(th)));
        const double Sp[9] = {0.0, -p2, p1, p2, 0.0, -p0, -p1, p0, 0.0};
        double S2[9];
        mat3_mul(Sp, Sp, S2);
#pragma unroll
        for (int i = 0; i < 9; i++) Jri[i] = 0.5 * Sp[i] + k * S2[i];
        Jri[0] += 1.0; Jri[4] += 1.0; Jri[8] += 1.0;
    }
    double R2tR1[9];
    mat3_tmul(R2, R1, R2tR1);
    mat3_mul(Jri, R2tR1, JR);
}

__device__ __forceinline__ void sym6_expand(const double *w, double *W) {     // upper triangle (21, row-major) -> full 6x6
    int p = 0;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = r; c < 6; c++) { W[6 * r + c] = w[p]; W[6 * c + r] = w[p]; p++; }
}

// Fused linearise + assemble + block-Jacobi setup for SE3 graphs (6x6 blocks); same structure as k_assemble_se2: one
// thread per block row walks its half edges, single writer per block, no atomics.
__global__ void __launch_bounds__(128) k_assemble_se3(LevelDev L, const double *__restrict__ poses, const double *__restrict__ hz,
                                                       double *__restrict__ rvec, int64_t anchor_row, double anchor_w, double lambda) {
    PDL_ENTER();
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int64_t slice = row >> 5;
    if (slice >= L.n_slices) return;
    const int mydeg = L.deg[row];
    const int maxdeg = __shfl_sync(0xffffffffu, mydeg, 0);
    double xi[8];
    ld_vec<8>(poses + row * 8, xi);
    // the diagonal block accumulates in shared memory (component-major: conflict-free), not in 72 registers: the edge loop's
    // 6x6 algebra needs them (255 registers + 0.5 KB of spills before)
    __shared__ double sHd[36][128];
    double g[6];
#pragma unroll
    for (int q = 0; q < 36; q++) sHd[q][threadIdx.x] = 0.0;
#pragma unroll
    for (int q = 0; q < 6; q++) g[q] = 0.0;
    const int64_t base = L.slice_ptr[slice];
    int64_t off = 0;
    for (int k = 0; k < maxdeg; k++) {
        const bool active = k < mydeg;
        const int cnt = __popc(__ballot_sync(0xffffffffu, active));
        if (active) {
            const uint32_t cw = __ldg(L.col + base + off + lane);
            const bool to_side = (cw & COL_ROLE_TO) != 0;
            double xj[8], z[7], wu[21];
            ld_vec<8>(poses + (int64_t)(cw & COL_LOCAL_MASK) * 8, xj);
            const double *m = hz + (base + off) * 28 + lane;
#pragma unroll
            for (int q = 0; q < 7; q++) z[q] = __ldg(m + (int64_t)q * cnt);
#pragma unroll
            for (int q = 0; q < 21; q++) wu[q] = __ldg(m + (int64_t)(7 + q) * cnt);
            double e[6], M[9], T[9], JR[9], Jri[9];
            se3_edge<true>(to_side ? xj : xi, to_side ? xi : xj, z, e, M, T, JR, Jri);
            // J = Jacobian w.r.t. this row's pose, K = w.r.t. the neighbour's (A / B of the edge, or B / A)
            double J[36], K[36];
#pragma unroll
            for (int q = 0; q < 36; q++) { J[q] = 0.0; K[q] = 0.0; }
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const double a_tt = -M[3 * i + j], a_tw = T[3 * i + j], a_ww = -JR[3 * i + j];
                    const double b_tt = M[3 * i + j], b_ww = Jri[3 * i + j];
                    J[6 * i + j] = to_side ? b_tt : a_tt;
                    J[6 * i + 3 + j] = to_side ? 0.0 : a_tw;
                    J[6 * (3 + i) + 3 + j] = to_side ? b_ww : a_ww;
                    K[6 * i + j] = to_side ? a_tt : b_tt;
                    K[6 * i + 3 + j] = to_side ? a_tw : 0.0;
                    K[6 * (3 + i) + 3 + j] = to_side ? a_ww : b_ww;
                }
            double W[36], WJ[36];
            sym6_expand(wu, W);
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    double s = 0.0;
#pragma unroll
                    for (int q = 0; q < 6; q++) s = fma(W[6 * r + q], J[6 * q + c], s);
                    WJ[6 * r + c] = s;
                }
            // own diagonal contribution J^T W J, gradient J^T W e, off-diagonal block J^T W K = (W J)^T K
#pragma unroll
            for (int r = 0; r < 6; r++) {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < 6; q++) s = fma(WJ[6 * q + r], e[q], s);
                g[r] += s;
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    double h = 0.0;
#pragma unroll
                    for (int q = 0; q < 6; q++) h = fma(J[6 * q + r], WJ[6 * q + c], h);
                    sHd[6 * r + c][threadIdx.x] += h;
                }
            }
            double *v = L.val + (base + off) * 36 + lane;
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    double h = 0.0;
#pragma unroll
                    for (int q = 0; q < 6; q++) h = fma(WJ[6 * q + r], K[6 * q + c], h);
                    v[(int64_t)(6 * r + c) * cnt] = h;
                    if (L.valf) L.valf[(base + off) * 36 + lane + (int64_t)(6 * r + c) * cnt] = (float)h;
                }
        }
        off += cnt;
    }
    const bool real = row < L.n;
    double Hd[36];
#pragma unroll
    for (int q = 0; q < 36; q++) Hd[q] = sHd[q][threadIdx.x];
    if (real) {
#pragma unroll
        for (int a = 0; a < 6; a++) Hd[7 * a] += lambda + (row == anchor_row ? anchor_w : 0.0);
    }
    double Di[36];
    if (real) inv_block<6>(Hd, Di);
    else {
#pragma unroll
        for (int q = 0; q < 36; q++) Di[q] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < 36; q++) {
        L.diag[(int64_t)q * L.n_pad + row] = Hd[q];
        L.dinv[(int64_t)q * L.n_pad + row] = Di[q];
        if (L.diagf) { L.diagf[(int64_t)q * L.n_pad + row] = (float)Hd[q]; L.dinvf[(int64_t)q * L.n_pad + row] = (float)Di[q]; }
    }
    double out[6] = {-g[0], -g[1], -g[2], -g[3], -g[4], -g[5]};
    st_vec<6>(rvec + row * 6, out);
    L.pos[row] = xi[0]; L.pos[L.n_pad + row] = xi[1]; L.pos[2 * L.n_pad + row] = xi[2];
}

// chi2 of an SE3 graph: ed = [28][n_edges] planes, ends as in k_chi2_se2
__global__ void __launch_bounds__(256) k_chi2_se3(int64_t n_edges, int64_t ed_stride, const uint2 *__restrict__ ends, const double *__restrict__ ed,
                                                   const double *__restrict__ poses, Scalars *S, double *partials) {
    PDL_ENTER();
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double c = 0.0;
    if (k < n_edges) {
        const uint2 en = ends[k];
        double x1[8], x2[8], z[7], wu[21], W[36], e[6];
        ld_vec<8>(poses + (int64_t)en.x * 8, x1);
        ld_vec<8>(poses + (int64_t)(en.y & COL_LOCAL_MASK) * 8, x2);
#pragma unroll
        for (int q = 0; q < 7; q++) z[q] = __ldg(ed + (int64_t)q * ed_stride + k);
#pragma unroll
        for (int q = 0; q < 21; q++) wu[q] = __ldg(ed + (int64_t)(7 + q) * ed_stride + k);
        se3_edge<false>(x1, x2, z, e, nullptr, nullptr, nullptr, nullptr);
        sym6_expand(wu, W);
#pragma unroll
        for (int cc = 0; cc < 6; cc++) {
            double t = 0.0;
#pragma unroll
            for (int r = 0; r < 6; r++) t = fma(e[r], W[6 * r + cc], t);
            c = fma(t, e[cc], c);
        }
    }
    reduce_and_finalize<256, FIN_CHI2>(&c, S, partials, 0, blockIdx.x, gridDim.x);
}

// t += dt ; q <- normalise(q Exp(dw)) ; ||dx||^2
__global__ void __launch_bounds__(256) k_retract_se3(LevelDev L, double *__restrict__ poses, const double *__restrict__ dx, double sign,
                                                      Scalars *S, double *partials) {
    PDL_ENTER();
    const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double n2 = 0.0;
    if (row < L.n) {
        double p[8], d[6];
        ld_vec<8>(poses + row * 8, p);
        ld_vec<6>(dx + row * 6, d);
#pragma unroll
        for (int a = 0; a < 6; a++) n2 = fma(d[a], d[a], n2);
        p[0] = fma(sign, d[0], p[0]); p[1] = fma(sign, d[1], p[1]); p[2] = fma(sign, d[2], p[2]);
        const double w0 = sign * d[3], w1 = sign * d[4], w2 = sign * d[5];
        const double th = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
        double sh, ch;
        sincos(0.5 * th, &sh, &ch);
        const double k = (th < 1e-12) ? 0.5 - th * th / 48.0 : sh / th;
        const double dq[4] = {ch, k * w0, k * w1, k * w2};
        double q[4];
        quat_mul(p + 4, dq, q);
        const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        p[4] = q[0] / n; p[5] = q[1] / n; p[6] = q[2] / n; p[7] = q[3] / n;
        st_vec<8>(poses + row * 8, p);
    }
    reduce_and_finalize<256, FIN_NORM>(&n2, S, partials, 0, blockIdx.x, gridDim.x);
}

// g2o-layout SE3 vertex values (x y z qx qy qz qw) <-> pose records (x, y, z, 0, qw, qx, qy, qz), quaternion normalised on the way in
__global__ void __launch_bounds__(256) k_import_poses_se3(int64_t n, const int64_t *__restrict__ row_valofs, const double *__restrict__ values,
                                                           double *__restrict__ poses) {
    PDL_ENTER();
    const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= n) return;
    const double *v = values + row_valofs[row];
    const double nq = sqrt(v[3] * v[3] + v[4] * v[4] + v[5] * v[5] + v[6] * v[6]);
    const double p[8] = {v[0], v[1], v[2], 0.0, v[6] / nq, v[3] / nq, v[4] / nq, v[5] / nq};
    st_vec<8>(poses + row * 8, p);
}
__global__ void __launch_bounds__(256) k_export_poses_se3(int64_t n, const int64_t *__restrict__ row_valofs, const double *__restrict__ poses,
                                                           double *__restrict__ values) {
    PDL_ENTER();
    const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= n) return;
    double p[8];
    ld_vec<8>(poses + row * 8, p);
    double *v = values + row_valofs[row];
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[5]; v[4] = p[6]; v[5] = p[7]; v[6] = p[4];
}

} // namespace pgo
