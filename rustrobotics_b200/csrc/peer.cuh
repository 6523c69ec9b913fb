// Cross-rank synchronisation over NVLink peer memory (sharded mode, one process per GPU).
//
// Every rank owns a Comm block at the start of its peer-visible arena (CUDA IPC).  A stage barrier with an optional
// all-reduce of up to three partial sums is ONE single-CTA kernel: thread t writes this rank's partials and then the
// current epoch into rank t's Comm (st.release.sys), then spins on its own Comm until rank t's epoch arrives
// (ld.acquire.sys); thread 0 adds the partials in rank order (identical bits on every rank) and runs the same scalar
// recurrence a single GPU runs in the "last block finalises" step.  Because kernels of one stream run in order, every
// kernel launched after it sees all peers' earlier writes, and no peer can be more than one epoch ahead (values are
// double-buffered by epoch parity).  A spin that exceeds ~20 s sets ST_COMM and ends the solve instead of hanging.
#pragma once
#include "kernels.cuh"

namespace pgo {

struct Comm {
    unsigned long long flags[MAX_RANKS];
    double vals[2][MAX_RANKS][4];
};
struct CommRef { Comm *p[MAX_RANKS]; };

template <int FIN>
__global__ void __launch_bounds__(32) k_xreduce(Comm *mine, CommRef peers, int rank, int world, Scalars *S, int lvl, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    constexpr int NV = fin_ndot(FIN);
    __shared__ unsigned long long ep;
    __shared__ int bad;
    const int t = threadIdx.x;
    if (t == 0) { ep = S->epoch + 1; S->epoch = ep; bad = 0; }
    __syncthreads();
    const unsigned long long epoch = ep;
    const int slot = (int)(epoch & 1);
    if (t < world) {
        Comm *p = peers.p[t];
#pragma unroll
        for (int k = 0; k < NV; k++) *((volatile double *)&p->vals[slot][rank][k]) = S->loc[k];
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&p->flags[rank]), "l"(epoch) : "memory");
        const long long t0 = clock64();
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(&mine->flags[t]) : "memory");
            if (v < epoch && clock64() - t0 > 40000000000ll) { bad = 1; break; }
        } while (v < epoch);
    }
    __syncthreads();
    if (t == 0) {
        if (bad) { S->status = ST_COMM; S->done = 1; __threadfence(); }
        else if (NV > 0) {
            double total[NV > 0 ? NV : 1];
#pragma unroll
            for (int k = 0; k < NV; k++) {
                double s = 0.0;
                for (int r = 0; r < world; r++) s += *((volatile double *)&mine->vals[slot][r][k]);
                total[k] = s;
            }
            finalize(FIN, S, total, lvl);
        }
    }
}

// {a.b (, b.c)} over the local rows -> FIN (used when the preconditioner itself has no kernel to fuse the dots into)
template <int D, int FIN>
__global__ void __launch_bounds__(128) k_dots(int64_t n_pad, const double *__restrict__ a, const double *__restrict__ b,
                                               const double *__restrict__ c, Scalars *S, double *partials, int lvl, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    constexpr int VS = VecStride<D>::value;
    const int64_t row = (int64_t)blockIdx.x * 128 + threadIdx.x;
    double dots[2] = {0.0, 0.0};
    if (row < n_pad) {
        double av[VS], bv[VS];
        ld_vec<VS>(a + row * VS, av);
        ld_vec<VS>(b + row * VS, bv);
#pragma unroll
        for (int q = 0; q < D; q++) dots[0] = fma(av[q], bv[q], dots[0]);
        if (FIN == FIN_RZ) {
            double cv[VS];
            ld_vec<VS>(c + row * VS, cv);
#pragma unroll
            for (int q = 0; q < D; q++) dots[1] = fma(cv[q], bv[q], dots[1]);
        }
    }
    reduce_and_finalize<128, FIN>(dots, S, partials, lvl, blockIdx.x, gridDim.x);
}

// v *= s
__global__ void __launch_bounds__(256) k_scale(int64_t n_doubles, double *__restrict__ v, double s) {
    PDL_ENTER();
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 2;
    if (i >= n_doubles) return;
    double2 t = *reinterpret_cast<double2 *>(v + i);
    t.x *= s; t.y *= s;
    *reinterpret_cast<double2 *>(v + i) = t;
}

} // namespace pgo
