// Cross-rank synchronisation over NVLink peer memory (sharded handles): every rank owns a Comm block at the start of its
// peer-visible arena.  All-reduces are fused into the kernel that produces the partial sums (kernels.cuh: reduce_and_finalize ->
// xrank_exchange, run by the kernel's last block); what is left here is the plain stage barrier.
#pragma once
#include "kernels.cuh"

namespace pgo {

// stage barrier between ranks (no sums): every rank's earlier writes are visible to every rank's later kernels
__global__ void __launch_bounds__(32) k_xbarrier(Scalars *S, int check_done) {
    PDL_ENTER();
    if (check_done && ld_done(S)) return;
    double none[1];
    xrank_exchange<0>(S, none);
}

// last node of the body of the device-side PCG loop (a CUDA-graph WHILE node): keep iterating until a kernel has set `done`
__global__ void __launch_bounds__(32) k_loop_cond(cudaGraphConditionalHandle hnd, const Scalars *S) {
    PDL_ENTER();
    if (threadIdx.x == 0) cudaGraphSetConditional(hnd, ld_done(S) ? 0u : 1u);
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
