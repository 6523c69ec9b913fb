// C ABI (include/pgo_b200.h) + device-resident Gauss-Newton / PCG driver.
//
// One pgo_gn_step = the body of the reference's optimisation loop
// (pose_graph_optimization.rs:271-274): build_linear_system + solve + update_nodes + global_error,
// with the UMFPACK factorisation replaced by a preconditioned (flexible) conjugate gradient that runs entirely
// on the GPU: the PCG iterations are captured once into a CUDA graph; every kernel tests a device
// `done` flag, so the host only polls a pinned copy of the scalars once per graph launch.
// Preconditioner: block-Jacobi, or an aggregation-AMG K-cycle (two Krylov-accelerated coarse solves per level,
// Notay) over rigid-motion coarse spaces with an explicit dense inverse on the coarsest level.
// Sharded mode (world > 1): every rank owns a contiguous vertex range of every level; neighbour rows are read
// directly from peer HBM over NVLink (CUDA IPC), dot products and stage barriers are one tiny kernel that
// exchanges partial sums through peer memory.
#include <cuda.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/pgo_b200.h"
#include "kernels.cuh"
#include "peer.cuh"

using namespace pgo;

namespace {

thread_local std::string g_create_error;
const int KDIM[3] = {3, 2, 6};     // scalar dimension of a vertex kind (g2o.rs:61,68,77)

#define NEED_DEVICE(h)                                                                             \
    do {                                                                                           \
        if (!(h)->stream) { (h)->err = "structure-only handle: no device state (there is no CPU fallback)"; return PGO_ERR_CUDA; } \
        if ((h)->world > 1 && !(h)->connected) { (h)->err = "sharded handle: call pgo_shard_connect first"; return PGO_ERR_ARG; } \
    } while (0)

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return PGO_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

struct LevelBuf {
    LevelDev d{};
    bool jds = false, kcycle = false;
    bool repl = false;                 // sharded handle: this level is replicated on every rank (no peer traffic inside it)
    bool first_repl = false;           // ... and the finer level is sharded: rhs / val / diag / pos rows are gathered from the peers
    SegMap src_rows{}, src_slots{};    // first_repl: row / stored-block ranges produced by each rank
    // sharded level: rows of other ranks that this rank's blocks reference (the halo).  Their vector records are pulled
    // from peer HBM into slots n_pad .. n_pad + n_halo of the local vector before a kernel gathers from it.
    uint32_t *halo_src = nullptr;      // [n_halo] (owner rank << COL_OWNER_SHIFT) | row local to the owner
    int64_t n_halo = 0;
    int64_t vec_rows = 0;              // rows (own, padded to the largest partition, + halo) every vector of the level has room for
    double *rhs = nullptr, *sol = nullptr, *xa = nullptr, *res = nullptr;      // cycle work vectors
    double *c1 = nullptr, *c2 = nullptr, *c3 = nullptr, *v1 = nullptr, *v2 = nullptr, *r1 = nullptr;   // K-cycle work vectors
    int ksteps = 2;                    // inner flexible-CG steps of the K-cycle at this level (2: Notay's; 3: one more, fully orthogonalised)
    double omega = 0.6;
    int grid128 = 0, gridw = 0, gridv = 0, grid8 = 0, grid4 = 0, lpr = 32;
};

struct ArenaReq { double **p; size_t count; };

} // namespace

struct MultiCtx;

struct pgo_handle {
    // the symbolic pass is built once per graph; the shards of a single-process multi-GPU handle share it (read-only)
    std::shared_ptr<Symbolic> symp;
    Symbolic &sym;
    explicit pgo_handle(std::shared_ptr<Symbolic> s = std::make_shared<Symbolic>()) : symp(std::move(s)), sym(*symp) {}
    MultiCtx *multi = nullptr;         // single-process multi-GPU handle (pgo_options.n_gpus > 1): this handle only dispatches to its shards
    bool ipc_peer[MAX_RANKS]{};        // peer_base[k] was opened with cudaIpcOpenMemHandle (process-per-GPU mode)
    bool shares_device = false;        // another shard of this handle's graph runs on the same GPU (a test vehicle)
    pgo_options opt{};
    int device = 0, world = 1, rank = 0;
    bool connected = false;
    cudaStream_t stream = nullptr;
    std::vector<LevelBuf> lv;
    std::vector<void *> allocs;
    std::vector<uint32_t> to_index;    // transient (pgo_create): index of every owned edge's `to` pose in the halo-extended pose array
    size_t device_bytes = 0;
    // peer-visible arena: identical layout on every rank
    char *arena = nullptr;
    size_t arena_bytes = 0;
    char *peer_base[MAX_RANKS]{};
    std::vector<ArenaReq> arena_reqs;
    Comm *comm = nullptr;
    // level-0 state (local rows)
    int64_t n_loc = 0, n_pad_loc = 0, n_edges_loc = 0, row0 = 0;
    int64_t ed_stride = 1;             // edges whose records `ed` holds (the n_edges_loc owned ones first): plane stride of ed
    double *poses = nullptr, *poses_saved = nullptr, *hz = nullptr, *ed = nullptr;
    double *vstage = nullptr;          // g2o-layout vertex values (n_values) for set/get_poses
    int64_t *row_valofs = nullptr;     // [n_loc] offset of each local row's values in vstage
    uint2 *ends = nullptr;
    double *x = nullptr, *r = nullptr, *p = nullptr, *q = nullptr, *z = nullptr;
    double *b0 = nullptr, *xs = nullptr;   // pgo_options.refine: the right-hand side b (the solve consumes h->r) and the first solution
    Scalars *S = nullptr, *hS = nullptr;   // device / pinned host (2 slots)
    double *partials = nullptr;
    double *gstage = nullptr;          // staging buffer of the deterministic Galerkin product (largest level)
    double *Ainv = nullptr, *Awork = nullptr;   // explicit inverse of the coarsest matrix; second buffer of the ping-pong inversion
    DenseMap dmap{};
    int dense_m = 0, invert_grid = 0;
    double omega_rho = 1.5;            // PGO_OMEGA_RHO: damping of the Jacobi smoother times the estimated spectral radius
    double *gj_pnext = nullptr;        // look-ahead pivot inverses [2][32 x 32]
    unsigned *gj_bar = nullptr;        // arrival counter of the inversion kernel's grid barrier (monotonic)
    unsigned gj_bar_base = 0;
    bool use_amg = false, omega_ready = false;
    int spmv_tma64 = 0, spmv_tma32 = 0; // PGO_SPMV_TMA64 / PGO_SPMV_TMA32: ring depth of the TMA-staged sliced SpMV (0: register-staged kernel)
    int64_t lpr4_min_rows = 16384;     // PGO_LPR4_MIN_ROWS
    bool pdl = true;                   // programmatic dependent launch of every kernel (PGO_PDL=0 disables)
    cudaError_t launch_err = cudaSuccess;
    bool lowp = false;                 // the cycle's SpMVs read fp32 copies of the stored blocks (opt.amg_fp64_storage == 0)
    int64_t anchor_row = -1;
    cudaGraphExec_t pcg_graph = nullptr;
    bool opt_while = true;             // the whole PCG loop is ONE graph launch, a WHILE conditional node iterating on the device (PGO_WHILE=0:
                                       // chunks of `chunk` iterations per launch, the host polling a pinned copy of the scalars)
    bool opt_while_sharded = false;    // PGO_WHILE=2
    bool pcg_while = false;            // ... and that is what pcg_graph holds
    int chk = 1;                       // `check_done` argument of the kernels of a PCG iteration: 0 while the WHILE-loop body is captured (nothing
                                       // runs behind the convergence flag there, and the flag's load is a dependent L2 round trip at the top of every kernel)
    int chunk = 8;
    int64_t launches_per_iter = 0;
    cudaEvent_t ev[PGO_NUM_PHASES + 2]{}, poll_ev[2]{};
    double ms[PGO_NUM_PHASES]{};
    int64_t launches[PGO_NUM_PHASES]{};
    int64_t launch_count = 0;
    bool have_step = false;
    std::string err;
};

namespace {

template <typename T> int dalloc(pgo_handle *h, T **p, size_t count, bool zero = true) {
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CK(cudaMalloc((void **)p, bytes));
    h->allocs.push_back((void *)*p);
    h->device_bytes += bytes;
    if (zero) CK(cudaMemsetAsync(*p, 0, bytes, h->stream));
    return PGO_OK;
}
template <typename T> int upload(pgo_handle *h, T **p, const std::vector<T> &v) {
    int rc = dalloc(h, p, v.size(), false);
    if (rc) return rc;
    if (!v.empty()) CK(cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return PGO_OK;
}

// Every kernel is launched through here.  With programmatic dependent launch (PGO_PDL, default on) the launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization: the grid is scheduled while its predecessor drains and blocks in
// griddepcontrol.wait (PDL_ENTER, the first statement of every kernel) until the predecessor has completed and flushed,
// so the ~2-3 us launch ramp of the ~70 small dependent kernels of one PCG iteration overlaps the previous kernel.
template <typename... KArgs, typename... Args>
inline void launch_k(pgo_handle *h, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = h->pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
    if (e != cudaSuccess && h->launch_err == cudaSuccess) h->launch_err = e;
}

inline int grid_for(int64_t n, int bs) { return (int)std::max<int64_t>(1, (n + bs - 1) / bs); }

// peer-visible vectors are carved from one arena whose layout is identical on every rank (sizes use the
// largest partition of the level), so that rank k's copy of a vector is peer_base[k] + the same offset
inline void arena_request(pgo_handle *h, double **p, size_t count) { h->arena_reqs.push_back({p, count}); }

int arena_commit(pgo_handle *h) {
    size_t off = (sizeof(Comm) + 255) / 256 * 256;
    std::vector<size_t> offs;
    for (auto &rq : h->arena_reqs) { offs.push_back(off); off += (rq.count * sizeof(double) + 255) / 256 * 256; }
    h->arena_bytes = off;
    CK(cudaMalloc((void **)&h->arena, off));
    h->allocs.push_back(h->arena);
    h->device_bytes += off;
    CK(cudaMemsetAsync(h->arena, 0, off, h->stream));
    for (size_t i = 0; i < offs.size(); i++) *h->arena_reqs[i].p = (double *)(h->arena + offs[i]);
    h->comm = (Comm *)h->arena;
    for (int k = 0; k < MAX_RANKS; k++) h->peer_base[k] = nullptr;
    h->peer_base[h->rank] = h->arena;
    return PGO_OK;
}

// the same vector on every rank (a vector of a replicated level is only ever read locally)
XRef xref(const pgo_handle *h, const double *local, bool repl = false) {
    XRef x{};
    const size_t off = (const char *)local - h->arena;
    for (int k = 0; k < MAX_RANKS; k++) x.p[k] = (k < h->world && !repl) ? (const double *)(h->peer_base[k] + off) : local;
    return x;
}

// ---- cross-rank stage barrier / all-reduce of the partial sums a kernel left in S->loc (world > 1 only)
// All-reduces (FIN != FIN_NONE) happen inside the kernel that produced the partial sums: its last block exchanges them with the
// peers' last blocks (kernels.cuh: reduce_and_finalize), which is also a barrier.  Only the plain barrier is a kernel of its own.
template <int FIN> void xreduce(pgo_handle *h, int lvl, int check_done) {
    if (FIN != FIN_NONE || h->world == 1 || h->lv[lvl].repl) return;
    launch_k(h, k_xbarrier, 1, 32, 0, h->S, check_done);
    h->launch_count += 1;
}
inline void xbarrier(pgo_handle *h, int check_done = 1) { xreduce<FIN_NONE>(h, 0, check_done); }
inline void lbarrier(pgo_handle *h, int lvl, int check_done = 1) { xreduce<FIN_NONE>(h, lvl, check_done); }   // no-op on replicated levels

// halo exchange of a sharded level's vector (stride doubles per row): ONE bulk kernel of independent peer reads over NVLink
// (latency-tolerant), after which every gather of the following kernel is local
void halo_pull(pgo_handle *h, int l, const double *v, int stride, int check_done) {
    LevelBuf &B = h->lv[l];
    if (h->world == 1 || B.repl || B.n_halo == 0) return;
    double *w = const_cast<double *>(v);
    launch_k(h, k_halo_pull, grid_for(B.n_halo, 128), 128, 0, w, xref(h, v), B.halo_src, B.n_halo, B.d.n_pad, stride, h->S, check_done);
    h->launch_count += 1;
}

// first replicated level: pull the rows the other ranks produced (planes of `stride` doubles, `comps` doubles per row/slot)
void gather_rows(pgo_handle *h, double *v, const SegMap &seg, int comps, int n_planes, int64_t plane_stride, int check_done) {
    const int64_t total = (int64_t)seg.off[h->world] * comps;
    if (total == 0) return;
    launch_k(h, k_gather_peer, grid_for(total, 256), 256, 0, v, xref(h, v), seg, h->rank, h->world, comps, n_planes, plane_stride, h->S, check_done);
    h->launch_count += 1;
}

// ---- SpMV launcher: sliced storage (level 0 and large coarse levels) or block CSR
template <int D, int MODE, int FIN, typename VT> void spmv_launch(pgo_handle *h, int l, const double *x, const double *r, double *y, double omega,
                                                                  const double *u1, const double *u2, int check) {
    LevelBuf &B = h->lv[l];
    const XRef xr = xref(h, x, true);
    if (B.jds) {
        const int ns = std::is_same<VT, float>::value ? h->spmv_tma32 : h->spmv_tma64;
        if (ns > 0) {                                // TMA-staged variant: ring of ns columns per warp in shared memory
            const size_t smem = spmv_tma_smem<D, VT>(ns);
            if (smem > 48 * 1024) cudaFuncSetAttribute(k_spmv_tma<D, MODE, FIN, VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            launch_k(h, k_spmv_tma<D, MODE, FIN, VT>, B.grid128, 128, smem, B.d, x, r, y, omega, u1, u2, h->S, h->partials, l, check, ns);
        } else launch_k(h, k_spmv<D, MODE, FIN, false, VT>, B.grid128, 128, 0, B.d, xr, x, r, y, omega, u1, u2, h->S, h->partials, l, check);
    }
    else if (B.lpr == 4) launch_k(h, k_spmv_csr<D, MODE, FIN, false, 4, VT>, B.grid4, 256, 0, B.d, xr, x, r, y, omega, u1, u2, h->S, h->partials, l, check);
    else if (B.lpr == 8) launch_k(h, k_spmv_csr<D, MODE, FIN, false, 8, VT>, B.grid8, 256, 0, B.d, xr, x, r, y, omega, u1, u2, h->S, h->partials, l, check);
    else launch_k(h, k_spmv_csr<D, MODE, FIN, false, 32, VT>, B.gridw, 256, 0, B.d, xr, x, r, y, omega, u1, u2, h->S, h->partials, l, check);
}
// CYC: the product belongs to the multigrid cycle (the preconditioner), which may read the fp32 copy of the blocks;
// the PCG operator product, the setup and the diagnostics always read the fp64 blocks
// Sharded levels: the product gathers neighbour rows of x from the halo slots behind the rank's own rows.  PULL: fetch them from the
// peers first (the caller's barrier / all-reduce made the peers' x final); false when the caller keeps the halo slots current itself.
// A fused all-reduce (FIN != FIN_NONE) is also a barrier between the ranks; with FIN_NONE the CALLER places the barriers its data
// hazards need (peers may still be pulling x when this returns).
template <int D, int MODE, int FIN, bool CYC = false, bool PULL = true> void spmv(pgo_handle *h, int l, const double *x, const double *r, double *y, double omega,
                                                                                  const double *u1, const double *u2, int check) {
    if (PULL) halo_pull(h, l, x, VecStride<D>::value, check);
    if (CYC && h->lowp) spmv_launch<D, MODE, FIN, float>(h, l, x, r, y, omega, u1, u2, check);
    else spmv_launch<D, MODE, FIN, double>(h, l, x, r, y, omega, u1, u2, check);
    h->launch_count += 1;
}
template <int D, int MODE, bool CYC = false> void spmv_any(pgo_handle *h, int l, const double *x, const double *r, double *y, double omega, int check) {
    spmv<D, MODE, FIN_NONE, CYC>(h, l, x, r, y, omega, nullptr, nullptr, check);
}

template <int D> void coarse_solve(pgo_handle *h, int l, const double *rhs, double *out, bool defer_combine = false);

template <int D> void dense_apply(pgo_handle *h, int l, const double *rhs, double *out) {
    LevelBuf &B = h->lv[l];      // always a local (single-GPU or replicated) level
    launch_k(h, k_dense_apply<D>, grid_for(B.d.n * D, 8), 256, sizeof(double) * h->dense_m, B.d.n, h->dmap, 0, 1, h->dense_m,
                                                                                             h->Ainv, xref(h, rhs, true), out, h->S, h->chk);
    h->launch_count += 1;
}

// ---- one multigrid cycle at level l: out = M_l(rhs).  FINK: dots fused into the last kernel (level 0 only).
// PRE: the pre-smoothing step xa = omega Dinv rhs was already done by the caller (fused into the PCG update kernel)
template <int D, int FINK, bool PRE = false> void cycle(pgo_handle *h, int l, const double *rhs, double *out) {
    LevelBuf &B = h->lv[l];
    const int last = (int)h->lv.size() - 1;
    if (l == last) {
        if (h->sym.dense_coarsest) dense_apply<D>(h, l, rhs, out);
        else {
            // no direct solve possible: a few damped block-Jacobi sweeps
            launch_k(h, k_dinv_apply<D, FIN_NONE>, B.grid128, 128, 0, B.d, rhs, B.xa, B.omega, nullptr, h->S, h->partials, h->chk);
            h->launch_count += 1;
            lbarrier(h, l);
            spmv_any<D, 2, true>(h, l, B.xa, rhs, B.res, B.omega, h->chk);
            lbarrier(h, l);
            spmv_any<D, 2, true>(h, l, B.res, rhs, out, B.omega, h->chk);
            lbarrier(h, l);
        }
        return;
    }
    LevelBuf &C = h->lv[l + 1];
    if (!PRE) {
        launch_k(h, k_dinv_apply<D, FIN_NONE>, B.grid128, 128, 0, B.d, rhs, B.xa, B.omega, nullptr, h->S, h->partials, h->chk);
        h->launch_count += 1;
    }
    lbarrier(h, l);                                  // the peers' xa rows are final before the halo pull
    // no barrier behind this product: xa is only overwritten by the prolongation, and every path to it crosses a barrier (the
    // gather of the first replicated level's right-hand side, or the barriers inside a sharded coarse solve)
    spmv_any<D, 1, true>(h, l, B.xa, rhs, B.res, 0.0, h->chk);
    launch_k(h, k_restrict<D>, grid_for(C.d.n_pad, 16), 256, 0, B.d, C.d, B.res, C.rhs, h->S, h->chk);
    h->launch_count += 1;
    if (C.first_repl) {                              // every rank restricted onto its own aggregates: all-gather the coarse rhs
        lbarrier(h, l);
        gather_rows(h, C.rhs, C.src_rows, VecStride<D>::value, 1, 0, 1);
    }
    // a K-cycle level leaves its two search directions in C.c1 / C.c2; their final combination is folded into the prolongation
    const bool kfold = l + 1 != last && C.kcycle;
    coarse_solve<D>(h, l + 1, C.rhs, C.sol, kfold);
    if (kfold) launch_k(h, k_prolong_k<D>, B.grid128, 128, 0, B.d, C.c1, C.c2, C.ksteps == 3 ? C.c3 : nullptr, B.xa, h->S, l + 1, h->chk);
    else launch_k(h, k_prolong<D>, B.grid128, 128, 0, B.d, C.sol, B.xa, h->S, h->chk);
    h->launch_count += 1;
    lbarrier(h, l);                                  // the peers' xa rows are final before the halo pull
    spmv<D, 2, FINK, true>(h, l, B.xa, rhs, out, B.omega, FINK == FIN_RZ ? h->q : nullptr, nullptr, h->chk);
    if (FINK == FIN_NONE) lbarrier(h, l);            // `out` is final everywhere (FINK != NONE: the fused all-reduce is the barrier)
}

// K-cycle: the coarse system of level l is solved by two flexible-CG steps preconditioned by the cycle of level l
template <int D> void coarse_solve(pgo_handle *h, int l, const double *rhs, double *out, bool defer_combine) {
    LevelBuf &B = h->lv[l];
    const int last = (int)h->lv.size() - 1;
    if (l == last || !B.kcycle) { cycle<D, FIN_NONE>(h, l, rhs, out); return; }
    cycle<D, FIN_NONE>(h, l, rhs, B.c1);
    spmv<D, 0, FIN_K1, true>(h, l, B.c1, nullptr, B.v1, 0.0, rhs, nullptr, h->chk);
    // r1 = rhs - alpha v1, fused with the pre-smoothing step of the second cycle
    launch_k(h, k_kresid_dinv<D, 1>, B.grid128, 128, 0, B.d, rhs, B.v1, nullptr, B.r1, B.xa, B.omega, h->S, l, h->chk);
    h->launch_count += 1;
    cycle<D, FIN_NONE, true>(h, l, B.r1, B.c2);
    spmv<D, 0, FIN_K2, true>(h, l, B.c2, nullptr, B.v2, 0.0, B.v1, B.r1, h->chk);
    const double *c3 = nullptr;
    if (B.ksteps == 3) {                             // third inner step: r2 = r1 - e2 v2 + e1 v1 (in place), c3 = M(r2), dots of c3
        launch_k(h, k_kresid_dinv<D, 2>, B.grid128, 128, 0, B.d, B.r1, B.v1, B.v2, B.r1, B.xa, B.omega, h->S, l, h->chk);
        h->launch_count += 1;
        cycle<D, FIN_NONE, true>(h, l, B.r1, B.c3);
        spmv<D, 0, FIN_K3, true>(h, l, B.c3, B.r1, B.res, 0.0, B.v1, B.v2, h->chk);
        c3 = B.c3;
    }
    if (!defer_combine) {                            // else: the caller's prolongation applies coef1 c1 + coef2 c2 (+ coef3 c3)
        if (c3) launch_k(h, k_kcombine<2>, B.gridv, 256, 0, B.d.n_pad * VecStride<D>::value, B.c1, B.c2, out, h->S, l, c3, h->chk);
        else launch_k(h, k_kcombine<1>, B.gridv, 256, 0, B.d.n_pad * VecStride<D>::value, B.c1, B.c2, out, h->S, l, nullptr, h->chk);
        h->launch_count += 1;
    }
}

template <int D, int FINK, bool PRE = false> void precondition(pgo_handle *h) {   // z = M^-1 r (+ r.z, z.q)
    LevelBuf &B = h->lv[0];
    if (h->use_amg && h->lv.size() > 1) cycle<D, FINK, PRE>(h, 0, h->r, h->z);
    else if (h->use_amg && h->sym.dense_coarsest) {      // the whole system fits the direct solve
        dense_apply<D>(h, 0, h->r, h->z);
        launch_k(h, k_dots<D, FINK>, B.grid128, 128, 0, B.d.n_pad, h->r, h->z, h->q, h->S, h->partials, 0, h->chk);
        h->launch_count += 1;
        xreduce<FINK>(h, 0, 1);
    } else {
        launch_k(h, k_dinv_apply<D, FINK>, B.grid128, 128, 0, B.d, h->r, h->z, 1.0, h->q, h->S, h->partials, h->chk);
        h->launch_count += 1;
        xreduce<FINK>(h, 0, 1);
    }
}

// p = z + beta p on the rank's own rows AND on the halo slots: the peers' z rows are pulled (z is final everywhere once the r.z
// all-reduce inside the preconditioner's last kernel has completed), so the halo of p never has to be exchanged and no barrier
// is needed between the update and the next product H p
template <int D> void update_p(pgo_handle *h) {
    LevelBuf &B = h->lv[0];
    halo_pull(h, 0, h->z, VecStride<D>::value, 1);
    const int64_t rows = B.d.n_pad + (h->world > 1 ? B.n_halo : 0);
    launch_k(h, k_update_p<D>, grid_for(rows * (VecStride<D>::value / 2), 256 * 4), 256, 0, rows, h->p, h->z, h->S, h->chk);
    h->launch_count += 1;
}

template <int D> void pcg_iteration(pgo_handle *h) {
    LevelBuf &B = h->lv[0];
    spmv<D, 0, FIN_PQ, false, false>(h, 0, h->p, nullptr, h->q, 0.0, nullptr, nullptr, h->chk);
    if (h->use_amg && h->lv.size() > 1) {
        launch_k(h, k_update_xr_dinv<D>, B.grid128, 128, 0, B.d, h->x, h->r, h->p, h->q, B.xa, B.omega, h->S, h->chk);
        precondition<D, FIN_RZ, true>(h);
    } else {
        launch_k(h, k_update_xr<D>, B.gridv, 256, 0, B.d.n_pad, h->x, h->r, h->p, h->q, h->S, h->chk);
        precondition<D, FIN_RZ>(h);
    }
    h->launch_count += 1;
    update_p<D>(h);
}

// The PCG loop as a device-side loop: a graph with one WHILE conditional node whose body is one PCG iteration followed by
// k_loop_cond (loop again unless a kernel has set `done`).  One graph launch per solve, no host polling, no iterations launched
// behind the convergence flag.  Returns false when this driver / toolkit cannot build it (the caller falls back to the chunked graph).
template <int D> bool build_pcg_while(pgo_handle *h) {
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ge = nullptr;
    const int64_t before = h->launch_count;
    bool capturing = false, ok = false;
    do {
        if (cudaGraphCreate(&g, 0) != cudaSuccess) break;
        cudaGraphConditionalHandle hnd;
        if (cudaGraphConditionalHandleCreate(&hnd, g, 1, cudaGraphCondAssignDefault) != cudaSuccess) break;
        // the loop condition is evaluated once in FRONT of the loop (no iteration at all when the initial residual already set `done`) ...
        if (cudaStreamBeginCaptureToGraph(h->stream, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) != cudaSuccess) break;
        capturing = true;
        launch_k(h, k_loop_cond, 1, 32, 0, hnd, (const Scalars *)h->S);
        cudaGraph_t out = nullptr;
        cudaError_t ce = cudaStreamEndCapture(h->stream, &out);
        capturing = false;
        if (ce != cudaSuccess || h->launch_err != cudaSuccess) break;
        cudaGraphNode_t head[4];
        size_t n_head = 4;
        if (cudaGraphGetNodes(g, head, &n_head) != cudaSuccess || n_head != 1) break;
        cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
        p.type = cudaGraphNodeTypeConditional;
        p.conditional.handle = hnd;
        p.conditional.type = cudaGraphCondTypeWhile;
        p.conditional.size = 1;
        cudaGraphNode_t node;
        if (cudaGraphAddNode(&node, g, head, 1, &p) != cudaSuccess) break;
        cudaGraph_t body = p.conditional.phGraph_out[0];
        if (cudaStreamBeginCaptureToGraph(h->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) != cudaSuccess) break;
        capturing = true;
        // ... and at the end of every iteration, so no kernel of the body ever runs behind the flag: they are captured without the
        // check (a dependent L2 round trip at the top of each of the ~110 kernels of an iteration)
        h->chk = 0;
        pcg_iteration<D>(h);
        h->chk = 1;
        launch_k(h, k_loop_cond, 1, 32, 0, hnd, (const Scalars *)h->S);
        h->launch_count += 1;
        out = nullptr;
        ce = cudaStreamEndCapture(h->stream, &out);
        capturing = false;
        if (ce != cudaSuccess || h->launch_err != cudaSuccess) break;
        if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) break;
        ok = true;
    } while (false);
    if (capturing) { cudaGraph_t out = nullptr; cudaStreamEndCapture(h->stream, &out); }
    h->chk = 1;
    if (g) cudaGraphDestroy(g);
    if (!ok) { (void)cudaGetLastError(); h->launch_err = cudaSuccess; h->launch_count = before; return false; }
    h->pcg_graph = ge;
    h->pcg_while = true;
    h->launches_per_iter = h->launch_count - before;
    h->launch_count = before;
    return true;
}

template <int D> int build_pcg_graph(pgo_handle *h) {
    if (h->pcg_graph) return PGO_OK;
    h->pcg_while = false;
    // Sharded handles use the device-side loop too when every shard has a GPU of its own and every coarse level is replicated (the
    // default up to ~2M poses per level-1 partition).  Shards that SHARE a GPU (tests on a one-GPU machine) stall with it -- WHILE body +
    // programmatic dependent launch + kernels spinning on another shard's kernels that need the same SMs (r03f / r04q: peer time-out;
    // either ingredient alone is fine) -- and with sharded coarse levels it has only been run for experiments (PGO_WHILE=2: green on two
    // real GPUs, profiles/r04p_*): both keep the chunked graph.
    bool sharded_coarse = false;
    for (size_t l = 1; l < h->lv.size(); l++) sharded_coarse |= h->world > 1 && !h->lv[l].repl;
    if (h->opt_while && ((!sharded_coarse && !h->shares_device) || h->opt_while_sharded) && build_pcg_while<D>(h)) return PGO_OK;
    cudaGraph_t g = nullptr;
    int64_t before = h->launch_count;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < h->chunk; i++) pcg_iteration<D>(h);
    cudaError_t ie = cudaStreamEndCapture(h->stream, &g);
    if (ie == cudaSuccess && h->launch_err != cudaSuccess) ie = h->launch_err;
    if (ie == cudaSuccess) ie = cudaGraphInstantiate(&h->pcg_graph, g, 0);
    if (g) cudaGraphDestroy(g);
    if (ie != cudaSuccess && h->pdl) {
        // programmatic dependent launch edges not accepted by this driver inside a captured graph: capture again without them
        (void)cudaGetLastError();
        h->pdl = false;
        h->launch_err = cudaSuccess; h->pcg_graph = nullptr;
        h->launch_count = before;
        return build_pcg_graph<D>(h);
    }
    CK(ie);
    h->launches_per_iter = (h->launch_count - before) / h->chunk;
    h->launch_count = before;
    return PGO_OK;
}

// ---- assemble the Gauss-Newton system at the current poses (H in lv[0], b in h->r)
template <int D> int assemble(pgo_handle *h, double lambda, int add_lambda) {
    LevelBuf &B = h->lv[0];
    xbarrier(h, 0);                                  // every rank's poses are final
    halo_pull(h, 0, h->poses, Dim<D>::PS, 0);
    if (D == 3) launch_k(h, k_assemble_se2, B.grid128, 128, 0, B.d, xref(h, h->poses, true), h->poses, h->hz, h->r, h->anchor_row, h->opt.anchor_weight,
                                                                  add_lambda ? lambda : 0.0);
    else launch_k(h, k_assemble_se3, B.grid128, 128, 0, B.d, h->poses, h->hz, h->r, h->anchor_row, h->opt.anchor_weight, add_lambda ? lambda : 0.0);
    h->launch_count += 1;
    CK(cudaGetLastError()); CK(h->launch_err);
    return PGO_OK;
}

// the peers' Comm blocks (= the start of their arenas) into the device-side scalars, once the shards are connected
int upload_comm_peers(pgo_handle *h) {
    Comm *peers[MAX_RANKS];
    for (int k = 0; k < MAX_RANKS; k++) peers[k] = (Comm *)h->peer_base[k < h->world ? k : h->rank];
    CK(cudaMemcpy(&h->S->comm_peer[0], peers, sizeof(peers), cudaMemcpyHostToDevice));
    return PGO_OK;
}

int comm_status(pgo_handle *h, const Scalars &s) {
    if (s.status == ST_COMM) { h->err = "peer synchronisation timed out (a rank is missing or out of step)"; return PGO_ERR_COMM; }
    return PGO_OK;
}

// power iteration for rho(Dinv H) on one level -> damping of the block-Jacobi smoother.  The norm is reduced over ALL
// ranks on the device (same value, hence the same omega, everywhere); the host only reads one scalar per iteration.
template <int D> int estimate_omega(pgo_handle *h, int l) {
    constexpr int VS = VecStride<D>::value;
    LevelBuf &B = h->lv[l];
    const int64_t nd = B.d.n_pad * VS;
    std::vector<double> v(nd, 0.0);
    // the start vector is a function of the GLOBAL row, so that it does not depend on the partition
    const int64_t g0 = B.repl ? 0 : h->sym.levels[l].part_off[h->rank];
    for (int64_t i = 0; i < B.d.n; i++)
        for (int c = 0; c < D; c++) {
            uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)((g0 + i) * 8 + c + 1);
            s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
            v[i * VS + c] = (double)(s >> 11) / 9007199254740992.0 - 0.5;
        }
    if (l == 0 && D == 3) {   // keep the landmark padding unknown out of it
        for (int64_t i = 0; i < B.d.n; i++) if (h->sym.vkind[h->sym.perm[h->row0 + i]] == 1) v[i * 4 + 2] = 0.0;
    }
    double *a = B.xa, *b = B.res;
    CK(cudaMemcpyAsync(a, v.data(), nd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    double rho = 1.0;
    for (int it = 0; it < 12; it++) {
        // b = H a ; a' = Dinv b ; rho ~ |a'| / |a|
        lbarrier(h, l, 0);
        spmv_any<D, 0>(h, l, a, nullptr, b, 0.0, 0);
        lbarrier(h, l, 0);                           // every peer has pulled its halo of `a` before it is overwritten
        launch_k(h, k_dinv_apply<D, FIN_NONE>, B.grid128, 128, 0, B.d, b, a, 1.0, nullptr, h->S, h->partials, 0);
        launch_k(h, k_dots<D, FIN_NORM>, B.grid128, 128, 0, B.d.n_pad, a, a, nullptr, h->S, h->partials, l, 0);
        xreduce<FIN_NORM>(h, l, 0);
        CK(cudaMemcpyAsync(&h->hS[0], h->S, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        int rc = comm_status(h, h->hS[0]);
        if (rc) return rc;
        const double nrm = std::sqrt(h->hS[0].norm2_dx);
        if (!(nrm > 0.0) || !std::isfinite(nrm)) { rho = 2.0; break; }
        if (it > 0) rho = nrm;                                   // |a| was normalised to 1 by the previous pass
        launch_k(h, k_scale, grid_for(nd / 2, 256), 256, 0, nd, a, 1.0 / nrm);
    }
    // damping of the block-Jacobi smoother: omega = c / rho(Dinv H).  c = 4/3.3 = 1.21 (the textbook 4/3 with a 10 % margin on the
    // power-iteration estimate) was round 1's choice; the PCG count keeps falling up to c ~ 1.5 and is flat from there to 1.75 on every
    // graph tried (SciPy prototype: manhattan 100k 49 -> 43, 300k 47 -> 43, intel 28 -> 26, M3500 54 -> 52; tools/research/), and the
    // smoother stays convergent / the cycle SPD up to c = 2, so c = 1.5 leaves a 33 % margin for an under-estimated or drifting rho.
    B.omega = h->omega_rho / std::max(rho, 1.0);
    if (B.omega > 1.0) B.omega = 1.0;
    return PGO_OK;
}

// numeric setup of the hierarchy for the current H: coarse positions, Galerkin products, inverses
template <int D> int amg_setup(pgo_handle *h) {
    constexpr int DD = D * D, NG = Dim<D>::NG, LS = Dim<D>::LS;
    if (!h->use_amg) return PGO_OK;
    const int last = (int)h->lv.size() - 1;
    auto to_float = [&](LevelBuf &B) {               // fp32 copy of the stored blocks for the cycle's SpMVs
        if (!h->lowp || !B.d.valf) return;
        const int64_t nv = (int64_t)DD * B.d.n_slots;
        if (nv == 0) return;
        launch_k(h, k_to_float, grid_for((nv + 1) / 2, 256), 256, 0, nv, B.d.val, B.d.valf);
        h->launch_count += 1;
        if (B.d.diagf) {                             // sliced levels: the diagonal blocks and their inverses too
            const int64_t nd = (int64_t)DD * B.d.n_pad;
            launch_k(h, k_to_float, grid_for((nd + 1) / 2, 256), 256, 0, nd, B.d.diag, B.d.diagf);
            launch_k(h, k_to_float, grid_for((nd + 1) / 2, 256), 256, 0, nd, B.d.dinv, B.d.dinvf);
            h->launch_count += 2;
        }
    };
    // level 0: the assembly kernel wrote the fp32 copies (blocks, diagonal blocks, their inverses) in the same pass
    for (int l = 0; l < last; l++) {
        LevelBuf &F = h->lv[l], &C = h->lv[l + 1];
        launch_k(h, k_coarse_pos<NG>, C.gridw, 256, 0, F.d, C.d);
        launch_k(h, k_lever<NG>, F.grid128, 128, 0, F.d, C.d);
        CK(cudaMemsetAsync(C.d.val, 0, sizeof(double) * DD * std::max<int64_t>(C.d.n_slots, 1), h->stream));
        CK(cudaMemsetAsync(C.d.diag, 0, sizeof(double) * DD * C.d.n_pad, h->stream));
        lbarrier(h, l, 0);                           // lever arms of neighbour rows on other ranks
        halo_pull(h, l, F.d.lev, LS, 0);
        if (F.d.gptr) {      // deterministic: project every fine block into the staging buffer, then one writer per coarse block
            if (F.jds) launch_k(h, k_galerkin_stage_jds<D>, F.grid128, 128, 0, F.d, xref(h, F.d.lev, true), h->gstage);
            else launch_k(h, k_galerkin_stage_csr<D>, F.gridw, 256, 0, F.d, xref(h, F.d.lev, true), h->gstage);
            launch_k(h, k_galerkin_reduce<D>, grid_for(F.d.n_gblk * GalerkinLanes<D>::value, 256), 256, 0, F.d, C.d, (const double *)h->gstage);
            h->launch_count += 1;
        } else if (F.jds) launch_k(h, k_galerkin_jds<D>, F.grid128, 128, 0, F.d, C.d, xref(h, F.d.lev, true));
        else launch_k(h, k_galerkin_csr<D>, F.gridw, 256, 0, F.d, C.d, xref(h, F.d.lev, true));
        h->launch_count += 3;
        if (C.first_repl) {                          // every rank built the coarse rows of its own aggregates: all-gather them
            lbarrier(h, l, 0);
            gather_rows(h, C.d.val, C.src_slots, DD, 1, 0, 0);
            gather_rows(h, C.d.diag, C.src_rows, 1, DD, C.d.n_pad, 0);
            gather_rows(h, C.d.pos, C.src_rows, 1, NG, C.d.n_pad, 0);
        }
        launch_k(h, k_invert_diag<D>, C.grid128, 128, 0, C.d);
        h->launch_count += 1;
        if (!(l + 1 == last && h->sym.dense_coarsest)) to_float(C);
    }
    if (h->sym.dense_coarsest) {
        LevelBuf &C = h->lv[last];
        const int m = h->dense_m;
        // ping-pong blocked Gauss-Jordan: the matrix is assembled into the buffer from which the last panel step writes Ainv
        const int n_panels = (m + GJ_W - 1) / GJ_W;
        double *first = (n_panels % 2 == 0) ? h->Ainv : h->Awork, *second = (n_panels % 2 == 0) ? h->Awork : h->Ainv;
        CK(cudaMemsetAsync(first, 0, sizeof(double) * (size_t)m * m, h->stream));
        if (C.jds) launch_k(h, k_dense_assemble<D, true>, C.grid128, 128, 0, C.d, h->dmap, 0, m, first);
        else launch_k(h, k_dense_assemble<D, false>, C.grid128, 128, 0, C.d, h->dmap, 0, m, first);
        h->launch_count += 1;
        void *args[] = {(void *)&h->dense_m, (void *)&first, (void *)&second, (void *)&h->gj_pnext, (void *)&h->gj_bar, (void *)&h->gj_bar_base};
        CK(cudaLaunchCooperativeKernel((void *)k_dense_invert_sym, dim3(h->invert_grid), dim3(256), args, GJ_SMEM, h->stream));
        h->gj_bar_base += (unsigned)n_panels * (unsigned)h->invert_grid;
        h->launch_count += 1;
    }
    CK(cudaGetLastError()); CK(h->launch_err);
    if (!h->omega_ready) {
        for (int l = 0; l < (int)h->lv.size(); l++) {
            if (l == last && h->sym.dense_coarsest) continue;
            int rc = estimate_omega<D>(h, l);
            if (rc) return rc;
        }
        h->omega_ready = true;
        if (h->pcg_graph) { cudaGraphExecDestroy(h->pcg_graph); h->pcg_graph = nullptr; }  // omega is baked into the graph
    }
    return PGO_OK;
}

int reset_scalars(pgo_handle *h) {
    Scalars s{};
    s.tol2 = h->opt.pcg_rtol * h->opt.pcg_rtol;
    s.max_iters = h->opt.pcg_max_iterations;
    // counters / epoch / world must survive (they are always consistent between kernels); everything before them is re-initialised
    CK(cudaMemcpyAsync(h->S, &s, offsetof(Scalars, counter), cudaMemcpyHostToDevice, h->stream));
    return PGO_OK;
}

// solve H x = b (b in h->r, destroyed) ; x in h->x
template <int D> int solve(pgo_handle *h, int32_t *iters_out) {
    LevelBuf &B = h->lv[0];
    const int64_t nd = B.d.n_pad * VecStride<D>::value;
    int rc = reset_scalars(h);
    if (rc) return rc;
    rc = build_pcg_graph<D>(h);
    if (rc) return rc;
    CK(cudaMemsetAsync(h->x, 0, nd * sizeof(double), h->stream));
    precondition<D, FIN_RZ_INIT>(h);
    halo_pull(h, 0, h->z, VecStride<D>::value, 1);   // p = z, halo slots included (pcg_iteration keeps them current from here on)
    CK(cudaMemcpyAsync(h->p, h->z, (nd + (h->world > 1 ? B.n_halo * VecStride<D>::value : 0)) * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (h->pcg_while) {                              // the loop runs on the device: one launch, the host only waits for the end
        CK(cudaGraphLaunch(h->pcg_graph, h->stream));
        CK(cudaMemcpyAsync(&h->hS[0], h->S, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->launch_count += h->launches_per_iter * std::max(h->hS[0].iters, 1);
    }
    // keep two graph launches in flight; poll the pinned scalars of the older one
    int slot = 0, inflight = 0;
    int64_t launched = 0;
    const int64_t max_graphs = (int64_t)h->opt.pcg_max_iterations / h->chunk + 2;
    bool done = h->pcg_while;
    while (!done) {
        if (launched < max_graphs) {
            CK(cudaGraphLaunch(h->pcg_graph, h->stream));
            h->launch_count += h->launches_per_iter * h->chunk;
            CK(cudaMemcpyAsync(&h->hS[slot], h->S, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaEventRecord(h->poll_ev[slot], h->stream));
            launched++; inflight++; slot ^= 1;
        }
        if (inflight == 2 || launched >= max_graphs) {
            const int old = (inflight == 2) ? slot : (slot ^ 1);
            CK(cudaEventSynchronize(h->poll_ev[old]));
            inflight--;
            if (h->hS[old].done) done = true;
            else if (launched >= max_graphs && inflight == 0) done = true;
        }
    }
    CK(cudaMemcpyAsync(&h->hS[0], h->S, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (iters_out) *iters_out = h->hS[0].iters;
    if ((rc = comm_status(h, h->hS[0]))) return rc;
    if (h->hS[0].status == ST_BREAKDOWN) {
        h->err = "PCG breakdown: H is not positive definite (isolated vertex, or graph without a pose-pose edge to anchor)";
        return PGO_ERR_SOLVER;
    }
    if (h->hS[0].status == ST_MAXIT) { h->err = "PCG did not converge within pcg_max_iterations"; return PGO_ERR_NOT_CONVERGED; }
    return PGO_OK;
}

// pgo_options.refine: dx = x1 + d with x1 the PCG solution and d the PCG solution of H d = b - H x1, the residual evaluated in
// double-double arithmetic (k_residual_dd) -- one round of iterative refinement with an extended-precision residual
template <int D> int solve_refined(pgo_handle *h, int32_t *iters_out) {
    if (!h->opt.refine) return solve<D>(h, iters_out);
    constexpr int VS = VecStride<D>::value;
    LevelBuf &B = h->lv[0];
    const int64_t nd = B.d.n_pad * VS;
    if (!h->b0) {
        int rc = dalloc(h, &h->b0, (size_t)nd); if (rc) return rc;
        rc = dalloc(h, &h->xs, (size_t)nd); if (rc) return rc;
    }
    CK(cudaMemcpyAsync(h->b0, h->r, nd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    int32_t it1 = 0, it2 = 0;
    int rc = solve<D>(h, &it1);
    if (iters_out) *iters_out = it1;
    if (rc != PGO_OK) return rc;
    xbarrier(h, 0);                                  // every shard's x is final before its halo rows are pulled
    halo_pull(h, 0, h->x, VS, 0);
    launch_k(h, k_residual_dd<D>, B.grid128, 128, 0, B.d, (const double *)h->x, (const double *)h->b0, h->r);
    h->launch_count += 1;
    CK(cudaMemcpyAsync(h->xs, h->x, nd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    const double rtol = h->opt.pcg_rtol;
    h->opt.pcg_rtol = h->opt.refine_rtol;
    rc = solve<D>(h, &it2);                          // H d = r ; d lands in h->x
    h->opt.pcg_rtol = rtol;
    launch_k(h, k_add_to, grid_for(nd / 2, 256), 256, 0, nd, h->x, (const double *)h->xs);
    h->launch_count += 1;
    CK(cudaGetLastError()); CK(h->launch_err);
    if (iters_out) *iters_out = it1 + it2;
    return rc;
}

template <int D> int retract(pgo_handle *h, double sign) {
    LevelBuf &B = h->lv[0];
    if (D == 3) launch_k(h, k_retract_se2, grid_for(B.d.n, 256), 256, 0, B.d, h->poses, h->x, sign, h->S, h->partials);
    else launch_k(h, k_retract_se3, grid_for(B.d.n, 256), 256, 0, B.d, h->poses, h->x, sign, h->S, h->partials);
    h->launch_count += 1;
    xreduce<FIN_NORM>(h, 0, 0);                      // also: every rank's poses are updated before anyone reads them
    CK(cudaGetLastError()); CK(h->launch_err);
    return PGO_OK;
}

template <int D> int chi2_launch(pgo_handle *h) {
    xbarrier(h, 0);
    halo_pull(h, 0, h->poses, Dim<D>::PS, 0);
    if (D == 3) launch_k(h, k_chi2_se2, grid_for(h->n_edges_loc, 256), 256, 0, h->n_edges_loc, h->ed_stride, h->ends, h->ed, h->poses, xref(h, h->poses, true), h->S, h->partials);
    else launch_k(h, k_chi2_se3, grid_for(h->n_edges_loc, 256), 256, 0, h->n_edges_loc, h->ed_stride, h->ends, h->ed, h->poses, h->S, h->partials);
    h->launch_count += 1;
    xreduce<FIN_CHI2>(h, 0, 0);
    CK(cudaGetLastError()); CK(h->launch_err);
    return PGO_OK;
}

// diagnostic: average time of ONE coarse solve at `level` (the K-cycle subtree below it), graph-launched like inside the PCG loop
template <int D> int time_coarse(pgo_handle *h, int level, int repeats, double *avg_ms) {
    if (!h->use_amg || level < 1 || level >= (int)h->lv.size()) { h->err = "pgo_time_coarse: no such coarse level"; return PGO_ERR_ARG; }
    if (h->world > 1 && !h->lv[level].repl) { h->err = "pgo_time_coarse: sharded level"; return PGO_ERR_UNSUPPORTED; }
    int rc = reset_scalars(h);
    if (rc) return rc;
    LevelBuf &B = h->lv[level];
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ge = nullptr;
    const int64_t before = h->launch_count;
    if (!h->pcg_graph) { int rc2 = build_pcg_graph<D>(h); if (rc2) return rc2; }
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    coarse_solve<D>(h, level, B.rhs, B.sol, false);
    CK(cudaStreamEndCapture(h->stream, &g));
    const int64_t per = h->launch_count - before;
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaGraphDestroy(g);
    for (int i = 0; i < 3; i++) CK(cudaGraphLaunch(ge, h->stream));
    CK(cudaEventRecord(h->ev[PGO_NUM_PHASES], h->stream));
    for (int i = 0; i < repeats; i++) CK(cudaGraphLaunch(ge, h->stream));
    CK(cudaEventRecord(h->ev[PGO_NUM_PHASES + 1], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[PGO_NUM_PHASES], h->ev[PGO_NUM_PHASES + 1]));
    cudaGraphExecDestroy(ge);
    *avg_ms = ms / repeats;
    h->launch_count = before + per * (repeats + 3);
    return PGO_OK;
}

// block dimension dispatch: 3 (SE2 / XY graphs) or 6 (SE3 graphs)
#define BY_D(h, fn, ...) ((h)->sym.D == 6 ? fn<6>(__VA_ARGS__) : fn<3>(__VA_ARGS__))

void multi_shutdown(pgo_handle *h);

int fail_create(pgo_handle *h, int rc, const std::string &msg) {
    g_create_error = msg.empty() ? h->err : msg;
    pgo_destroy(h);
    return rc;
}

} // namespace

// ================================================================================================
extern "C" {

const char *pgo_version(void) { return "pgo_b200 0.2 (sm_100a, fp64)"; }

void pgo_default_options(pgo_options *o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->anchor_weight = 1e7;
    o->pcg_rtol = 1e-10;
    o->pcg_max_iterations = 200000;
    o->preconditioner = PGO_PRECOND_AMG;
    o->sort_window = 2048;
    o->amg_max_levels = 12;
    o->device = -1;
    o->world = 1;
    o->rank = 0;
    o->amg_dense_max = 640;
    o->amg_aggregate_size = 0;        // auto: 16 (single GPU), 24 (sharded)
    o->amg_kcycle = MAX_LEVELS;
    o->amg_kcycle3 = -1;
}

const char *pgo_last_error(const pgo_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void pgo_destroy(pgo_handle *h) {
    if (!h) return;
    if (h->multi) multi_shutdown(h);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->pcg_graph) cudaGraphExecDestroy(h->pcg_graph);
    for (int k = 0; k < MAX_RANKS; k++)
        if (h->peer_base[k] && h->ipc_peer[k]) cudaIpcCloseMemHandle(h->peer_base[k]);
    for (void *p : h->allocs) cudaFree(p);
    if (h->hS) cudaFreeHost(h->hS);
    for (auto &e : h->ev) if (e) cudaEventDestroy(e);
    for (auto &e : h->poll_ev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

} // extern "C"

// device-resident state of one handle (a single-GPU handle, one rank of a process-per-GPU job, or one shard of a single-process
// multi-GPU handle): h->sym / h->opt / world / rank are set.  Runs on the thread that will own the shard's CUDA device.
static int create_shard(pgo_handle *h, const double *vval, int64_t ne, const uint8_t *ekind, const double *emeas, const double *einfo) {
    Symbolic &S = h->sym;
    double t_last = now_s();
    // per-dimension record sizes (kernels.cuh: Dim<D>)
    const int D = S.D, DD = D * D, VS = D == 6 ? 6 : 4, PS = D == 6 ? 8 : 4, NG = D == 6 ? 3 : 2, LS = D == 6 ? 4 : 2, NM = D == 6 ? 28 : 10;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { h->err = "no CUDA device available (this library has no CPU fallback)"; return PGO_ERR_CUDA; }
    if (h->opt.device >= 0) { if (cudaSetDevice(h->opt.device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return PGO_ERR_CUDA; } }
    cudaGetDevice(&h->device);
#define CKC(call) do { int rc_ = (call); if (rc_) return rc_; } while (0)
#define CKU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return PGO_ERR_CUDA; } } while (0)
    CKU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto &e : h->ev) CKU(cudaEventCreate(&e));
    for (auto &e : h->poll_ev) CKU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CKU(cudaHostAlloc((void **)&h->hS, 2 * sizeof(Scalars), cudaHostAllocDefault));
    std::memset(h->hS, 0, 2 * sizeof(Scalars));
    TICK("shard: context + stream");

    const int rank = h->rank, world = h->world;
    const int nl = (int)S.levels.size();
    h->lv.resize(nl);
    int64_t max_grid = 1, max_stage = 0;
    // ---- pass 1: local structure of every level + arena requests (sizes from the largest partition: same layout on all ranks)
    for (int l = 0; l < nl; l++) {
        HostLevel &H = S.levels[l];
        LevelBuf &B = h->lv[l];
        LevelDev &d = B.d;
        // a replicated level (sharded handles only) has ONE partition that every rank holds completely
        B.repl = world > 1 && H.repl;
        B.first_repl = B.repl && !S.levels[l - 1].repl;
        if (B.repl && H.jds) { h->err = "replicated levels must be block CSR"; return PGO_ERR_UNSUPPORTED; }
        const int pk = B.repl ? 0 : rank;              // partition of this level held by this rank
        const int lworld = (int)H.part_real.size();
        const int64_t r0 = H.part_off[pk], r1 = H.part_off[pk + 1];
        const int64_t s0 = H.part_slot[pk], s1 = H.part_slot[pk + 1];
        B.jds = H.jds;
        d.n = H.part_real[pk]; d.n_pad = r1 - r0; d.n_slots = s1 - s0; d.n_slices = H.jds ? d.n_pad / 32 : 0;
        int64_t max_pad = 32, max_slots = 1;
        for (int k = 0; k < lworld; k++) {
            max_pad = std::max(max_pad, H.part_off[k + 1] - H.part_off[k]);
            max_slots = std::max(max_slots, H.part_slot[k + 1] - H.part_slot[k]);
        }
        if (B.first_repl) {
            for (int k = 0; k <= MAX_RANKS; k++) {
                const int64_t row = H.src_off[std::min(k, world)];
                B.src_rows.off[k] = row; B.src_slots.off[k] = H.adj_ptr[row];
            }
        }
        B.grid128 = grid_for(d.n_pad, 128);
        B.gridw = grid_for(d.n_pad, 8);
        B.grid8 = grid_for(d.n_pad, 32);
        B.grid4 = grid_for(d.n_pad, 64);
        B.lpr = (!H.jds && d.n > 0 && d.n_slots <= 12 * d.n) ? 8 : 32;     // short rows: 8 lanes per row
        // a big level with short rows: 4 lanes per row, so that the whole level is (about) one wave of CTAs
        if (B.lpr == 8 && d.n >= h->lpr4_min_rows) B.lpr = 4;
        B.gridv = grid_for(d.n_pad * (VS / 2), 256);
        max_grid = std::max<int64_t>(max_grid, std::max(B.grid128, B.gridw));
        // row pointers
        std::vector<int64_t> rp;
        if (H.jds) { rp.resize(d.n_slices + 1); for (int64_t i = 0; i <= d.n_slices; i++) rp[i] = H.slice_ptr[r0 / 32 + i] - s0; }
        else { rp.resize(d.n_pad + 1); for (int64_t i = 0; i <= d.n_pad; i++) rp[i] = H.adj_ptr[r0 + i] - s0; }
        // halo of every partition of a sharded level: sorted unique rows of other partitions referenced by its blocks
        int64_t max_halo = 0;
        std::vector<int64_t> halo;                     // this rank's, as global padded rows
        if (lworld > 1) {
            std::vector<int64_t> tmp;
            for (int k = 0; k < lworld; k++) {
                tmp.clear();
                for (int64_t q = H.adj_ptr[H.part_off[k]]; q < H.adj_ptr[H.part_off[k + 1]]; q++) {
                    const int64_t nb = H.adj_nbr[q];
                    if (nb < H.part_off[k] || nb >= H.part_off[k + 1]) tmp.push_back(nb);
                }
                std::sort(tmp.begin(), tmp.end());
                tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
                max_halo = std::max<int64_t>(max_halo, (int64_t)tmp.size());
                if (k == pk) halo = tmp;
            }
            B.n_halo = (int64_t)halo.size();
            std::vector<uint32_t> hs(std::max<size_t>(halo.size(), 1), 0);
            for (size_t i = 0; i < halo.size(); i++) { const int ow = H.part_of(halo[i]); hs[i] = ((uint32_t)ow << COL_OWNER_SHIFT) | (uint32_t)(halo[i] - H.part_off[ow]); }
            CKC(upload(h, &B.halo_src, hs));
            if (d.n_pad + max_halo > (int64_t)COL_LOCAL_MASK) { h->err = "level too large for the column word"; return PGO_ERR_ARG; }
        }
        max_pad += max_halo;                           // every vector of the level has room for the halo records behind its own rows
        auto local_index = [&](int64_t nb) -> uint32_t {   // index of a neighbour row in this rank's (halo-extended) vectors
            if (nb >= r0 && nb < r1) return (uint32_t)(nb - r0);
            return (uint32_t)(d.n_pad + (std::lower_bound(halo.begin(), halo.end(), nb) - halo.begin()));
        };
        std::vector<int32_t> dg(d.n_pad, 0);
        std::vector<uint32_t> cl(std::max<int64_t>(d.n_slots, 1), 0);
        parallel_for(r1 - r0, 8192, [&](int64_t a0, int64_t a1) {
            for (int64_t r = r0 + a0; r < r0 + a1; r++) {
                dg[r - r0] = (int32_t)(H.adj_ptr[r + 1] - H.adj_ptr[r]);
                for (int64_t q = H.adj_ptr[r]; q < H.adj_ptr[r + 1]; q++)
                    cl[H.adj_slot[q] - s0] = local_index(H.adj_nbr[q]) | (H.adj_flags.empty() ? 0u : H.adj_flags[q]);
            }
        });
        if (l == 0) {   // the chi2 edge list addresses the `to` pose the same way
            h->to_index.resize(S.n_edges);
            parallel_for(S.n_edges, 65536, [&](int64_t e0, int64_t e1) {
                for (int64_t e = e0; e < e1; e++) {
                    const int64_t a = S.iperm[S.efrom[e]], b = S.iperm[S.eto[e]];
                    h->to_index[e] = (a >= r0 && a < r1) ? local_index(b) : 0u;
                }
            });
        }
        int64_t *drp; int32_t *ddg; uint32_t *dcl;
        CKC(upload(h, &drp, rp)); CKC(upload(h, &ddg, dg)); CKC(upload(h, &dcl, cl));
        d.slice_ptr = drp; d.deg = ddg; d.col = dcl;
        if (B.first_repl) {   // gathered from the peers after the Galerkin product: peer-visible
            arena_request(h, &d.val, (size_t)DD * max_slots);
            arena_request(h, &d.diag, (size_t)DD * max_pad);
            arena_request(h, &d.pos, (size_t)NG * max_pad);
        } else {
            CKC(dalloc(h, &d.val, (size_t)DD * std::max<int64_t>(d.n_slots, 1) + 2));      // + 16 bytes: the TMA-staged SpMV copies 16-byte-aligned windows
            CKC(dalloc(h, &d.diag, (size_t)DD * d.n_pad));
            CKC(dalloc(h, &d.pos, (size_t)NG * d.n_pad));
        }
        CKC(dalloc(h, &d.dinv, (size_t)DD * d.n_pad));
        if (h->lowp && H.jds) {
            CKC(dalloc(h, &d.diagf, (size_t)DD * d.n_pad));
            CKC(dalloc(h, &d.dinvf, (size_t)DD * d.n_pad));
        }
        if (h->lowp && !(l == nl - 1 && S.dense_coarsest)) CKC(dalloc(h, &d.valf, (size_t)DD * std::max<int64_t>(d.n_slots, 1) + 4));
        if (!H.agg.empty()) {
            HostLevel &Cn = S.levels[l + 1];
            const int64_t c0 = Cn.part_off[(world > 1 && Cn.repl) ? 0 : rank];
            std::vector<int32_t> ag(d.n_pad, -1);
            for (int64_t r = r0; r < r1; r++) if (H.agg[r] >= 0) ag[r - r0] = (int32_t)(H.agg[r] - c0);
            int32_t *dag;
            CKC(upload(h, &dag, ag));
            d.agg = dag;
            if ((int)H.gal_ptr.size() > pk && !H.gal_ptr[pk].empty()) {     // deterministic Galerkin product: contributor lists
                int32_t *dgp, *dgs;
                CKC(upload(h, &dgp, H.gal_ptr[pk])); CKC(upload(h, &dgs, H.gal_src[pk]));
                d.gptr = dgp; d.gsrc = dgs; d.n_gblk = (int64_t)H.gal_ptr[pk].size() - 1;
                max_stage = std::max<int64_t>(max_stage, (int64_t)DD * (d.n_slots + d.n_pad));
            } else {                                                        // coarse level in sliced storage: atomic scatter
                std::vector<int32_t> ct(H.ctgt.begin() + s0, H.ctgt.begin() + s1), cs(H.cstr.begin() + s0, H.cstr.begin() + s1);
                if (ct.empty()) { ct.push_back(0); cs.push_back(1); }
                int32_t *dct, *dcs;
                CKC(upload(h, &dct, ct)); CKC(upload(h, &dcs, cs));
                d.ctgt = dct; d.cstr = dcs;
            }
        }
        if (!H.mem_ptr.empty()) {
            HostLevel &Fn = S.levels[l - 1];
            // first replicated level: only the rows built from this rank's finer rows have (local) members here; the
            // other rows are gathered from the rank that owns their members
            const int64_t o0 = B.first_repl ? H.src_off[rank] : r0, o1 = B.first_repl ? H.src_off[rank + 1] : r1;
            std::vector<int64_t> mp(d.n_pad + 1);
            const int64_t m0 = H.mem_ptr[o0];
            for (int64_t i = 0; i <= d.n_pad; i++) mp[i] = H.mem_ptr[std::min(std::max(r0 + i, o0), o1)] - m0;
            std::vector<int32_t> mi(H.mem_idx.begin() + m0, H.mem_idx.begin() + H.mem_ptr[o1]);
            for (auto &v : mi) v -= (int32_t)Fn.part_off[(world > 1 && Fn.repl) ? 0 : rank];
            if (mi.empty()) mi.push_back(0);
            int64_t *dmp; int32_t *dmi;
            CKC(upload(h, &dmp, mp)); CKC(upload(h, &dmi, mi));
            d.mem_ptr = dmp; d.mem_idx = dmi;
        }
        // peer-visible vectors of this level
        const size_t vec = (size_t)VS * max_pad;
        arena_request(h, &d.lev, (size_t)LS * max_pad);
        if (h->use_amg) {
            arena_request(h, &B.xa, vec); arena_request(h, &B.res, vec);
            if (l > 0) {
                arena_request(h, &B.rhs, vec); arena_request(h, &B.sol, vec);
                arena_request(h, &B.c1, vec); arena_request(h, &B.c2, vec); arena_request(h, &B.v1, vec);
                arena_request(h, &B.v2, vec); arena_request(h, &B.r1, vec); arena_request(h, &B.c3, vec);
            }
        }
        B.kcycle = l > 0 && l <= h->opt.amg_kcycle;
        // default (-1): level 1 of SE2 / XY graphs.  Measured on B200 (profiles/r01s_kcycle3.log): 1M-pose Manhattan graph 53 -> 41
        // PCG iterations and 40.3 -> 36.6 ms; on the 250k-pose SE3 sphere 38 -> 32 iterations but 29.4 -> 31.5 ms (its 6x6
        // coarse levels are the larger share of an iteration), so SE3 graphs keep two steps
        // Hierarchies of five or more levels (SE2 graphs beyond ~1.5M poses) lose convergence through the recursion -- the two-grid rate of
        // this aggregation (~0.6) is outside the regime where two inner steps keep a K-cycle level-independent: 47 / 117 / 368 PCG
        // iterations at 1M / 2M / 4M poses -- so there every K-cycle level runs three steps: 66 / 178 iterations, 182 -> 126 ms and
        // 971 -> 539 ms per GN step on one GPU (profiles/r03h_size_scaling.log).  Four levels: level 1 only (46 vs 47 iterations, slower).
        const int k3 = h->opt.amg_kcycle3 >= 0 ? h->opt.amg_kcycle3 : (D == 3 ? (nl >= 5 ? nl : 1) : 0);
        B.ksteps = (B.kcycle && l <= k3) ? 3 : 2;
        B.vec_rows = max_pad;
    }
    TICK("shard: level structures");
    {
        const size_t vec = (size_t)VS * h->lv[0].vec_rows;
        arena_request(h, &h->poses, (size_t)PS * h->lv[0].vec_rows);
        arena_request(h, &h->x, vec); arena_request(h, &h->r, vec); arena_request(h, &h->p, vec);
        arena_request(h, &h->q, vec); arena_request(h, &h->z, vec);
    }
    if (h->use_amg && S.dense_coarsest) {
        const HostLevel &HL = S.levels[nl - 1];     // single GPU, or replicated: one partition
        h->dmap.off[0] = 0;
        for (int k = 0; k < MAX_RANKS; k++) h->dmap.off[k + 1] = (int32_t)HL.n;
        h->dense_m = (int)HL.n * D;
    }
    CKC(arena_commit(h));
    TICK("shard: arena");
    if (max_stage > 0) CKC(dalloc(h, &h->gstage, (size_t)max_stage, false));
    if (h->use_amg && S.dense_coarsest) {
        const int m = h->dense_m;
        CKC(dalloc(h, &h->Ainv, (size_t)m * m));
        CKC(dalloc(h, &h->Awork, (size_t)m * m));

        int per_sm = 0, sms = 0;
        CKC(dalloc(h, &h->gj_pnext, (size_t)2 * GJ_W * GJ_W));
        CKC(dalloc(h, &h->gj_bar, 1));
        CKU(cudaFuncSetAttribute(k_dense_invert_sym, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GJ_SMEM));
        CKU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dense_invert_sym, 256, GJ_SMEM));
        CKU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
        const int nt = (m + GJ_T - 1) / GJ_T;
        h->invert_grid = std::max(2, std::min(std::max(per_sm, 1) * sms, gj_num_tiles(nt) + 1));      // + 1: the CTA that produces the pivot inverses
        CKU(cudaFuncSetAttribute(k_dense_apply<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (6 * 1024 + 8))));
        CKU(cudaFuncSetAttribute(k_dense_apply<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (6 * 1024 + 8))));
    }
    // PCG iterations per captured graph: the host keeps two launches in flight, so up to ~2 chunks of early-exit kernels run after
    // convergence; measured at config 4 (41 AMG iterations): chunk 1 / 2 / 4 / 8 -> PCG 34.17 / 34.19 / 34.70 / 35.79 ms
    h->chunk = h->use_amg ? 2 : 16;
    if (const char *e = std::getenv("PGO_CHUNK")) h->chunk = std::max(1, std::atoi(e));   // PCG iterations per captured graph

    // ---- level-0 vertex data of the owned rows, in storage order
    const HostLevel &H0 = S.levels[0];
    const int64_t r0 = H0.part_off[rank], r1 = H0.part_off[rank + 1], s0 = H0.part_slot[rank], s1 = H0.part_slot[rank + 1];
    h->row0 = r0; h->n_loc = H0.part_real[rank]; h->n_pad_loc = r1 - r0;
    {
        std::vector<uint8_t> vk(h->n_pad_loc, 0);
        std::vector<int64_t> rvo(h->n_pad_loc, 0);
        for (int64_t r = r0; r < r0 + h->n_loc; r++) { const int64_t v = S.perm[r]; vk[r - r0] = S.vkind[v]; rvo[r - r0] = S.vvalofs[v]; }
        uint8_t *dvk;
        CKC(upload(h, &dvk, vk));
        h->lv[0].d.vkind = dvk;
        CKC(upload(h, &h->row_valofs, rvo));
        CKC(dalloc(h, &h->vstage, (size_t)S.n_values));
        CKU(cudaMemcpyAsync(h->vstage, vval, S.n_values * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        if (D == 6) launch_k(h, k_import_poses_se3, grid_for(h->n_loc, 256), 256, 0, h->n_loc, h->row_valofs, h->vstage, h->poses);
        else launch_k(h, k_import_poses, grid_for(h->n_loc, 256), 256, 0, h->n_loc, h->row_valofs, dvk, h->vstage, h->poses);
        CKU(cudaGetLastError());
        if (D == 6) h->lv[0].d.quat = h->poses;
    }
    TICK("shard: vertex data");
    {
        const int64_t ag = S.anchor >= 0 ? S.iperm[S.anchor] : -1;
        h->anchor_row = (ag >= r0 && ag < r1) ? ag - r0 : -1;
    }
    // ---- measurements.  Edges incident to the rows of this rank: first the ones it owns (owner = the rank of `from`; these are the
    // edges its chi2 kernel sums), then the ones it only sees from the `to` side.  The caller's packed arrays go to the device as they
    // are and the edge-ordered records `ed` ([NM][n_inc] planes) are built there (k_build_ed); the half-edge stream hz the assembly
    // kernel reads is then gathered from them, also on the device (k_build_hz).
    {
        static const int NMEAS[3] = {3, 2, 7}, NINFO[3] = {6, 3, 21};
        int64_t nkind[3] = {0, 0, 0};
        for (int64_t k = 0; k < ne; k++) nkind[ekind[k]]++;
        const bool mixed = (nkind[0] > 0) + (nkind[1] > 0) + (nkind[2] > 0) > 1;
        const int64_t n_meas = nkind[0] * NMEAS[0] + nkind[1] * NMEAS[1] + nkind[2] * NMEAS[2];
        const int64_t n_info = nkind[0] * NINFO[0] + nkind[1] * NINFO[1] + nkind[2] * NINFO[2];
        std::vector<int64_t> mofs, iofs;               // packed offsets: only when the edge kinds are mixed (SE2 graphs with landmarks)
        if (mixed) {
            mofs.resize(ne); iofs.resize(ne);
            int64_t mo = 0, io = 0;
            for (int64_t k = 0; k < ne; k++) { mofs[k] = mo; iofs[k] = io; mo += NMEAS[ekind[k]]; io += NINFO[ekind[k]]; }
        }
        std::vector<int32_t> inc, inc_of;
        int64_t nm = ne, n_inc = ne;
        if (world > 1) {
            inc_of.assign(ne, -1);
            for (int64_t k = 0; k < ne; k++) { const int64_t a = S.iperm[S.efrom[k]]; if (a >= r0 && a < r1) inc.push_back((int32_t)k); }
            nm = (int64_t)inc.size();
            for (int64_t k = 0; k < ne; k++) {
                const int64_t a = S.iperm[S.efrom[k]], b = S.iperm[S.eto[k]];
                if (!(a >= r0 && a < r1) && b >= r0 && b < r1) inc.push_back((int32_t)k);
            }
            n_inc = (int64_t)inc.size();
            for (int64_t i = 0; i < n_inc; i++) inc_of[inc[i]] = (int32_t)i;
        }
        h->n_edges_loc = nm;
        h->ed_stride = std::max<int64_t>(n_inc, 1);
        const int64_t stride = h->ed_stride;
        std::vector<uint2> ends(std::max<int64_t>(nm, 1));
        parallel_for(nm, 65536, [&](int64_t i0, int64_t i1) {
            for (int64_t i = i0; i < i1; i++) {
                const int64_t k = world > 1 ? inc[i] : i;
                const int64_t a = S.iperm[S.efrom[k]];
                ends[i] = make_uint2((uint32_t)(a - r0), h->to_index[k] | (ekind[k] == 1 ? COL_EDGE_XY : 0u));
            }
        });
        CKC(upload(h, &h->ends, ends));
        CKC(dalloc(h, &h->ed, (size_t)NM * stride));
        {
            struct Tmp {                                // device copies of the caller's arrays: only needed until `ed` is built
                void *p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
                ~Tmp() { for (void *q : p) if (q) cudaFree(q); }
            } tmp;
            auto push = [&](int slot, const void *src, size_t bytes) -> cudaError_t {
                if (bytes == 0) return cudaSuccess;
                cudaError_t e = cudaMalloc(&tmp.p[slot], bytes);
                if (e == cudaSuccess) e = cudaMemcpyAsync(tmp.p[slot], src, bytes, cudaMemcpyHostToDevice, h->stream);
                return e;
            };
            CKU(push(0, emeas, (size_t)n_meas * sizeof(double)));
            CKU(push(1, einfo, (size_t)n_info * sizeof(double)));
            CKU(push(2, ekind, (size_t)ne));
            CKU(push(3, mofs.data(), mofs.size() * sizeof(int64_t)));
            CKU(push(4, iofs.data(), iofs.size() * sizeof(int64_t)));
            CKU(push(5, inc.data(), world > 1 ? inc.size() * sizeof(int32_t) : 0));
            if (n_inc > 0) {
                if (D == 6) launch_k(h, k_build_ed<28>, grid_for(n_inc, 256), 256, 0, n_inc, (const int32_t *)tmp.p[5], (const uint8_t *)tmp.p[2], (const int64_t *)tmp.p[3],
                                     (const int64_t *)tmp.p[4], (const double *)tmp.p[0], (const double *)tmp.p[1], h->ed, stride);
                else launch_k(h, k_build_ed<10>, grid_for(n_inc, 256), 256, 0, n_inc, (const int32_t *)tmp.p[5], (const uint8_t *)tmp.p[2], (const int64_t *)tmp.p[3],
                              (const int64_t *)tmp.p[4], (const double *)tmp.p[0], (const double *)tmp.p[1], h->ed, stride);
            }
            CKU(cudaStreamSynchronize(h->stream));
            CKU(h->launch_err);
        }
        TICK("shard: edge records");
        h->to_index.clear(); h->to_index.shrink_to_fit();
        max_grid = std::max<int64_t>(max_grid, grid_for(nm, 256));
        // slot -> index into `ed` of the edge the stored block comes from
        const int64_t nsl = std::max<int64_t>(s1 - s0, 1);
        std::vector<int32_t> se(nsl, 0);
        parallel_for(s1 - s0, 65536, [&](int64_t a, int64_t b) {
            for (int64_t q = a; q < b; q++) {
                const int32_t e = S.slot_edge[s0 + q];                     // -1: padding slot of the sliced storage (never read)
                se[q] = e < 0 ? 0 : (world > 1 ? inc_of[e] : e);
            }
        });
        int32_t *dse = nullptr;
        CKU(cudaMalloc((void **)&dse, nsl * sizeof(int32_t)));
        cudaError_t ce = cudaMemcpyAsync(dse, se.data(), nsl * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
        if (ce == cudaSuccess) {
            int rc = dalloc(h, &h->hz, (size_t)NM * nsl);
            if (rc) { cudaFree(dse); return rc; }
            if (D == 6) launch_k(h, k_build_hz<28>, h->lv[0].grid128, 128, 0, h->lv[0].d, (const int32_t *)dse, (const double *)h->ed, stride, h->hz);
            else launch_k(h, k_build_hz<10>, h->lv[0].grid128, 128, 0, h->lv[0].d, (const int32_t *)dse, (const double *)h->ed, stride, h->hz);
            ce = cudaStreamSynchronize(h->stream);
            if (ce == cudaSuccess) ce = h->launch_err;
        }
        cudaFree(dse);
        CKU(ce);
    }
    TICK("shard: half-edge stream");
    CKC(dalloc(h, &h->S, 1));
    CKC(dalloc(h, &h->partials, (size_t)4 * max_grid + 8));
    {
        Scalars s{};
        s.world = world;
        s.rank = rank;
        s.comm_mine = (Comm *)h->arena;
        s.spin_limit = 40000000000ll;
        if (const char *e = std::getenv("PGO_COMM_TIMEOUT_S")) s.spin_limit = (long long)(std::max(0.01, std::atof(e)) * 2.0e9);
        for (int k = 0; k < MAX_RANKS; k++) s.comm_peer[k] = (Comm *)h->arena;    // the peers' blocks are filled in when the shards connect
        s.repl_from = nl;
        for (int l = nl - 1; l >= 1; l--) if (h->lv[l].repl) s.repl_from = l;
        CKU(cudaMemcpyAsync(h->S, &s, sizeof(Scalars), cudaMemcpyHostToDevice, h->stream));
    }
    CKU(cudaStreamSynchronize(h->stream));
#undef CKC
#undef CKU
    return PGO_OK;
}

// ---- single-process multi-GPU handle (pgo_options.n_gpus > 1): one worker thread per shard, bound to the shard's device.
// Every entry point hands the same call to all workers and waits for them: the kernels of different shards wait for each
// other on the device (peer.cuh), so the shards' host sides must be able to block independently -- but the caller still sees
// ONE blocking call on ONE handle from ONE thread, like the reference's `&mut self` methods (pose_graph_optimization.rs:215, :247).
struct MultiCtx {
    std::vector<pgo_handle *> shard;
    std::vector<int> dev;
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    std::function<int(pgo_handle *, int)> job;
    uint64_t gen = 0;
    int pending = 0;
    bool stop = false;
    std::vector<int> rc;
};

// CUDA loads a kernel's code lazily, at its first launch, and that load may synchronise the whole context.  Shards that share
// ONE device also share one context, and their kernels wait for each other on the device (peer.cuh): a shard's host thread
// blocked in such a load, behind a peer's spinning kernel that waits for the very kernel being loaded, is a deadlock (the
// "concurrent kernels" caveat of lazy loading).  So a multi-GPU handle loads every kernel of this library up front, through the
// driver's module enumeration (libcuda is only dlopen'ed: no link-time dependency, nothing happens on a machine without a driver).
static bool preload_all_kernels(std::string &err) {
    void *lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = "dlopen(libcuda.so.1) failed"; return false; }
    typedef CUresult (*fn_get_module)(CUmodule *, CUfunction);
    typedef CUresult (*fn_count)(unsigned int *, CUmodule);
    typedef CUresult (*fn_enum)(CUfunction *, unsigned int, CUmodule);
    typedef CUresult (*fn_load)(CUfunction);
    fn_get_module get_module = (fn_get_module)dlsym(lib, "cuFuncGetModule");
    fn_count count = (fn_count)dlsym(lib, "cuModuleGetFunctionCount");
    fn_enum enumerate = (fn_enum)dlsym(lib, "cuModuleEnumerateFunctions");
    fn_load load = (fn_load)dlsym(lib, "cuFuncLoad");
    if (!get_module || !count || !enumerate || !load) { err = "this CUDA driver has no cuModuleEnumerateFunctions / cuFuncLoad"; return false; }
    cudaFunction_t f = nullptr;
    if (cudaGetFuncBySymbol(&f, (const void *)k_scale) != cudaSuccess || !f) { (void)cudaGetLastError(); err = "cudaGetFuncBySymbol failed"; return false; }
    CUmodule mod = nullptr;
    unsigned int n = 0;
    if (get_module(&mod, (CUfunction)f) != CUDA_SUCCESS || count(&n, mod) != CUDA_SUCCESS || n == 0) { err = "module enumeration failed"; return false; }
    std::vector<CUfunction> fs(n);
    if (enumerate(fs.data(), n, mod) != CUDA_SUCCESS) { err = "cuModuleEnumerateFunctions failed"; return false; }
    for (CUfunction g : fs) if (load(g) != CUDA_SUCCESS) { err = "cuFuncLoad failed"; return false; }
    return true;
}

static void multi_worker(MultiCtx *M, int k) {
    cudaSetDevice(M->dev[k]);
    uint64_t seen = 0;
    for (;;) {
        std::function<int(pgo_handle *, int)> job;
        {
            std::unique_lock<std::mutex> lk(M->m);
            M->cv_go.wait(lk, [&] { return M->stop || M->gen != seen; });
            if (M->stop) return;
            seen = M->gen;
            job = M->job;
        }
        const int rc = job(M->shard[k], k);
        {
            std::lock_guard<std::mutex> lk(M->m);
            M->rc[k] = rc;
            if (--M->pending == 0) M->cv_done.notify_all();
        }
    }
}

// run fn(shard, k) on every shard's worker; returns the first non-zero status (and copies that shard's message)
template <typename F> static int multi_run(pgo_handle *h, F &&fn) {
    MultiCtx *M = h->multi;
    const int n = (int)M->shard.size();
    {
        std::lock_guard<std::mutex> lk(M->m);
        M->job = std::forward<F>(fn);
        M->pending = n;
        M->gen++;
    }
    M->cv_go.notify_all();
    {
        std::unique_lock<std::mutex> lk(M->m);
        M->cv_done.wait(lk, [&] { return M->pending == 0; });
        M->job = nullptr;
    }
    for (int k = 0; k < n; k++)
        if (M->rc[k]) { if (M->shard[k]) h->err = "GPU shard " + std::to_string(k) + ": " + M->shard[k]->err; return M->rc[k]; }
    return PGO_OK;
}

namespace { void multi_shutdown(pgo_handle *h) {
    MultiCtx *M = h->multi;
    if (!M) return;
    if (!M->th.empty()) {
        multi_run(h, [](pgo_handle *s, int) { pgo_destroy(s); return 0; });
        { std::lock_guard<std::mutex> lk(M->m); M->stop = true; }
        M->cv_go.notify_all();
        for (auto &t : M->th) t.join();
    } else for (pgo_handle *s : M->shard) pgo_destroy(s);
    delete M;
    h->multi = nullptr;
} }

static void normalise_options(pgo_options &o, int world) {
    pgo_options dflt; pgo_default_options(&dflt);
    if (o.pcg_rtol <= 0) o.pcg_rtol = dflt.pcg_rtol;
    if (o.pcg_max_iterations <= 0) o.pcg_max_iterations = dflt.pcg_max_iterations;
    if (o.sort_window <= 0) o.sort_window = dflt.sort_window;
    if (o.amg_max_levels <= 0) o.amg_max_levels = dflt.amg_max_levels;
    if (o.amg_max_levels > MAX_LEVELS) o.amg_max_levels = MAX_LEVELS;
    if (o.anchor_weight == 0) o.anchor_weight = dflt.anchor_weight;
    if (o.amg_dense_max <= 0) o.amg_dense_max = dflt.amg_dense_max;
    if (o.amg_dense_max > 1024) o.amg_dense_max = 1024;
    if (o.refine_rtol <= 0) o.refine_rtol = 2e-4;
    // default upper bound on the members of an aggregate: 16 on one GPU (measured optimum at config 4, profiles/r02f_knob_sweep.log).
    // Sharded handles use 24: the partition-local level-0 aggregation leaves a larger, less regular level 1, and with 16 the
    // multilevel K-cycle needs 51 PCG iterations at world 2 and 8 where one GPU needs 41; with 24 the scipy prototype fed with
    // the library's own aggregates (tools/research/hierarchy_study.py, which reproduces 42 / 51 for the old default) gives
    // 43 (world 2) and 45 (world 8).  On one GPU 24 measures the same as 16 (41 iterations, 34.26 vs 34.19 ms).
    if (o.amg_aggregate_size <= 1) o.amg_aggregate_size = world > 1 ? 24 : 16;
}

// options + environment knobs -> handle fields (h->opt, world, rank set)
static SymbolicOptions configure_handle(pgo_handle *h) {
    h->use_amg = h->opt.preconditioner == PGO_PRECOND_AMG;
    SymbolicOptions so;
    so.world = h->world;
    so.sort_window = h->opt.sort_window;
    so.max_levels = h->opt.amg_max_levels;
    so.agg_size = h->opt.amg_aggregate_size;
    so.dense_max = h->opt.amg_dense_max;
    so.build_amg = h->use_amg;
    if (const char *e = std::getenv("PGO_REPL_MAX_ROWS")) so.repl_max_rows = std::atoll(e);      // tuning knobs
    if (const char *e = std::getenv("PGO_SPMV_TMA64")) h->spmv_tma64 = std::max(0, std::min(16, std::atoi(e)));
    if (const char *e = std::getenv("PGO_SPMV_TMA32")) h->spmv_tma32 = std::max(0, std::min(16, std::atoi(e)));
    if (const char *e = std::getenv("PGO_LPR4_MIN_ROWS")) h->lpr4_min_rows = std::atoll(e);
    if (const char *e = std::getenv("PGO_PDL")) h->pdl = std::atoi(e) != 0;
    if (const char *e = std::getenv("PGO_OMEGA_RHO")) h->omega_rho = std::min(1.9, std::max(0.5, std::atof(e)));
    if (const char *e = std::getenv("PGO_WHILE")) { h->opt_while = std::atoi(e) != 0; h->opt_while_sharded = std::atoi(e) == 2; }
    h->lowp = h->use_amg && h->opt.amg_fp64_storage == 0;
    return so;
}

// the packed value arrays must hold exactly what the per-kind counts say (the device copies read that many)
static bool check_graph_arrays(int64_t nv, const uint8_t *vkind, int64_t ne, const uint8_t *ekind, std::string &err) {
    for (int64_t i = 0; i < nv; i++) if (vkind[i] > 2) { err = "pgo_create: vertex_kind[" + std::to_string(i) + "] = " + std::to_string(vkind[i]) + " (must be 0, 1 or 2)"; return false; }
    for (int64_t k = 0; k < ne; k++) if (ekind[k] > 2) { err = "pgo_create: edge_kind[" + std::to_string(k) + "] = " + std::to_string(ekind[k]) + " (must be 0, 1 or 2)"; return false; }
    return true;
}

static int create_multi(pgo_handle **out, const pgo_options &opt_in, int64_t nv, const uint32_t *vid, const uint8_t *vkind, const double *vval,
                        int64_t ne, const uint8_t *ekind, const uint32_t *efrom, const uint32_t *eto, const double *emeas, const double *einfo) {
    const int n = opt_in.n_gpus;
    if (n > MAX_RANKS) { g_create_error = "pgo_create: n_gpus > 8"; return PGO_ERR_ARG; }
    if (opt_in.world > 1 || opt_in.rank != 0) { g_create_error = "pgo_create: n_gpus (single-process multi-GPU) and world / rank (process per GPU) are exclusive"; return PGO_ERR_ARG; }
    if (opt_in.device == -2) { g_create_error = "pgo_create: a structure-only handle has no GPUs (use world / rank for partition queries)"; return PGO_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_create_error = "no CUDA device available (this library has no CPU fallback)"; return PGO_ERR_CUDA; }
    std::vector<int> dev(n);
    for (int k = 0; k < n; k++) {
        dev[k] = opt_in.device_ids ? opt_in.device_ids[k] : k;
        if (dev[k] < 0 || dev[k] >= ndev) { g_create_error = "pgo_create: device_ids[" + std::to_string(k) + "] = " + std::to_string(dev[k]) + " but the machine has " + std::to_string(ndev) + " CUDA device(s)"; return PGO_ERR_ARG; }
    }
    for (int a = 0; a < n; a++)
        for (int b = 0; b < n; b++) {
            int ok = 1;
            if (dev[a] != dev[b] && (cudaDeviceCanAccessPeer(&ok, dev[a], dev[b]) != cudaSuccess || !ok)) {
                g_create_error = "pgo_create: GPU " + std::to_string(dev[a]) + " has no peer access to GPU " + std::to_string(dev[b]);
                return PGO_ERR_COMM;
            }
        }
    pgo_handle *h = new pgo_handle();
    h->opt = opt_in;
    h->opt.device_ids = nullptr;                    // borrowed during this call only
    h->opt.world = n; h->opt.rank = 0;
    normalise_options(h->opt, n);
    h->world = n; h->rank = 0;
    const SymbolicOptions so = configure_handle(h);
    if (!build_symbolic(h->sym, so, nv, vid, vkind, ne, ekind, efrom, eto)) return fail_create(h, PGO_ERR_ARG, h->sym.error);
    MultiCtx *M = new MultiCtx();
    h->multi = M;
    M->dev = dev;
    M->rc.assign(n, 0);
    for (int k = 0; k < n; k++) {
        pgo_handle *s = new pgo_handle(h->symp);
        s->opt = h->opt; s->opt.rank = k; s->opt.device = dev[k];
        s->world = n; s->rank = k;
        configure_handle(s);
        M->shard.push_back(s);
    }
    for (int k = 0; k < n; k++) M->th.emplace_back(multi_worker, M, k);
    bool shared_device = false;
    for (int a = 0; a < n; a++) for (int b = a + 1; b < n; b++) shared_device |= dev[a] == dev[b];
    for (int k = 0; k < n; k++) M->shard[k]->shares_device = shared_device;
    int rc = multi_run(h, [&](pgo_handle *s, int k) {
        const int rc1 = create_shard(s, vval, ne, ekind, emeas, einfo);
        if (rc1 != PGO_OK) return rc1;
        bool first_on_device = true;                     // one load per context
        for (int j = 0; j < k; j++) first_on_device &= dev[j] != dev[k];
        std::string perr;
        if (first_on_device && !preload_all_kernels(perr) && shared_device) {
            s->err = "shards share GPU " + std::to_string(dev[k]) + " and the kernels cannot be pre-loaded (" + perr + "): set CUDA_MODULE_LOADING=EAGER";
            return (int)PGO_ERR_CUDA;
        }
        return (int)PGO_OK;
    });
    if (rc == PGO_OK)
        rc = multi_run(h, [&](pgo_handle *s, int k) {            // peer access from this shard's device to the others'
            for (int j = 0; j < n; j++) {
                if (dev[j] == dev[k]) continue;
                const cudaError_t e = cudaDeviceEnablePeerAccess(dev[j], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { s->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return (int)PGO_ERR_COMM; }
                (void)cudaGetLastError();
            }
            return (int)PGO_OK;
        });
    if (rc != PGO_OK) return fail_create(h, rc, h->err);
    for (int k = 0; k < n; k++) {                     // one address space: a peer's arena is simply its pointer
        pgo_handle *s = M->shard[k];
        for (int j = 0; j < MAX_RANKS; j++) s->peer_base[j] = j < n ? M->shard[j]->arena : nullptr;
    }
    rc = multi_run(h, [&](pgo_handle *s, int) { const int r = upload_comm_peers(s); s->connected = r == PGO_OK; return r; });
    if (rc != PGO_OK) return fail_create(h, rc, h->err);
    for (auto &L : h->sym.levels) { L.ctgt.clear(); L.ctgt.shrink_to_fit(); L.cstr.clear(); L.cstr.shrink_to_fit(); L.gal_ptr.clear(); L.gal_ptr.shrink_to_fit(); L.gal_src.clear(); L.gal_src.shrink_to_fit(); }
    *out = h;
    return PGO_OK;
}

extern "C" {

int pgo_create(pgo_handle **out, const pgo_options *opt_in,
               int64_t nv, const uint32_t *vid, const uint8_t *vkind, const double *vval,
               int64_t ne, const uint8_t *ekind, const uint32_t *efrom, const uint32_t *eto,
               const double *emeas, const double *einfo) {
    if (!out) return PGO_ERR_ARG;
    *out = nullptr;
    if (!vid || !vkind || !vval || nv <= 0 || ne < 0 || (ne > 0 && (!ekind || !efrom || !eto || !emeas || !einfo))) {
        g_create_error = "pgo_create: null or empty input";
        return PGO_ERR_ARG;
    }
    if (!check_graph_arrays(nv, vkind, ne, ekind, g_create_error)) return PGO_ERR_ARG;
    if (opt_in && opt_in->n_gpus > 1) return create_multi(out, *opt_in, nv, vid, vkind, vval, ne, ekind, efrom, eto, emeas, einfo);
    pgo_handle *h = new pgo_handle();
    if (opt_in) h->opt = *opt_in; else pgo_default_options(&h->opt);
    h->opt.device_ids = nullptr;
    if (h->opt.world <= 0) h->opt.world = 1;
    normalise_options(h->opt, h->opt.world);
    h->world = h->opt.world; h->rank = h->opt.rank;
    if (h->rank < 0 || h->rank >= h->world) return fail_create(h, PGO_ERR_ARG, "pgo_create: rank out of range");
    if (h->world > MAX_RANKS) return fail_create(h, PGO_ERR_ARG, "pgo_create: world > 8");
    const SymbolicOptions so = configure_handle(h);
    if (!build_symbolic(h->sym, so, nv, vid, vkind, ne, ekind, efrom, eto)) return fail_create(h, PGO_ERR_ARG, h->sym.error);
    if (h->opt.device != -2) {                        // -2: structure-only handle (no device): symbolic-pass queries only
        const int rc = create_shard(h, vval, ne, ekind, emeas, einfo);
        if (rc != PGO_OK) return fail_create(h, rc, h->err);
        // the big transient host arrays are not needed any more
        for (auto &L : h->sym.levels) { L.ctgt.clear(); L.ctgt.shrink_to_fit(); L.cstr.clear(); L.cstr.shrink_to_fit(); L.gal_ptr.clear(); L.gal_ptr.shrink_to_fit(); L.gal_src.clear(); L.gal_src.shrink_to_fit(); }
    }
    *out = h;
    return PGO_OK;
}

// ---- sharded handles: exchange of the peer-memory handles (the caller moves the bytes, e.g. with torch.distributed.all_gather)
// a shard's blob: the CUDA IPC handle of its peer arena + the UUID of its GPU (shards that share a GPU must know it)
struct ShardBlob { cudaIpcMemHandle_t mem; unsigned char uuid[16]; };
int pgo_shard_handle_bytes(void) { return (int)sizeof(ShardBlob); }

int pgo_shard_export(pgo_handle *h, void *buf, int64_t cap) {
    if (!h || !buf || cap < (int64_t)sizeof(ShardBlob)) return PGO_ERR_ARG;
    if (h->multi) { h->err = "pgo_shard_export: a single-process multi-GPU handle (n_gpus) connects its shards itself"; return PGO_ERR_ARG; }
    if (!h->stream) { h->err = "structure-only handle"; return PGO_ERR_CUDA; }
    ShardBlob blob;
    std::memset(&blob, 0, sizeof(blob));
    CK(cudaIpcGetMemHandle(&blob.mem, h->arena));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    std::memcpy(blob.uuid, prop.uuid.bytes, sizeof(blob.uuid));
    std::memcpy(buf, &blob, sizeof(blob));
    return PGO_OK;
}

int pgo_shard_connect(pgo_handle *h, const void *all_handles, int64_t n_handles) {
    if (!h || !all_handles || n_handles != h->world) return PGO_ERR_ARG;
    if (h->multi) { h->err = "pgo_shard_connect: a single-process multi-GPU handle (n_gpus) connects its shards itself"; return PGO_ERR_ARG; }
    if (!h->stream) { h->err = "structure-only handle"; return PGO_ERR_CUDA; }
    ShardBlob mine;
    std::memcpy(&mine, (const char *)all_handles + (size_t)h->rank * sizeof(ShardBlob), sizeof(ShardBlob));
    for (int k = 0; k < h->world; k++) {
        if (k == h->rank) continue;
        ShardBlob blob;
        std::memcpy(&blob, (const char *)all_handles + (size_t)k * sizeof(blob), sizeof(blob));
        if (std::memcmp(blob.uuid, mine.uuid, sizeof(mine.uuid)) == 0) h->shares_device = true;
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, blob.mem, cudaIpcMemLazyEnablePeerAccess));
        h->peer_base[k] = (char *)p;
        h->ipc_peer[k] = true;
    }
    { int rc = upload_comm_peers(h); if (rc) return rc; }
    h->connected = true;
    return PGO_OK;
}

int pgo_get_partition(const pgo_handle *h, int32_t *world, int32_t *rank, int64_t *vertex_range, int64_t *n_remote_blocks) {
    if (!h) return PGO_ERR_ARG;
    const Symbolic &S = h->sym;
    if (world) *world = S.world;
    if (rank) *rank = h->rank;
    if (vertex_range) for (int k = 0; k <= S.world; k++) vertex_range[k] = S.vrange[k];
    if (n_remote_blocks) {
        const HostLevel &H = S.levels[0];
        for (int k = 0; k < S.world; k++) {
            int64_t c = 0;
            for (int64_t r = H.part_off[k]; r < H.part_off[k + 1]; r++)
                for (int64_t q = H.adj_ptr[r]; q < H.adj_ptr[r + 1]; q++) c += H.part_of(H.adj_nbr[q]) != k;
            n_remote_blocks[k] = c;
        }
    }
    return PGO_OK;
}

int pgo_get_sizes(const pgo_handle *h, int64_t *nv, int64_t *ne, int64_t *len, int64_t *nval) {
    if (!h) return PGO_ERR_ARG;
    if (nv) *nv = h->sym.n;
    if (ne) *ne = h->sym.n_edges;
    if (len) *len = h->sym.len;
    if (nval) *nval = h->sym.n_values;
    return PGO_OK;
}

int pgo_chi2(pgo_handle *h, double *chi2) {
    if (!h || !chi2) return PGO_ERR_ARG;
    if (h->multi) {
        double v[MAX_RANKS];
        const int rc = multi_run(h, [&](pgo_handle *s, int k) { return pgo_chi2(s, &v[k]); });
        *chi2 = v[0];
        return rc;
    }
    NEED_DEVICE(h);
    int rc = BY_D(h, chi2_launch, h);
    if (rc) return rc;
    CK(cudaMemcpyAsync(&h->hS[0], h->S, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if ((rc = comm_status(h, h->hS[0]))) return rc;
    *chi2 = h->hS[0].chi2;
    return PGO_OK;
}

int pgo_gn_step(pgo_handle *h, double lambda, int add_lambda, double *norm_dx, double *chi2, int32_t *pcg_iterations) {
    if (!h) return PGO_ERR_ARG;
    if (h->multi) {                                  // scalars come back identical on every shard
        double nd[MAX_RANKS], c2[MAX_RANKS]; int32_t it[MAX_RANKS];
        const int rc = multi_run(h, [&](pgo_handle *s, int k) { return pgo_gn_step(s, lambda, add_lambda, &nd[k], &c2[k], &it[k]); });
        if (rc != PGO_OK && rc != PGO_ERR_NOT_CONVERGED) return rc;
        if (norm_dx) *norm_dx = nd[0];
        if (chi2) *chi2 = c2[0];
        if (pcg_iterations) *pcg_iterations = it[0];
        h->have_step = true;
        return rc;
    }
    NEED_DEVICE(h);
    auto mark = [&](int i) { cudaEventRecord(h->ev[i], h->stream); };
    int64_t lc[6];
    mark(0); lc[0] = h->launch_count;
    int rc = BY_D(h, assemble, h, lambda, add_lambda);
    if (rc) return rc;
    mark(1); lc[1] = h->launch_count;
    rc = BY_D(h, amg_setup, h);
    if (rc) return rc;
    mark(2); lc[2] = h->launch_count;
    int32_t iters = 0;
    int src = BY_D(h, solve_refined, h, &iters);
    if (src == PGO_ERR_SOLVER && h->use_amg && h->omega_ready) {
        // The smoother dampings were estimated once, on the first H of this handle.  H is re-linearised every step (and the caller may
        // have moved the poses far away with pgo_set_poses): if rho(Dinv H) has grown past the margin the cycle stops being positive
        // definite and PCG reports a breakdown.  Re-estimate on the current H and solve once more before giving up.
        h->omega_ready = false;
        rc = BY_D(h, assemble, h, lambda, add_lambda);          // the solve consumed b
        if (rc) return rc;
        rc = BY_D(h, amg_setup, h);
        if (rc) return rc;
        src = BY_D(h, solve_refined, h, &iters);
    }
    if (src != PGO_OK && src != PGO_ERR_NOT_CONVERGED) return src;
    mark(3); lc[3] = h->launch_count;
    rc = BY_D(h, retract, h, 1.0);
    if (rc) return rc;
    mark(4); lc[4] = h->launch_count;
    rc = BY_D(h, chi2_launch, h);
    if (rc) return rc;
    mark(5); lc[5] = h->launch_count;
    CK(cudaMemcpyAsync(&h->hS[0], h->S, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 5; i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
        h->ms[i] = ms; h->launches[i] = lc[i + 1] - lc[i];
    }
    h->have_step = true;
    if ((rc = comm_status(h, h->hS[0]))) return rc;
    if (norm_dx) *norm_dx = std::sqrt(h->hS[0].norm2_dx);
    if (chi2) *chi2 = h->hS[0].chi2;
    if (pcg_iterations) *pcg_iterations = iters;
    return src;
}

int pgo_undo_last_step(pgo_handle *h) {
    if (!h) return PGO_ERR_ARG;
    if (h->multi) return multi_run(h, [&](pgo_handle *s, int) { return pgo_undo_last_step(s); });
    NEED_DEVICE(h);
    if (!h->have_step) { h->err = "pgo_undo_last_step: no step to undo"; return PGO_ERR_ARG; }
    int rc = BY_D(h, retract, h, -1.0);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->have_step = false;                            // a step can be undone once
    return PGO_OK;
}

int pgo_linearize_and_solve(pgo_handle *h, int32_t *pcg_iterations) {
    if (!h) return PGO_ERR_ARG;
    if (h->multi) {
        int32_t it[MAX_RANKS];
        const int rc = multi_run(h, [&](pgo_handle *s, int k) { return pgo_linearize_and_solve(s, &it[k]); });
        if (pcg_iterations) *pcg_iterations = it[0];
        return rc;
    }
    NEED_DEVICE(h);
    h->have_step = false;                            // h->x is overwritten: the last step's dx is gone
    int rc = BY_D(h, assemble, h, 0.0, 0);
    if (rc) return rc;
    rc = BY_D(h, amg_setup, h);
    if (rc) return rc;
    return BY_D(h, solve_refined, h, pcg_iterations);
}

// vertex values of the rows this rank owns are a contiguous span of the packed array (contiguous vertex ranges)
static void owned_span(const pgo_handle *h, int64_t *o0, int64_t *o1) {
    const Symbolic &S = h->sym;
    const int64_t a = S.vrange[h->rank], b = S.vrange[h->rank + 1];
    *o0 = S.vvalofs[a]; *o1 = b < S.n ? S.vvalofs[b] : S.n_values;
}

int pgo_get_poses(pgo_handle *h, double *out, int64_t n_values) {
    if (!h || !out) return PGO_ERR_ARG;
    if (h->multi) return multi_run(h, [&](pgo_handle *s, int) { return pgo_get_poses(s, out, n_values); });   // disjoint spans of `out`
    NEED_DEVICE(h);
    const Symbolic &S = h->sym;
    if (n_values != S.n_values) { h->err = "pgo_get_poses: wrong buffer length"; return PGO_ERR_ARG; }
    int64_t o0, o1;
    owned_span(h, &o0, &o1);
    if (S.D == 6) launch_k(h, k_export_poses_se3, grid_for(h->n_loc, 256), 256, 0, h->n_loc, h->row_valofs, h->poses, h->vstage);
    else launch_k(h, k_export_poses, grid_for(h->n_loc, 256), 256, 0, h->n_loc, h->row_valofs, h->lv[0].d.vkind, h->poses, h->vstage);
    h->launch_count += 1;
    CK(cudaGetLastError()); CK(h->launch_err);
    CK(cudaMemcpyAsync(out + o0, h->vstage + o0, (o1 - o0) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PGO_OK;
}

int pgo_set_poses(pgo_handle *h, const double *in, int64_t n_values) {
    if (!h || !in) return PGO_ERR_ARG;
    if (h->multi) return multi_run(h, [&](pgo_handle *s, int) { return pgo_set_poses(s, in, n_values); });
    NEED_DEVICE(h);
    const Symbolic &S = h->sym;
    if (n_values != S.n_values) { h->err = "pgo_set_poses: wrong buffer length"; return PGO_ERR_ARG; }
    int64_t o0, o1;
    owned_span(h, &o0, &o1);
    CK(cudaMemcpyAsync(h->vstage + o0, in + o0, (o1 - o0) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (S.D == 6) launch_k(h, k_import_poses_se3, grid_for(h->n_loc, 256), 256, 0, h->n_loc, h->row_valofs, h->vstage, h->poses);
    else launch_k(h, k_import_poses, grid_for(h->n_loc, 256), 256, 0, h->n_loc, h->row_valofs, h->lv[0].d.vkind, h->vstage, h->poses);
    h->launch_count += 1;
    CK(cudaGetLastError()); CK(h->launch_err);
    CK(cudaStreamSynchronize(h->stream));   // `in` is borrowed for the duration of the call only
    h->have_step = false;
    return PGO_OK;
}

int pgo_snapshot_poses(pgo_handle *h) {
    if (!h) return PGO_ERR_ARG;
    if (h->multi) return multi_run(h, [&](pgo_handle *s, int) { return pgo_snapshot_poses(s); });
    NEED_DEVICE(h);
    const size_t cnt = (size_t)(h->sym.D == 6 ? 8 : 4) * h->n_pad_loc;
    if (!h->poses_saved) { int rc = dalloc(h, &h->poses_saved, cnt, false); if (rc) return rc; }
    CK(cudaMemcpyAsync(h->poses_saved, h->poses, cnt * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PGO_OK;
}

int pgo_restore_poses(pgo_handle *h) {
    if (!h) return PGO_ERR_ARG;
    if (h->multi) return multi_run(h, [&](pgo_handle *s, int) { return pgo_restore_poses(s); });
    NEED_DEVICE(h);
    if (!h->poses_saved) { h->err = "pgo_restore_poses: no snapshot"; return PGO_ERR_ARG; }
    CK(cudaMemcpyAsync(h->poses, h->poses_saved, (size_t)(h->sym.D == 6 ? 8 : 4) * h->n_pad_loc * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->have_step = false;
    return PGO_OK;
}

int pgo_get_dx(pgo_handle *h, double *out, int64_t len) {
    if (!h || !out) return PGO_ERR_ARG;
    if (h->multi) return multi_run(h, [&](pgo_handle *s, int) { return pgo_get_dx(s, out, len); });           // disjoint entries of `out`
    NEED_DEVICE(h);
    const Symbolic &S = h->sym;
    if (len != S.len) { h->err = "pgo_get_dx: wrong buffer length"; return PGO_ERR_ARG; }
    const int VS = S.D == 6 ? 6 : 4;
    std::vector<double> xs((size_t)VS * h->n_pad_loc);
    CK(cudaMemcpyAsync(xs.data(), h->x, xs.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int64_t r = 0; r < h->n_loc; r++) {
        const int64_t v = S.perm[h->row0 + r];
        const int d = KDIM[S.vkind[v]];
        for (int c = 0; c < d; c++) out[S.voffset[v] + c] = xs[VS * r + c];
    }
    return PGO_OK;
}

int pgo_get_pattern(const pgo_handle *hc, int64_t *n, int64_t *nnz, int32_t *col_ptr, int32_t *row_idx) {
    pgo_handle *h = const_cast<pgo_handle *>(hc);
    if (!h) return PGO_ERR_ARG;
    if (!build_csc_pattern(h->sym)) { h->err = h->sym.error; return PGO_ERR_UNSUPPORTED; }
    if (n) *n = h->sym.len;
    if (nnz) *nnz = (int64_t)h->sym.csc_row.size();
    if (col_ptr) std::memcpy(col_ptr, h->sym.csc_ptr.data(), h->sym.csc_ptr.size() * sizeof(int32_t));
    if (row_idx) std::memcpy(row_idx, h->sym.csc_row.data(), h->sym.csc_row.size() * sizeof(int32_t));
    return PGO_OK;
}

int pgo_get_block_structure(const pgo_handle *hc, int64_t *n_blocks, int64_t *brow, int32_t *bcol, int64_t *eslots) {
    pgo_handle *h = const_cast<pgo_handle *>(hc);
    if (!h) return PGO_ERR_ARG;
    build_canonical(h->sym);
    if (n_blocks) *n_blocks = (int64_t)h->sym.bcol.size();
    if (brow) std::memcpy(brow, h->sym.brow_ptr.data(), h->sym.brow_ptr.size() * sizeof(int64_t));
    if (bcol) std::memcpy(bcol, h->sym.bcol.data(), h->sym.bcol.size() * sizeof(int32_t));
    if (eslots) std::memcpy(eslots, h->sym.edge_slots.data(), h->sym.edge_slots.size() * sizeof(int64_t));
    return PGO_OK;
}

int pgo_get_anchor(const pgo_handle *h, int64_t *v) {
    if (!h || !v) return PGO_ERR_ARG;
    *v = h->sym.anchor;
    return PGO_OK;
}

// Sharded handles return the block rows they own: entries of other ranks' block rows are left untouched,
// so the caller can merge the ranks' outputs (they are disjoint).
int pgo_get_system(pgo_handle *h, double lambda, int add_lambda, double *csc_values, double *b) {
    if (!h) return PGO_ERR_ARG;
    if (h->multi) {
        // the lazily built structures live in the shared symbolic pass: build them once, here, before the shards read them
        if (!build_csc_pattern(h->sym)) { h->err = h->sym.error; return PGO_ERR_UNSUPPORTED; }
        return multi_run(h, [&](pgo_handle *s, int) { return pgo_get_system(s, lambda, add_lambda, csc_values, b); });   // disjoint block rows
    }
    NEED_DEVICE(h);
    Symbolic &S = h->sym;
    if (!build_csc_pattern(S)) { h->err = S.error; return PGO_ERR_UNSUPPORTED; }
    int rc = BY_D(h, assemble, h, lambda, add_lambda);
    if (rc) return rc;
    HostLevel &H = S.levels[0];
    const int D = S.D, DD = D * D, VS = D == 6 ? 6 : 4;
    const int64_t r0 = h->row0, s0 = H.part_slot[h->rank], s1 = H.part_slot[h->rank + 1], npl = h->n_pad_loc;
    std::vector<double> val((size_t)DD * std::max<int64_t>(s1 - s0, 1)), diag((size_t)DD * npl), rv((size_t)VS * npl);
    CK(cudaMemcpyAsync(val.data(), h->lv[0].d.val, val.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(diag.data(), h->lv[0].d.diag, diag.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(rv.data(), h->r, rv.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    // canonical block values of the owned block rows (duplicate edges between a vertex pair sum into one block)
    std::vector<double> blk((size_t)DD * S.bcol.size(), 0.0);
    auto find = [&](int32_t rr, int32_t cc) -> int64_t {
        auto bb = S.bcol.begin() + S.brow_ptr[rr], ee = S.bcol.begin() + S.brow_ptr[rr + 1];
        return std::lower_bound(bb, ee, cc) - S.bcol.begin();
    };
    for (int64_t r = 0; r < h->n_loc; r++) {
        const int32_t u = S.perm[r0 + r];
        const int lane = (int)(r & 31);
        double *d = &blk[DD * find(u, u)];
        for (int c = 0; c < DD; c++) d[c] += diag[(size_t)c * npl + r];
        for (int64_t qi = H.adj_ptr[r0 + r]; qi < H.adj_ptr[r0 + r + 1]; qi++) {
            const int64_t slot = H.adj_slot[qi] - s0, cnt = H.adj_cnt[qi];
            const int32_t v = S.perm[H.adj_nbr[qi]];
            double *o = &blk[DD * find(u, v)];
            const double *src = val.data() + (slot - lane) * DD + lane;
            for (int c = 0; c < DD; c++) o[c] += src[c * cnt];
        }
    }
    if (csc_values) {
        int64_t o = 0;
        for (int64_t v = 0; v < S.n; v++) {
            const int dv = KDIM[S.vkind[v]];
            for (int c = 0; c < dv; c++)
                for (int64_t pp = S.brow_ptr[v]; pp < S.brow_ptr[v + 1]; pp++) {
                    const int32_t u = S.bcol[pp];
                    const int du = KDIM[S.vkind[u]];
                    if (u >= S.vrange[h->rank] && u < S.vrange[h->rank + 1]) {
                        const int64_t bi = find(u, (int32_t)v);                    // block (row u, col v)
                        for (int rr = 0; rr < du; rr++) csc_values[o + rr] = blk[DD * bi + D * rr + c];
                    }
                    o += du;
                }
        }
    }
    if (b) {
        for (int64_t r = 0; r < h->n_loc; r++) {
            const int64_t v = S.perm[r0 + r];
            const int d = KDIM[S.vkind[v]];
            for (int c = 0; c < d; c++) b[S.voffset[v] + c] = rv[VS * r + c];
        }
    }
    return PGO_OK;
}

int pgo_get_timings(pgo_handle *h, double *ms, int64_t *launches, int32_t n) {
    if (!h) return PGO_ERR_ARG;
    if (h->multi) {                                  // per phase: the slowest shard's time, the launches of all shards
        for (int i = 0; i < n && i < PGO_NUM_PHASES; i++) {
            double m = 0; int64_t c = 0;
            for (pgo_handle *s : h->multi->shard) { m = std::max(m, s->ms[i]); c += s->launches[i]; }
            if (ms) ms[i] = m;
            if (launches) launches[i] = c;
        }
        return PGO_OK;
    }
    for (int i = 0; i < n && i < PGO_NUM_PHASES; i++) {
        if (ms) ms[i] = h->ms[i];
        if (launches) launches[i] = h->launches[i];
    }
    return PGO_OK;
}

int pgo_time_spmv(pgo_handle *h, int32_t repeats, double *avg_ms) {
    if (!h || !avg_ms || repeats <= 0) return PGO_ERR_ARG;
    if (h->multi) {
        double v[MAX_RANKS] = {0};
        const int rc = multi_run(h, [&](pgo_handle *s, int k) { return pgo_time_spmv(s, repeats, &v[k]); });
        *avg_ms = *std::max_element(v, v + h->multi->shard.size());
        return rc;
    }
    NEED_DEVICE(h);
    LevelBuf &B = h->lv[0];
    // p -> q with the PCG SpMV; done-flag test disabled so the launches always do the work, no cross-rank reduction
    halo_pull(h, 0, h->p, h->sym.D == 6 ? 6 : 4, 0);
    const bool f32 = h->lowp && std::getenv("PGO_TIME_SPMV_F32") != nullptr;     // diagnostic: time the fp32-storage residual product instead
    auto launch = [&]() {
        if (h->sym.D == 6) { if (f32) spmv_launch<6, 1, FIN_NONE, float>(h, 0, h->p, h->r, h->q, 0.0, nullptr, nullptr, 0); else spmv_launch<6, 0, FIN_NONE, double>(h, 0, h->p, nullptr, h->q, 0.0, nullptr, nullptr, 0); }
        else { if (f32) spmv_launch<3, 1, FIN_NONE, float>(h, 0, h->p, h->r, h->q, 0.0, nullptr, nullptr, 0); else spmv_launch<3, 0, FIN_NONE, double>(h, 0, h->p, nullptr, h->q, 0.0, nullptr, nullptr, 0); }
    };
    for (int i = 0; i < 3; i++) launch();
    CK(cudaEventRecord(h->ev[PGO_NUM_PHASES], h->stream));
    for (int i = 0; i < repeats; i++) launch();
    CK(cudaEventRecord(h->ev[PGO_NUM_PHASES + 1], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[PGO_NUM_PHASES], h->ev[PGO_NUM_PHASES + 1]));
    *avg_ms = ms / repeats;
    h->ms[PGO_PHASE_SPMV_FINE] = *avg_ms;
    h->launch_count += repeats + 3;
    return PGO_OK;
}

int pgo_time_coarse(pgo_handle *h, int32_t level, int32_t repeats, double *avg_ms) {
    if (!h || !avg_ms || repeats <= 0) return PGO_ERR_ARG;
    if (h->multi) {
        double v[MAX_RANKS] = {0};
        const int rc = multi_run(h, [&](pgo_handle *s, int k) { return pgo_time_coarse(s, level, repeats, &v[k]); });
        *avg_ms = *std::max_element(v, v + h->multi->shard.size());
        return rc;
    }
    NEED_DEVICE(h);
    if (!h->have_step) { h->err = "pgo_time_coarse: run a Gauss-Newton step first (the hierarchy must be set up)"; return PGO_ERR_ARG; }
    return BY_D(h, time_coarse, h, level, repeats, avg_ms);
}

int pgo_get_stats(const pgo_handle *h, int64_t *rows, int64_t *offdiag, int64_t *levels, int64_t *bytes) {
    if (!h) return PGO_ERR_ARG;
    if (h->multi) {                                  // totals over the shards
        if (rows) *rows = h->sym.n;
        if (offdiag) *offdiag = 2 * h->sym.n_edges;
        if (levels) *levels = (int64_t)h->sym.levels.size();
        if (bytes) { *bytes = 0; for (pgo_handle *s : h->multi->shard) *bytes += (int64_t)s->device_bytes; }
        return PGO_OK;
    }
    if (rows) *rows = h->stream ? h->n_loc : h->sym.n;
    if (offdiag) {
        const HostLevel &H = h->sym.levels[0];
        *offdiag = h->stream ? H.adj_ptr[H.part_off[h->rank + 1]] - H.adj_ptr[H.part_off[h->rank]] : 2 * h->sym.n_edges;
    }
    if (levels) *levels = (int64_t)h->sym.levels.size();
    if (bytes) *bytes = (int64_t)h->device_bytes;
    return PGO_OK;
}

int pgo_get_level_sizes(const pgo_handle *h, int32_t max_levels, int64_t *rows, int64_t *blocks) {
    if (!h) return PGO_ERR_ARG;
    const int nl = (int)h->sym.levels.size();
    for (int l = 0; l < nl && l < max_levels; l++) {
        if (rows) rows[l] = h->sym.levels[l].n;
        if (blocks) blocks[l] = h->sym.levels[l].adj_ptr[h->sym.levels[l].n_pad];
    }
    return nl;
}

int pgo_get_aggregates(const pgo_handle *h, int32_t level, int32_t *coarse_row, int64_t n) {
    if (!h || !coarse_row) return PGO_ERR_ARG;
    const Symbolic &S = h->sym;
    if (level < 0 || level + 1 >= (int)S.levels.size() || S.levels[level].agg.empty()) return PGO_ERR_ARG;
    const HostLevel &L = S.levels[level];
    if (level == 0) {                                // per vertex, lut order
        if (n != S.n) return PGO_ERR_ARG;
        for (int64_t v = 0; v < S.n; v++) coarse_row[v] = L.agg[S.iperm[v]];
    } else {                                         // per row of the level (global padded numbering, -1 for padding rows)
        if (n != L.n_pad) return PGO_ERR_ARG;
        for (int64_t r = 0; r < L.n_pad; r++) coarse_row[r] = L.agg[r];
    }
    return PGO_OK;
}

} // extern "C"
