// C ABI (include/pgo_b200.h) + device-resident Gauss-Newton / PCG driver.
//
// One pgo_gn_step = the body of the reference's optimisation loop
// (pose_graph_optimization.rs:271-274): build_linear_system + solve + update_nodes + global_error,
// with the UMFPACK factorisation replaced by a preconditioned conjugate gradient that runs entirely
// on the GPU: the PCG iterations are captured once into a CUDA graph; every kernel tests a device
// `done` flag, so the host only polls a pinned copy of the scalars once per graph launch.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pgo_b200.h"
#include "kernels.cuh"

using namespace pgo;

namespace {

thread_local std::string g_create_error;

#define NEED_DEVICE(h)                                                                             \
    do {                                                                                           \
        if (!(h)->stream) { (h)->err = "structure-only handle: no device state (there is no CPU fallback)"; return PGO_ERR_CUDA; } \
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
    double *r = nullptr, *xa = nullptr, *res = nullptr, *e = nullptr;  // V-cycle work vectors
    double omega = 0.6;
    int grid128 = 0;
};

} // namespace

struct pgo_handle {
    Symbolic sym;
    pgo_options opt{};
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<LevelBuf> lv;
    std::vector<void *> allocs;
    size_t device_bytes = 0;
    // level-0 state
    double *poses = nullptr, *poses_saved = nullptr, *hz = nullptr, *ed = nullptr;
    double *vstage = nullptr;          // g2o-layout vertex values (n_values) for set/get_poses
    int64_t *row_valofs = nullptr;     // [n] offset of each storage row's values in vstage
    int2 *ends = nullptr;
    double *x = nullptr, *r = nullptr, *p = nullptr, *q = nullptr, *z = nullptr;
    Scalars *S = nullptr, *hS = nullptr;   // device / pinned host (2 slots)
    double *partials = nullptr;
    double *Ainv = nullptr;
    bool dense_coarsest = false, use_amg = false, omega_ready = false;
    int64_t anchor_row = -1;
    cudaGraphExec_t pcg_graph = nullptr;
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

inline int grid_for(int64_t n, int bs) { return (int)((n + bs - 1) / bs); }

// ---- V-cycle: z = M^-1 r, all launches on h->stream.  FINK: finalize kind of the last kernel on level 0.
template <int FINK> void launch_post(pgo_handle *h, int l, const double *r_l, double *out) {
    LevelBuf &B = h->lv[l];
    k_spmv<3, 2, FINK><<<B.grid128, 128, 0, h->stream>>>(B.d, B.xa, r_l, out, B.omega, h->S, h->partials, 1);
}

template <int FINK> void vcycle(pgo_handle *h, int l, const double *r_l, double *out) {
    LevelBuf &B = h->lv[l];
    const int last = (int)h->lv.size() - 1;
    if (l == last) {
        if (h->dense_coarsest) {
            k_dense_apply<3><<<1, 256, 0, h->stream>>>(B.d, h->Ainv, r_l, out, h->S);
            h->launch_count += 1;
        } else {
            // no dense solve possible: a few damped block-Jacobi sweeps
            k_dinv_apply<3, FIN_NONE><<<B.grid128, 128, 0, h->stream>>>(B.d, r_l, B.xa, B.omega, h->S, h->partials, 1);
            k_spmv<3, 2, FIN_NONE><<<B.grid128, 128, 0, h->stream>>>(B.d, B.xa, r_l, B.res, B.omega, h->S, h->partials, 1);
            k_spmv<3, 2, FIN_NONE><<<B.grid128, 128, 0, h->stream>>>(B.d, B.res, r_l, out, B.omega, h->S, h->partials, 1);
            h->launch_count += 3;
        }
        return;
    }
    LevelBuf &C = h->lv[l + 1];
    k_dinv_apply<3, FIN_NONE><<<B.grid128, 128, 0, h->stream>>>(B.d, r_l, B.xa, B.omega, h->S, h->partials, 1);
    k_spmv<3, 1, FIN_NONE><<<B.grid128, 128, 0, h->stream>>>(B.d, B.xa, r_l, B.res, 0.0, h->S, h->partials, 1);
    k_restrict3<<<C.grid128, 128, 0, h->stream>>>(B.d, C.d, B.res, C.r, h->S);
    vcycle<FIN_NONE>(h, l + 1, C.r, C.e);
    k_prolong3<<<B.grid128, 128, 0, h->stream>>>(B.d, C.d, C.e, B.xa, h->S);
    if (l == 0) launch_post<FINK>(h, l, r_l, out);
    else launch_post<FIN_NONE>(h, l, r_l, out);
    h->launch_count += 5;
}

template <int FINK> void precondition(pgo_handle *h) {   // z = M^-1 r (+ r.z)
    if (h->use_amg && h->lv.size() > 1) vcycle<FINK>(h, 0, h->r, h->z);
    else {
        LevelBuf &B = h->lv[0];
        k_dinv_apply<3, FINK><<<B.grid128, 128, 0, h->stream>>>(B.d, h->r, h->z, 1.0, h->S, h->partials, 1);
        h->launch_count += 1;
    }
}

void pcg_iteration(pgo_handle *h) {
    LevelBuf &B = h->lv[0];
    const int64_t nd = B.d.n_pad * 4;
    const int g256 = grid_for(nd / 2, 256);
    k_spmv<3, 0, FIN_PQ><<<B.grid128, 128, 0, h->stream>>>(B.d, h->p, nullptr, h->q, 0.0, h->S, h->partials, 1);
    k_update_xr<3><<<g256, 256, 0, h->stream>>>(B.d.n_pad, h->x, h->r, h->p, h->q, h->S);
    precondition<FIN_RZ>(h);
    k_update_p<3><<<g256, 256, 0, h->stream>>>(B.d.n_pad, h->p, h->z, h->S);
    h->launch_count += 3;
}

int build_pcg_graph(pgo_handle *h) {
    if (h->pcg_graph) return PGO_OK;
    cudaGraph_t g = nullptr;
    int64_t before = h->launch_count;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < h->chunk; i++) pcg_iteration(h);
    CK(cudaStreamEndCapture(h->stream, &g));
    h->launches_per_iter = (h->launch_count - before) / h->chunk;
    h->launch_count = before;
    CK(cudaGraphInstantiate(&h->pcg_graph, g, 0));
    cudaGraphDestroy(g);
    return PGO_OK;
}

// ---- assemble the Gauss-Newton system at the current poses (H in lv[0], b in h->r)
int assemble(pgo_handle *h, double lambda, int add_lambda) {
    LevelBuf &B = h->lv[0];
    k_assemble_se2<<<B.grid128, 128, 0, h->stream>>>(B.d, h->poses, h->hz, h->r, h->anchor_row, h->opt.anchor_weight,
                                                      add_lambda ? lambda : 0.0);
    h->launch_count += 1;
    CK(cudaGetLastError());
    return PGO_OK;
}

// power iteration for rho(Dinv H) on one level -> damping of the block-Jacobi smoother
int estimate_omega(pgo_handle *h, int l) {
    LevelBuf &B = h->lv[l];
    const int64_t nd = B.d.n_pad * 4;
    std::vector<double> v(nd, 0.0);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (int64_t i = 0; i < B.d.n; i++)
        for (int c = 0; c < 3; c++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v[i * 4 + c] = (double)(s >> 11) / 9007199254740992.0 - 0.5; }
    if (B.d.vkind) {   // keep the landmark padding unknown out of it
        for (int64_t i = 0; i < B.d.n; i++) if (h->sym.vkind[h->sym.perm[i]] == 1) v[i * 4 + 2] = 0.0;
    }
    double *a = B.xa, *b = B.res;
    CK(cudaMemcpyAsync(a, v.data(), nd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    double rho = 1.0;
    for (int it = 0; it < 12; it++) {
        // b = H a ; a' = Dinv b ; rho ~ |a'| / |a|
        k_spmv<3, 0, FIN_NONE><<<B.grid128, 128, 0, h->stream>>>(B.d, a, nullptr, b, 0.0, h->S, h->partials, 0);
        k_dinv_apply<3, FIN_NONE><<<B.grid128, 128, 0, h->stream>>>(B.d, b, a, 1.0, h->S, h->partials, 0);
        CK(cudaMemcpyAsync(v.data(), a, nd * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        double nrm = 0.0;
        for (double t : v) nrm += t * t;
        nrm = std::sqrt(nrm);
        if (!(nrm > 0.0) || !std::isfinite(nrm)) { rho = 2.0; break; }
        rho = nrm;                       // |a| was normalised to 1
        for (double &t : v) t /= nrm;
        CK(cudaMemcpyAsync(a, v.data(), nd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    // first iterate is not normalised: rho from the last step only
    B.omega = 4.0 / (3.0 * 1.1 * rho);
    if (B.omega > 1.0) B.omega = 1.0;
    return PGO_OK;
}

// numeric setup of the hierarchy for the current H: coarse positions, Galerkin products, inverses
int amg_setup(pgo_handle *h) {
    if (!h->use_amg || h->lv.size() < 2) return PGO_OK;
    const int last = (int)h->lv.size() - 1;
    for (int l = 0; l < last; l++) {
        LevelBuf &F = h->lv[l], &C = h->lv[l + 1];
        k_coarse_pos<<<C.grid128, 128, 0, h->stream>>>(F.d, C.d);
        CK(cudaMemsetAsync(C.d.val, 0, sizeof(double) * 9 * C.d.n_slots, h->stream));
        CK(cudaMemsetAsync(C.d.diag, 0, sizeof(double) * 9 * C.d.n_pad, h->stream));
        k_galerkin3<<<F.grid128, 128, 0, h->stream>>>(F.d, C.d);
        k_invert_diag3<<<C.grid128, 128, 0, h->stream>>>(C.d);
        h->launch_count += 3;
    }
    if (h->dense_coarsest) {
        LevelBuf &C = h->lv[last];
        const int m = (int)C.d.n * 3;
        k_dense_invert<3><<<1, 256, sizeof(double) * m * m, h->stream>>>(C.d, h->Ainv);
        h->launch_count += 1;
    }
    CK(cudaGetLastError());
    if (!h->omega_ready) {
        for (int l = 0; l < (int)h->lv.size(); l++) {
            if (l == last && h->dense_coarsest) continue;
            int rc = estimate_omega(h, l);
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
    // counters must survive (they are always 0 between kernels); everything else is re-initialised
    CK(cudaMemcpyAsync(h->S, &s, offsetof(Scalars, counter), cudaMemcpyHostToDevice, h->stream));
    return PGO_OK;
}

// solve H x = b (b in h->r, destroyed) ; x in h->x
int solve(pgo_handle *h, int32_t *iters_out) {
    LevelBuf &B = h->lv[0];
    const int64_t nd = B.d.n_pad * 4;
    int rc = reset_scalars(h);
    if (rc) return rc;
    rc = build_pcg_graph(h);
    if (rc) return rc;
    CK(cudaMemsetAsync(h->x, 0, nd * sizeof(double), h->stream));
    precondition<FIN_RZ_INIT>(h);
    CK(cudaMemcpyAsync(h->p, h->z, nd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    // keep two graph launches in flight; poll the pinned scalars of the older one
    int slot = 0, inflight = 0;
    int64_t launched = 0;
    const int64_t max_graphs = (int64_t)h->opt.pcg_max_iterations / h->chunk + 2;
    bool done = false;
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
    if (h->hS[0].status == ST_BREAKDOWN) {
        h->err = "PCG breakdown: H is not positive definite (isolated vertex, or graph without a pose-pose edge to anchor)";
        return PGO_ERR_SOLVER;
    }
    if (h->hS[0].status == ST_MAXIT) { h->err = "PCG did not converge within pcg_max_iterations"; return PGO_ERR_NOT_CONVERGED; }
    return PGO_OK;
}

int retract(pgo_handle *h, double sign) {
    LevelBuf &B = h->lv[0];
    k_retract_se2<<<grid_for(B.d.n, 256), 256, 0, h->stream>>>(B.d, h->poses, h->x, sign, h->S, h->partials);
    h->launch_count += 1;
    CK(cudaGetLastError());
    return PGO_OK;
}

int chi2_launch(pgo_handle *h) {
    k_chi2_se2<<<std::max(1, grid_for(h->sym.n_edges, 256)), 256, 0, h->stream>>>(h->sym.n_edges, h->ends, h->ed, h->poses, h->S, h->partials);
    h->launch_count += 1;
    CK(cudaGetLastError());
    return PGO_OK;
}

int fail_create(pgo_handle *h, int rc, const std::string &msg) {
    g_create_error = msg.empty() ? h->err : msg;
    pgo_destroy(h);
    return rc;
}

} // namespace

// ================================================================================================
extern "C" {

const char *pgo_version(void) { return "pgo_b200 0.1 (sm_100a, fp64)"; }

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
}

const char *pgo_last_error(const pgo_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void pgo_destroy(pgo_handle *h) {
    if (!h) return;
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->pcg_graph) cudaGraphExecDestroy(h->pcg_graph);
    for (void *p : h->allocs) cudaFree(p);
    if (h->hS) cudaFreeHost(h->hS);
    for (auto &e : h->ev) if (e) cudaEventDestroy(e);
    for (auto &e : h->poll_ev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

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
    pgo_handle *h = new pgo_handle();
    if (opt_in) h->opt = *opt_in; else pgo_default_options(&h->opt);
    if (h->opt.pcg_rtol <= 0) h->opt.pcg_rtol = 1e-10;
    if (h->opt.pcg_max_iterations <= 0) h->opt.pcg_max_iterations = 200000;
    if (h->opt.sort_window <= 0) h->opt.sort_window = 2048;
    if (h->opt.amg_max_levels <= 0) h->opt.amg_max_levels = 12;
    if (h->opt.anchor_weight == 0) h->opt.anchor_weight = 1e7;
    h->use_amg = h->opt.preconditioner == PGO_PRECOND_AMG;

    SymbolicOptions so;
    so.sort_window = h->opt.sort_window;
    so.amg_max_levels = h->opt.amg_max_levels;
    so.coarsest_max = 48;
    so.build_amg = h->use_amg;
    if (!build_symbolic(h->sym, so, nv, vid, vkind, ne, ekind, efrom, eto)) return fail_create(h, PGO_ERR_ARG, h->sym.error);
    Symbolic &S = h->sym;
    if (S.D != 3) return fail_create(h, PGO_ERR_UNSUPPORTED, "SE3 graphs are not supported by this build yet");

    if (h->opt.device == -2) { *out = h; return PGO_OK; }   // structure-only handle (no device): symbolic-pass queries only
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail_create(h, PGO_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (h->opt.device >= 0) { if (cudaSetDevice(h->opt.device) != cudaSuccess) return fail_create(h, PGO_ERR_CUDA, "cudaSetDevice failed"); }
    cudaGetDevice(&h->device);
#define CKC(call) do { int rc_ = (call); if (rc_) return fail_create(h, rc_, ""); } while (0)
#define CKU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail_create(h, PGO_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
    CKU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto &e : h->ev) CKU(cudaEventCreate(&e));
    for (auto &e : h->poll_ev) CKU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CKU(cudaHostAlloc((void **)&h->hS, 2 * sizeof(Scalars), cudaHostAllocDefault));
    std::memset(h->hS, 0, 2 * sizeof(Scalars));

    // ---- level structures
    const int nl = (int)S.levels.size();
    h->lv.resize(nl);
    int64_t max_grid = 1;
    for (int l = 0; l < nl; l++) {
        HostLevel &H = S.levels[l];
        LevelBuf &B = h->lv[l];
        LevelDev &d = B.d;
        d.n = H.n; d.n_pad = H.n_pad; d.n_slices = H.n_slices; d.n_slots = H.n_slots;
        B.grid128 = grid_for(H.n_pad, 128);
        max_grid = std::max<int64_t>(max_grid, B.grid128);
        int64_t *sp; int32_t *dg; uint32_t *cl;
        CKC(upload(h, &sp, H.slice_ptr)); CKC(upload(h, &dg, H.deg)); CKC(upload(h, &cl, H.col));
        d.slice_ptr = sp; d.deg = dg; d.col = cl;
        CKC(dalloc(h, &d.val, (size_t)9 * H.n_slots));
        CKC(dalloc(h, &d.diag, (size_t)9 * H.n_pad));
        CKC(dalloc(h, &d.dinv, (size_t)9 * H.n_pad));
        CKC(dalloc(h, &d.pos, (size_t)2 * H.n_pad));
        if (!H.agg.empty()) {
            int32_t *ag; int64_t *ct; int32_t *cs;
            CKC(upload(h, &ag, H.agg)); CKC(upload(h, &ct, H.ctgt)); CKC(upload(h, &cs, H.cstr));
            d.agg = ag; d.ctgt = ct; d.cstr = cs;
        }
        if (!H.mem_ptr.empty()) {
            int64_t *mp; int32_t *mi;
            CKC(upload(h, &mp, H.mem_ptr)); CKC(upload(h, &mi, H.mem_idx));
            d.mem_ptr = mp; d.mem_idx = mi;
        }
        if (h->use_amg && nl > 1) {
            CKC(dalloc(h, &B.xa, (size_t)4 * H.n_pad)); CKC(dalloc(h, &B.res, (size_t)4 * H.n_pad));
            if (l > 0) { CKC(dalloc(h, &B.r, (size_t)4 * H.n_pad)); CKC(dalloc(h, &B.e, (size_t)4 * H.n_pad)); }
        }
    }
    max_grid = std::max<int64_t>(max_grid, grid_for(std::max<int64_t>(ne, 1), 256));
    h->dense_coarsest = h->use_amg && nl > 1 && S.levels[nl - 1].n * 3 <= 150;
    if (h->dense_coarsest) {
        const int m = (int)S.levels[nl - 1].n * 3;
        CKC(dalloc(h, &h->Ainv, (size_t)m * m));
        CKU(cudaFuncSetAttribute(k_dense_invert<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 150 * 150)));
    }
    h->chunk = (h->use_amg && nl > 1) ? 4 : 16;

    // ---- level-0 vertex data in storage order
    const int64_t n = S.n, n_pad = S.levels[0].n_pad;
    {
        std::vector<uint8_t> vk(n_pad, 0);
        std::vector<int64_t> rvo(n_pad, 0);
        for (int64_t r = 0; r < n; r++) { const int64_t v = S.perm[r]; vk[r] = S.vkind[v]; rvo[r] = S.vvalofs[v]; }
        uint8_t *dvk;
        CKC(upload(h, &dvk, vk));
        h->lv[0].d.vkind = dvk;
        CKC(upload(h, &h->row_valofs, rvo));
        CKC(dalloc(h, &h->poses, (size_t)4 * n_pad));
        CKC(dalloc(h, &h->vstage, (size_t)S.n_values));
        CKU(cudaMemcpyAsync(h->vstage, vval, S.n_values * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        k_import_poses<<<grid_for(n, 256), 256, 0, h->stream>>>(n, h->row_valofs, dvk, h->vstage, h->poses);
        CKU(cudaGetLastError());
    }
    h->anchor_row = S.anchor >= 0 ? S.iperm[S.anchor] : -1;
    // ---- measurements: per-edge packed offsets
    std::vector<int64_t> mofs(ne + 1, 0), iofs(ne + 1, 0);
    for (int64_t k = 0; k < ne; k++) { mofs[k + 1] = mofs[k] + (ekind[k] == 0 ? 3 : 2); iofs[k + 1] = iofs[k] + (ekind[k] == 0 ? 6 : 3); }
    auto edge_rec = [&](int64_t k, double *o) {   // z: x y cos sin ; Omega upper (6)
        const double *m = emeas + mofs[k], *w = einfo + iofs[k];
        for (int c = 0; c < 10; c++) o[c] = 0.0;
        o[0] = m[0]; o[1] = m[1];
        if (ekind[k] == 0) { o[2] = std::cos(m[2]); o[3] = std::sin(m[2]); for (int c = 0; c < 6; c++) o[4 + c] = w[c]; }
        else { for (int c = 0; c < 3; c++) o[4 + c] = w[c]; }
    };
    {   // half-edge stream, laid out like val with 10 components
        HostLevel &H = S.levels[0];
        std::vector<double> hz((size_t)10 * H.n_slots, 0.0);
        double rec[10];
        for (int64_t r = 0; r < n; r++) {
            const int lane = (int)(r & 31);
            for (int64_t qi = H.adj_ptr[r]; qi < H.adj_ptr[r + 1]; qi++) {
                const int64_t slot = H.adj_slot[qi];
                const int64_t cnt = H.adj_cnt[qi];
                edge_rec(S.slot_edge[slot], rec);
                double *dst = hz.data() + (slot - lane) * 10 + lane;
                for (int c = 0; c < 10; c++) dst[c * cnt] = rec[c];
            }
        }
        CKC(upload(h, &h->hz, hz));
    }
    {   // edge-ordered copy for chi2
        std::vector<int2> ends(std::max<int64_t>(ne, 1));
        std::vector<double> ed((size_t)10 * std::max<int64_t>(ne, 1), 0.0);
        double rec[10];
        for (int64_t k = 0; k < ne; k++) {
            const int a = S.iperm[S.efrom[k]], b = S.iperm[S.eto[k]];
            ends[k] = make_int2(a, ekind[k] == 1 ? ~b : b);
            edge_rec(k, rec);
            for (int c = 0; c < 10; c++) ed[(size_t)c * ne + k] = rec[c];
        }
        CKC(upload(h, &h->ends, ends));
        CKC(upload(h, &h->ed, ed));
    }
    // ---- solver vectors
    CKC(dalloc(h, &h->x, (size_t)4 * n_pad)); CKC(dalloc(h, &h->r, (size_t)4 * n_pad)); CKC(dalloc(h, &h->p, (size_t)4 * n_pad));
    CKC(dalloc(h, &h->q, (size_t)4 * n_pad)); CKC(dalloc(h, &h->z, (size_t)4 * n_pad));
    CKC(dalloc(h, &h->S, 1));
    CKC(dalloc(h, &h->partials, (size_t)max_grid + 8));
    CKU(cudaStreamSynchronize(h->stream));
#undef CKC
#undef CKU
    // host copies only needed for the structure queries stay in h->sym; drop the big transient ones
    S.levels[0].ctgt.clear(); S.levels[0].ctgt.shrink_to_fit();
    S.levels[0].cstr.clear(); S.levels[0].cstr.shrink_to_fit();
    *out = h;
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
    NEED_DEVICE(h);
    int rc = chi2_launch(h);
    if (rc) return rc;
    CK(cudaMemcpyAsync(&h->hS[0], h->S, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *chi2 = h->hS[0].chi2;
    return PGO_OK;
}

int pgo_gn_step(pgo_handle *h, double lambda, int add_lambda, double *norm_dx, double *chi2, int32_t *pcg_iterations) {
    if (!h) return PGO_ERR_ARG;
    NEED_DEVICE(h);
    int64_t l0 = h->launch_count;
    auto mark = [&](int i) { cudaEventRecord(h->ev[i], h->stream); };
    int64_t lc[6];
    mark(0); lc[0] = h->launch_count;
    int rc = assemble(h, lambda, add_lambda);
    if (rc) return rc;
    mark(1); lc[1] = h->launch_count;
    rc = amg_setup(h);
    if (rc) return rc;
    mark(2); lc[2] = h->launch_count;
    int32_t iters = 0;
    int src = solve(h, &iters);
    if (src != PGO_OK && src != PGO_ERR_NOT_CONVERGED) return src;
    mark(3); lc[3] = h->launch_count;
    rc = retract(h, 1.0);
    if (rc) return rc;
    mark(4); lc[4] = h->launch_count;
    rc = chi2_launch(h);
    if (rc) return rc;
    mark(5); lc[5] = h->launch_count;
    CK(cudaMemcpyAsync(&h->hS[0], h->S, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 5; i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
        h->ms[i] = ms; h->launches[i] = lc[i + 1] - lc[i];
    }
    (void)l0;
    h->have_step = true;
    if (norm_dx) *norm_dx = std::sqrt(h->hS[0].norm2_dx);
    if (chi2) *chi2 = h->hS[0].chi2;
    if (pcg_iterations) *pcg_iterations = iters;
    return src;
}

int pgo_undo_last_step(pgo_handle *h) {
    if (!h) return PGO_ERR_ARG;
    NEED_DEVICE(h);
    if (!h->have_step) { h->err = "pgo_undo_last_step: no step to undo"; return PGO_ERR_ARG; }
    int rc = retract(h, -1.0);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return PGO_OK;
}

int pgo_linearize_and_solve(pgo_handle *h, int32_t *pcg_iterations) {
    if (!h) return PGO_ERR_ARG;
    NEED_DEVICE(h);
    int rc = assemble(h, 0.0, 0);
    if (rc) return rc;
    rc = amg_setup(h);
    if (rc) return rc;
    return solve(h, pcg_iterations);
}

int pgo_get_poses(pgo_handle *h, double *out, int64_t n_values) {
    if (!h || !out) return PGO_ERR_ARG;
    NEED_DEVICE(h);
    const Symbolic &S = h->sym;
    if (n_values != S.n_values) { h->err = "pgo_get_poses: wrong buffer length"; return PGO_ERR_ARG; }
    k_export_poses<<<grid_for(S.n, 256), 256, 0, h->stream>>>(S.n, h->row_valofs, h->lv[0].d.vkind, h->poses, h->vstage);
    h->launch_count += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, h->vstage, S.n_values * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PGO_OK;
}

int pgo_set_poses(pgo_handle *h, const double *in, int64_t n_values) {
    if (!h || !in) return PGO_ERR_ARG;
    NEED_DEVICE(h);
    const Symbolic &S = h->sym;
    if (n_values != S.n_values) { h->err = "pgo_set_poses: wrong buffer length"; return PGO_ERR_ARG; }
    CK(cudaMemcpyAsync(h->vstage, in, S.n_values * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    k_import_poses<<<grid_for(S.n, 256), 256, 0, h->stream>>>(S.n, h->row_valofs, h->lv[0].d.vkind, h->vstage, h->poses);
    h->launch_count += 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));   // `in` is borrowed for the duration of the call only
    h->have_step = false;
    return PGO_OK;
}

int pgo_snapshot_poses(pgo_handle *h) {
    if (!h) return PGO_ERR_ARG;
    NEED_DEVICE(h);
    const size_t cnt = (size_t)4 * h->sym.levels[0].n_pad;
    if (!h->poses_saved) { int rc = dalloc(h, &h->poses_saved, cnt, false); if (rc) return rc; }
    CK(cudaMemcpyAsync(h->poses_saved, h->poses, cnt * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PGO_OK;
}

int pgo_restore_poses(pgo_handle *h) {
    if (!h) return PGO_ERR_ARG;
    NEED_DEVICE(h);
    if (!h->poses_saved) { h->err = "pgo_restore_poses: no snapshot"; return PGO_ERR_ARG; }
    CK(cudaMemcpyAsync(h->poses, h->poses_saved, (size_t)4 * h->sym.levels[0].n_pad * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->have_step = false;
    return PGO_OK;
}

int pgo_get_dx(pgo_handle *h, double *out, int64_t len) {
    if (!h || !out) return PGO_ERR_ARG;
    NEED_DEVICE(h);
    const Symbolic &S = h->sym;
    if (len != S.len) { h->err = "pgo_get_dx: wrong buffer length"; return PGO_ERR_ARG; }
    std::vector<double> xs((size_t)4 * S.n);
    CK(cudaMemcpyAsync(xs.data(), h->x, xs.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int64_t r = 0; r < S.n; r++) {
        const int64_t v = S.perm[r];
        const int d = S.vkind[v] == 0 ? 3 : 2;
        for (int c = 0; c < d; c++) out[S.voffset[v] + c] = xs[4 * r + c];
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

int pgo_get_system(pgo_handle *h, double lambda, int add_lambda, double *csc_values, double *b) {
    if (!h) return PGO_ERR_ARG;
    NEED_DEVICE(h);
    Symbolic &S = h->sym;
    if (!build_csc_pattern(S)) { h->err = S.error; return PGO_ERR_UNSUPPORTED; }
    int rc = assemble(h, lambda, add_lambda);
    if (rc) return rc;
    HostLevel &H = S.levels[0];
    std::vector<double> val((size_t)9 * H.n_slots), diag((size_t)9 * H.n_pad), rv((size_t)4 * H.n_pad);
    CK(cudaMemcpyAsync(val.data(), h->lv[0].d.val, val.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(diag.data(), h->lv[0].d.diag, diag.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(rv.data(), h->r, rv.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    // canonical block values (duplicate edges between a vertex pair sum into one block)
    std::vector<double> blk((size_t)9 * S.bcol.size(), 0.0);
    auto find = [&](int32_t rr, int32_t cc) -> int64_t {
        auto bb = S.bcol.begin() + S.brow_ptr[rr], ee = S.bcol.begin() + S.brow_ptr[rr + 1];
        return std::lower_bound(bb, ee, cc) - S.bcol.begin();
    };
    for (int64_t r = 0; r < S.n; r++) {
        const int32_t u = S.perm[r];
        const int lane = (int)(r & 31);
        double *d = &blk[9 * find(u, u)];
        for (int c = 0; c < 9; c++) d[c] += diag[(size_t)c * H.n_pad + r];
        for (int64_t qi = H.adj_ptr[r]; qi < H.adj_ptr[r + 1]; qi++) {
            const int64_t slot = H.adj_slot[qi], cnt = H.adj_cnt[qi];
            const int32_t v = S.perm[H.adj_nbr[qi]];
            double *o = &blk[9 * find(u, v)];
            const double *src = val.data() + (slot - lane) * 9 + lane;
            for (int c = 0; c < 9; c++) o[c] += src[c * cnt];
        }
    }
    if (csc_values) {
        int64_t o = 0;
        for (int64_t v = 0; v < S.n; v++) {
            const int dv = S.vkind[v] == 0 ? 3 : 2;
            for (int c = 0; c < dv; c++)
                for (int64_t pp = S.brow_ptr[v]; pp < S.brow_ptr[v + 1]; pp++) {
                    const int32_t u = S.bcol[pp];
                    const int du = S.vkind[u] == 0 ? 3 : 2;
                    const int64_t bi = find(u, (int32_t)v);                    // block (row u, col v)
                    for (int rr = 0; rr < du; rr++) csc_values[o++] = blk[9 * bi + 3 * rr + c];
                }
        }
    }
    if (b) {
        for (int64_t r = 0; r < S.n; r++) {
            const int64_t v = S.perm[r];
            const int d = S.vkind[v] == 0 ? 3 : 2;
            for (int c = 0; c < d; c++) b[S.voffset[v] + c] = rv[4 * r + c];
        }
    }
    return PGO_OK;
}

int pgo_get_timings(pgo_handle *h, double *ms, int64_t *launches, int32_t n) {
    if (!h) return PGO_ERR_ARG;
    for (int i = 0; i < n && i < PGO_NUM_PHASES; i++) {
        if (ms) ms[i] = h->ms[i];
        if (launches) launches[i] = h->launches[i];
    }
    return PGO_OK;
}

int pgo_time_spmv(pgo_handle *h, int32_t repeats, double *avg_ms) {
    if (!h || !avg_ms || repeats <= 0) return PGO_ERR_ARG;
    NEED_DEVICE(h);
    LevelBuf &B = h->lv[0];
    // p -> q with the PCG SpMV; done-flag test disabled so the launches always do the work
    for (int i = 0; i < 3; i++) k_spmv<3, 0, FIN_NONE><<<B.grid128, 128, 0, h->stream>>>(B.d, h->p, nullptr, h->q, 0.0, h->S, h->partials, 0);
    CK(cudaEventRecord(h->ev[PGO_NUM_PHASES], h->stream));
    for (int i = 0; i < repeats; i++) k_spmv<3, 0, FIN_NONE><<<B.grid128, 128, 0, h->stream>>>(B.d, h->p, nullptr, h->q, 0.0, h->S, h->partials, 0);
    CK(cudaEventRecord(h->ev[PGO_NUM_PHASES + 1], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->ev[PGO_NUM_PHASES], h->ev[PGO_NUM_PHASES + 1]));
    *avg_ms = ms / repeats;
    h->ms[PGO_PHASE_SPMV_FINE] = *avg_ms;
    h->launch_count += repeats + 3;
    return PGO_OK;
}

int pgo_get_stats(const pgo_handle *h, int64_t *rows, int64_t *offdiag, int64_t *levels, int64_t *bytes) {
    if (!h) return PGO_ERR_ARG;
    if (rows) *rows = h->sym.n;
    if (offdiag) *offdiag = 2 * h->sym.n_edges;
    if (levels) *levels = (int64_t)h->lv.size();
    if (bytes) *bytes = (int64_t)h->device_bytes;
    return PGO_OK;
}

} // extern "C"
