/*
 * pgo_b200.h -- C ABI of the B200-native pose-graph-optimization hot path.
 *
 * Drop-in boundary for jgsimard/RustRobotics src/mapping/pose_graph_optimization.rs.
 * The reference has no FFI of its own (pure Rust calling russell_sparse); the seam this ABI
 * replaces is the body of the Gauss-Newton loop,
 *     let dx = self.build_linear_system(lambda)?.solve()?;     (pose_graph_optimization.rs:271)
 *     self.update_nodes(&dx);  dx.norm();  global_error(self)   (:272-274)
 * i.e. one call maps (poses, edges, lambda) -> (poses', |dx|, chi2) with all state resident in
 * HBM.  A Rust `extern "C"` block binds exactly these symbols (INTEGRATION.md shows it).
 *
 * Conventions: plain pointers and sizes only; every function returns a status code of enum pgo_status, 0 = ok, and
 * never throws or aborts; pgo_last_error gives the message of the last failure on the handle
 * (or of the last failed pgo_create when handle == NULL).  A handle is single-owner and not
 * thread-safe (PoseGraph::optimize takes &mut self, :247): one host thread at a time, every call
 * blocks until its result is there.  Input arrays are borrowed for the duration of the call only.
 * Multi-GPU: set pgo_options.n_gpus (and optionally device_ids) and the SAME single handle drives
 * all GPUs of the box from the one calling thread's point of view (one worker thread per GPU
 * lives inside the library) -- the reference's `PoseGraph::new(path, solver)?.optimize(n, ..)`
 * needs no other change.  There is no CPU fallback: without a CUDA device every entry
 * point that computes returns PGO_ERR_CUDA.
 */
#ifndef PGO_B200_H
#define PGO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgo_handle pgo_handle;

typedef enum {
    PGO_OK = 0,
    PGO_ERR_ARG = 1,        /* bad argument / malformed graph (unknown vertex id, kind mismatch) */
    PGO_ERR_CUDA = 2,       /* CUDA runtime failure, or no device */
    PGO_ERR_COMM = 3,       /* multi-GPU handles: peer access unavailable, or a cross-GPU synchronisation timed out */
    PGO_ERR_SOLVER = 4,     /* PCG breakdown (p^T H p <= 0 or NaN): H not positive definite */
    PGO_ERR_NOT_CONVERGED = 5, /* PCG hit pcg_max_iterations; dx of that step was still applied */
    PGO_ERR_UNSUPPORTED = 6
} pgo_status;

/* vertex kinds = Node variants (pose_graph_optimization.rs:149-154); values per vertex as on a
 * g2o VERTEX line (g2o.rs:54-78): SE2 = x y theta ; XY = x y ; SE3 = x y z qx qy qz qw */
enum { PGO_VERTEX_SE2 = 0, PGO_VERTEX_XY = 1, PGO_VERTEX_SE3 = 2 };
/* edge kinds = Edge variants (:21-26); measurement as on a g2o EDGE line (g2o.rs:79-137):
 * SE2 = dx dy dtheta + 6 upper-triangular information entries (row-major),
 * SE2_XY = x y + 3, SE3 = x y z qx qy qz qw + 21 */
enum { PGO_EDGE_SE2 = 0, PGO_EDGE_SE2_XY = 1, PGO_EDGE_SE3 = 2 };

enum { PGO_PRECOND_BLOCK_JACOBI = 0, PGO_PRECOND_AMG = 1 };

typedef struct {
    double anchor_weight;      /* 1e7: added to the diagonal of the `from` vertex of the first pose-pose edge (:330-336) */
    double pcg_rtol;           /* stop when sqrt(r^T M^-1 r) <= pcg_rtol * sqrt(b^T M^-1 b) */
    int32_t pcg_max_iterations;
    int32_t preconditioner;    /* PGO_PRECOND_* */
    int32_t sort_window;       /* rows per degree-sorting window of the sliced storage (multiple of 32; 0 = default) */
    int32_t amg_max_levels;    /* 0 = default */
    int32_t device;            /* CUDA device ordinal, -1 = current; -2 = structure-only handle (symbolic-pass queries, no device) */
    int32_t world;             /* number of ranks (one process per GPU) sharing the graph; 1 = single GPU */
    int32_t rank;              /* this process' rank, 0 .. world-1: it owns a contiguous vertex range */
    int32_t amg_dense_max;     /* a level with at most this many block rows is solved directly (explicit inverse); 0 = default */
    int32_t amg_aggregate_size;/* upper bound on the members of an aggregate; 0 = default (16 on one GPU, 24 for sharded handles) */
    int32_t amg_kcycle;        /* coarse levels 1..amg_kcycle are solved by a K-cycle (two Krylov-accelerated cycles per
                                * visit, Notay), deeper ones by a V-cycle; 0 = plain V-cycle; default: all levels */
    int32_t amg_kcycle3;       /* coarse levels 1..amg_kcycle3 run THREE inner flexible-CG steps per K-cycle visit instead of two
                                * (a more accurate coarse solve: fewer PCG iterations, 1.5x the coarse work); 0 = none;
                                * -1 (default) = level 1 for SE2 / XY graphs, none for SE3 graphs */
    int32_t amg_fp64_storage;  /* 0 (default): the SpMVs INSIDE the multigrid cycle (smoother, residual, K-cycle products) read
                                * fp32 copies of the stored blocks of every level and accumulate in fp64 -- the preconditioner
                                * is an approximate operator anyway, and half the bytes is half the time of an HBM-bound
                                * kernel; the PCG operator H p, the residual recurrence, all dot products and the result stay
                                * fp64, so the solve converges to the same pcg_rtol.  1: everything reads the fp64 blocks */
    int32_t n_gpus;            /* 0 / 1 (default): one GPU (`device`).  2 .. 8: SINGLE-PROCESS multi-GPU handle -- the graph's block rows
                                * are split into n_gpus contiguous vertex ranges (lut order), one shard per GPU, and this one handle
                                * drives all of them; every entry point keeps its single-GPU meaning (world / rank must stay 1 / 0) */
    const int32_t *device_ids; /* n_gpus CUDA device ordinals, NULL = 0 .. n_gpus-1.  The devices need peer access to each other
                                * (NVLink / NVSwitch on a B200 box).  An ordinal may appear more than once: those shards then share
                                * that GPU (how the multi-GPU path is tested on a one-GPU machine).  Borrowed during pgo_create only */
    int32_t refine;            /* 0 (default): dx = the PCG solution.  1: one round of iterative refinement on top of it, with the residual
                                * b - H dx evaluated in double-double arithmetic and the correction solved to refine_rtol.  On graphs of
                                * ~1M poses the first Gauss-Newton step cannot be pinned below ~1e-5 m by ANY plain-fp64 solver (direct
                                * ones included); this mode reaches the 1e-6 m of small graphs there, for ~40 % more time */
    double refine_rtol;        /* relative tolerance of the correction solve; 0 = default 2e-4 */
} pgo_options;

/* pgo_options grows at its END between versions of the library (n_gpus / device_ids and refine / refine_rtol are round-2 additions):
 * a caller must be compiled against the header that ships with the library it loads (pgo_default_options writes sizeof(pgo_options)
 * bytes); pgo_version() lets a binding check. */
/* fills *opt with the defaults (anchor 1e7, rtol 1e-10, max 200000 iterations, AMG K-cycle, fp32 preconditioner storage, single GPU) */
void pgo_default_options(pgo_options *opt);

/*
 * Build the device-resident problem: uploads poses/edges, runs the one-time symbolic pass
 * (block structure of H, edge -> block slot map, sliced storage, AMG hierarchy structure).
 * Vertices are given in lut order = VERTEX line order (g2o.rs:60,67,76); edges in file order;
 * edge endpoints are g2o ids resolved through vertex_id like the reference's lut (:312-313).
 * vertex_values / edge_measurement / edge_information_upper are packed back to back with the
 * per-kind counts above: the caller guarantees sum(values per vertex kind), sum(measurement values
 * per edge kind) and sum(information values per edge kind) doubles respectively (the host mirrors
 * check this before calling; kinds outside 0..2 are rejected here with PGO_ERR_ARG).
 * SE2/XY graphs and SE3 graphs cannot be mixed in one handle.
 */
int pgo_create(pgo_handle **out, const pgo_options *opt,
               int64_t n_vertices, const uint32_t *vertex_id, const uint8_t *vertex_kind,
               const double *vertex_values,
               int64_t n_edges, const uint8_t *edge_kind, const uint32_t *edge_from_id,
               const uint32_t *edge_to_id, const double *edge_measurement,
               const double *edge_information_upper);
void pgo_destroy(pgo_handle *h);
const char *pgo_last_error(const pgo_handle *h);

/* --- process-per-GPU sharding (world > 1; an alternative to n_gpus for launchers that already run one process per GPU,
 * e.g. `torchrun bench.py --gpus N`): every rank passes the WHOLE graph to pgo_create and keeps the
 * block rows of its contiguous vertex range.  Neighbour rows on other ranks are read straight from peer HBM over
 * NVLink, so the ranks must exchange one CUDA IPC handle each before the first computing call: every rank exports
 * pgo_shard_handle_bytes() bytes, the caller all-gathers them in rank order (e.g. torch.distributed.all_gather) and
 * hands the concatenation to pgo_shard_connect.  All computing entry points are then COLLECTIVE: every rank must
 * make the same calls in the same order.  Scalars (chi2, |dx|, iterations) come back identical on every rank;
 * pgo_get_poses / pgo_get_dx / pgo_get_system fill only the part the rank owns (a contiguous span). */
int pgo_shard_handle_bytes(void);
int pgo_shard_export(pgo_handle *h, void *buf, int64_t capacity);
int pgo_shard_connect(pgo_handle *h, const void *all_handles, int64_t n_handles);
/* partition: vertex_range[world + 1] (lut order), n_remote_blocks[world] = off-diagonal blocks of H whose column
 * vertex lives on another rank (the halo the SpMV reads over NVLink).  Any argument may be NULL. */
int pgo_get_partition(const pgo_handle *h, int32_t *world, int32_t *rank, int64_t *vertex_range, int64_t *n_remote_blocks);

/* sizes: len = total scalar dimension returned by parse_g2o (g2o.rs:142) */
int pgo_get_sizes(const pgo_handle *h, int64_t *n_vertices, int64_t *n_edges, int64_t *len,
                  int64_t *n_vertex_values);

/* global_error (:537-574): sum over edges of e^T Omega e at the current poses */
int pgo_chi2(pgo_handle *h, double *chi2);

/*
 * One Gauss-Newton / Levenberg-Marquardt step (:271-274): linearise + assemble H dx = -b
 * (:305-369; lambda is added to every diagonal iff add_lambda != 0, :362-366), solve by PCG,
 * retract (:229-245), chi2 of the new poses.  norm_dx = ||dx||_2 (:273).
 * Device-resident; only the three scalars cross the bus.
 */
int pgo_gn_step(pgo_handle *h, double lambda, int add_lambda,
                double *norm_dx, double *chi2, int32_t *pcg_iterations);

/* LM rejection, update_nodes(&(-dx)) (:277): retract the last step's dx with the opposite sign */
int pgo_undo_last_step(pgo_handle *h);

/* poses in the layout of pgo_create's vertex_values (theta = atan2(im, re) of the stored unit complex) */
int pgo_get_poses(pgo_handle *h, double *vertex_values_out, int64_t n_values);
int pgo_set_poses(pgo_handle *h, const double *vertex_values, int64_t n_values);
/* device-side checkpoint of the poses (the reference's "call optimize again from the current nodes" state, :157-160):
 * snapshot copies the current poses to a second HBM buffer, restore copies them back.  No host traffic. */
int pgo_snapshot_poses(pgo_handle *h);
int pgo_restore_poses(pgo_handle *h);
/* the dx of the last pgo_gn_step / pgo_linearize_and_solve, length len, lut (scalar-offset) order */
int pgo_get_dx(pgo_handle *h, double *dx_out, int64_t len);

/* linearize_and_solve (:371-373): build with lambda = 0 and solve, without updating the poses */
int pgo_linearize_and_solve(pgo_handle *h, int32_t *pcg_iterations);

/* --- symbolic-pass outputs, for the bit-exact structure checks ------------------------------ */
/* scalar CSC pattern of H (sorted rows per column, duplicates merged) = what the reference's
 * COO (36 / 25 puts per edge, zeros included, :194-206) becomes inside russell_sparse/UMFPACK.
 * Call with col_ptr == NULL to query n and nnz. */
int pgo_get_pattern(const pgo_handle *h, int64_t *n, int64_t *nnz, int32_t *col_ptr, int32_t *row_idx);
/* block CSR of H over vertices in lut order (sorted unique column vertices, diagonal included)
 * and the edge -> block slot map: for edge k, slots[4k..4k+3] = positions in block_col of the
 * blocks (from,from), (from,to), (to,from), (to,to) that update_linear_system (:184-187) writes.
 * Call with block_row_ptr == NULL to query n_blocks. */
int pgo_get_block_structure(const pgo_handle *h, int64_t *n_blocks, int64_t *block_row_ptr,
                            int32_t *block_col, int64_t *edge_slots);
/* index (in lut order) of the anchored vertex, -1 if the graph has no pose-pose edge (:330) */
int pgo_get_anchor(const pgo_handle *h, int64_t *vertex_index);

/* --- debug / parity ------------------------------------------------------------------------- */
/* assemble at the current poses and return the CSC values (in pgo_get_pattern order) and b */
int pgo_get_system(pgo_handle *h, double lambda, int add_lambda, double *csc_values, double *b);

/* --- measurement ---------------------------------------------------------------------------- */
enum { PGO_PHASE_ASSEMBLE = 0, PGO_PHASE_PRECOND_SETUP = 1, PGO_PHASE_PCG = 2, PGO_PHASE_RETRACT = 3,
       PGO_PHASE_CHI2 = 4, PGO_PHASE_SPMV_FINE = 5, PGO_NUM_PHASES = 6 };
/* per-phase CUDA-event time (ms) and kernel-launch count of the last pgo_gn_step;
 * PGO_PHASE_SPMV_FINE is filled only when timing of the dominant kernel is enabled */
int pgo_get_timings(pgo_handle *h, double *ms_per_phase, int64_t *launches_per_phase, int32_t n_phases);
/* time `repeats` back-to-back launches of the fine-level BSR SpMV (the dominant kernel) on the
 * handle's stream with CUDA events; returns the average ms per launch */
int pgo_time_spmv(pgo_handle *h, int32_t repeats, double *avg_ms);
/* diagnostic: average time (ms) of ONE coarse solve of AMG level `level` >= 1 (the whole K-cycle subtree below it), launched as
 * a CUDA graph like inside the PCG loop; needs a previous pgo_gn_step / pgo_linearize_and_solve */
int pgo_time_coarse(pgo_handle *h, int32_t level, int32_t repeats, double *avg_ms);
/* structure statistics for the roofline accounting: block rows, off-diagonal blocks, stored slots */
int pgo_get_stats(const pgo_handle *h, int64_t *n_block_rows, int64_t *n_offdiag_blocks,
                  int64_t *n_levels, int64_t *device_bytes);

/* rows and stored off-diagonal blocks of every level of the hierarchy; returns the number of levels */
int pgo_get_level_sizes(const pgo_handle *h, int32_t max_levels, int64_t *rows, int64_t *blocks);

/* library / device identification, e.g. "pgo_b200 0.1 sm_100a" */
/* diagnostic (structure-only handles included): the aggregate = row of level `level + 1` (global padded numbering) of every
 * vertex in lut order (level 0, n = n_vertices) or of every padded row of `level` (level >= 1, n = padded rows, -1 = padding) */
int pgo_get_aggregates(const pgo_handle *h, int32_t level, int32_t *coarse_row, int64_t n);

const char *pgo_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PGO_B200_H */
