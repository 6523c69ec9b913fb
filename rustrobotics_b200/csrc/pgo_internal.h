// Internal data structures shared by the host symbolic pass and the CUDA solver.
// Not part of the ABI (include/pgo_b200.h is).
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

namespace pgo {

constexpr int MAX_RANKS = 8;
// column word of a stored block: [31] own row is the edge's `to` vertex, [30] pose-landmark edge (level 0 only),
// [29:27] owner rank of the neighbour row, [26:0] the neighbour's row index local to its owner.
constexpr uint32_t COL_ROLE_TO = 0x80000000u;
constexpr uint32_t COL_EDGE_XY = 0x40000000u;
constexpr int COL_OWNER_SHIFT = 27;
constexpr uint32_t COL_LOCAL_MASK = (1u << COL_OWNER_SHIFT) - 1u;
constexpr uint32_t COL_FLAG_MASK = COL_ROLE_TO | COL_EDGE_XY;

// One level of the block-sparse hierarchy, in GLOBAL PADDED row numbering: the rows of partition (rank) k are
// [part_off[k], part_off[k+1]) with part_real[k] real rows first and padding rows (deg 0) behind them, so that every
// partition starts on a 32-row slice boundary.  A rank's device arrays are contiguous sub-ranges of these.
//
// Level 0 is stored as sliced jagged storage ("slice-contiguous JDS"): rows in slices of 32 (one warp, one thread per
// block row); inside a slice the rows are ordered by decreasing off-diagonal count, so in "column" k of the slice the
// active rows are lanes 0..cnt_k-1 and their entries are stored densely:  slot(k, lane) = slice_ptr[s] + off_k + lane,
// off_k = sum_{k'<k} cnt_k'.  Values of a column are component-major, val[(slice_ptr[s] + off_k) * DD + c * cnt_k + lane],
// so a whole slice is ONE contiguous blob that a warp streams front to back with coalesced loads.
// Coarse levels (small, L2-resident, latency-bound) are plain block CSR, one warp per row: slot = adj_ptr[row] + k,
// val[slot * DD + c].
struct HostLevel {
    bool jds = false;
    int64_t n = 0;                    // real rows (all partitions)
    int64_t n_pad = 0;                // global padded rows
    int64_t n_slots = 0;
    std::vector<int64_t> part_off;    // [world + 1], multiples of 32
    std::vector<int64_t> part_real;   // [world]
    std::vector<int64_t> part_slot;   // [world + 1] slot range of each partition
    std::vector<uint8_t> real;        // [n_pad] 1 for real rows
    // sharded handles: a small coarse level is REPLICATED (every rank holds all of it: part_off = {0, n_pad}); on the
    // first replicated level src_off[k] .. src_off[k+1] are the rows whose members (finer rows) live on rank k
    bool repl = false;
    std::vector<int64_t> src_off;     // [world + 1], first replicated level only
    // JDS only
    int64_t n_slices = 0;
    std::vector<int64_t> slice_ptr;   // [n_slices + 1]
    std::vector<int32_t> deg;         // [n_pad]
    // adjacency in storage order (both formats); for CSR levels slot == position
    std::vector<int64_t> adj_ptr;     // [n_pad + 1]
    std::vector<int64_t> adj_slot;    // slot index of each entry
    std::vector<int32_t> adj_cnt;     // JDS: cnt_k (component stride) of each entry; CSR: 1
    std::vector<int32_t> adj_nbr;     // neighbour global padded row
    std::vector<uint32_t> adj_flags;  // level 0: COL_ROLE_TO / COL_EDGE_XY
    // aggregation towards the next coarser level (empty on the coarsest)
    std::vector<int32_t> agg;         // [n_pad] coarse global padded row of each row (-1 for padding rows)
    std::vector<int32_t> ctgt;        // [n_slots] Galerkin target of each stored block, LOCAL to the owning partition:
                                      //   >= 0: element offset of component 0 in the coarse val array ; < 0: diagonal of coarse row (-1 - row)
    std::vector<int32_t> cstr;        // [n_slots] stride between the 9 components of that target (1: CSR ; cnt: sliced storage)
    // deterministic Galerkin product (coarse level stored as block CSR): per partition k, the contributors of every coarse block in a
    // FIXED order.  Coarse block ids are local to the coarse partition: [0, rows) = diagonal blocks, rows + s = stored block s;
    // a contributor is an index into the fine partition's staging buffer: [0, slots) = stored block, slots + r = diagonal block of row r.
    std::vector<std::vector<int32_t>> gal_ptr, gal_src;
    // this level seen as the coarse side of the finer level: members (finer global padded rows) of each row
    std::vector<int64_t> mem_ptr;     // [n_pad + 1]
    std::vector<int32_t> mem_idx;

    int part_of(int64_t row) const { int k = 0; while (row >= part_off[k + 1]) k++; return k; }
};

// Output of the one-time symbolic pass over the graph.
struct Symbolic {
    int D = 3;                         // block dimension (3: SE2/XY, 6: SE3)
    int world = 1;
    int64_t n = 0, n_edges = 0, len = 0, n_values = 0;
    std::vector<uint8_t> vkind;        // lut order
    std::vector<int64_t> voffset;      // lut: scalar offset of each vertex (g2o.rs:60,67,76)
    std::vector<int64_t> vvalofs;      // offset into the packed vertex_values array
    std::vector<int32_t> efrom, eto;   // edge endpoints as vertex indices (lut order)
    std::vector<uint8_t> ekind;
    int64_t anchor = -1;               // lut index of the anchored vertex (:330-336)
    std::vector<int64_t> vrange;       // [world + 1] contiguous vertex (lut) range of each rank
    // canonical block CSR (lut order) + edge -> block slots + scalar CSC pattern
    std::vector<int64_t> brow_ptr; std::vector<int32_t> bcol; std::vector<int64_t> edge_slots;
    std::vector<int32_t> csc_ptr, csc_row;
    // internal ordering of level 0
    std::vector<int32_t> perm;         // [levels[0].n_pad] global padded row -> lut index (-1 for padding rows)
    std::vector<int32_t> iperm;        // [n] lut index -> global padded row
    std::vector<HostLevel> levels;     // levels[0] = the Gauss-Newton system
    std::vector<int32_t> slot_edge;    // per stored slot of level 0: the edge it comes from
    bool dense_coarsest = false;       // the last level is solved directly (explicit inverse)
    std::string error;
};

struct SymbolicOptions {
    int world = 1;
    int sort_window = 2048;
    int max_levels = 12;
    int agg_size = 16;
    int dense_max = 640;               // a level with at most this many rows is solved directly
    int64_t repl_max_rows = 131072;    // world > 1: coarse levels up to this many rows are replicated on every rank (no cross-rank sync inside them)
    int64_t jds_min_rows = INT64_MAX;     // coarse levels at least this large use the sliced storage (thread per row), smaller ones block CSR
    bool build_amg = true;
};

// Builds everything above. Returns false and sets sym.error on malformed input.
bool build_symbolic(Symbolic &sym, const SymbolicOptions &opt,
                    int64_t n_vertices, const uint32_t *vertex_id, const uint8_t *vertex_kind,
                    int64_t n_edges, const uint8_t *edge_kind, const uint32_t *edge_from_id,
                    const uint32_t *edge_to_id);

// Host-side set-up is seconds of index manipulation at 1M poses next to Gauss-Newton steps of 40 ms: loops whose iterations are
// independent run on the host's cores.  f(begin, end) is called on disjoint contiguous ranges; results do not depend on the
// thread count (PGO_HOST_THREADS overrides it).
template <typename F> inline void parallel_for(int64_t n, int64_t grain, F f) {
    unsigned nt = std::thread::hardware_concurrency();
    if (const char *e = std::getenv("PGO_HOST_THREADS")) nt = (unsigned)std::max(1, std::atoi(e));
    nt = std::min<unsigned>(std::max(1u, nt), 16u);
    nt = (unsigned)std::min<int64_t>(nt, std::max<int64_t>(1, n / std::max<int64_t>(grain, 1)));
    if (nt <= 1) { f((int64_t)0, n); return; }
    std::vector<std::thread> th;
    th.reserve(nt);
    for (unsigned t = 0; t < nt; t++) {
        const int64_t b = n * t / nt, e = n * (t + 1) / nt;
        th.emplace_back([=]() { f(b, e); });
    }
    for (auto &x : th) x.join();
}

// PGO_SYM_TIMING=1: wall time of the phases of pgo_create on stderr (needs a `double t_last = now_s();` in scope)
inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define TICK(name) do { if (std::getenv("PGO_SYM_TIMING")) { double t_ = pgo::now_s(); std::fprintf(stderr, "[sym] %-28s %.3f s\n", name, t_ - t_last); t_last = t_; } } while (0)

// lazily computed (only the structure checks and pgo_get_system need them)
bool build_canonical(Symbolic &sym);      // brow_ptr / bcol / edge_slots
bool build_csc_pattern(Symbolic &sym);    // csc_ptr / csc_row

} // namespace pgo
