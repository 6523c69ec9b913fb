// Internal data structures shared by the host symbolic pass and the CUDA solver.
// Not part of the ABI (include/pgo_b200.h is).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pgo {

constexpr uint32_t COL_MASK = 0x3FFFFFFFu;   // neighbour row index
constexpr uint32_t COL_ROLE_TO = 0x80000000u; // own row is the edge's `to` vertex
constexpr uint32_t COL_EDGE_XY = 0x40000000u; // pose-landmark edge (EDGE_SE2_XY)

// Sliced jagged storage ("slice-contiguous JDS") of one block-sparse level.
// Rows are grouped in slices of 32 (one warp, one thread per block row); inside a slice the rows
// are ordered by decreasing off-diagonal count, so in "column" k of the slice the active rows are
// lanes 0..cnt_k-1 and their entries are stored densely:  slot(k, lane) = slice_ptr[s] + off_k + lane,
// off_k = sum_{k'<k} cnt_k'.  Values of a column are component-major,
//   val[(slice_ptr[s] + off_k) * DD + c * cnt_k + lane],
// so a whole slice is ONE contiguous blob that a warp streams front to back with coalesced loads.
struct HostLevel {
    int64_t n = 0, n_pad = 0, n_slices = 0, n_slots = 0;
    std::vector<int64_t> slice_ptr;   // [n_slices + 1]
    std::vector<int32_t> deg;         // [n_pad]
    std::vector<uint32_t> col;        // [n_slots]  neighbour row | flags (level 0 only)
    // adjacency in storage order, for host-side bookkeeping
    std::vector<int64_t> adj_ptr;     // [n + 1]
    std::vector<int64_t> adj_slot;    // slot index of each entry
    std::vector<int32_t> adj_cnt;     // cnt_k (component stride) of each entry
    std::vector<int32_t> adj_nbr;     // neighbour row
    // aggregation towards the next coarser level (empty on the coarsest)
    std::vector<int32_t> agg;         // [n_pad] coarse row of each row (-1 for padding rows)
    std::vector<int64_t> ctgt;        // [n_slots] Galerkin target: component-0 offset in coarse val, or
                                      //           (coarse row | DIAG_FLAG) when both ends share the aggregate
    std::vector<int32_t> cstr;        // [n_slots] component stride at the target (cnt_k of the coarse column)
    // this level seen as the coarse side of the finer level: members of each row
    std::vector<int64_t> mem_ptr;     // [n + 1]
    std::vector<int32_t> mem_idx;     // finer-level rows
};
constexpr int64_t CTGT_DIAG = int64_t(1) << 62;

// Output of the one-time symbolic pass over the graph.
struct Symbolic {
    int D = 3;                         // block dimension (3: SE2/XY, 6: SE3)
    int64_t n = 0, n_edges = 0, len = 0, n_values = 0;
    std::vector<uint8_t> vkind;        // lut order
    std::vector<int64_t> voffset;      // lut: scalar offset of each vertex (g2o.rs:60,67,76)
    std::vector<int64_t> vvalofs;      // offset into the packed vertex_values array
    std::vector<int32_t> efrom, eto;   // edge endpoints as vertex indices (lut order)
    std::vector<uint8_t> ekind;
    int64_t anchor = -1;               // lut index of the anchored vertex (:330-336)
    // canonical block CSR (lut order) + edge -> block slots + scalar CSC pattern
    std::vector<int64_t> brow_ptr; std::vector<int32_t> bcol; std::vector<int64_t> edge_slots;
    std::vector<int32_t> csc_ptr, csc_row;
    // internal ordering
    std::vector<int32_t> perm, iperm;  // perm[internal row] = lut index ; iperm = inverse
    std::vector<HostLevel> levels;     // levels[0] = the Gauss-Newton system
    // per stored slot of level 0: the edge it comes from (measurement scatter at create time)
    std::vector<int32_t> slot_edge;
    std::string error;
};

struct SymbolicOptions { int sort_window = 4096; int amg_max_levels = 12; int coarsest_max = 64; bool build_amg = true; };

// Builds everything above. Returns false and sets sym.error on malformed input.
bool build_symbolic(Symbolic &sym, const SymbolicOptions &opt,
                    int64_t n_vertices, const uint32_t *vertex_id, const uint8_t *vertex_kind,
                    int64_t n_edges, const uint8_t *edge_kind, const uint32_t *edge_from_id,
                    const uint32_t *edge_to_id);

// lazily computed (only the structure checks and pgo_get_system need them)
bool build_canonical(Symbolic &sym);      // brow_ptr / bcol / edge_slots
bool build_csc_pattern(Symbolic &sym);    // csc_ptr / csc_row

} // namespace pgo
