// Host-side g2o loader / writer: C++ mirror of RustRobotics src/mapping/g2o.rs (parse_g2o, :35-143).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace robotics { namespace mapping {

// The graph as flat arrays, in the packing the C ABI takes (include/pgo_b200.h: pgo_create).
// vertices are in lut order = VERTEX line order (g2o.rs:60,67,76); edges in file order.
struct G2oGraph {
    int64_t len = 0;                       // total scalar dimension (g2o.rs:142)
    std::vector<uint32_t> vertex_id;
    std::vector<uint8_t> vertex_kind;      // 0 SE2, 1 XY, 2 SE3
    std::vector<double> vertex_values;     // x y theta | x y | x y z qx qy qz qw
    std::vector<uint8_t> edge_kind;        // 0 EDGE_SE2, 1 EDGE_SE2_XY, 2 EDGE_SE3:QUAT
    std::vector<uint32_t> edge_from, edge_to;
    std::vector<double> edge_meas;         // 3 | 2 | 7 per edge
    std::vector<double> edge_info_upper;   // 6 | 3 | 21 per edge (row-major upper triangle, g2o.rs:88-93)
};

// parse_g2o (g2o.rs:35-143).  Returns false and fills `error` where the reference returns Err or
// panics: unreadable file, unknown tag (unimplemented!, :138), blank line (:53), bad id or number,
// wrong field count (the todo!() arms).
bool parse_g2o(const std::string &filename, G2oGraph &out, std::string &error);

// the packed value arrays hold exactly what the per-kind counts say, kinds are 0..2 (what pgo_create reads on the device side)
bool validate_graph(const G2oGraph &g, std::string &error);
// the same check on bare arrays (lengths of the value arrays against what the kinds need)
bool validate_graph_arrays(size_t n_vertices, const uint8_t *vertex_kind, size_t n_values, size_t n_edges, const uint8_t *edge_kind,
                           size_t n_meas, size_t n_info, std::string &error);

// writes the graph back in g2o text form with round-trip precision (%.17g)
bool write_g2o(const std::string &filename, const G2oGraph &g, std::string &error);

}} // namespace robotics::mapping
