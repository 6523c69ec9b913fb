"""g2o loader / writer: host-side mirror of reference src/mapping/g2o.rs (parse_g2o, :35-143), implemented in
C++ (csrc/host/g2o.cpp) and bound here."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._lib import lib, ptr

KEYS = ("vertex_id", "vertex_kind", "vertex_values", "edge_kind", "edge_from", "edge_to", "edge_meas", "edge_info_upper")


def parse_g2o(filename):
    """-> (len, graph) where graph is a dict of the flat arrays `pgo_create` takes.

    Mirrors `parse_g2o(filename) -> (len, edges, lut, nodes)` (g2o.rs:35-46): vertices come back in lut
    (VERTEX line) order, edges in file order; malformed input raises ValueError where the reference returns
    Err or panics."""
    L = lib()
    h = L.pg_parse_g2o(os.fsencode(str(filename)))
    if not h:
        raise ValueError("parse_g2o: " + L.pg_last_error().decode())
    try:
        s = [C.c_int64() for _ in range(6)]
        L.pg_graph_sizes(h, *[C.byref(x) for x in s])
        nv, ne, ln, nval, nmeas, ninfo = [x.value for x in s]
        g = dict(vertex_id=np.empty(nv, np.uint32), vertex_kind=np.empty(nv, np.uint8), vertex_values=np.empty(nval),
                 edge_kind=np.empty(ne, np.uint8), edge_from=np.empty(ne, np.uint32), edge_to=np.empty(ne, np.uint32),
                 edge_meas=np.empty(nmeas), edge_info_upper=np.empty(ninfo))
        L.pg_graph_fill(h, *[ptr(g[k]) for k in KEYS])
    finally:
        L.pg_graph_free(h)
    return ln, g


def write_g2o(filename, graph):
    """Write a graph dict as g2o text with round-trip (%.17g) precision."""
    L = lib()
    a = {k: np.ascontiguousarray(graph[k], dt) for k, dt in zip(KEYS, (np.uint32, np.uint8, np.float64, np.uint8, np.uint32,
                                                                       np.uint32, np.float64, np.float64))}
    rc = L.pg_write_g2o(os.fsencode(str(filename)), len(a["vertex_id"]), ptr(a["vertex_id"]), ptr(a["vertex_kind"]),
                        ptr(a["vertex_values"]), len(a["vertex_values"]), len(a["edge_kind"]), ptr(a["edge_kind"]),
                        ptr(a["edge_from"]), ptr(a["edge_to"]), ptr(a["edge_meas"]), len(a["edge_meas"]),
                        ptr(a["edge_info_upper"]), len(a["edge_info_upper"]))
    if rc != 0:
        msg = L.pg_last_error().decode()
        raise (ValueError if msg.startswith("graph arrays") else OSError)("write_g2o: " + msg)
