"""CPU ORACLE driver (test infrastructure, NOT product code).

Drives oracle/pgo_oracle.c (a plain-C restatement of RustRobotics'
src/mapping/g2o.rs and src/mapping/pose_graph_optimization.rs) and supplies the one
piece that lives in a third-party dependency of the reference: the sparse direct solve
(russell_sparse 0.7.1 -> SuiteSparse UMFPACK, pose_graph_optimization.rs:130-141,
Cargo.lock:1582-1585; not vendored, not installable here).  SciPy's SuperLU -- an exact
sparse LU like UMFPACK -- stands in for it.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this module.  Parity is pinned by tests/test_oracle_kat.py against every
known-answer value in the reference's own tests (g2o.rs:149-175,
pose_graph_optimization.rs:580-739).  SE(3): repo-defined semantics, parity unpinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

SE2, XY, SE3 = 0, 1, 2
KIND_NVAL = (3, 2, 7)
KIND_DIM = (3, 2, 6)


def build(force: bool = False) -> Path:
    so = _HERE / "libpgo_oracle.so"
    src = _HERE / "pgo_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-B", "libpgo_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        P = C.c_void_p
        L.og_parse_g2o.restype = P
        L.og_parse_g2o.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.og_create.restype = P
        L.og_create.argtypes = [C.c_int64, P, P, P, C.c_int64, P, P, P, P, P]
        L.og_free.argtypes = [P]
        for f in ("og_num_vertices", "og_num_edges", "og_len", "og_num_vertex_values"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [P]
        L.og_get_vertices.argtypes = [P, P, P, P, P]
        L.og_get_state.argtypes = [P, P]
        L.og_set_state.argtypes = [P, P]
        L.og_get_edges.argtypes = [P, P, P, P]
        L.og_get_edge_data.argtypes = [P, P, P, P, P]
        L.og_edge_linearize.argtypes = [P, C.c_int64, P, P, P]
        L.og_global_error.restype = C.c_double
        L.og_global_error.argtypes = [P]
        L.og_put_count.restype = C.c_int64
        L.og_put_count.argtypes = [P, C.c_int]
        L.og_build_linear_system.restype = C.c_int64
        L.og_build_linear_system.argtypes = [P, C.c_double, C.c_int, P, P, P, P]
        L.og_coo_to_csc.restype = C.c_int64
        L.og_coo_to_csc.argtypes = [C.c_int64, C.c_int64, P, P, P, P, P, P]
        L.og_update_nodes.argtypes = [P, P, C.c_double]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class LinearSystem:
    """H (CSC, duplicates summed, rows sorted) and b, as the reference hands them to UMFPACK."""

    def __init__(self, n, col_ptr, row_idx, vals, b, puts):
        self.n, self.col_ptr, self.row_idx, self.vals, self.b, self.puts = n, col_ptr, row_idx, vals, b, puts

    def csc(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.vals, self.row_idx, self.col_ptr), shape=(self.n, self.n))

    def solve(self):
        """SparseLinearSystem::solve, pose_graph_optimization.rs:124-144 (UMFPACK -> SuperLU)."""
        import scipy.sparse.linalg as spla
        A = self.csc()
        try:
            # H is symmetric with a positive diagonal: UMFPACK's default strategy picks its
            # symmetric path (AMD on A+A^T, diagonal pivots preferred); this is SuperLU's analogue.
            lu = spla.splu(A, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                           options=dict(SymmetricMode=True))
        except RuntimeError:
            lu = spla.splu(A)
        return lu.solve(self.b)


class OraclePoseGraph:
    """Mirror of `PoseGraph` (pose_graph_optimization.rs:155-163, 214-432) on the C restatement."""

    GAUSS_NEWTON, LEVENBERG_MARQUARDT = 0, 1

    def __init__(self, handle, solver=0):
        if not handle:
            raise ValueError("oracle: graph construction failed")
        self._h = handle
        self.solver = solver
        self.iteration = 0
        L = lib()
        self.len = L.og_len(self._h)
        self.n_vertices = L.og_num_vertices(self._h)
        self.n_edges = L.og_num_edges(self._h)

    # -- constructors ---------------------------------------------------------------
    @classmethod
    def from_g2o(cls, path, solver=0):
        """PoseGraph::new, :215-227 (parse_g2o, g2o.rs:35-143)."""
        err = C.create_string_buffer(256)
        h = lib().og_parse_g2o(os.fsencode(str(path)), err, 256)
        if not h:
            raise ValueError("oracle parse_g2o: " + err.value.decode())
        return cls(h, solver)

    @classmethod
    def from_arrays(cls, vertex_id, vertex_kind, vertex_values, edge_kind, edge_from, edge_to,
                    edge_meas, edge_info_upper, solver=0):
        a = [np.ascontiguousarray(vertex_id, np.uint32), np.ascontiguousarray(vertex_kind, np.uint8),
             np.ascontiguousarray(vertex_values, np.float64), np.ascontiguousarray(edge_kind, np.uint8),
             np.ascontiguousarray(edge_from, np.uint32), np.ascontiguousarray(edge_to, np.uint32),
             np.ascontiguousarray(edge_meas, np.float64), np.ascontiguousarray(edge_info_upper, np.float64)]
        h = lib().og_create(len(a[0]), _p(a[0]), _p(a[1]), _p(a[2]), len(a[3]), _p(a[3]), _p(a[4]), _p(a[5]),
                            _p(a[6]), _p(a[7]))
        if not h:
            raise ValueError("oracle: edge references unknown vertex id")
        return cls(h, solver)

    def __del__(self):
        try:
            if self._h:
                lib().og_free(self._h)
                self._h = None
        except Exception:
            pass

    # -- accessors -------------------------------------------------------------------
    def vertices(self):
        n = self.n_vertices
        vid = np.empty(n, np.uint32); kind = np.empty(n, np.uint8); off = np.empty(n, np.int64)
        val = np.empty(lib().og_num_vertex_values(self._h), np.float64)
        lib().og_get_vertices(self._h, _p(vid), _p(kind), _p(off), _p(val))
        return vid, kind, off, val

    def arrays(self):
        """The graph as the flat arrays `from_arrays` (and the product C ABI) take."""
        vid, kind, _, val = self.vertices()
        ne = self.n_edges
        ek = np.empty(ne, np.uint8)
        lib().og_get_edges(self._h, _p(ek), None, None)
        nm = int(np.sum(np.array(KIND_NVAL)[ek])); d = np.array(KIND_DIM)[ek]; ni = int(np.sum(d * (d + 1) // 2))
        ef = np.empty(ne, np.uint32); et = np.empty(ne, np.uint32)
        meas = np.empty(nm, np.float64); info = np.empty(ni, np.float64)
        lib().og_get_edge_data(self._h, _p(ef), _p(et), _p(meas), _p(info))
        return dict(vertex_id=vid, vertex_kind=kind, vertex_values=val, edge_kind=ek, edge_from=ef, edge_to=et,
                    edge_meas=meas, edge_info_upper=info)

    def edge_endpoints(self):
        ne = self.n_edges
        ek = np.empty(ne, np.uint8); fi = np.empty(ne, np.int64); ti = np.empty(ne, np.int64)
        lib().og_get_edges(self._h, _p(ek), _p(fi), _p(ti))
        return ek, fi, ti

    def state(self):
        s = np.empty((self.n_vertices, 8), np.float64)
        lib().og_get_state(self._h, _p(s))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, np.float64)
        assert s.shape == (self.n_vertices, 8)
        lib().og_set_state(self._h, _p(s))

    def edge_linearize(self, k):
        ek, fi, ti = self.edge_endpoints()
        _, kind, _, _ = self.vertices()
        d = KIND_DIM[ek[k]]; di = KIND_DIM[kind[fi[k]]]; dj = KIND_DIM[kind[ti[k]]]
        e = np.zeros(6); A = np.zeros(36); B = np.zeros(36)
        lib().og_edge_linearize(self._h, k, _p(e), _p(A), _p(B))
        return e[:d].copy(), A[:d * di].reshape(d, di).copy(), B[:d * dj].reshape(d, dj).copy()

    # -- the hot path ----------------------------------------------------------------
    def global_error(self):
        """global_error, :537-574."""
        return float(lib().og_global_error(self._h))

    def build_linear_system(self, lam=0.0, coo=False):
        """build_linear_system, :305-369, then the COO->CSC russell_sparse does before UMFPACK."""
        L = lib()
        lm = int(self.solver == self.LEVENBERG_MARQUARDT)
        cap = L.og_put_count(self._h, lm)
        ci = np.empty(cap, np.int32); cj = np.empty(cap, np.int32); cv = np.empty(cap, np.float64)
        b = np.empty(self.len, np.float64)
        puts = L.og_build_linear_system(self._h, lam, lm, _p(ci), _p(cj), _p(cv), _p(b))
        assert puts == cap
        col_ptr = np.empty(self.len + 1, np.int32); ri = np.empty(cap, np.int32); vals = np.empty(cap, np.float64)
        nnz = L.og_coo_to_csc(self.len, puts, _p(ci), _p(cj), _p(cv), _p(col_ptr), _p(ri), _p(vals))
        sls = LinearSystem(self.len, col_ptr, ri[:nnz].copy(), vals[:nnz].copy(), b, puts)
        if coo:
            sls.coo = (ci, cj, cv)
        return sls

    def linearize_and_solve(self):
        """linearize_and_solve, :371-373."""
        return self.build_linear_system(0.0).solve()

    def update_nodes(self, dx, sign=1.0):
        """update_nodes, :229-245."""
        dx = np.ascontiguousarray(dx, np.float64)
        assert dx.shape == (self.len,)
        lib().og_update_nodes(self._h, _p(dx), float(sign))

    def optimize(self, num_iterations, log=False, plot=False, return_norms=False):
        """PoseGraph::optimize, :247-303 (chi2 history; stop when |dx| < 1e-4; LM bookkeeping
        including the quirk that a rejected step's error is pushed and becomes last_error)."""
        tolerance = 1e-4
        lam = 0.01
        norms = []
        last_error = self.global_error()
        errors = [last_error]
        if log:
            print(f"Loaded graph with {self.n_vertices} nodes and {self.n_edges} edges")
            print(f"initial error :{errors[-1]:.5f}")
        for i in range(num_iterations):
            self.iteration += 1
            dx = self.build_linear_system(lam).solve()
            self.update_nodes(dx)
            norm_dx = float(np.linalg.norm(dx))
            error = self.global_error()
            if self.solver == self.LEVENBERG_MARQUARDT:
                if last_error < error:
                    self.update_nodes(dx, -1.0)
                    lam *= 2.0
                else:
                    lam /= 2.0
            last_error = error
            norms.append(norm_dx)
            errors.append(error)
            if log:
                print(f"step {i:3} : |dx| = {norm_dx:3.5f}, error = {errors[-1]:3.5f}")
            if norm_dx < tolerance:
                break
        return (errors, norms) if return_norms else errors
