"""Mirror of reference src/mapping/pose_graph_optimization.rs' public API.

`PoseGraph(file_path, solver)` = `PoseGraph::new` (:215), `optimize(num_iterations, log, plot)` (:247) returns the
chi2 history, `plot()` (:375).  The loop lives in the C++ host mirror (csrc/host/pose_graph.cpp) and every
Gauss-Newton step runs on the GPU through the C ABI (include/pgo_b200.h); this module only marshals arrays.
"""
from __future__ import annotations

import ctypes as C
import enum
import os

import numpy as np

from ._lib import lib, pgo_options, ptr
from .g2o import KEYS


class PgoError(RuntimeError):
    """What the reference reports as Err(Box<dyn Error>)."""


class PoseGraphSolver(enum.IntEnum):   # :28-32
    GaussNewton = 0
    LevenbergMarquardt = 1


STATUS = {0: "ok", 1: "bad argument", 2: "CUDA error", 3: "multi-GPU communication error", 4: "solver breakdown", 5: "PCG not converged",
          6: "unsupported"}
BLOCK_JACOBI, AMG = 0, 1


def Options(**kw) -> pgo_options:
    """pgo_options with the library defaults, overridden by keyword (anchor_weight, pcg_rtol,
    pcg_max_iterations, preconditioner, sort_window, amg_max_levels, device, world, rank, amg_dense_max,
    amg_aggregate_size, amg_kcycle, amg_kcycle3, amg_fp64_storage, n_gpus, device_ids, refine, refine_rtol).

    Multi-GPU from ONE process: `Options(n_gpus=8)` (devices 0..7) or `Options(n_gpus=2, device_ids=[0, 0])`
    (two shards sharing GPU 0: the multi-GPU path on a one-GPU machine)."""
    o = pgo_options()
    lib().pgo_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k}")
        if k == "device_ids" and v is not None:
            ids = (C.c_int32 * len(v))(*[int(d) for d in v])
            o._device_ids_keepalive = ids          # pgo_create borrows the array
            v = C.cast(ids, C.POINTER(C.c_int32))
            if "n_gpus" not in kw:
                o.n_gpus = len(ids)
        setattr(o, k, v)
    return o


class PoseGraph:
    def __init__(self, file_path=None, solver=PoseGraphSolver.GaussNewton, *, graph=None, name="graph", options=None,
                 comm=None):
        """`comm` is only for the process-per-GPU mode (options.world > 1, e.g. under torchrun): an object with
        `all_gather_bytes(bytes) -> bytes` (rank order), `all_reduce_sum(ndarray) -> ndarray` and `barrier()`, supplied by
        the launcher (bench.py's TorchComm).  This package itself imports no torch; the single-process multi-GPU mode
        (options.n_gpus) needs no comm at all."""
        L = lib()
        self._pg = None
        self.solver = PoseGraphSolver(solver)
        opt = C.byref(options) if options is not None else None
        if file_path is not None:
            self._pg = L.pg_new(os.fsencode(str(file_path)), int(self.solver), opt)
        elif graph is not None:
            a = {k: np.ascontiguousarray(graph[k], dt) for k, dt in zip(
                KEYS, (np.uint32, np.uint8, np.float64, np.uint8, np.uint32, np.uint32, np.float64, np.float64))}
            self._pg = L.pg_from_arrays(name.encode(), int(self.solver), opt, len(a["vertex_id"]), ptr(a["vertex_id"]),
                                        ptr(a["vertex_kind"]), ptr(a["vertex_values"]), len(a["vertex_values"]),
                                        len(a["edge_kind"]), ptr(a["edge_kind"]), ptr(a["edge_from"]), ptr(a["edge_to"]),
                                        ptr(a["edge_meas"]), len(a["edge_meas"]), ptr(a["edge_info_upper"]),
                                        len(a["edge_info_upper"]))
        else:
            raise TypeError("PoseGraph needs a g2o file path or graph=<arrays>")
        if not self._pg:
            raise PgoError(L.pg_last_error().decode())
        self._h = L.pg_handle(self._pg)
        self.norms: list[float] = []
        self.pcg_iterations: list[int] = []
        self.world = int(options.world) if options is not None else 1
        self.rank = int(options.rank) if options is not None else 0
        self._comm = comm
        if self.world > 1 and options.device != -2:
            self.connect_shards()

    @classmethod
    def new(cls, file_path, solver=PoseGraphSolver.GaussNewton, **kw):
        return cls(file_path, solver, **kw)

    def close(self):
        if getattr(self, "_pg", None):
            lib().pg_free(self._pg)
            self._pg = None
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- process-per-GPU sharding (options.world > 1; the launcher supplies `comm`) -----------------------------
    def connect_shards(self):
        """Exchange the ranks' peer-memory handles (one all-gather of 80 bytes per rank at start-up, moved by the launcher's
        `comm`) and connect the shards.  Afterwards every computing call is collective."""
        if self._comm is None:
            raise PgoError("process-per-GPU PoseGraph (options.world > 1) needs comm=<launcher's communicator>; "
                           "for multi-GPU from one process use options.n_gpus instead")
        n = lib().pgo_shard_handle_bytes()
        buf = C.create_string_buffer(n)
        self._check(lib().pgo_shard_export(self._h, buf, n), "pgo_shard_export")
        blob = self._comm.all_gather_bytes(buf.raw)
        if len(blob) != n * self.world:
            raise PgoError("shard handle exchange failed")
        self._check(lib().pgo_shard_connect(self._h, blob, self.world), "pgo_shard_connect")
        self._comm.barrier()           # every rank has opened every arena before the first peer read

    def _merge_owned(self, out):
        """process-per-GPU getters fill the span this rank owns: sum the (disjoint, zero elsewhere) spans over ranks"""
        if self.world == 1:
            return out
        return self._comm.all_reduce_sum(out)

    # ---- reference API ----------------------------------------------------------------------
    def optimize(self, num_iterations, log=False, plot=False):
        """-> chi2 history [chi2(x0), chi2(x1), ...] (:247-303)."""
        cap = int(num_iterations) + 1
        err = np.empty(cap); nrm = np.empty(cap); its = np.empty(cap, np.int32)
        n = lib().pg_optimize(self._pg, int(num_iterations), int(bool(log)), int(bool(plot)), ptr(err), cap, ptr(nrm), ptr(its))
        if n < 0:
            raise PgoError(lib().pg_last_error().decode())
        self.norms += nrm[:n - 1].tolist()
        self.pcg_iterations += its[:n - 1].tolist()
        return err[:n].tolist()

    def plot(self):
        if lib().pg_plot(self._pg) != 0:
            raise PgoError(lib().pg_last_error().decode())

    # ---- additions the north star asks for / parity hooks -------------------------------------
    @property
    def num_nodes(self): return lib().pg_num_nodes(self._pg)
    @property
    def num_edges(self): return lib().pg_num_edges(self._pg)
    @property
    def len(self): return lib().pg_len(self._pg)

    def _check(self, rc, what):
        if rc != 0:
            raise PgoError(f"{what}: {STATUS.get(rc, rc)}: {lib().pgo_last_error(self._h).decode()}")

    def global_error(self):
        v = C.c_double()
        self._check(lib().pgo_chi2(self._h, C.byref(v)), "pgo_chi2")
        return v.value

    def sizes(self):
        s = [C.c_int64() for _ in range(4)]
        lib().pgo_get_sizes(self._h, *[C.byref(x) for x in s])
        return tuple(x.value for x in s)   # n_vertices, n_edges, len, n_vertex_values

    def poses(self):
        """vertex values in the input packing (x y theta | x y), lut order."""
        out = np.zeros(self.sizes()[3])
        self._check(lib().pgo_get_poses(self._h, ptr(out), len(out)), "pgo_get_poses")
        return self._merge_owned(out)

    def set_poses(self, values):
        v = np.ascontiguousarray(values, np.float64)
        self._check(lib().pgo_set_poses(self._h, ptr(v), len(v)), "pgo_set_poses")

    def snapshot_poses(self):
        self._check(lib().pgo_snapshot_poses(self._h), "pgo_snapshot_poses")

    def restore_poses(self):
        self._check(lib().pgo_restore_poses(self._h), "pgo_restore_poses")

    def gn_step(self, lam=0.0, add_lambda=False, allow_not_converged=True):
        nd, c2, it = C.c_double(), C.c_double(), C.c_int32()
        rc = lib().pgo_gn_step(self._h, lam, int(add_lambda), C.byref(nd), C.byref(c2), C.byref(it))
        if rc != 0 and not (allow_not_converged and rc == 5):
            self._check(rc, "pgo_gn_step")
        return nd.value, c2.value, it.value

    def undo_last_step(self):
        self._check(lib().pgo_undo_last_step(self._h), "pgo_undo_last_step")

    def linearize_and_solve(self):
        it = C.c_int32()
        self._check(lib().pgo_linearize_and_solve(self._h, C.byref(it)), "pgo_linearize_and_solve")
        return self.dx(), it.value

    def dx(self):
        dx = np.zeros(self.len)
        self._check(lib().pgo_get_dx(self._h, ptr(dx), len(dx)), "pgo_get_dx")
        return self._merge_owned(dx)

    def pattern(self):
        n, nnz = C.c_int64(), C.c_int64()
        self._check(lib().pgo_get_pattern(self._h, C.byref(n), C.byref(nnz), None, None), "pgo_get_pattern")
        cp = np.empty(n.value + 1, np.int32); ri = np.empty(nnz.value, np.int32)
        self._check(lib().pgo_get_pattern(self._h, C.byref(n), C.byref(nnz), ptr(cp), ptr(ri)), "pgo_get_pattern")
        return cp, ri

    def block_structure(self):
        nb = C.c_int64()
        self._check(lib().pgo_get_block_structure(self._h, C.byref(nb), None, None, None), "pgo_get_block_structure")
        nv, ne = self.sizes()[:2]
        rp = np.empty(nv + 1, np.int64); bc = np.empty(nb.value, np.int32); es = np.empty(4 * ne, np.int64)
        self._check(lib().pgo_get_block_structure(self._h, C.byref(nb), ptr(rp), ptr(bc), ptr(es)), "pgo_get_block_structure")
        return rp, bc, es.reshape(ne, 4)

    def anchor(self):
        v = C.c_int64()
        self._check(lib().pgo_get_anchor(self._h, C.byref(v)), "pgo_get_anchor")
        return v.value

    def system(self, lam=0.0, add_lambda=False):
        cp, ri = self.pattern()
        vals = np.zeros(len(ri)); b = np.zeros(self.len)
        self._check(lib().pgo_get_system(self._h, lam, int(add_lambda), ptr(vals), ptr(b)), "pgo_get_system")
        return cp, ri, self._merge_owned(vals), self._merge_owned(b)

    def timings(self):
        ms = np.zeros(6); ln = np.zeros(6, np.int64)
        lib().pgo_get_timings(self._h, ptr(ms), ptr(ln), 6)
        names = ("assemble", "precond_setup", "pcg", "retract", "chi2", "spmv_fine")
        return {k: (float(m), int(c)) for k, m, c in zip(names, ms, ln)}

    def time_spmv(self, repeats=20):
        v = C.c_double()
        self._check(lib().pgo_time_spmv(self._h, repeats, C.byref(v)), "pgo_time_spmv")
        return v.value

    def time_coarse(self, level, repeats=20):
        v = C.c_double()
        self._check(lib().pgo_time_coarse(self._h, level, repeats, C.byref(v)), "pgo_time_coarse")
        return v.value

    def level_sizes(self):
        rows = np.zeros(16, np.int64); blocks = np.zeros(16, np.int64)
        nl = lib().pgo_get_level_sizes(self._h, 16, ptr(rows), ptr(blocks))
        return rows[:nl].tolist(), blocks[:nl].tolist()

    def aggregates(self, level=0, n=None):
        """aggregate (row of level + 1, global padded numbering) of every vertex in lut order (level 0) or of every padded
        row of `level` (n = number of padded rows; -1 marks padding rows)"""
        if level == 0:
            nv = C.c_int64(); ne = C.c_int64(); ln = C.c_int64(); nval = C.c_int64()
            lib().pgo_get_sizes(self._h, C.byref(nv), C.byref(ne), C.byref(ln), C.byref(nval))
            n = nv.value
        out = np.zeros(n, np.int32)
        self._check(lib().pgo_get_aggregates(self._h, level, ptr(out), n), "pgo_get_aggregates")
        return out

    def partition(self):
        w, r = C.c_int32(), C.c_int32()
        lib().pgo_get_partition(self._h, C.byref(w), C.byref(r), None, None)
        vr = np.zeros(w.value + 1, np.int64); rb = np.zeros(w.value, np.int64)
        lib().pgo_get_partition(self._h, C.byref(w), C.byref(r), ptr(vr), ptr(rb))
        return dict(world=w.value, rank=r.value, vertex_range=vr.tolist(), remote_blocks=rb.tolist())

    def stats(self):
        s = [C.c_int64() for _ in range(4)]
        lib().pgo_get_stats(self._h, *[C.byref(x) for x in s])
        return dict(block_rows=s[0].value, offdiag_blocks=s[1].value, levels=s[2].value, device_bytes=s[3].value)
