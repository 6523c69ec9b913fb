"""ctypes binding of libpgo_b200.so: the C ABI of include/pgo_b200.h plus the C wrappers (pg_*) over the
C++ host mirror (csrc/host/pose_graph.hpp)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB = None
P = C.c_void_p

# every symbol include/pgo_b200.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "pgo_default_options", "pgo_create", "pgo_destroy", "pgo_last_error", "pgo_get_sizes", "pgo_chi2", "pgo_gn_step",
    "pgo_undo_last_step", "pgo_get_poses", "pgo_set_poses", "pgo_get_dx", "pgo_linearize_and_solve", "pgo_get_pattern",
    "pgo_get_block_structure", "pgo_get_anchor", "pgo_get_system", "pgo_get_timings", "pgo_time_spmv", "pgo_time_coarse", "pgo_get_stats",
    "pgo_version", "pgo_snapshot_poses", "pgo_restore_poses", "pgo_shard_handle_bytes", "pgo_shard_export",
    "pgo_shard_connect", "pgo_get_partition", "pgo_get_level_sizes", "pgo_get_aggregates",
]


class pgo_options(C.Structure):
    _fields_ = [("anchor_weight", C.c_double), ("pcg_rtol", C.c_double), ("pcg_max_iterations", C.c_int32),
                ("preconditioner", C.c_int32), ("sort_window", C.c_int32), ("amg_max_levels", C.c_int32),
                ("device", C.c_int32), ("world", C.c_int32), ("rank", C.c_int32), ("amg_dense_max", C.c_int32),
                ("amg_aggregate_size", C.c_int32), ("amg_kcycle", C.c_int32), ("amg_kcycle3", C.c_int32), ("amg_fp64_storage", C.c_int32),
                ("n_gpus", C.c_int32), ("device_ids", C.POINTER(C.c_int32)), ("refine", C.c_int32), ("refine_rtol", C.c_double)]


def lib_path() -> Path:
    return Path(__file__).resolve().parent.parent / "libpgo_b200.so"


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not path.exists():
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = C.CDLL(str(path))
    i64, i32, dbl = C.c_int64, C.c_int32, C.c_double
    L.pgo_version.restype = C.c_char_p
    L.pgo_last_error.restype = C.c_char_p
    L.pgo_last_error.argtypes = [P]
    L.pgo_default_options.argtypes = [C.POINTER(pgo_options)]
    L.pgo_create.argtypes = [C.POINTER(P), C.POINTER(pgo_options), i64, P, P, P, i64, P, P, P, P, P]
    L.pgo_destroy.argtypes = [P]
    L.pgo_get_sizes.argtypes = [P, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    L.pgo_chi2.argtypes = [P, C.POINTER(dbl)]
    L.pgo_gn_step.argtypes = [P, dbl, C.c_int, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(i32)]
    L.pgo_undo_last_step.argtypes = [P]
    L.pgo_get_poses.argtypes = [P, P, i64]
    L.pgo_set_poses.argtypes = [P, P, i64]
    L.pgo_get_dx.argtypes = [P, P, i64]
    L.pgo_snapshot_poses.argtypes = [P]
    L.pgo_shard_handle_bytes.restype = C.c_int
    L.pgo_shard_export.argtypes = [P, P, i64]
    L.pgo_shard_connect.argtypes = [P, P, i64]
    L.pgo_get_partition.argtypes = [P, C.POINTER(i32), C.POINTER(i32), P, P]
    L.pgo_get_level_sizes.argtypes = [P, i32, P, P]
    L.pgo_get_aggregates.argtypes = [P, i32, P, i64]
    L.pgo_restore_poses.argtypes = [P]
    L.pgo_linearize_and_solve.argtypes = [P, C.POINTER(i32)]
    L.pgo_get_pattern.argtypes = [P, C.POINTER(i64), C.POINTER(i64), P, P]
    L.pgo_get_block_structure.argtypes = [P, C.POINTER(i64), P, P, P]
    L.pgo_get_anchor.argtypes = [P, C.POINTER(i64)]
    L.pgo_get_system.argtypes = [P, dbl, C.c_int, P, P]
    L.pgo_get_timings.argtypes = [P, P, P, i32]
    L.pgo_time_spmv.argtypes = [P, i32, C.POINTER(dbl)]
    L.pgo_time_coarse.argtypes = [P, i32, i32, C.POINTER(dbl)]
    L.pgo_get_stats.argtypes = [P, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    # C++ host mirror
    L.pg_last_error.restype = C.c_char_p
    L.pg_new.restype = P
    L.pg_new.argtypes = [C.c_char_p, C.c_int, C.POINTER(pgo_options)]
    L.pg_from_arrays.restype = P
    L.pg_from_arrays.argtypes = [C.c_char_p, C.c_int, C.POINTER(pgo_options), i64, P, P, P, i64, i64, P, P, P, P, i64, P, i64]
    L.pg_free.argtypes = [P]
    L.pg_handle.restype = P
    L.pg_handle.argtypes = [P]
    for f in ("pg_num_nodes", "pg_num_edges", "pg_len"):
        getattr(L, f).restype = i64
        getattr(L, f).argtypes = [P]
    L.pg_optimize.restype = i64
    L.pg_optimize.argtypes = [P, i64, C.c_int, C.c_int, P, i64, P, P]
    L.pg_global_error.argtypes = [P, C.POINTER(dbl)]
    L.pg_plot.argtypes = [P]
    L.pg_parse_g2o.restype = P
    L.pg_parse_g2o.argtypes = [C.c_char_p]
    L.pg_graph_free.argtypes = [P]
    L.pg_graph_sizes.argtypes = [P] + [C.POINTER(i64)] * 6
    L.pg_graph_fill.argtypes = [P] * 9
    L.pg_write_g2o.argtypes = [C.c_char_p, i64, P, P, P, i64, i64, P, P, P, P, i64, P, i64]
    _LIB = L
    return L


def ptr(a):
    return a.ctypes.data_as(P) if a is not None else None
