"""Deterministic synthetic pose graphs for the BASELINE.json benchmark configs.

The reference ships no generator; `manhattan.c` is this repo's definition.  Graphs come back as
the flat arrays the C ABI (`include/pgo_b200.h: pgo_create`) takes.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so, src = _HERE / "libpgo_synth.so", _HERE / "manhattan.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=gnu11", "-shared", "-o", str(so), str(src), "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        P = C.c_void_p
        L.pgo_synth_manhattan_se2.restype = C.c_int64
        L.pgo_synth_manhattan_se2.argtypes = [C.c_int64, C.c_int64, C.c_uint64, P, P, P, P, P, P]
        L.pgo_synth_sphere_se3.restype = C.c_int64
        L.pgo_synth_sphere_se3.argtypes = [C.c_int64, C.c_int64, C.c_double, C.c_uint64, P, P, P, P, P, P]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def manhattan_se2(n_poses: int, n_edges: int | None = None, seed: int = 42, with_ground_truth: bool = False):
    """SE(2) Manhattan-world graph: `n_poses` poses, `n_edges` (default 4*n_poses) edges."""
    n = int(n_poses)
    target = 4 * n if n_edges is None else int(n_edges)
    vv = np.empty(3 * n); gt = np.empty(3 * n)
    ef = np.empty(target, np.uint32); et = np.empty(target, np.uint32)
    em = np.empty(3 * target); ei = np.empty(6 * target)
    ne = _lib().pgo_synth_manhattan_se2(n, target, seed, _p(vv), _p(gt), _p(ef), _p(et), _p(em), _p(ei))
    if ne < 0:
        raise ValueError("manhattan_se2: bad arguments")
    g = dict(vertex_id=np.arange(n, dtype=np.uint32), vertex_kind=np.zeros(n, np.uint8), vertex_values=vv,
             edge_kind=np.zeros(ne, np.uint8), edge_from=ef[:ne].copy(), edge_to=et[:ne].copy(),
             edge_meas=em[:3 * ne].copy(), edge_info_upper=ei[:6 * ne].copy())
    if with_ground_truth:
        g["ground_truth"] = gt
    return g


def sphere_se3(levels: int, per_level: int, radius: float = 100.0, seed: int = 42, with_ground_truth: bool = False):
    """SE(3) sphere-spiral graph: levels*per_level poses, ~4x as many edges."""
    n = int(levels) * int(per_level)
    cap = 4 * n
    vv = np.empty(7 * n); gt = np.empty(7 * n)
    ef = np.empty(cap, np.uint32); et = np.empty(cap, np.uint32)
    em = np.empty(7 * cap); ei = np.empty(21 * cap)
    ne = _lib().pgo_synth_sphere_se3(levels, per_level, radius, seed, _p(vv), _p(gt), _p(ef), _p(et), _p(em), _p(ei))
    if ne < 0:
        raise ValueError("sphere_se3: bad arguments")
    g = dict(vertex_id=np.arange(n, dtype=np.uint32), vertex_kind=np.full(n, 2, np.uint8), vertex_values=vv,
             edge_kind=np.full(ne, 2, np.uint8), edge_from=ef[:ne].copy(), edge_to=et[:ne].copy(),
             edge_meas=em[:7 * ne].copy(), edge_info_upper=ei[:21 * ne].copy())
    if with_ground_truth:
        g["ground_truth"] = gt
    return g
