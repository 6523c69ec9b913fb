"""In-tree build of libpgo_b200.so (nvcc, sm_100a) and the synthetic-graph helper library."""
from __future__ import annotations

import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent
LIB = ROOT / "libpgo_b200.so"


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    srcs = list((ROOT / "csrc").rglob("*.cu")) + list((ROOT / "csrc").rglob("*.cuh")) + list((ROOT / "csrc").rglob("*.cpp")) + \
        list((ROOT / "csrc").rglob("*.h")) + list((ROOT / "csrc").rglob("*.hpp")) + [ROOT.parent / "include" / "pgo_b200.h"]
    return any(s.stat().st_mtime > t for s in srcs)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a (`-gencode arch=compute_100a,code=sm_100a -lineinfo`)."""
    if force or _stale():
        cmd = ["make", "-C", str(ROOT / "csrc")] + (["-B"] if force else [])
        subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.DEVNULL)
    from . import synthetic
    synthetic.build(force)
    return LIB
