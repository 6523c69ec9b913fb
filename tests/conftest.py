import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
SE2_GRAPHS = ["simulation-pose-pose", "simulation-pose-landmark", "intel", "dlr", "input_M3500_g2o"]
KEYS = ("vertex_id", "vertex_kind", "vertex_values", "edge_kind", "edge_from", "edge_to", "edge_meas", "edge_info_upper")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz")
    return {k: z[k] for k in z.files}


def graph_of(gold):
    return {k: gold[k] for k in KEYS}


@pytest.fixture(scope="session")
def built():
    """build (if stale) and return the product library; the oracle is built on first use."""
    import shutil
    from rustrobotics_b200 import _build
    if not _build.LIB.exists() and not (shutil.which("nvcc") or Path("/usr/local/cuda/bin/nvcc").exists()):
        pytest.skip("libpgo_b200.so is not built and there is no nvcc to build it: the host-side tests (g2o loader, symbolic pass, C ABI "
                    "surface) live in the same CUDA library as the kernels -- there is no CPU-only build, as there is no CPU fallback")
    _build.build()
    from rustrobotics_b200.mapping import _lib
    return _lib.lib()


@pytest.fixture(scope="session")
def g2o_files(tmp_path_factory, built):
    """the bundled reference graphs, re-written as g2o text from the committed fixtures"""
    from rustrobotics_b200 import write_g2o
    d = tmp_path_factory.mktemp("g2o")
    out = {}
    for name in SE2_GRAPHS:
        p = d / f"{name}.g2o"
        write_g2o(p, graph_of(load_golden(name)))
        out[name] = p
    return out
