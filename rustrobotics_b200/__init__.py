"""pgo-b200: B200-native (CUDA sm_100a, fp64) pose-graph optimisation behind the API of
jgsimard/RustRobotics' `robotics::mapping::{PoseGraph, PoseGraphSolver}`.

The compute path is the in-tree shared library `libpgo_b200.so` (C ABI: include/pgo_b200.h).
There is no CPU fallback: constructing a PoseGraph without a CUDA device raises.
"""
from .mapping import PoseGraph, PoseGraphSolver, PgoError, Options, parse_g2o, write_g2o  # noqa: F401

__version__ = "0.1.0"
