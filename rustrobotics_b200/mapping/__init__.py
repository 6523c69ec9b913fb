"""Mirror of `robotics::mapping` (reference src/mapping/mod.rs:6 re-exports PoseGraph, PoseGraphSolver)."""
from .g2o import parse_g2o, write_g2o  # noqa: F401
from .pose_graph_optimization import Options, PgoError, PoseGraph, PoseGraphSolver  # noqa: F401
