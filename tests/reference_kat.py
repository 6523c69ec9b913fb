"""Known-answer values held by the reference's own tests (the parity anchors)."""

# g2o.rs:149-175  from_g2o: (nodes, edges, len)
FROM_G2O = {
    "simulation-pose-pose": (400, 1773, 1200),
    "simulation-pose-landmark": (77, 297, 195),
    "intel": (1728, 4830, 5184),
    "dlr": (3873, 17605, 11043),
}
# pose_graph_optimization.rs:580-598  initial_global_error: (value, epsilon)
INITIAL_ERROR = {
    "simulation-pose-pose": (138862234.0, 10.0),
    "simulation-pose-landmark": (3030.0, 1.0),
    "intel": (1795139.0, 1e-2),
    "dlr": (369655336.0, 10.0),
}
# :600-631  final_global_error after optimize(100) with Gauss-Newton: (value, epsilon)
FINAL_ERROR = {
    "simulation-pose-pose": (8269.0, 1.0),
    "simulation-pose-landmark": (474.0, 1.0),
    "intel": (360.0, 1.0),
    "dlr": (56860.0, 1.0),
}
# :633-690  linearize_pose_pose_constraint_correct on simulation-pose-landmark, edges[0] and edges[10] (eps 1e-3)
POSE_POSE_JAC = {
    0: ([[0.0, 1.0, 0.113], [-1.0, 0.0, 0.024], [0.0, 0.0, -1.0]], [[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]]),
    10: ([[0.037, 0.999, 0.138], [-0.999, 0.037, -0.982], [0.0, 0.0, -1.0]],
         [[-0.037, -0.999, 0.0], [0.999, -0.037, 0.0], [0.0, 0.0, 1.0]]),
}
# :692-722  linearize_pose_landmark_constraint_correct, edges[1] (eps 1e-3)
POSE_LANDMARK_JAC = {1: ([[0.0, 1.0, 0.358], [-1.0, 0.0, -0.051]], [[0.0, -1.0], [1.0, 0.0]])}
# :724-739  linearize_and_solve_correct: first five entries of dx (eps 1e-3)
FIRST_DX = [1.68518905e-01, 5.74311089e-01, -5.08805168e-02, -3.67482151e-02, 8.89458085e-01]
