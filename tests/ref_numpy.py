"""Independent numpy/scipy cross-checks for the oracle (NOT the oracle itself).

Different algorithms on purpose: FK by plain 4x4 matrix products, exact polytope distance as the
min-norm point of the Minkowski-difference vertex set (NNLS), exact penetration depth from qhull's
facets of the Minkowski difference.
"""
import numpy as np
from scipy.optimize import nnls
from scipy.spatial import ConvexHull

SCENARIO_IDS = [11, 12, 13, 14, 15, 21, 22, 23, 24, 25, 31, 32, 33, 34, 35, 41, 42, 43, 44, 45]


def uniform_configs(robot, n, seed):
    """q = lo + (hi-lo)*rng.random((n,9)), float32."""
    lim = robot.joint_limits[robot.arm_joint_ids]
    rng = np.random.default_rng(seed)
    return (lim[:, 0] + (lim[:, 1] - lim[:, 0]) * rng.random((n, len(lim)))).astype(np.float32)


def fk_numpy(robot, q, finger_open=0.04):
    """All link frames [L,4,4] in the robot base frame for one configuration."""
    val = np.zeros(robot.n_links)
    for k, j in enumerate(robot.arm_joint_ids):
        val[j] = q[k]
    for j in robot.finger_joint_ids:
        val[j] = finger_open
    out = np.zeros((robot.n_links, 4, 4))
    for i in range(robot.n_links):
        M = np.eye(4)
        a = robot.joint_axis[i]
        if robot.joint_type[i] == 1:
            a = a / np.linalg.norm(a)
            K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
            M[:3, :3] = np.eye(3) + np.sin(val[i]) * K + (1 - np.cos(val[i])) * (K @ K)
        elif robot.joint_type[i] == 2:
            M[:3, 3] = a * val[i]
        T = robot.joint_origin[i] @ M
        p = robot.joint_parent[i]
        out[i] = out[p] @ T if p >= 0 else T
    return out


def md_points(A, B):
    return (A[:, None, :] - B[None, :, :]).reshape(-1, 3)


def md_distance_nnls(A, B, big=1e3):
    """Distance between conv(A) and conv(B): min-norm point of conv(A-B) via NNLS on [P^T; big*1^T] l = [0; big]."""
    P = md_points(A, B)
    M = np.vstack([P.T, big * np.ones((1, len(P)))])
    rhs = np.array([0, 0, 0, big], dtype=float)
    lam, _ = nnls(M, rhs, maxiter=20 * len(P))
    lam = lam / lam.sum()
    return float(np.linalg.norm(P.T @ lam))


def md_penetration_qhull(A, B):
    """Exact penetration depth of overlapping polytopes (min facet distance of the Minkowski difference), or
    None if the origin is outside."""
    hull = ConvexHull(md_points(A, B))
    off = hull.equations[:, 3]       # n.x + off <= 0 inside
    if off.max() > 0:
        return None
    return float((-off).min())


def transform(T, v):
    return v @ T[:3, :3].T + T[:3, 3]
