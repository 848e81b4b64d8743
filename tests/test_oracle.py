"""Pins the CPU oracle (oracle/oracle.c): golden artefacts G3/G4/G5 of the reference + independent numerics."""
import numpy as np
import pytest

import oracle
from oracle import Oracle
from ref_numpy import (SCENARIO_IDS, fk_numpy, md_distance_nnls, md_penetration_qhull, transform,
                       uniform_configs)

ALL = oracle.SCENE | oracle.PLANE | oracle.SELF | oracle.LIMITS
HOME = np.array([0.0, 0.0, -0.017792060227770554, -0.7601235411041661, 0.019782607023391807, -2.342050140544315,
                 0.029840531355804868, 1.5411935298621688, 0.7534486589746342])


def test_fk_matches_numpy(robot):
    o = Oracle(robot)
    q = uniform_configs(robot, 50, 1).astype(np.float64)
    T = o.fk(q)
    for i in range(len(q)):
        assert np.abs(T[i] - fk_numpy(robot, q[i])[:, :3, :]).max() < 1e-12


def test_g3_fk_known_answers(robot, golden, scene_loader):
    """Every accepted IK row reaches one of the 200 grasps: pos_err + deg/1000 <= 0.015 (simulation.py:306).
    Pins joint order [base_y, base_x, j1..j7], axes, origins and the EE frame (link 14)."""
    o = Oracle(robot)
    n = 0
    worst = 0.0
    for sid in SCENARIO_IDS:
        ik = golden[f'ik_{sid:03d}']
        if not len(ik):
            continue
        T = o.fk(ik)[:, robot.end_effector_id]
        gr = golden[f'grasps_csv_{sid:03d}'].reshape(-1, 4, 4)
        for t in T:
            pe = np.linalg.norm(gr[:, :3, 3] - t[:, 3], axis=1)
            Rrel = np.einsum('nij,kj->nik', gr[:, :3, :3], t[:, :3])
            ang = np.degrees(np.arccos(np.clip((np.trace(Rrel, axis1=1, axis2=2) - 1) / 2, -1, 1)))
            e = (pe + ang / 1000).min()
            worst = max(worst, e)
            assert e <= 0.015
            n += 1
    assert n == 89
    assert worst < 0.0083        # worst combined error 0.00827


@pytest.mark.parametrize('sid', SCENARIO_IDS)
def test_g4_known_negatives(sid, robot, golden, scene_loader):
    """Each golden IK row passed joint limits, in_self_collision()==False and in_collision()==False under
    pyBullet (scenario.py:109-119)."""
    ik = golden[f'ik_{sid:03d}']
    if not len(ik):
        pytest.skip('no accepted IK rows for this scenario')
    o = Oracle(robot, scene_loader(sid))
    bits, dist, reason = o.check(ik, flags=ALL, want_reason=True)
    assert not bits.any(), (reason, dist)


def test_g4_smallest_clearances(robot, golden, scene_loader):
    """smallest accepted raw clearances sit just above the three cutoffs
    (hull-hull 1.22 mm vs 1 mm, box-hull 0.78 mm vs ~0, hull-plane 0.61 mm vs 0)."""
    hull_min, box_min, plane_min, self_min = np.inf, np.inf, np.inf, np.inf
    for sid in SCENARIO_IDS:
        ik = golden[f'ik_{sid:03d}']
        if not len(ik):
            continue
        sc = scene_loader(sid)
        o = Oracle(robot, sc)
        for q in ik:
            pd = o.pair_distances(q)
            hull_min = min(hull_min, pd[1:, :sc.n_pieces].min() + 0.002)   # raw hull distance
            box_min = min(box_min, pd[0, :sc.n_pieces].min() + 0.001)      # ~raw box distance (face contact)
            plane_min = min(plane_min, pd[1:, sc.n_pieces].min() + 0.001)  # lowest hull vertex above the plane
            self_min = min(self_min, o.self_distances(q).min() + 0.002)
    assert 0.00120 < hull_min < 0.00125
    assert 0.00070 < box_min < 0.00085
    assert 0.00055 < plane_min < 0.00065
    assert self_min > 0.0028


def test_q1_single_hull_gate_rejects_a_golden_row(robot, golden, scene_loader):
    """With the stale one-piece gate the solid block fills the opening and a golden row collides (Q1)."""
    o = Oracle(robot, scene_loader(45, '_singlehull'))
    bits, _ = o.check(golden['ik_045'])
    assert bits.sum() == 1


def test_g5_home_is_free(robot, scene_loader):
    for sid in SCENARIO_IDS:
        o = Oracle(robot, scene_loader(sid))
        bits, dist = o.check(HOME[None], flags=ALL)
        assert not bits[0] and dist[0] == pytest.approx(0.01)   # nothing within the query distance


def test_gjk_against_nnls(robot, scene_loader):
    """GJK core distance == min-norm point of the Minkowski difference (independent NNLS solve)."""
    sc = scene_loader(11)
    o = Oracle(robot, sc)
    q = uniform_configs(robot, 400, 11).astype(np.float64)
    checked = 0
    for i in range(len(q)):
        pd = o.pair_distances(q[i], flags=oracle.SCENE | oracle.NO_EPA)
        T = o.fk(q[i])[0]
        for s, p in zip(*np.nonzero(np.isfinite(pd[:, :sc.n_pieces]))):
            if pd[s, p] <= -0.002 + 1e-12:
                continue     # cores overlap
            A = transform(T[robot.shape_link[s]], robot.shape_verts(s))
            ref = md_distance_nnls(A, sc.piece(p)) - 0.002
            assert abs(ref - pd[s, p]) < 2e-7, (i, s, p, ref, pd[s, p])
            checked += 1
            if checked >= 60:
                return
    assert checked >= 20


def test_epa_against_qhull(robot, scene_loader):
    """EPA depth == min facet distance of the Minkowski-difference hull (qhull)."""
    sc = scene_loader(11)
    o = Oracle(robot, sc)
    q = uniform_configs(robot, 600, 12).astype(np.float64)
    checked = 0
    for i in range(len(q)):
        pd = o.pair_distances(q[i], flags=oracle.SCENE)
        T = o.fk(q[i])[0]
        for s, p in zip(*np.nonzero(pd[:, :sc.n_pieces] < -0.002)):
            A = transform(T[robot.shape_link[s]], robot.shape_plane_verts(s))
            ref = md_penetration_qhull(A, sc.piece(p))
            assert ref is not None
            em = robot.shape_plane_margin[s] + 0.001
            assert abs(-(ref + em) - pd[s, p]) < 1e-7, (i, s, p, ref, pd[s, p])
            checked += 1
            if checked >= 60:
                return
    assert checked >= 20


def test_raw_kernels_simple_shapes():
    cube = np.array([[x, y, z] for x in (-.5, .5) for y in (-.5, .5) for z in (-.5, .5)])
    T = np.eye(4)
    T[0, 3] = 1.5
    d, _ = oracle.gjk_distance(cube, cube, Tb=T)
    assert d == pytest.approx(0.5, abs=1e-12)
    T[0, 3] = 0.75
    dep, _ = oracle.epa_depth(cube, cube, Tb=T)
    assert dep == pytest.approx(0.25, abs=1e-9)
    T[:3, 3] = [1.5, 1.5, 0]
    d, _ = oracle.gjk_distance(cube, cube, Tb=T)
    assert d == pytest.approx(np.sqrt(0.5), abs=1e-12)
    tet = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    T = np.eye(4)
    T[:3, 3] = [0.1, 0.1, 0.1]
    dep, _ = oracle.epa_depth(tet, tet * 0.01, Tb=T)   # tiny tetra inside: exit through the nearest face
    assert dep == pytest.approx(0.11, abs=1e-9)


def test_threshold_and_query_distance_semantics(robot, scene_loader):
    """dist is clamped at the query distance; verdict == (dist < threshold); monotone in the threshold."""
    sc = scene_loader(34)
    o = Oracle(robot, sc)
    q = uniform_configs(robot, 2000, 34).astype(np.float64)
    bits, dist = o.check(q)
    assert dist.max() <= 0.01 and (dist == 0.01).any()
    assert np.array_equal(bits, dist < -0.001)
    b2, _ = o.check(q, threshold=0.0)
    assert (b2 >= bits).all()
    assert 0.2 < bits.mean() < 0.5     # obstacle+plane collision rate ~0.3-0.4 on uniform configs


def test_edges_and_poses_consistency(robot, scene_loader):
    sc = scene_loader(34)
    o = Oracle(robot, sc)
    q = uniform_configs(robot, 40, 5).astype(np.float64)
    qb = np.clip(q + 0.1, robot.joint_limits[robot.arm_joint_ids][:, 0], robot.joint_limits[robot.arm_joint_ids][:, 1])
    be, de = o.check_edges(q, qb, substeps=4)
    for e in range(len(q)):
        sub = np.array([q[e] + k / 3.0 * (qb[e] - q[e]) for k in range(4)])
        b, d = o.check(sub)
        assert be[e] == b.any() and de[e] == pytest.approx(d.min())
    # free-floating gripper at the FK pose of the EE == hand + fingers of the full robot at q
    T = o.fk(q)
    poses = T[:, robot.end_effector_id]
    bp, dp = o.check_poses(poses, links=[11, 12, 13])
    for i in range(len(q)):
        pd = o.pair_distances(q[i])
        ref = min(pd[9:12].min(), 0.01)
        assert dp[i] == pytest.approx(ref, abs=1e-12)
