"""Property tests of the oracle itself (hypothesis): symmetric finger geometry, rigid-motion invariance of the raw
GJK / EPA kernels, agreement of GJK and EPA at the contact boundary."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle
from oracle import Oracle


def _rot(axis, ang):
    axis = np.asarray(axis, float)
    axis /= np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


CUBE = np.array([[x, y, z] for x in (-.5, .5) for y in (-.5, .5) for z in (-.5, .5)])
finite = st.floats(-2, 2, allow_nan=False, allow_infinity=False)


@settings(max_examples=60, deadline=None)
@given(st.tuples(finite, finite, finite), st.tuples(finite, finite, finite), st.floats(0, 3.1), st.floats(0.2, 2.0))
def test_gjk_epa_rigid_motion_invariance(t, axis, ang, scale):
    if np.linalg.norm(axis) < 1e-3:
        axis = (0, 0, 1)
    A, B = CUBE * scale, CUBE * 0.7
    Tb = np.eye(4)
    Tb[:3, :3] = _rot(axis, ang)
    Tb[:3, 3] = t
    d0, _ = oracle.gjk_distance(A, B, Tb=Tb)
    # apply the same rigid motion G to both bodies: distances are unchanged
    G = np.eye(4)
    G[:3, :3] = _rot((1, 2, 3), 0.7)
    G[:3, 3] = (0.3, -1.0, 0.5)
    d1, _ = oracle.gjk_distance(A, B, Ta=G, Tb=G @ Tb)
    assert abs(d0 - d1) < 1e-9
    # symmetry
    d2, _ = oracle.gjk_distance(B, A, Ta=Tb)
    assert abs(d0 - d2) < 1e-9
    if d0 == 0.0:
        p0, _ = oracle.epa_depth(A, B, Tb=Tb)
        p1, _ = oracle.epa_depth(A, B, Ta=G, Tb=G @ Tb)
        assert p0 >= 0 and abs(p0 - p1) < 1e-7


@settings(max_examples=40, deadline=None)
@given(st.floats(0.0, 0.49), st.floats(-0.3, 0.3), st.floats(-0.3, 0.3))
def test_gjk_and_epa_meet_at_contact(pen, dy, dz):
    """Axis-aligned cubes pushed into each other by `pen` along x: separation distance and penetration depth are the
    same linear function of the offset on either side of contact."""
    T = np.eye(4)
    T[:3, 3] = (1.0 - pen, dy, dz)
    if pen > 0:
        dep, _ = oracle.epa_depth(CUBE, CUBE, Tb=T)
        assert abs(dep - pen) < 1e-9
    T[0, 3] = 1.0 + pen
    d, _ = oracle.gjk_distance(CUBE, CUBE, Tb=T)
    assert abs(d - pen) < 1e-12


def test_finger_symmetry(robot):
    """The right finger is the left finger rotated by pi about z (mobile_panda.urdf:376-382): mirrored hull."""
    left, right = robot.shape_verts(10), robot.shape_verts(11)
    mirrored = left * np.array([-1.0, -1.0, 1.0])
    assert sorted(map(tuple, np.round(mirrored, 9))) == sorted(map(tuple, np.round(right, 9)))
    o = Oracle(robot)
    T = o.fk(np.zeros((1, 9)))[0]
    # both fingers open symmetrically about the hand's x-z plane
    lf, rf, hand = T[12], T[13], T[11]
    rel_l = hand[:, :3].T @ (lf[:, 3] - hand[:, 3])
    rel_r = hand[:, :3].T @ (rf[:, 3] - hand[:, 3])
    assert np.allclose(rel_l * np.array([1, -1, 1]), rel_r, atol=1e-12) and abs(rel_l[1] - 0.04) < 1e-12


def test_non_finite_inputs_are_rejected(robot, scene_loader):
    """A NaN / infinite joint value (or pose entry) has no geometry: bit set, reason 1, distance = query distance, whatever
    the flags say -- the convention the CUDA path follows too (include/jogramop_b200.h, jg_check)."""
    import oracle
    o = oracle.Oracle(robot, scene_loader(11))
    q = np.tile(np.array([0.0, 0.0, -0.0178, -0.76, 0.0198, -2.342, 0.0298, 1.541, 0.753]), (4, 1))
    q[1, 3] = np.nan
    q[2, 0] = np.inf
    b, d, r = o.check(q, flags=oracle.SCENE | oracle.PLANE, want_reason=True)
    assert b.tolist() == [False, True, True, False] and r.tolist() == [0, 1, 1, 0]
    assert d[1] == d[2] == 0.01 and d[0] == d[3]
    T = np.tile(np.eye(4), (3, 1, 1))
    T[:, :3, 3] = [2.0, 0.0, 1.0]
    T[1, 0, 3] = np.nan
    pb, pd = o.check_poses(T, links=[11, 12, 13])
    assert pb.tolist() == [False, True, False] and pd[1] == 0.01
    eb, ed = o.check_edges(q[[0, 1]], q[[3, 3]], substeps=8)
    assert eb.tolist() == [False, True]
