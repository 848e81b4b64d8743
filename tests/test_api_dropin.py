"""Drop-in surface: same names / arguments / error behaviour as the reference's scenario.py and simulation.py
(SURVEY.md §8b).  CPU part: signatures, state handling, exceptions.  GPU part: single-shot semantics vs the oracle."""
import ast
import inspect
import os

import numpy as np
import pytest

from jogramop_framework_b200 import ingest, util
from jogramop_framework_b200.scenario import Scenario
from jogramop_framework_b200.simulation import FrankaRobot, GraspingSimulator

REFERENCE = '/root/reference'

EXPECTED = {   # name -> positional argument names (without self), from the reference
    (Scenario, '__init__'): ['scenario_id', 'robot_pose', 'with_platform'],
    (Scenario, 'select_n_grasps'): ['n', 'seed'],
    (Scenario, 'get_robot_and_sim'): ['with_gui'],
    (Scenario, 'default_robot_pose'): [],
    (Scenario, 'get_ik_solutions'): [],
    (GraspingSimulator, '__init__'): ['verbose'],
    (GraspingSimulator, 'links_in_collision'): ['body1', 'link1', 'body2', 'link2', 'threshold'],
    (GraspingSimulator, 'body_in_collision'): ['query_body'],
    (GraspingSimulator, 'add_scene_from_file'): ['scene_fn'],
    (FrankaRobot, '__init__'): ['simulator', 'pose', 'with_platform'],
    (FrankaRobot, 'reset_arm_joints'): ['joint_config'],
    (FrankaRobot, 'reset_gripper'): ['open_scale'],
    (FrankaRobot, 'forward_kinematics'): ['joint_conf', 'as_matrix'],
    (FrankaRobot, 'inverse_kinematics'): ['pos', 'orn', 'combined_threshold', 'null_space_control'],
    (FrankaRobot, 'get_arm_joint_conf_from_motor_joint_conf'): ['ik_joint_config'],
    (FrankaRobot, 'arm_joint_limits'): [], (FrankaRobot, 'arm_joints_pos'): [], (FrankaRobot, 'joint_pos'): [],
    (FrankaRobot, 'in_self_collision'): [], (FrankaRobot, 'in_collision'): [], (FrankaRobot, 'end_effector_pose'): [],
    (FrankaRobot, 'reset_home_pose'): [],
}


def test_signatures_match_the_reference_surface():
    for (cls, name), args in EXPECTED.items():
        params = [p for p in inspect.signature(getattr(cls, name)).parameters if p != 'self']
        assert params[:len(args)] == args, (cls.__name__, name, params)
    assert inspect.signature(GraspingSimulator.links_in_collision).parameters['threshold'].default == -0.001
    assert inspect.signature(FrankaRobot.inverse_kinematics).parameters['combined_threshold'].default == 0.015
    assert np.array_equal(FrankaRobot.tf_grasp2ee, [[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]])


@pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, 'simulation.py')), reason='reference checkout not present')
def test_signatures_against_reference_source():
    """When the reference checkout is mounted: every method of EXPECTED exists there with the same argument names."""
    mods = {'Scenario': 'scenario.py', 'GraspingSimulator': 'simulation.py', 'FrankaRobot': 'simulation.py'}
    for (cls, name), args in EXPECTED.items():
        tree = ast.parse(open(os.path.join(REFERENCE, mods[cls.__name__])).read())
        cdef = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls.__name__)
        fdef = next((n for n in cdef.body if isinstance(n, ast.FunctionDef) and n.name == name), None)
        if fdef is None:
            assert cls is GraspingSimulator and name == '__init__' or name in ('add_scene',), (cls.__name__, name)
            continue
        ref_args = [a.arg for a in fdef.args.args if a.arg != 'self']
        assert ref_args == args, (cls.__name__, name, ref_args)


def test_scenario_loading_and_grasp_selection():
    s = Scenario(11)
    assert s.id == 11 and s.with_platform and s.select_indices is None
    assert np.array_equal(s.robot_pose, Scenario.default_robot_pose())
    assert len(s.gs) == 200 and s.grasp_poses.shape == (200, 4, 4)
    assert len(s.scene.objects) == 1 and len(s.scene.bg_objects) == 2
    s.select_n_grasps(10, seed=3)
    idx = np.random.default_rng(3).choice(200, 10, replace=False)        # scenario.py:56-57
    assert np.array_equal(s.select_indices, idx) and s.grasp_poses.shape == (10, 4, 4)
    s.select_n_grasps()
    assert s.select_indices is None
    with pytest.raises(ValueError):
        s.select_n_grasps(201)
    with pytest.raises(FileNotFoundError):
        Scenario(99)


def test_grasps_are_converted_to_the_franka_frame(golden):
    """scenario.py:31 + create_cpp_export.py:60-72 (G1) through the drop-in class."""
    s = Scenario(21)
    mine = np.linalg.inv(s.robot_pose) @ s.gs.poses
    assert np.abs(mine.reshape(200, 16) - golden['grasps_csv_021']).max() < 1e-6


def test_robot_state_and_error_conventions():
    s = Scenario(34)
    robot, sim = s.get_robot_and_sim()
    assert isinstance(robot, FrankaRobot) and isinstance(sim, GraspingSimulator)
    assert robot.end_effector_id == 14 and robot.arm_joint_ids == [0, 1, 3, 4, 5, 6, 7, 8, 9]
    assert robot.finger_joint_ids == [12, 13] and robot.home_conf[:2] == [0.0, 0.0] and len(robot.home_conf) == 9
    assert np.allclose(robot.arm_joints_pos(), robot.home_conf)
    assert np.allclose(robot.joint_pos()[-2:], 0.04) and len(robot.joint_pos()) == 11
    lim = robot.arm_joint_limits()
    assert lim.shape == (9, 2) and np.allclose(lim[0], [-1, 1]) and np.allclose(lim[7], [-0.0873, 3.8223])
    with pytest.raises(AssertionError):
        robot.reset_arm_joints([0.0] * 7)              # simulation.py:210
    with pytest.raises(ValueError):
        robot.forward_kinematics([0.0] * 5)            # simulation.py:237
    with pytest.raises(AssertionError):
        robot.reset_gripper(1.5)                       # simulation.py:314
    with pytest.raises(DeprecationWarning):
        robot.get_ee_pose_for_grasp(None)              # simulation.py:189
    with pytest.raises(NotImplementedError):
        robot.move_to([0] * 9)
    assert 'plane' in sim._env_bodies and len(sim._env_bodies) == 3 and len(sim._moving_bodies) == 1
    assert robot.body_id not in list(sim._env_bodies.values()) + list(sim._moving_bodies.values())
    robot.reset_arm_joints()
    assert robot.get_arm_joint_conf_from_motor_joint_conf(list(range(11))).tolist() == list(range(9))
    r2, _ = Scenario(34, with_platform=False).get_robot_and_sim()
    assert r2.end_effector_id == 11 and r2.arm_joint_ids == [0, 1, 2, 3, 4, 5, 6] and len(r2.home_conf) == 7
    assert r2.model.self_pairs.shape == (25, 2)


def test_trajectory_planner_ptp_profile():
    """simulation.py:480-535: synchronised trapezoids -- all joints arrive together, monotone, velocity and
    acceleration limits respected, dt spacing, degenerate cases."""
    from jogramop_framework_b200.simulation import TrajectoryPlanner
    q0 = np.array([0.0, 0.0, 0.1, -0.5, 0.0, -2.0, 0.0, 1.5, 0.7])
    q1 = np.array([0.5, -0.2, 0.1, 0.4, 1.0, -1.0, -0.3, 1.5, 2.0])
    tr = TrajectoryPlanner.ptp(q0, q1)
    wp, t = tr.joint_pos, tr.time_steps
    assert np.allclose(wp[0], q0) and np.allclose(wp[-1], q1)
    d_max = np.abs(q1 - q0).max()
    T = (d_max * 1.5 + 0.7 ** 2) / (0.7 * 1.5)
    assert t[-1] == pytest.approx(T) and len(t) == int(T // (1 / 20)) + 1
    prog = (wp - q0) * np.sign(q1 - q0)
    assert (np.diff(prog, axis=0) >= -1e-12).all()                      # monotone towards the target
    vel = np.abs(np.diff(wp, axis=0)) / np.diff(t)[:, None]
    assert vel.max() <= 0.7 + 1e-9
    assert np.allclose(wp[:, 2], 0.1) and np.allclose(wp[:, 7], 1.5)    # joints that do not move stay put
    # the slowest joint cruises at v_max in the middle of the motion
    j = int(np.argmax(np.abs(q1 - q0)))
    assert vel[len(vel) // 2, j] == pytest.approx(0.7, rel=1e-6)
    short = TrajectoryPlanner.ptp(q0, q0 + 0.01)
    assert np.allclose(short.joint_pos[-1], q0 + 0.01) and len(short) >= 2
    same = TrajectoryPlanner.ptp(q0, q0)
    assert len(same) == 2 and np.allclose(same.joint_pos, q0)
    steps = list(tr)
    assert len(steps) == len(tr) and len(steps[0]) == 4


def test_timer_and_quaternion_helpers():
    t = util.Timer()
    t.start('a'); t.stop('a'); t.count('x'); t.count('x')
    assert t.counters['x'] == 2 and t.timers['a'] >= 0
    with pytest.raises(ValueError):
        t.stop('never started')
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        R = util.rotation_matrix_from_quaternion(q)
        q2 = util.quaternion_from_rotation_matrix(R)
        assert util.angle_between_quaternions(q, q2) < 1e-6
        pos, quat = util.position_and_quaternion_from_tf(util.tf_from_pos_quat([1, 2, 3], q))
        assert np.allclose(pos, [1, 2, 3]) and util.angle_between_quaternions(q, quat) < 1e-6
    assert util.angle_between_quaternions([0, 0, 0, 1], [0, 0, np.sin(0.25), np.cos(0.25)], as_degree=True) == \
        pytest.approx(np.degrees(0.5))
    assert util.SCENARIO_IDS[0] == 11 and len(util.SCENARIO_IDS) == 20


def test_retarget_matches_rebuild(scene_loader):
    """Moving the robot re-expresses the baked robot-frame scene exactly like rebuilding it from the world poses."""
    m = scene_loader(11)
    pose = ingest.default_robot_pose()
    pose[:3, 3] += [0.1, -0.2, 0.0]
    r = m.retarget(pose)
    back = r.retarget(m.robot_pose)
    assert np.abs(back.piece_verts - m.piece_verts).max() < 1e-12
    assert np.allclose(r.plane, [0, 0, 1, 0.05])
    sub = m.subset([1], with_plane=False)
    assert sub.n_pieces == 4 and not sub.has_plane


# ------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_single_shot_queries_match_oracle(robot, scene_loader, golden):
    import oracle
    from ref_numpy import uniform_configs
    s = Scenario(11)
    rob, sim = s.get_robot_and_sim()
    o = oracle.Oracle(robot, scene_loader(11))
    # home pose: free (G5); golden IK row: free (G4), through the reference's own call sequence (README usage)
    assert not rob.in_collision() and not rob.in_self_collision()
    rob.reset_arm_joints(golden['ik_011'][0])
    assert not rob.in_self_collision() and not rob.in_collision()
    # FK: world pose of panda_grasptarget (simulation.py:223-249)
    pos, quat = rob.forward_kinematics(golden['ik_011'][0])
    T = o.fk(golden['ik_011'][:1])[0, 14]
    Tw = s.robot_pose[:3, :3] @ T[:, :3], s.robot_pose[:3, :3] @ T[:, 3] + s.robot_pose[:3, 3]
    assert np.abs(pos - Tw[1]).max() < 1e-5
    assert util.angle_between_quaternions(quat, util.quaternion_from_rotation_matrix(Tw[0])) < 1e-4
    M = rob.forward_kinematics(list(golden['ik_011'][0]) + [0.04, 0.04], as_matrix=True)     # len 11 accepted
    assert M.shape == (4, 4) and np.abs(M[:3, 3] - Tw[1]).max() < 1e-5
    q14 = np.zeros(14)
    q14[rob.arm_joint_ids] = golden['ik_011'][0]
    assert np.abs(rob.forward_kinematics(q14)[0] - Tw[1]).max() < 1e-5                        # len 14 accepted
    # random configurations: stateful single-shot calls == oracle verdicts (outside the 1 mm band)
    q = uniform_configs(robot, 60, 4).astype(np.float64)
    rb, rd = o.check(q, flags=oracle.SCENE | oracle.PLANE)
    sb, sd = o.check(q, flags=oracle.SELF)
    pairs = o.self_distances
    for i in range(len(q)):
        rob.reset_arm_joints(q[i])
        if abs(rd[i] + 0.001) > 1e-3:
            assert rob.in_collision() == bool(rb[i])
            assert sim.body_in_collision(rob.body_id) == bool(rb[i])       # the reference's per-body walk: same verdict
        if abs(sd[i] + 0.001) > 1e-3:
            assert rob.in_self_collision() == bool(sb[i])
    # one link pair (links_in_collision, simulation.py:82-95)
    rob.reset_arm_joints(q[0])
    d = pairs(q[0])
    k = int(np.argmin(d))
    la, lb = robot.self_pairs[k]
    if np.isfinite(d[k]) and abs(d[k] + 0.001) > 1e-3:
        assert sim.links_in_collision(rob.body_id, la, rob.body_id, lb) == bool(d[k] < -0.001)
    sim.dismiss()


@pytest.mark.gpu
def test_batched_variants_and_filter(robot, scene_loader, golden):
    import oracle
    import torch
    from ref_numpy import uniform_configs
    s = Scenario(34)
    rob, sim = s.get_robot_and_sim()
    o = oracle.Oracle(robot, scene_loader(34))
    q = uniform_configs(robot, 5000, 8)
    state = list(rob.arm_joints_pos())
    bits, dist = rob.in_collision_batch(q)
    rb, rd = o.check(q.astype(np.float64))
    assert np.abs(dist.cpu().numpy() - rd).max() < 1e-4
    assert rob.arm_joints_pos().tolist() == state             # batched calls do not touch the robot state
    pos, quat = rob.forward_kinematics_batch(q[:50])
    assert pos.shape == (50, 3) and quat.shape == (50, 4)
    allT = rob.forward_kinematics_batch(q[:50], all_links=True)
    assert allT.shape == (50, 15, 4, 4)
    ref = o.fk(q[:50].astype(np.float64))
    W = np.einsum('ij,nljk->nlik', s.robot_pose[:3, :3], ref)
    W[..., 3] += s.robot_pose[:3, 3]
    assert np.abs(allT[:, :, :3, :].cpu().numpy() - W).max() < 1e-5
    # the get_ik_solutions filter as one call: golden rows all pass, random rows get the oracle's reason codes
    accepted, reason, count = s.filter_configurations(golden['ik_034'], rob)
    assert len(accepted) == 15 and count.counters == {'IK solution is OK': 15}
    _, _, rr = o.check(q.astype(np.float64), flags=15, want_reason=True)
    _, reason, count = s.filter_configurations(q, rob)
    assert (reason == rr).mean() > 0.995
    # gripper poses in the WORLD frame == hand+fingers of the arm at the same FK pose
    T = rob.forward_kinematics_batch(q[:500], as_matrix=True).cpu().numpy()
    gb, gd = sim.gripper_in_collision_batch(T)
    for i in range(0, 500, 25):
        pd = o.pair_distances(q[i].astype(np.float64))
        assert abs(float(gd[i]) - min(pd[9:12].min(), 0.01)) < 1e-4
    be, de = rob.edges_in_collision_batch(q[:100], q[100:200], substeps=8)
    assert be.shape == (100,) and de.dtype == torch.float32
    # the batched queries' engine may spread a call of several passes over 4 arenas; per-body / per-pair helpers keep one
    full = sim._engine('full')
    assert sim.lanes == 4 and full._L.jg_engine_lanes(full._e) == 4
    # swept validation of a point-to-point trajectory (simulation.py:480-535 waypoints) with one batched check
    from jogramop_framework_b200.simulation import TrajectoryPlanner
    tr = TrajectoryPlanner.ptp(np.asarray(rob.home_conf), golden['ik_034'][0])
    hit, first, dist = rob.trajectory_in_collision(tr)
    rb2, rd2 = o.check(tr.joint_pos, flags=7)
    assert np.abs(dist - rd2).max() < 1e-4 and (first == -1) == (not rb2[np.abs(rd2 + 0.001) > 1e-3].any() and not hit)
    sim.dismiss()
