""""Next" rows: export writers (create_cpp_export.py:69-82) against the golden export files, and the
batched inverse kinematics (simulation.py:251-310) through its own acceptance criterion."""
import warnings

import numpy as np
import pytest

from jogramop_framework_b200 import export, ingest
from jogramop_framework_b200.scenario import Scenario
from ref_numpy import uniform_configs


@pytest.mark.parametrize('sid', [11, 34, 45])
def test_export_files_match_golden(sid, tmp_path, golden):
    s = Scenario(sid)
    ik = golden[f'ik_{sid:03d}']
    out = export.export_for_cpp(s, out_dir=str(tmp_path), ik_solutions=ik)
    # obstacles.obj: same triangles in the same order, 6 significant digits (G2)
    (v, f), = ingest.read_obj_shapes(f'{out}/obstacles.obj')
    head = open(f'{out}/obstacles.obj').read().split('\n')[:4]
    assert head[0].startswith('# Created by Open3D') and head[3] == f'# number of triangles: {len(f)}'
    assert np.abs(v[f] - golden[f'obstacles_tris_{sid:03d}']).max() < 1.1e-5
    # grasps.csv (G1), robot_start_conf.csv (G5), grasp_IK_solutions.csv round trip
    assert np.abs(np.loadtxt(f'{out}/grasps.csv', delimiter=',') - golden[f'grasps_csv_{sid:03d}']).max() < 1e-6
    assert np.array_equal(np.loadtxt(f'{out}/robot_start_conf.csv', delimiter=','), golden[f'start_conf_{sid:03d}'])
    assert np.array_equal(np.loadtxt(f'{out}/grasp_IK_solutions.csv', delimiter=',', ndmin=2), ik)


def test_export_empty_ik_file(tmp_path):
    """Scenarios without accepted IK rows produce an empty csv (013, 015, 041 in the reference)."""
    out = export.export_for_cpp(Scenario(13), out_dir=str(tmp_path), ik_solutions=np.zeros((0, 9)))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        assert np.loadtxt(f'{out}/grasp_IK_solutions.csv', delimiter=',', ndmin=2).size == 0


@pytest.mark.gpu
def test_ik_reaches_reachable_poses(robot):
    """Targets produced by FK of random in-limit configurations are reachable: the DLS IK must converge for most of
    them from the home pose, its error output must be the true error of its output configuration, and every returned
    configuration respects the joint limits."""
    import oracle
    s = Scenario(34)
    rob, sim = s.get_robot_and_sim()
    o = oracle.Oracle(robot)
    q_true = uniform_configs(robot, 4000, 3)
    T = rob.forward_kinematics_batch(q_true, as_matrix=True).cpu().numpy().astype(np.float64)
    q, err = rob.inverse_kinematics_batch(T)
    q, err = q.cpu().numpy().astype(np.float64), err.cpu().numpy()
    lim = rob.arm_joint_limits()
    assert (q >= lim[:, 0] - 1e-6).all() and (q <= lim[:, 1] + 1e-6).all()
    ok = err <= 0.015                                   # the reference's acceptance test (simulation.py:306)
    assert ok.mean() > 0.6, ok.mean()
    # err_out == error recomputed with the double-precision FK of the oracle
    Tr = o.fk(q)[:, robot.end_effector_id]
    Tw = np.einsum('ij,njk->nik', s.robot_pose[:3, :3], Tr)
    Tw[:, :, 3] += s.robot_pose[:3, 3]
    perr = np.linalg.norm(Tw[:, :, 3] - T[:, :3, 3], axis=1)
    Rrel = np.einsum('nij,nkj->nik', T[:, :3, :3], Tw[:, :, :3])
    ang = np.degrees(np.arccos(np.clip((np.trace(Rrel, axis1=1, axis2=2) - 1) / 2, -1, 1)))
    assert np.abs((perr + ang / 1000) - err)[ok].max() < 2e-4
    # restarts from the true configuration converge immediately
    q2, err2 = rob.inverse_kinematics_batch(T, q_init=q_true)
    assert (err2.cpu().numpy() < 1e-3).mean() > 0.999
    # position-only IK (orn=None in the reference)
    rob.reset_arm_joints()
    sol = rob.inverse_kinematics(T[0, :3, 3], None)
    if sol is not None:
        assert np.linalg.norm(rob.forward_kinematics(sol)[0] - T[0, :3, 3]) <= 0.015
    sim.dismiss()


@pytest.mark.gpu
def test_get_ik_solutions_end_to_end(robot, golden, scene_loader):
    """scenario.py:82-129 end to end on the GPU (soft target G6: the reference accepted 15 of 200 grasps in 034 with
    pyBullet's IK, which is not reproducible): every returned row passes the oracle's limits + self + scene check and
    reaches one of the grasps within the acceptance threshold."""
    import oracle
    s = Scenario(34)
    sols = s.get_ik_solutions()
    o = oracle.Oracle(robot, scene_loader(34))
    if len(sols):
        bits, dist = o.check(sols, flags=7)
        assert (np.abs(dist[bits] + 0.001) < 1e-3).all()
        T = o.fk(sols)[:, robot.end_effector_id]
        gr = golden['grasps_csv_034'].reshape(-1, 4, 4)
        for t in T:
            pe = np.linalg.norm(gr[:, :3, 3] - t[:, 3], axis=1)
            Rrel = np.einsum('nij,kj->nik', gr[:, :3, :3], t[:, :3])
            ang = np.degrees(np.arccos(np.clip((np.trace(Rrel, axis1=1, axis2=2) - 1) / 2, -1, 1)))
            assert (pe + ang / 1000).min() <= 0.0151
    sols_b, gi = s.get_ik_solutions_batch(restarts=8)
    print(f'scenario 034: sequential loop {len(sols)} solutions, batched x8 restarts {len(sols_b)} rows for '
          f'{len(set(gi.tolist()))} grasps (reference with pyBullet IK: 15)')
    if len(sols_b):      # float32 rows on the GPU vs float64 here: disagreements only inside the 1 mm band
        # (limits are checked separately: rows clamped onto a limit are float32 images of it)
        bits, dist = o.check(sols_b.astype(np.float32).astype(np.float64), flags=7)
        assert (np.abs(dist[bits] + 0.001) < 1e-3).all()
        assert bits.mean() < 0.05
        lim = robot.joint_limits[robot.arm_joint_ids]
        assert (sols_b >= lim[:, 0] - 1e-6).all() and (sols_b <= lim[:, 1] + 1e-6).all()


REF_G6 = {11: 1, 12: 3, 13: 0, 14: 1, 15: 0, 21: 1, 22: 4, 23: 2, 24: 1, 25: 1, 31: 10, 32: 11, 33: 15, 34: 15, 35: 8,
          41: 0, 42: 5, 43: 2, 44: 5, 45: 4}    # rows of scenarios/*/export/grasp_IK_solutions.csv


def _reached(o, robot, rows, grasps_robot):
    T = o.fk(rows)[:, robot.end_effector_id]
    pe = np.linalg.norm(grasps_robot[None, :, :3, 3] - T[:, None, :, 3], axis=2)
    Rrel = np.einsum('gij,nkj->ngik', grasps_robot[:, :3, :3], T[:, :, :3])
    ang = np.degrees(np.arccos(np.clip((np.trace(Rrel, axis1=2, axis2=3) - 1) / 2, -1, 1)))
    e = pe + ang / 1000
    return e.argmin(1), e.min(1), e


@pytest.mark.gpu
def test_ik_outcomes_against_the_references_recorded_outcomes_g6(robot, golden, scene_loader):
    """G6 as a pin (simulation.py:251-310 + scenario.py:82-129): for all 20 scenarios the batched IK + filter must
    (a) return only rows that the oracle accepts (limits, self-collision, scene) and that reach one of the 200 grasps
    within the reference's acceptance threshold, and (b) accept at least as many distinct grasps as the reference's
    grasp_IK_solutions.csv holds; most of the very grasps the reference reached are recovered as well."""
    import oracle
    found = n_ref = 0
    for sid, n_rows in REF_G6.items():
        s = Scenario(sid)
        o = oracle.Oracle(robot, scene_loader(sid))
        gr = np.linalg.inv(s.robot_pose) @ s.gs.poses
        ref_rows = golden[f'ik_{sid:03d}'].reshape(-1, 9)
        assert len(ref_rows) == n_rows
        ref_g = set(_reached(o, robot, ref_rows, gr)[0].tolist()) if n_rows else set()
        q, gi = s.get_ik_solutions_batch(restarts=256)
        got = set(gi.tolist())
        if len(q):
            rows = q.astype(np.float32).astype(np.float64)
            bits, dist = o.check(rows, flags=7)
            assert (np.abs(dist[bits] + 0.001) < 1e-3).all() and bits.mean() < 0.05     # fp32 vs fp64 only inside the band
            lim = robot.joint_limits[robot.arm_joint_ids]
            assert (q >= lim[:, 0] - 1e-6).all() and (q <= lim[:, 1] + 1e-6).all()
            e_target = _reached(o, robot, rows, gr)[2][np.arange(len(rows)), gi]    # error to the grasp each row was solved for
            assert (e_target <= 0.0152).all()
        print(f'{sid:03d}: reference {n_rows} rows / {len(ref_g)} grasps; batched x256: {len(q)} rows / {len(got)} grasps, '
              f'{len(ref_g & got)} of the reference\'s grasps')
        assert len(got) >= len(ref_g), sid
        found += len(ref_g & got)
        n_ref += len(ref_g)
    assert n_ref == 89 and found >= 80, found
