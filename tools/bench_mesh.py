#!/usr/bin/env python
"""Mesh-mode throughput (LBVH over a triangle mesh, group-cooperative traversal) on one GPU:
  robot vs export/obstacles.obj (scenario 011 / 045, 1 M configurations; verdict + clamped distance, and verdict only),
  Panda hand hulls and the burg gripper vs obstacles.obj / the visual meshes (2 M poses).  JSON lines."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from jogramop_framework_b200.engine import CollisionEngine, SCENE, PLANE, NO_EPA  # noqa: E402
from jogramop_framework_b200.graspset import burg_gripper_model  # noqa: E402
from jogramop_framework_b200.ingest import RobotModel, SceneModel  # noqa: E402
from jogramop_framework_b200.scenario import Scenario  # noqa: E402

DATA = os.path.join(ROOT, 'jogramop_framework_b200', 'data')


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    robot = RobotModel.load(os.path.join(DATA, 'robot_mobile_panda.npz'))
    lim = robot.joint_limits[robot.arm_joint_ids]
    n = 1_000_000
    for sid in (11, 45):
        scene = SceneModel.load(os.path.join(DATA, f'scene_{sid:03d}.npz'))
        q = torch.from_numpy((lim[:, 0] + (lim[:, 1] - lim[:, 0]) * np.random.default_rng(sid).random((n, 9))).astype(np.float32)).cuda()
        em = CollisionEngine(robot, scene, mesh_mode=True)
        ms = timed(lambda: em.check(q, flags=SCENE | PLANE | NO_EPA, packed=True))
        st = em.stats()
        print(json.dumps({'case': f'robot vs obstacles.obj, scenario {sid:03d} ({len(scene.tris)} triangles), bit + clamped distance', 'n': n,
                          'ms': ms, 'checks_per_s': n / ms * 1e3, 'triangle_candidates_per_check': st['pairs'] / n,
                          'gjk_run_per_check': st['gjk_run'] / n, 'launches': st['kernel_launches']}), flush=True)
        ms = timed(lambda: em.check(q, flags=SCENE | PLANE, packed=True, want_dist=False))
        print(json.dumps({'case': f'robot vs obstacles.obj, scenario {sid:03d}, verdict only', 'n': n, 'ms': ms, 'checks_per_s': n / ms * 1e3}), flush=True)
        em.close()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        s = Scenario(11)
    scene = s._scene_model
    vis = s.visual_scene_model()
    m = 2_000_000
    rng = np.random.default_rng(5011)
    tgt = scene.piece(0)
    lo, hi = tgt.min(0) - 0.25, tgt.max(0) + 0.25
    pq = np.concatenate([(lo + (hi - lo) * rng.random((m, 3))), rng.normal(size=(m, 4))], 1).astype(np.float32)
    pq_d = torch.from_numpy(pq).cuda()
    links = [robot.end_effector_id - 3, *robot.finger_joint_ids]
    for name, model, sc, lk, ee, thr in (('Panda hand hulls vs obstacles.obj, zero triangle margin (boolean via GJK)', robot, scene, links, None, -0.001),
                                         ('burg gripper boxes vs obstacles.obj (boolean, exact box-triangle tests)', burg_gripper_model(), scene, [0], 0, 1e-6),
                                         ('burg gripper boxes vs the visual meshes (the reference filter)', burg_gripper_model(), vis, [0], 0, 1e-6)):
        eng = CollisionEngine(model, sc, mesh_mode=True, mesh_margin=0.0, threshold=thr, finger_open=0.04 if model is robot else 0.0)
        if model is robot:
            fn = lambda: eng.check_poses(pq_d, links=lk, flags=SCENE | PLANE | NO_EPA, packed=True)   # noqa: E731
        else:
            fn = lambda: eng.check_poses(pq_d, links=lk, ee_link=ee, want_dist=False, packed=True)   # noqa: E731
        ms = timed(fn, reps=2, warm=1)
        st = eng.stats()
        print(json.dumps({'case': name, 'triangles': int(len(sc.tris)), 'poses': m, 'ms': ms, 'poses_per_s': m / ms * 1e3,
                          'triangle_candidates_per_pose': st['pairs'] / m, 'launches': st['kernel_launches']}), flush=True)
        eng.close()


if __name__ == '__main__':
    main()
