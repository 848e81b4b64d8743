#!/usr/bin/env python
"""How the reference's grasp annotations (burg sampler + burg's procedural gripper mesh, not in the reference tree) fare
under this repo's gripper model (Panda hand + finger hulls on panda_grasptarget): which part collides, with what."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jogramop_framework_b200.scenario import Scenario  # noqa: E402
from jogramop_framework_b200.simulation import FrankaRobot  # noqa: E402

for sid in (11, 21, 34, 45):
    s = Scenario(sid)
    robot, sim = s.get_robot_and_sim()
    ee = s.gs.poses                                   # already in the end-effector convention
    T = (np.linalg.inv(robot.pose) @ ee)[:, :3, :].astype(np.float32)
    eng = sim._engine('full')
    hand, f1, f2 = robot.end_effector_id - 3, *robot.finger_joint_ids
    out = {}
    for name, links in (('all', [hand, f1, f2]), ('hand', [hand]), ('fingers', [f1, f2])):
        hit, d = eng.check_poses(T, links=links)
        out[name] = (float(hit.float().mean()), float(d.min()), float(d.median()))
    print(sid, {k: tuple(round(x, 4) for x in v) for k, v in out.items()})
