#!/usr/bin/env python
"""Kernel-category split (jg_profile sections) of one 1 M pass of scenario 011 for the flag sets users run: scene + plane,
self-collision only, the full get_ik_solutions filter (limits + self + scene) with distances and verdict-only."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from jogramop_framework_b200.engine import CollisionEngine, SCENE, PLANE, SELF, LIMITS
from jogramop_framework_b200.ingest import RobotModel, SceneModel
DATA = os.path.join(ROOT, 'jogramop_framework_b200', 'data')
robot = RobotModel.load(os.path.join(DATA, 'robot_mobile_panda.npz'))
scene = SceneModel.load(os.path.join(DATA, 'scene_011.npz'))
lim = robot.joint_limits[robot.arm_joint_ids]
n=1_000_000
q = torch.from_numpy((lim[:, 0] + (lim[:, 1] - lim[:, 0]) * np.random.default_rng(11).random((n, 9))).astype(np.float32)).cuda()
eng = CollisionEngine(robot, scene)
CASES = [('scene',3,True),('self only',4,True),('filter',15,True),('filter bits',15,False)]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if c[0] == sys.argv[1]]
for name, fl, wd in CASES:
    for _ in range(3): eng.check(q, flags=fl, packed=True, want_dist=wd)
    torch.cuda.synchronize()
    eng.profile(True)
    for _ in range(5): eng.check(q, flags=fl, packed=True, want_dist=wd)
    p = eng.get_profile(); eng.profile(False)
    st = eng.stats()
    print(name, {k: round(v[0]/5,3) for k,v in p.items()}, 'total', round(sum(v[0] for v in p.values())/5,3), {k: st[k] for k in ('pairs','self_pairs','gjk_run','epa_run','overflow_run')})

if len(sys.argv) <= 1 or sys.argv[1] == 'mesh':
    em = CollisionEngine(robot, scene, mesh_mode=True)
    for _ in range(3): em.check(q, flags=3, packed=True)
    torch.cuda.synchronize()
    em.profile(True)
    for _ in range(5): em.check(q, flags=3, packed=True)
    p = em.get_profile(); em.profile(False)
    st = em.stats()
    print('mesh mode', {k: round(v[0]/5,3) for k,v in p.items()}, 'total', round(sum(v[0] for v in p.values())/5,3), {k: st[k] for k in ('pairs','gjk_run','kernel_launches')})

if len(sys.argv) <= 1 or sys.argv[1] == 'edges':
    sc34 = SceneModel.load(os.path.join(DATA, 'scene_034.npz'))
    ee = CollisionEngine(robot, sc34)
    m = 1 << 14
    rng = np.random.default_rng(34)
    qa = (lim[:, 0] + (lim[:, 1] - lim[:, 0]) * rng.random((m, 9))).astype(np.float32)
    qb = np.clip(qa + (np.random.default_rng(3400).random((m, 9)) - 0.5).astype(np.float32) * 0.5, lim[:, 0], lim[:, 1]).astype(np.float32)
    qa, qb = torch.from_numpy(qa).cuda(), torch.from_numpy(qb).cuda()
    for want in (True, False):
        for _ in range(3): ee.check_edges(qa, qb, substeps=64, packed=True, want_dist=want)
        torch.cuda.synchronize()
        ee.profile(True)
        for _ in range(5): ee.check_edges(qa, qb, substeps=64, packed=True, want_dist=want)
        p = ee.get_profile(); ee.profile(False)
        st = ee.stats()
        print('edges 2^14 x 64' + ('' if want else ' (bits only)'), {k: round(v[0]/5,3) for k,v in p.items()}, 'total', round(sum(v[0] for v in p.values())/5,3), {k: st[k] for k in ('pairs','gjk_run','epa_run','kernel_launches')})
