"""GPU-vs-oracle sweep of the other entry points (a script, not collected by pytest): for scenarios 011, 024, 034, 045 --
40 000 edges x 64 substeps (with and without distances), 300 000 free-floating gripper poses, 300 000 configurations
in the clamped JG_NO_EPA mode with all flags.   python tests/parity_sweep_paths.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import oracle
from oracle import Oracle
from ref_numpy import uniform_configs
from jogramop_framework_b200.engine import CollisionEngine
from jogramop_framework_b200.ingest import RobotModel, SceneModel
DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'jogramop_framework_b200', 'data')
robot = RobotModel.load(os.path.join(DATA, 'robot_mobile_panda.npz'))
lim = robot.joint_limits[robot.arm_joint_ids]
def rep(tag, gb, gd, rb, rd):
    err = np.abs(gd - rd); outside = np.abs(rd + 0.001) > 1e-3
    print(tag, 'max err %.3e' % err.max(), 'n>1e-4', int((err > 1e-4).sum()), 'mismatch outside', int((gb[outside] != rb[outside]).sum()), 'inside', int((gb[~outside] != rb[~outside]).sum()), 'rate %.3f' % rb.mean(), flush=True)
    for i in np.nonzero(err > 1e-4)[0][:3]: print('    idx', int(i), gd[i], rd[i], flush=True)
for sid in (11, 24, 34, 45):
    scene = SceneModel.load(os.path.join(DATA, f'scene_{sid:03d}.npz'))
    e = CollisionEngine(robot, scene); o = Oracle(robot, scene)
    # edges
    m = 40000
    qa = uniform_configs(robot, m, 8000 + sid)
    qb = np.clip(qa + np.random.default_rng(8100 + sid).uniform(-0.25, 0.25, qa.shape), lim[:, 0], lim[:, 1]).astype(np.float32)
    rb, rd = o.check_edges(qa.astype(np.float64), qb.astype(np.float64), substeps=64)
    b, d = e.check_edges(qa, qb, substeps=64)
    rep(f'{sid} edges {m}x64', b.cpu().numpy().astype(bool), d.cpu().numpy().astype(np.float64), rb, rd)
    b2, _ = e.check_edges(qa, qb, substeps=64, want_dist=False)
    outside = np.abs(rd + 0.001) > 1e-3
    print(f'{sid} edges bits-only mismatches outside band', int((b2.cpu().numpy().astype(bool)[outside] != rb[outside]).sum()), flush=True)
    # poses
    n = 300000
    rng = np.random.default_rng(8200 + sid)
    tgt = scene.piece(0); lo, hi = tgt.min(0) - 0.25, tgt.max(0) + 0.25
    pos = lo + (hi - lo) * rng.random((n, 3)); quat = rng.normal(size=(n, 4)); quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    x, y, z, w = quat.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).reshape(n, 3, 3)
    T = np.concatenate([R, pos[:, :, None]], 2)
    rb, rd = o.check_poses(T, links=[11, 12, 13], flags=3)
    b, d = e.check_poses(T.astype(np.float32), links=[11, 12, 13], flags=3)
    rep(f'{sid} poses {n}', b.cpu().numpy().astype(bool), d.cpu().numpy().astype(np.float64), rb, rd)
    # clamped mode
    q = uniform_configs(robot, 300000, 8300 + sid)
    rb, rd = o.check(q.astype(np.float64), flags=15 | oracle.NO_EPA)
    b, d, _ = e.check(q, flags=15 | 16)
    rep(f'{sid} NO_EPA 300000', b.cpu().numpy().astype(bool), d.cpu().numpy().astype(np.float64), rb, rd)
    e.close()
