#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on ONE GPU (bench.py measures the headline config 2):
  config 3  every scenario, 1 M uniform configurations each (seed = scenario id), bit + distance
  config 4  scenario 034: 1 M edges x 64 substeps (seeds 34 / 3400), early-exit rate reported
  config 5  scenario 011: 10 M free-floating gripper poses (seed 5011), bit + distance
  extra     self-collision + limits + scene (the full get_ik_solutions filter), mesh mode (LBVH)
Prints one JSON line per measurement.  CUDA events, 3 warm-up + 5 timed repetitions each."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jogramop_framework_b200.engine import CollisionEngine, SCENE, PLANE, SELF, LIMITS  # noqa: E402
from jogramop_framework_b200.ingest import RobotModel, SceneModel  # noqa: E402

DATA = os.path.join(ROOT, 'jogramop_framework_b200', 'data')
IDS = [11, 12, 13, 14, 15, 21, 22, 23, 24, 25, 31, 32, 33, 34, 35, 41, 42, 43, 44, 45]


def timed(fn, reps=5, warm=3):
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


def uniform(robot, n, seed):
    lim = robot.joint_limits[robot.arm_joint_ids]
    return (lim[:, 0] + (lim[:, 1] - lim[:, 0]) * np.random.default_rng(seed).random((n, 9))).astype(np.float32)


def main():
    robot = RobotModel.load(os.path.join(DATA, 'robot_mobile_panda.npz'))
    lim = robot.joint_limits[robot.arm_joint_ids]
    n = 1_000_000
    for sid in IDS:
        scene = SceneModel.load(os.path.join(DATA, f'scene_{sid:03d}.npz'))
        eng = CollisionEngine(robot, scene)
        q = torch.from_numpy(uniform(robot, n, sid)).cuda()
        ms = timed(lambda: eng.check(q, flags=SCENE | PLANE, packed=True))
        bits, _, _ = eng.check(q, flags=SCENE | PLANE)
        print(json.dumps({'config': 3, 'scenario': sid, 'n': n, 'ms': ms, 'checks_per_s': n / ms * 1e3,
                          'collision_rate': bits.float().mean().item(), 'pieces': scene.n_pieces}), flush=True)
        if sid == 11:
            ms = timed(lambda: eng.check(q, flags=SCENE | PLANE | SELF | LIMITS, packed=True, want_reason=True))
            print(json.dumps({'config': 'filter (limits+self+scene)', 'scenario': sid, 'n': n, 'ms': ms,
                              'checks_per_s': n / ms * 1e3}), flush=True)
            ms = timed(lambda: eng.check(q, flags=SCENE | PLANE | 16, packed=True))
            print(json.dumps({'config': 'verdict + clearance only (JG_NO_EPA)', 'scenario': sid, 'n': n, 'ms': ms,
                              'checks_per_s': n / ms * 1e3}), flush=True)
            ms = timed(lambda: eng.check(q, flags=SCENE | PLANE, packed=True, want_dist=False))
            print(json.dumps({'config': 'verdict bits only (min_dist = NULL)', 'scenario': sid, 'n': n, 'ms': ms,
                              'checks_per_s': n / ms * 1e3}), flush=True)
            ms = timed(lambda: eng.check(q, flags=SCENE | PLANE | SELF | LIMITS, packed=True, want_dist=False, want_reason=True))
            print(json.dumps({'config': 'filter (limits+self+scene), verdict bits + reason only', 'scenario': sid, 'n': n,
                              'ms': ms, 'checks_per_s': n / ms * 1e3}), flush=True)
            # config 5: free-floating gripper poses
            m = 10_000_000
            rng = np.random.default_rng(5000 + sid)
            tgt = scene.piece(0)
            lo, hi = tgt.min(0) - 0.25, tgt.max(0) + 0.25
            pos = (lo + (hi - lo) * rng.random((m, 3))).astype(np.float32)
            quat = rng.normal(size=(m, 4)).astype(np.float32)
            pq = torch.from_numpy(np.concatenate([pos, quat], 1)).cuda()
            ms = timed(lambda: eng.check_poses(pq, links=[11, 12, 13], packed=True), reps=3, warm=2)
            bits, _ = eng.check_poses(pq, links=[11, 12, 13])
            print(json.dumps({'config': 5, 'scenario': sid, 'n': m, 'ms': ms, 'checks_per_s': m / ms * 1e3,
                              'collision_rate': bits.float().mean().item()}), flush=True)
            # mesh mode (LBVH over export/obstacles.obj)
            em = CollisionEngine(robot, scene, mesh_mode=True)
            ms = timed(lambda: em.check(q, flags=SCENE | PLANE | 16, packed=True))
            print(json.dumps({'config': 'mesh mode (LBVH, 254 triangles)', 'scenario': sid, 'n': n, 'ms': ms,
                              'checks_per_s': n / ms * 1e3}), flush=True)
            em.close()
        if sid == 34:
            # config 4: edges
            m = 1_000_000
            qa = uniform(robot, m, 34)
            qb = np.clip(qa + np.random.default_rng(3400).uniform(-0.25, 0.25, qa.shape), lim[:, 0], lim[:, 1]).astype(np.float32)
            qa_d, qb_d = torch.from_numpy(qa).cuda(), torch.from_numpy(qb).cuda()
            ms = timed(lambda: eng.check_edges(qa_d, qb_d, substeps=64, packed=True), reps=3, warm=2)
            bits, _ = eng.check_edges(qa_d, qb_d, substeps=64)
            print(json.dumps({'config': 4, 'scenario': sid, 'edges': m, 'substeps': 64, 'ms': ms,
                              'checks_per_s': m * 64 / ms * 1e3, 'edges_per_s': m / ms * 1e3,
                              'edge_collision_rate': bits.float().mean().item()}), flush=True)
            ms = timed(lambda: eng.check_edges(qa_d, qb_d, substeps=64, packed=True, want_dist=False), reps=3, warm=2)
            print(json.dumps({'config': '4, verdict bits only (what an RRT edge validation needs)', 'scenario': sid, 'edges': m,
                              'substeps': 64, 'ms': ms, 'checks_per_s': m * 64 / ms * 1e3, 'edges_per_s': m / ms * 1e3}), flush=True)
        eng.close()


if __name__ == '__main__':
    main()
