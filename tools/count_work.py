#!/usr/bin/env python
"""Counted work of one pass with the INSTRUMENTED library (libjogramop_b200_count.so, -DJG_COUNT_WORK): executed support
dot products, GJK iterations and polytope expansions per check -> the counted flops-per-check figure that audits the
nominal one of the roofline.  Run by bench.py in a subprocess, outside every timed region.

  JG_B200_LIB_VARIANT=count python tools/count_work.py [--scenario 11] [--n 1000000] [--seed 11] [--flags 3]
"""
import argparse
import ctypes as C
import json
import os
import sys

os.environ['JG_B200_LIB_VARIANT'] = 'count'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from jogramop_framework_b200 import _lib  # noqa: E402
from jogramop_framework_b200.engine import CollisionEngine  # noqa: E402
from jogramop_framework_b200.ingest import RobotModel, SceneModel  # noqa: E402

# flops per counted unit: a support dot product = 3 multiply-adds (5 flops) + 1 compare; one GJK iteration's simplex
# update and one polytope expansion (closest face, visibility tests, new faces) ~100 / ~150 flops (SURVEY §8d's constants)
FLOPS_PER_DOT, FLOPS_PER_GJK_ITER, FLOPS_PER_EPA_ITER = 6, 100, 150


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scenario', type=int, default=11)
    ap.add_argument('--n', type=int, default=1_000_000)
    ap.add_argument('--seed', type=int, default=11)
    ap.add_argument('--flags', type=int, default=3)
    a = ap.parse_args()
    data = os.path.join(ROOT, 'jogramop_framework_b200', 'data')
    robot = RobotModel.load(os.path.join(data, 'robot_mobile_panda.npz'))
    scene = SceneModel.load(os.path.join(data, f'scene_{a.scenario:03d}.npz'))
    lim = robot.joint_limits[robot.arm_joint_ids]
    q = (lim[:, 0] + (lim[:, 1] - lim[:, 0]) * np.random.default_rng(a.seed).random((a.n, 9))).astype(np.float32)
    eng = CollisionEngine(robot, scene, device=0)
    qd = torch.from_numpy(q).cuda()
    L = _lib.load()
    _lib.check(L.jg_work_counters_reset(0), 'jg_work_counters_reset')
    eng.check(qd, flags=a.flags, packed=True)
    out = (C.c_uint64 * 4)()
    _lib.check(L.jg_work_counters(0, out), 'jg_work_counters')
    dots, gjk_it, epa_it = int(out[0]), int(out[1]), int(out[2])
    S, P = robot.n_shapes, scene.n_pieces
    fk = 15 * 63 + S * 33 + S * (P + 1) * 6       # FK + shape boxes + box tests (deterministic, SURVEY §8d W_fk + W_bv + W_bp)
    narrow = (dots * FLOPS_PER_DOT + gjk_it * FLOPS_PER_GJK_ITER + epa_it * FLOPS_PER_EPA_ITER) / a.n
    print(json.dumps({'n': a.n, 'scenario': a.scenario, 'flags': a.flags,
                      'support_dots_per_check': dots / a.n, 'gjk_iterations_per_check': gjk_it / a.n,
                      'polytope_expansions_per_check': epa_it / a.n,
                      'counted_narrow_flops_per_check': narrow, 'broad_flops_per_check': fk,
                      'flops_per_unit': {'dot': FLOPS_PER_DOT, 'gjk_iter': FLOPS_PER_GJK_ITER, 'epa_iter': FLOPS_PER_EPA_ITER}}))
    eng.close()


if __name__ == '__main__':
    main()
