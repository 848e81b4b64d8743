#!/usr/bin/env python
"""IK outcome per scenario beside the reference's recorded outcome (rows of
``scenarios/*/export/grasp_IK_solutions.csv``, produced by ``Scenario.get_ik_solutions`` with pyBullet's IK).

    python tools/ik_g6_counts.py [restarts ...]   (GPU)   > profiles/r02_ik_g6_counts.txt

For every scenario: which grasps the reference's accepted rows reach (golden FK), how many rows / distinct grasps the
sequential drop-in loop (``get_ik_solutions``) and the batched variant (``get_ik_solutions_batch``) accept, and how
many of the reference's grasps the batched variant recovers.  Every accepted row is re-checked with the oracle.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from jogramop_framework_b200.ingest import RobotModel, SceneModel  # noqa: E402
from jogramop_framework_b200.scenario import Scenario  # noqa: E402
from jogramop_framework_b200.util import SCENARIO_IDS  # noqa: E402

REF = {11: 1, 12: 3, 13: 0, 14: 1, 15: 0, 21: 1, 22: 4, 23: 2, 24: 1, 25: 1, 31: 10, 32: 11, 33: 15, 34: 15, 35: 8,
       41: 0, 42: 5, 43: 2, 44: 5, 45: 4}


def reached_grasps(o, robot, rows, grasps_robot):
    """index of the grasp each configuration reaches (combined error pos + deg/1000), and that error"""
    if len(rows) == 0:
        return np.zeros(0, int), np.zeros(0)
    T = o.fk(rows)[:, robot.end_effector_id]
    pe = np.linalg.norm(grasps_robot[None, :, :3, 3] - T[:, None, :, 3], axis=2)
    Rrel = np.einsum('gij,nkj->ngik', grasps_robot[:, :3, :3], T[:, :, :3])
    ang = np.degrees(np.arccos(np.clip((np.trace(Rrel, axis1=2, axis2=3) - 1) / 2, -1, 1)))
    e = pe + ang / 1000
    return e.argmin(1), e.min(1)


def main():
    restarts = [int(x) for x in sys.argv[1:]] or [1, 8, 32]
    data = os.path.join(ROOT, 'jogramop_framework_b200', 'data')
    gold = np.load(os.path.join(ROOT, 'tests', 'golden', 'export_golden.npz'))
    robot = RobotModel.load(os.path.join(data, 'robot_mobile_panda.npz'))
    import contextlib
    import io
    print('scenario  reference rows/grasps | sequential loop rows/grasps (s) | ' +
          ' | '.join(f'batched x{r} rows/grasps/ref-grasps-found (s)' for r in restarts))
    tot = np.zeros(3 + 3 * len(restarts), int)
    for sid in SCENARIO_IDS:
        with contextlib.redirect_stdout(io.StringIO()):
            s = Scenario(sid)
        o = oracle.Oracle(robot, SceneModel.load(os.path.join(data, f'scene_{sid:03d}.npz')))
        gr = np.linalg.inv(s.robot_pose) @ s.gs.poses
        ref_rows = gold[f'ik_{sid:03d}'].reshape(-1, 9)
        ref_g, ref_e = reached_grasps(o, robot, ref_rows, gr)
        assert len(ref_rows) == REF[sid] and (ref_e <= 0.015).all()
        t0 = time.time()
        with contextlib.redirect_stdout(io.StringIO()):
            seq = s.get_ik_solutions()
        t_seq = time.time() - t0
        seq_g, seq_e = reached_grasps(o, robot, seq.reshape(-1, 9), gr)
        assert (seq_e <= 0.0151).all()
        if len(seq):
            b, d = o.check(seq.reshape(-1, 9), flags=15)
            assert (np.abs(d[b] + 0.001) < 1e-3).all()
        cols = [f'{sid:03d}      {len(ref_rows):3d}/{len(set(ref_g.tolist())):3d}',
                f'{len(seq):3d}/{len(set(seq_g.tolist())):3d} ({t_seq:.1f})']
        row = [len(ref_rows), len(seq), len(set(seq_g.tolist()))]
        for r in restarts:
            t0 = time.time()
            with contextlib.redirect_stdout(io.StringIO()):
                q, gi = s.get_ik_solutions_batch(restarts=r)
            t_b = time.time() - t0
            if len(q):
                b, d = o.check(q.astype(np.float32).astype(np.float64), flags=7)
                assert (np.abs(d[b] + 0.001) < 1e-3).all()
                g2, e2 = reached_grasps(o, robot, q.astype(np.float32).astype(np.float64), gr)
                assert (e2 <= 0.0152).all()
            found = len(set(ref_g.tolist()) & set(gi.tolist()))
            cols.append(f'{len(q):4d}/{len(set(gi.tolist())):3d}/{found:2d} of {len(set(ref_g.tolist())):2d} ({t_b:.2f})')
            row += [len(q), len(set(gi.tolist())), found]
        tot += np.array(row)
        print('  |  '.join(cols), flush=True)
    print('total   ', tot.tolist())


if __name__ == '__main__':
    main()
