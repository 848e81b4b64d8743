"""Large GPU-vs-oracle parity sweep (a script, not collected by pytest): every scenario, N configurations each (default
300 000), limits + self + scene + plane.  Prints per scenario: max |d - d_oracle|, distances off by more than 1e-4 m,
verdict mismatches outside / inside the 1 mm band.   python tests/parity_sweep.py [N]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import oracle
from oracle import Oracle
from ref_numpy import uniform_configs, SCENARIO_IDS
from jogramop_framework_b200.engine import CollisionEngine
from jogramop_framework_b200.ingest import RobotModel, SceneModel
DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'jogramop_framework_b200', 'data')
robot = RobotModel.load(os.path.join(DATA, 'robot_mobile_panda.npz'))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
worst = []
for sid in SCENARIO_IDS:
    scene = SceneModel.load(os.path.join(DATA, f'scene_{sid:03d}.npz'))
    e = CollisionEngine(robot, scene)
    o = Oracle(robot, scene)
    q = uniform_configs(robot, n, 7000 + sid)
    t0 = time.time()
    rb, rd = o.check(q.astype(np.float64), flags=15)
    t1 = time.time()
    bits, dist, _ = e.check(q, flags=15)
    gb, gd = bits.cpu().numpy().astype(bool), dist.cpu().numpy().astype(np.float64)
    err = np.abs(gd - rd)
    outside = np.abs(rd + 0.001) > 1e-3
    bad = np.nonzero(err > 1e-4)[0]
    print(sid, 'max err %.3e' % err.max(), 'n>1e-4', len(bad), 'verdict mismatches outside band', int((gb[outside] != rb[outside]).sum()), 'inside', int((gb[~outside] != rb[~outside]).sum()), 'oracle s %.1f' % (t1 - t0), flush=True)
    for i in bad[:3]:
        print('   idx', int(i), 'gpu', gd[i], 'ref', rd[i], 'q', q[i].tolist(), flush=True)
    e.close()
