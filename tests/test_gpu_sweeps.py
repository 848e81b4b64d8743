"""GPU-vs-oracle parity sweeps over every scenario and every entry point (driver-run: ``pytest -m gpu``).

Round 1 ran these as hand-started scripts with 300 000 samples per case (profiles/r01_parity_sweep.txt); here they are
tests at 50 000 per scenario so that the driver's GPU run carries the evidence.  Bars (BASELINE.json north_star):
verdicts identical wherever the reference distance is more than 1 mm from the decision surface, |d - d_ref| <= 1e-4 m.
"""
import os

import numpy as np
import pytest

import oracle
from oracle import Oracle
from ref_numpy import SCENARIO_IDS, uniform_configs

pytestmark = pytest.mark.gpu

ALL = oracle.SCENE | oracle.PLANE | oracle.SELF | oracle.LIMITS
N = int(os.environ.get('JG_SWEEP_N', 50000))


def _cmp(tag, gb, gd, rb, rd, band=1e-3, tol=1e-4):
    gb, gd = gb.cpu().numpy().astype(bool), gd.cpu().numpy().astype(np.float64)
    err = np.abs(gd - rd)
    outside = np.abs(rd + 0.001) > band
    mism_out = int((gb[outside] != rb[outside]).sum())
    print(f'{tag}: max |d - d_oracle| {err.max():.2e}, verdict mismatches outside the band {mism_out}, inside '
          f'{int((gb[~outside] != rb[~outside]).sum())}, collision rate {rb.mean():.3f}')
    assert err.max() <= tol, (tag, float(err.max()), int(np.argmax(err)))
    assert mism_out == 0, tag


@pytest.mark.parametrize('sid', SCENARIO_IDS)
def test_sweep_all_flags_every_scenario(sid, robot, scene_loader):
    from jogramop_framework_b200.engine import CollisionEngine
    scene = scene_loader(sid)
    e, o = CollisionEngine(robot, scene), Oracle(robot, scene)
    q = uniform_configs(robot, N, 7000 + sid)
    rb, rd, rr = o.check(q.astype(np.float64), flags=ALL, want_reason=True)
    b, d, r = e.check(q, flags=ALL, want_reason=True)
    _cmp(f'{sid:03d} all flags {N}', b, d, rb, rd)
    # reason codes (scenario.py:109-119 order) wherever the deciding distance is outside the band
    r = r.cpu().numpy()
    clear = np.abs(rd + 0.001) > 1e-3
    assert (r[clear] != rr[clear]).mean() < 2e-3     # self vs scene may swap when both are inside the band
    e.close()


@pytest.mark.parametrize('sid', [11, 24, 34, 45])
def test_sweep_edges_poses_clamped(sid, robot, scene_loader):
    from jogramop_framework_b200.engine import CollisionEngine
    scene = scene_loader(sid)
    e, o = CollisionEngine(robot, scene), Oracle(robot, scene)
    lim = robot.joint_limits[robot.arm_joint_ids]
    # edges (BASELINE config 4 semantics): 64 substeps, qb = clip(qa + U(-0.25, 0.25))
    m = max(N // 10, 100)
    qa = uniform_configs(robot, m, 8000 + sid)
    qb = np.clip(qa + np.random.default_rng(8100 + sid).uniform(-0.25, 0.25, qa.shape), lim[:, 0], lim[:, 1]).astype(np.float32)
    rb, rd = o.check_edges(qa.astype(np.float64), qb.astype(np.float64), substeps=64)
    b, d = e.check_edges(qa, qb, substeps=64)
    _cmp(f'{sid:03d} edges {m}x64', b, d, rb, rd)
    b2, _ = e.check_edges(qa, qb, substeps=64, want_dist=False)
    outside = np.abs(rd + 0.001) > 1e-3
    assert np.array_equal(b2.cpu().numpy().astype(bool)[outside], rb[outside])
    # free-floating Panda hand + fingers (BASELINE config 5 sampling: target box + 0.25 m, uniform rotations)
    rng = np.random.default_rng(8200 + sid)
    tgt = scene.piece(0)
    lo, hi = tgt.min(0) - 0.25, tgt.max(0) + 0.25
    pos = lo + (hi - lo) * rng.random((N, 3))
    quat = rng.normal(size=(N, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    x, y, z, w = quat.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w),
                  1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w),
                  1 - 2 * (x * x + y * y)], 1).reshape(N, 3, 3)
    T = np.concatenate([R, pos[:, :, None]], 2)
    links = [robot.end_effector_id - 3, *robot.finger_joint_ids]
    rb, rd = o.check_poses(T, links=links, flags=3)
    b, d = e.check_poses(T.astype(np.float32), links=links, flags=3)
    _cmp(f'{sid:03d} poses {N}', b, d, rb, rd)
    # clamped mode (penetrations reported as -(margins)), all flags
    q = uniform_configs(robot, N, 8300 + sid)
    rb, rd = o.check(q.astype(np.float64), flags=ALL | oracle.NO_EPA)
    b, d, _ = e.check(q, flags=ALL | 16)
    _cmp(f'{sid:03d} clamped {N}', b, d, rb, rd)
    e.close()


def test_links_in_collision_all_42_pairs(robot):
    """``GraspingSimulator.links_in_collision`` (simulation.py:82-95) for every ordered pair ``in_self_collision``
    walks (simulation.py:448-463): a one-pair engine per pair, 1000 configurations each against the oracle's per-pair
    contact distance; plus the stateful single-shot call on a handful of them, and a non-default threshold."""
    from jogramop_framework_b200 import engine as E
    from jogramop_framework_b200.scenario import Scenario
    s = Scenario(11)
    rob, sim = s.get_robot_and_sim()
    o = Oracle(robot, None)
    q = uniform_configs(robot, 1000, 4242)
    ref = np.stack([o.self_distances(qi.astype(np.float64)) for qi in q])       # [1000, 42], +inf beyond 0.01
    pairs = [tuple(int(x) for x in p) for p in robot.self_pairs]
    assert len(pairs) == 42
    worst, n_close = 0.0, 0
    for k, (la, lb) in enumerate(pairs):
        eng = sim._engine(('links', la, lb), bodies=[], with_plane=False, self_pairs=[(la, lb)], small=True)
        b, d, _ = eng.check(q, flags=E.SELF)
        d, b = d.cpu().numpy().astype(np.float64), b.cpu().numpy().astype(bool)
        r = np.minimum(ref[:, k], 0.01)
        worst = max(worst, float(np.abs(d - r).max()))
        assert np.abs(d - r).max() <= 1e-4, (la, lb, float(np.abs(d - r).max()))
        outside = np.abs(r + 0.001) > 1e-3
        assert np.array_equal(b[outside], (r < -0.001)[outside]), (la, lb)
        n_close += int((r < 0.01).sum())
        # the stateful single-shot form, default and custom threshold (a free parameter of the reference call)
        for i in np.argsort(r)[:2].tolist() + [0]:
            rob.reset_arm_joints(q[i].astype(np.float64))
            got = sim.links_in_collision(rob.body_id, la, rob.body_id, lb)
            if abs(r[i] + 0.001) > 1e-3:
                assert got == bool(r[i] < -0.001), (la, lb, i)
            if abs(r[i] - 0.005) > 1e-4:
                assert sim.links_in_collision(rob.body_id, la, rob.body_id, lb, threshold=0.005) == bool(r[i] < 0.005)
    print(f'links_in_collision: 42 pairs x 1000 configurations, max |d - d_oracle| {worst:.2e}, {n_close} pair samples within 1 cm')
    assert n_close > 500
    sim.dismiss()


def test_low_threshold_verdict_only_matches_distance_mode(robot, scene_loader):
    """with a threshold below -(margin sum) the verdict-only shortcut (overlap => collision) is not exact;
    the engine must keep the penetration stages then.  threshold = -0.005: bits with and without distances agree."""
    from jogramop_framework_b200.engine import CollisionEngine
    scene = scene_loader(11)
    e = CollisionEngine(robot, scene, threshold=-0.005)
    o = Oracle(robot, scene)
    q = uniform_configs(robot, 20000, 515)
    b_d, d, _ = e.check(q, flags=ALL)
    b_v, _, _ = e.check(q, flags=ALL, want_dist=False)
    rb, rd = o.check(q.astype(np.float64), flags=ALL, threshold=-0.005)
    b_d, b_v, d = b_d.cpu().numpy(), b_v.cpu().numpy(), d.cpu().numpy()
    outside = np.abs(rd + 0.005) > 1e-4
    assert np.array_equal(b_d[outside], rb[outside])
    assert np.array_equal(b_v[outside], rb[outside])
    assert 0.02 < rb.mean() < 0.5 and (rd < -0.002).any()
    e.close()


def test_calls_on_different_streams_share_the_arena_safely(robot, scene_loader):
    """a device call on the torch stream followed (without a sync) by a host-path call on the library's own
    stream, and the other way round, must not corrupt each other (one scratch arena per engine: begin_call orders them)."""
    import torch
    from jogramop_framework_b200.engine import CollisionEngine
    scene = scene_loader(34)
    e = CollisionEngine(robot, scene)
    q1, q2 = uniform_configs(robot, 300000, 1), uniform_configs(robot, 300000, 2)
    ref1 = e.check(q1, flags=ALL)
    ref2 = e.check(q2, flags=ALL)
    torch.cuda.synchronize()
    r1b, r1d = ref1[0].clone(), ref1[1].clone()
    r2b, r2d = ref2[0].clone(), ref2[1].clone()
    dq1 = torch.from_numpy(q1).cuda()
    side = torch.cuda.Stream()
    for _ in range(3):
        a = e.check(dq1, flags=ALL)                                  # async on the current stream
        hb, hd, _ = e.check_host(q2, flags=ALL)                      # library stream, returns when done
        with torch.cuda.stream(side):
            c = e.check(dq1, flags=ALL)                              # third stream, right after
        torch.cuda.synchronize()
        assert torch.equal(a[0], r1b) and torch.equal(a[1], r1d)
        assert torch.equal(c[0], r1b) and torch.equal(c[1], r1d)
        hb_bits = np.unpackbits(hb.numpy().view(np.uint8), bitorder='little')[:len(q2)].astype(bool)
        assert np.array_equal(hb_bits, r2b.cpu().numpy()) and np.array_equal(hd.numpy(), r2d.cpu().numpy())
    e.close()


def test_simplified_geometry_profile(scene_loader):
    """mobile_panda_fingersSmallMesh.urdf (boxes around the links: the sister planners' profile; 13 shapes, panda_link7
    carries TWO of them, so its self pairs cover several shape combinations): all flags vs the oracle, and the drop-in
    classes with ``geometry='simplified'``."""
    from jogramop_framework_b200.engine import CollisionEngine
    from jogramop_framework_b200.ingest import RobotModel
    from jogramop_framework_b200.scenario import Scenario
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    small = RobotModel.load(os.path.join(ROOT, 'jogramop_framework_b200', 'data', 'robot_mobile_panda_simplified.npz'))
    assert small.n_shapes == 13 and (small.shape_link == 9).sum() == 2
    for sid in (11, 34):
        scene = scene_loader(sid)
        e, o = CollisionEngine(small, scene), Oracle(small, scene)
        q = uniform_configs(small, N, 9100 + sid)
        rb, rd = o.check(q.astype(np.float64), flags=ALL)
        b, d, _ = e.check(q, flags=ALL)
        _cmp(f'{sid:03d} simplified geometry, all flags {N}', b, d, rb, rd)
        e.close()
    s = Scenario(11)
    rob, sim = s.get_robot_and_sim(geometry='simplified')
    full, sim2 = s.get_robot_and_sim()
    q = uniform_configs(small, 20000, 9200)
    bs, _ = rob.in_collision_batch(q)
    bf, _ = full.in_collision_batch(q)
    # the boxes contain the links: whatever collides with the full meshes (well inside) collides with the boxes
    assert bs.float().mean() > bf.float().mean()
    assert (bs | ~bf).float().mean() > 0.995
    sim.dismiss()
    sim2.dismiss()
