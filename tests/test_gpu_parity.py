"""GPU parity tests: the CUDA path (through the C-ABI) against the double-precision CPU oracle on the same seeded
inputs.  Bars (BASELINE.json north_star): verdicts identical wherever the reference distance is more than 1 mm
from the decision surface; |d - d_ref| <= 1e-4 m; FK link poses within 1e-5."""
import numpy as np
import pytest
import torch

import oracle
from oracle import Oracle
from ref_numpy import SCENARIO_IDS, uniform_configs

pytestmark = pytest.mark.gpu

DIST_TOL = 1e-4     # north_star: distances must match within 1e-4 m
BAND = 1e-3         # north_star: verdicts must agree outside 1 mm of the decision surface
FK_TOL = 1e-5       # north_star: FK link poses within 1e-5


@pytest.fixture(scope='module')
def engines(robot, scene_loader):
    from jogramop_framework_b200.engine import CollisionEngine
    cache = {}

    def get(sid, **kw):
        key = (sid, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = CollisionEngine(robot, scene_loader(sid) if sid is not None else None, **kw)
        return cache[key]
    yield get
    for e in cache.values():
        e.close()


def compare(bits, dist, rbits, rdist, threshold=-0.001):
    bits, dist = bits.cpu().numpy(), dist.cpu().numpy().astype(np.float64)
    err = np.abs(dist - rdist)
    assert err.max() <= DIST_TOL, f'max |d - d_ref| = {err.max()}'
    outside = np.abs(rdist - threshold) > BAND
    assert np.array_equal(bits[outside], rbits[outside])
    return int((bits != rbits).sum()), float(err.max())


def test_library_is_loaded_and_no_cpu_fallback():
    from jogramop_framework_b200 import _lib
    L = _lib.load()
    assert L.jg_device_count() >= 1
    with open('/proc/self/maps') as fh:
        assert 'libjogramop_b200' in fh.read()      # (.so, or the _bounds / _count twin a debug run selected)


def test_fk_parity(robot, engines):
    e = engines(None)
    o = Oracle(robot)
    q = uniform_configs(robot, 20000, 7)
    ref = o.fk(q.astype(np.float64))
    got = e.fk(q).cpu().numpy()
    assert np.abs(got - ref).max() <= FK_TOL
    # world-frame EE pose through base_pose (forward_kinematics returns link 14 in WORLD, simulation.py:223-249)
    from jogramop_framework_b200.ingest import default_robot_pose
    pose = default_robot_pose()
    got = e.fk(q[:100], link_ids=[14], base_pose=pose).cpu().numpy()[:, 0]
    refw = np.einsum('ij,njk->nik', pose[:3, :3], ref[:100, 14])
    refw[:, :, 3] += pose[:3, 3]
    assert np.abs(got - refw).max() <= FK_TOL


@pytest.mark.parametrize('sid', [11, 21, 34, 43, 45])
@pytest.mark.parametrize('no_epa', [0, oracle.NO_EPA])
def test_scene_parity(sid, no_epa, robot, scene_loader, engines):
    """Full semantics (signed distance incl. penetration depth) and the clamped JG_NO_EPA mode."""
    e = engines(sid)
    o = Oracle(robot, scene_loader(sid))
    q = uniform_configs(robot, 30000, sid)
    rb, rd = o.check(q.astype(np.float64), flags=oracle.SCENE | oracle.PLANE | no_epa)
    bits, dist, _ = e.check(q, flags=oracle.SCENE | oracle.PLANE | no_epa)
    nd, me = compare(bits, dist, rb, rd)
    print(f'scenario {sid}: collision rate {rb.mean():.3f}, disagreements in band {nd}, max |dd| {me:.2e}')
    assert 0.15 < rb.mean() < 0.6
    if not no_epa:
        assert rd.min() < -0.05 and e.stats()['epa_pairs'] > 1000     # deep penetrations were really measured


def test_overflow_kernel_parity(robot, scene_loader, engines):
    """The warp-cooperative overflow pass (large polytopes) on EVERY penetrating pair (JG_DEBUG_FORCE_OVERFLOW)."""
    e = engines(11)
    o = Oracle(robot, scene_loader(11))
    q = uniform_configs(robot, 20000, 17)
    rb, rd = o.check(q.astype(np.float64), flags=oracle.SCENE | oracle.PLANE | oracle.SELF)
    bits, dist, _ = e.check(q, flags=oracle.SCENE | oracle.PLANE | oracle.SELF | 0x40000000)
    compare(bits, dist, rb, rd)
    bits2, dist2, _ = e.check(q, flags=oracle.SCENE | oracle.PLANE | oracle.SELF)
    assert (dist - dist2).abs().max().item() < 2e-5
    # a smaller pass capacity has fewer hand-over records (capacity / 4 = 16384): the first overflows of the pass continue
    # the compact polytope, the later ones restart from the GJK simplex -- both inside one pass.  (The overflow list
    # itself holds `capacity` entries: with every polytope forced through it, the pass must stay below that.)
    es = engines(11, chunk=65536)
    bits3, dist3, _ = es.check(q, flags=oracle.SCENE | oracle.PLANE | oracle.SELF | 0x40000000)
    compare(bits3, dist3, rb, rd)
    assert 65536 // 4 < es.stats()['overflow_run'] < 65536


def _tri_scene(sc, margin=0.001):
    import copy
    t = copy.copy(sc)
    t.piece_verts = sc.tri_verts[sc.tris].reshape(-1, 3)
    t.piece_vert_offset = (np.arange(len(sc.tris) + 1) * 3).astype(np.int32)
    t.piece_body = sc.tri_body
    t.piece_margin = np.full(len(sc.tris), margin)
    return t


def test_mesh_mode_edges_host_path_and_deep_trees(robot, scene_loader):
    """Mesh mode beyond single configurations: edge validation (substeps share the LBVH queries), the host-buffer entry
    points (blocking and ticket form), scenario 045, and a degenerate mesh whose LBVH is a long chain (many coincident
    centroids -> Morton ties) plus one far larger than a lane's old 64-entry stack could hold a frontier of."""
    import copy
    from jogramop_framework_b200.engine import CollisionEngine
    sc = scene_loader(45)
    o = Oracle(robot, _tri_scene(sc))
    e = CollisionEngine(robot, sc, mesh_mode=True)
    fl = oracle.SCENE | oracle.PLANE | oracle.NO_EPA
    lim = robot.joint_limits[robot.arm_joint_ids]
    qa = uniform_configs(robot, 1500, 451)
    qb = np.clip(qa + np.random.default_rng(452).uniform(-0.25, 0.25, qa.shape), lim[:, 0], lim[:, 1]).astype(np.float32)
    rb, rd = o.check_edges(qa.astype(np.float64), qb.astype(np.float64), substeps=16, flags=fl)
    b, d = e.check_edges(qa, qb, substeps=16, flags=fl)
    compare(b, d, rb, rd)
    q = uniform_configs(robot, 30000, 453)
    rb, rd = o.check(q.astype(np.float64), flags=fl)
    b, d, _ = e.check(q, flags=fl)
    compare(b, d, rb, rd)
    hb, hd, _ = e.check_host(q, flags=fl)
    assert np.array_equal(np.unpackbits(hb.numpy().view(np.uint8), bitorder='little')[:len(q)].astype(bool), b.cpu().numpy())
    assert np.array_equal(hd.numpy(), d.cpu().numpy())
    tk = e.check_host_submit(q, flags=fl)
    tb, td, _ = e.check_host_wait(tk)
    assert np.array_equal(tb.numpy(), hb.numpy()) and np.array_equal(td.numpy(), hd.numpy())
    e.close()
    # degenerate mesh: 3000 copies of a few triangles at one place (identical Morton codes: the hierarchy is decided by
    # the index tie-break) + a dense fan of slivers around the robot's workspace
    rng = np.random.default_rng(7)
    base = np.array([[0.5, -0.2, 0.3], [0.7, -0.2, 0.3], [0.6, 0.0, 0.45]])
    tv = [base + 1e-7 * rng.normal(size=(3, 3)) for _ in range(3000)]
    ang = np.linspace(0, 2 * np.pi, 2000, endpoint=False)
    tv += [np.array([[0.6, 0.3, 0.2], [0.6 + 0.5 * np.cos(a), 0.3 + 0.5 * np.sin(a), 0.6], [0.6 + 0.5 * np.cos(a + 0.004), 0.3 + 0.5 * np.sin(a + 0.004), 0.6]]) for a in ang]
    deg = copy.copy(sc)
    deg.tri_verts = np.concatenate(tv)
    deg.tris = np.arange(len(deg.tri_verts)).reshape(-1, 3).astype(np.int32)
    deg.tri_body = np.zeros(len(deg.tris), np.int32)
    od = Oracle(robot, _tri_scene(deg))
    ed = CollisionEngine(robot, deg, mesh_mode=True)
    q = uniform_configs(robot, 4000, 454)
    q[:, 0] = np.random.default_rng(455).uniform(-0.3, 0.3, len(q))     # base near the slivers / the stack of copies
    q[:, 1] = np.random.default_rng(456).uniform(-0.2, 0.6, len(q))
    rb, rd = od.check(q.astype(np.float64), flags=fl)
    b, d, _ = ed.check(q, flags=fl)
    compare(b, d, rb, rd)
    assert 0.05 < rb.mean() < 0.95
    ed.close()


@pytest.mark.parametrize('sid', [11, 34])
def test_mesh_mode_lbvh_parity(sid, robot, scene_loader):
    """Mesh mode: link hulls vs the merged triangle mesh (export/obstacles.obj) through the GPU-built LBVH ==
    the oracle run on a scene whose convex pieces are the individual triangles."""
    import copy
    from jogramop_framework_b200.engine import CollisionEngine
    sc = scene_loader(sid)
    tri_scene = copy.copy(sc)
    tri_scene.piece_verts = sc.tri_verts[sc.tris].reshape(-1, 3)
    tri_scene.piece_vert_offset = (np.arange(len(sc.tris) + 1) * 3).astype(np.int32)
    tri_scene.piece_body = sc.tri_body
    tri_scene.piece_margin = np.full(len(sc.tris), 0.001)
    o = Oracle(robot, tri_scene)
    e = CollisionEngine(robot, sc, mesh_mode=True)
    q = uniform_configs(robot, 20000, 100 + sid)
    fl = oracle.SCENE | oracle.PLANE | oracle.NO_EPA
    rb, rd = o.check(q.astype(np.float64), flags=fl)
    bits, dist, _ = e.check(q, flags=fl)
    compare(bits, dist, rb, rd)
    # the surface of the mesh is what collides: a mesh-mode hit implies a hit of the solid convex pieces
    eb, ed, _ = Oracle(robot, sc).check(q.astype(np.float64), flags=fl), None, None
    assert (eb[0] | ~rb).mean() > 0.999
    assert 0.1 < rb.mean() < 0.6
    e.close()


def test_verdict_only_mode_gives_the_same_bits(robot, scene_loader, engines):
    """min_dist == NULL: the engine only proves / disproves 'some contact distance < threshold'.  Bits must be
    identical to the full query's bits (same fp32 arithmetic up to early exits), reasons outside the band too."""
    e = engines(11)
    q = uniform_configs(robot, 200000, 23)
    for fl in (3, 15, 4):
        b_full, d_full, r_full = e.check(q, flags=fl, want_reason=True)
        b_fast, d_none, r_fast = e.check(q, flags=fl, want_dist=False, want_reason=True)
        assert d_none is None
        clear = (d_full + 0.001).abs() > 1e-5
        assert torch.equal(b_full[clear], b_fast[clear])
        assert (b_full != b_fast).sum().item() <= 3
        assert (r_full[clear] == r_fast[clear]).float().mean().item() > 0.9999
    be, _ = e.check_edges(q[:2000], q[2000:4000], substeps=16, want_dist=False)
    bf, df = e.check_edges(q[:2000], q[2000:4000], substeps=16)
    assert (be != bf).sum().item() <= 1


def test_fixed_base_panda_without_platform(scene_loader):
    """panda.urdf (with_platform=False, simulation.py:136-139,161-164,453-455): 7 dof, a collision shape on the BASE
    link (index -1), 25 self pairs, end effector = link 11."""
    import os
    from conftest import ROOT
    from jogramop_framework_b200.engine import CollisionEngine
    from jogramop_framework_b200.ingest import RobotModel
    panda = RobotModel.load(os.path.join(ROOT, 'jogramop_framework_b200', 'data', 'robot_panda.npz'))
    assert panda.shape_link[0] == -1 and len(panda.arm_joint_ids) == 7
    sc = scene_loader(21)
    e = CollisionEngine(panda, sc)
    o = Oracle(panda, sc)
    q = uniform_configs(panda, 20000, 5)
    assert q.shape == (20000, 7)
    fl = oracle.SCENE | oracle.PLANE | oracle.SELF | oracle.LIMITS
    rb, rd = o.check(q.astype(np.float64), flags=fl)
    bits, dist, _ = e.check(q, flags=fl)
    compare(bits, dist, rb, rd)
    assert np.abs(e.fk(q[:2000]).cpu().numpy() - o.fk(q[:2000].astype(np.float64))).max() <= FK_TOL
    e.close()


def test_degenerate_scenes(robot, scene_loader):
    """No obstacle pieces (plane only), no plane (pieces only), and no scene at all (self-collision / FK only)."""
    from jogramop_framework_b200.engine import CollisionEngine
    sc = scene_loader(34)
    q = uniform_configs(robot, 5000, 6)
    plane_only = sc.subset([], with_plane=True)
    e = CollisionEngine(robot, plane_only)
    rb, rd = Oracle(robot, plane_only).check(q.astype(np.float64), flags=3)
    bits, dist, _ = e.check(q, flags=3)
    compare(bits, dist, rb, rd)
    e.close()
    no_plane = sc.subset([0, 1, 2], with_plane=False)
    e = CollisionEngine(robot, no_plane)
    rb, rd = Oracle(robot, no_plane).check(q.astype(np.float64), flags=3)
    bits, dist, _ = e.check(q, flags=3)
    compare(bits, dist, rb, rd)
    e.close()
    e = CollisionEngine(robot, None)
    rb, rd = Oracle(robot, None).check(q.astype(np.float64), flags=oracle.SELF)
    bits, dist, _ = e.check(q, flags=oracle.SELF | oracle.SCENE | oracle.PLANE)     # scene flags are no-ops without a scene
    compare(bits, dist, rb, rd)
    e.close()


def test_scene_with_more_than_31_pieces(robot, scene_loader):
    """The broad phase reserves the slots of a shape's keys with lanes 0..P: scenes with more than 31 pieces take
    several 32-key chunks.  40 pieces = scenario 034's obstacles copied onto a grid around the robot."""
    import copy
    from jogramop_framework_b200.engine import CollisionEngine
    sc = scene_loader(34)
    parts, margins, bodies = [], [], []
    shifts = [(0.0, 0.0)] + [(0.8 * i - 2.4, 1.2 * j - 1.2) for i in range(5) for j in range(3)]
    for dx, dy in shifts:
        for pi in range(sc.n_pieces):
            parts.append(sc.piece(pi) + np.array([dx, dy, 0.0]))
            margins.append(sc.piece_margin[pi])
            bodies.append(sc.piece_body[pi])
            if len(parts) == 40:
                break
        if len(parts) == 40:
            break
    assert len(parts) == 40
    big = copy.copy(sc)
    big.piece_verts = np.concatenate(parts)
    big.piece_vert_offset = np.cumsum([0] + [len(x) for x in parts]).astype(np.int32)
    big.piece_margin = np.asarray(margins, dtype=np.float64)
    big.piece_body = np.asarray(bodies, dtype=np.int32)
    q = uniform_configs(robot, 20000, 77)
    e = CollisionEngine(robot, big)
    rb, rd = Oracle(robot, big).check(q.astype(np.float64), flags=3)
    bits, dist, _ = e.check(q, flags=3)
    compare(bits, dist, rb, rd)
    assert 0.2 < rb.mean() < 0.95
    e.close()


def test_invariances(robot, scene_loader):
    """Properties the geometry must have: moving robot base and scene together changes nothing; the verdict is
    monotone in the threshold; a smaller query distance only clamps the reported distance."""
    from jogramop_framework_b200.engine import CollisionEngine
    from jogramop_framework_b200.ingest import default_robot_pose
    sc = scene_loader(43)
    q = uniform_configs(robot, 20000, 43)
    e = CollisionEngine(robot, sc)
    b0, d0, _ = e.check(q, flags=3)
    # shift the platform joints by (dy, dx) and the scene by the same vector in the robot frame
    dy, dx = 0.125, -0.25
    q2 = q.copy()
    q2[:, 0] += dy
    q2[:, 1] += dx
    import copy
    sc2 = copy.copy(sc)
    sc2.piece_verts = sc.piece_verts + np.array([dx, dy, 0.0])
    e2 = CollisionEngine(robot, sc2)
    b2, d2, _ = e2.check(q2, flags=3)
    assert (d0 - d2).abs().max().item() < 2e-5
    assert (b0 != b2).sum().item() <= 2
    e2.close()
    # threshold monotonicity and query-distance clamping
    e3 = CollisionEngine(robot, sc, threshold=0.004, query_distance=0.005)
    b3, d3, _ = e3.check(q, flags=3)
    assert (b3 | ~b0).all().item()
    assert torch.allclose(d3, torch.clamp(d0, max=0.005), atol=2e-6)
    assert torch.equal(b3, d3 < 0.004)
    e3.close()
    e.close()


def test_scene_only_and_plane_only(robot, scene_loader, engines):
    e = engines(11)
    o = Oracle(robot, scene_loader(11))
    q = uniform_configs(robot, 8000, 3)
    for fl in (oracle.SCENE | oracle.NO_EPA, oracle.PLANE):
        rb, rd = o.check(q.astype(np.float64), flags=fl)
        bits, dist, _ = e.check(q, flags=fl)
        compare(bits, dist, rb, rd)


def test_self_collision_and_reason_codes(robot, scene_loader, engines):
    e = engines(34)
    o = Oracle(robot, scene_loader(34))
    lim = robot.joint_limits[robot.arm_joint_ids]
    q = uniform_configs(robot, 12000, 5)
    q[::7, 3] = lim[3, 1] + 0.2       # limit violations
    q[5::11, 0] = np.nan              # NaN never passes the limit test
    fl = oracle.SCENE | oracle.PLANE | oracle.SELF | oracle.LIMITS | oracle.NO_EPA
    rb, rd, rr = o.check(q.astype(np.float64), flags=fl, want_reason=True)
    bits, dist, reason = e.check(q, flags=fl, want_reason=True)
    ok = ~np.isnan(q).any(1)
    nd, me = compare(bits[torch.from_numpy(ok)], dist[torch.from_numpy(ok)], rb[ok], rd[ok])
    assert bits.cpu().numpy()[~ok].all()
    reason = reason.cpu().numpy()
    assert (reason[~ok] == 1).all()
    clear = ok & (np.abs(rd + 0.001) > BAND)
    agree = (reason[clear] == rr[clear]).mean()
    assert agree > 0.999          # reason may differ only when self and scene are both within the band
    assert (rr == 2).sum() > 100 and (rr == 1).sum() > 1000
    # self-collision alone (in_self_collision, simulation.py:439-464), with and without penetration depth
    for extra in (oracle.NO_EPA, 0):
        rb, rd = o.check(q[ok].astype(np.float64), flags=oracle.SELF | extra)
        bits, dist, _ = e.check(q[ok], flags=oracle.SELF | extra)
        compare(bits, dist, rb, rd)
    assert 0.05 < rb.mean() < 0.4


def test_golden_ik_rows_are_free_on_gpu(robot, golden, scene_loader, engines):
    """G4 on the GPU: all 89 accepted IK solutions pass limits + self + scene."""
    fl = oracle.SCENE | oracle.PLANE | oracle.SELF | oracle.LIMITS
    total = 0
    for sid in SCENARIO_IDS:
        ik = golden[f'ik_{sid:03d}']
        if not len(ik):
            continue
        e = engines(sid)
        bits, dist, reason = e.check(ik.astype(np.float32), flags=fl, want_reason=True)
        assert not bits.any().item(), (sid, reason.cpu().numpy())
        total += len(ik)
    assert total == 89


def test_edges_parity(robot, scene_loader, engines):
    e = engines(34)
    o = Oracle(robot, scene_loader(34))
    lim = robot.joint_limits[robot.arm_joint_ids]
    qa = uniform_configs(robot, 3000, 34)
    rng = np.random.default_rng(3400)
    qb = np.clip(qa + rng.uniform(-0.25, 0.25, qa.shape), lim[:, 0], lim[:, 1]).astype(np.float32)
    for sub, fl in ((64, oracle.SCENE | oracle.PLANE | oracle.NO_EPA), (5, oracle.SCENE | oracle.PLANE)):
        # the oracle interpolates in double from the same float32 endpoints; the kernel interpolates in float32
        rb, rd = o.check_edges(qa.astype(np.float64), qb.astype(np.float64), substeps=sub, flags=fl)
        bits, dist = e.check_edges(qa, qb, substeps=sub, flags=fl)
        compare(bits, dist, rb, rd)


def test_gripper_poses_parity(robot, scene_loader, engines):
    sc = scene_loader(11)
    e = engines(11)
    o = Oracle(robot, sc)
    rng = np.random.default_rng(5011)
    n = 20000
    tgt = sc.piece(0)        # target object is body 0 / piece 0
    lo, hi = tgt.min(0) - 0.25, tgt.max(0) + 0.25
    pos = lo + (hi - lo) * rng.random((n, 3))
    quat = rng.normal(size=(n, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    pq = np.concatenate([pos, quat], 1).astype(np.float32)
    x, y, z, w = (pq[:, 3 + i].astype(np.float64) for i in range(4))
    nrm = np.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / nrm, y / nrm, z / nrm, w / nrm
    Rm = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                   2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                   2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).reshape(n, 3, 3)
    T = np.concatenate([Rm, pq[:, :3, None].astype(np.float64)], 2)
    fl = oracle.SCENE | oracle.PLANE
    rb, rd = o.check_poses(T, links=[11, 12, 13], flags=fl)
    bits, dist = e.check_poses(pq, links=[11, 12, 13], flags=fl)
    compare(bits, dist, rb, rd)
    bits2, dist2 = e.check_poses(T.astype(np.float32), links=[11, 12, 13], flags=fl)
    compare(bits2, dist2, rb, rd)
    rb, rd = o.check_poses(T, links=[11, 12, 13], flags=fl | oracle.NO_EPA)
    bits, dist = e.check_poses(pq, links=[11, 12, 13], flags=fl | oracle.NO_EPA)
    compare(bits, dist, rb, rd)
    assert 0.02 < rb.mean() < 0.9


def test_chunking_tails_and_empty(robot, scene_loader, engines):
    """Ragged sizes: n not a multiple of 32 / of the pass size, n = 0, n = 1."""
    e = engines(34, chunk=4096)
    big = engines(34)
    q = uniform_configs(robot, 4096 * 2 + 37, 9)
    b1, d1, _ = e.check(q, flags=3 | 16)
    b2, d2, _ = big.check(q, flags=3 | 16)
    assert torch.equal(b1, b2) and torch.equal(d1, d2)
    b0, d0, _ = e.check(q[:0], flags=3)
    assert b0.numel() == 0 and d0.numel() == 0
    b, d, _ = e.check(q[:1], flags=3 | 16)
    assert b.shape == (1,) and bool(b[0]) == bool(b2[0])
    packed, _, _ = e.check(q[:33], flags=3 | 16, packed=True)
    assert packed.numel() == 2


def test_host_path_matches_device_path(robot, scene_loader, engines):
    e = engines(11, chunk=8192)
    q = uniform_configs(robot, 8192 * 3 + 100, 21)
    from jogramop_framework_b200.engine import unpack_bits
    b1, d1, r1 = e.check(q, flags=15 | 16, want_reason=True)
    qp = torch.from_numpy(q).pin_memory()
    wb, wd, wr = e.check_host(qp, flags=15 | 16, want_reason=True)
    assert torch.equal(unpack_bits(wb.cuda(), len(q)), b1)
    assert torch.equal(wd.cuda(), d1) and torch.equal(wr.cuda(), r1)
    wb2, wd2, _ = e.check_host(q, flags=15 | 16)       # pageable numpy input
    assert torch.equal(wb2, wb) and torch.equal(wd2, wd)


def test_async_host_path_two_batches_in_flight(robot, scene_loader, engines):
    """jg_check_host_submit / _wait: results identical to the device path, tickets alternate buffer sets, a third
    submit before a wait is refused, pageable buffers are drained at wait time."""
    from jogramop_framework_b200._lib import JogramopError
    from jogramop_framework_b200.engine import unpack_bits
    e = engines(11, chunk=8192)
    batches = [uniform_configs(robot, n, 30 + i) for i, n in enumerate([8192, 5000, 8192, 33, 7000])]
    ref = [e.check(q, flags=15, want_reason=True) for q in batches]
    pinned = [torch.from_numpy(q).pin_memory() for q in batches]
    tickets = [e.check_host_submit(pinned[0], flags=15, want_reason=True), e.check_host_submit(pinned[1], flags=15, want_reason=True)]
    with pytest.raises(JogramopError):
        e.check_host_submit(pinned[2], flags=15, want_reason=True)
    with pytest.raises(JogramopError):
        e.check_host(pinned[2], flags=15)                  # the synchronous call refuses to share the buffers
    got = []
    for i in range(len(batches)):
        got.append(e.check_host_wait(tickets[i]))
        if i + 2 < len(batches):
            src = pinned[i + 2] if i % 2 == 0 else batches[i + 2]      # pinned and pageable inputs
            tickets.append(e.check_host_submit(src, flags=15, want_reason=True))
    for (wb, wd, wr), (b1, d1, r1), q in zip(got, ref, batches):
        assert torch.equal(unpack_bits(wb.cuda(), len(q)), b1)
        assert torch.equal(wd.cuda(), d1) and torch.equal(wr.cuda(), r1)
    assert e.check_host_wait(tickets[0])[0] is got[0][0]   # waiting twice is harmless
    with pytest.raises(JogramopError):
        e.check_host_submit(uniform_configs(robot, 8193, 1), flags=3)    # more than one pass: use check_host
    wb, wd, _ = e.check_host(pinned[0], flags=15)          # synchronous path works again once everything is collected
    assert torch.equal(unpack_bits(wb.cuda(), len(batches[0])), ref[0][0])


def test_engine_pool_batches_in_flight(robot, scene_loader):
    """EnginePool: whole batches side by side on several arenas / streams give exactly what one engine gives batch by
    batch -- inputs produced on the caller's stream just before the submit, preallocated outputs, ragged sizes, more
    batches than lanes, the host-buffer form, and the oracle on one of them."""
    from jogramop_framework_b200._lib import JogramopError
    from jogramop_framework_b200.engine import CollisionEngine, EnginePool, unpack_bits
    scene = scene_loader(11)
    single = CollisionEngine(robot, scene, chunk=8192)
    pool = EnginePool(robot, scene, lanes=3, chunk=8192)
    sizes = [8192, 20000, 33, 5000, 8192 * 2 + 1, 1, 7000, 12345, 64, 30000]
    host = [uniform_configs(robot, n, 700 + i) for i, n in enumerate(sizes)]
    ref = [single.check(q, flags=15, want_reason=True) for q in host]
    tickets = []
    for q in host:
        qd = torch.from_numpy(q).cuda() * 2.0
        qd = qd * 0.5                                   # produced on the current stream right before the submit
        tickets.append(pool.submit(qd, flags=15, want_reason=True))
        del qd
    for tk, (b1, d1, r1) in zip(tickets, ref):
        b, d, r = pool.wait(tk)
        assert torch.equal(b, b1) and torch.equal(d, d1) and torch.equal(r, r1)
    # check_many + preallocated outputs
    outs = pool.check_many(host[:4], flags=3, packed=True)
    for (w, d, _), q in zip(outs, host[:4]):
        b1, d1, _ = single.check(q, flags=3)
        assert torch.equal(unpack_bits(w, len(q)), b1) and torch.equal(d, d1)
    words = torch.zeros((4, 1024), dtype=torch.int32, device='cuda')
    dists = torch.zeros((4, 20000), dtype=torch.float32, device='cuda')
    tks = [pool.submit(host[i], flags=3, packed=True, bits_out=words[i], dist_out=dists[i]) for i in range(4)]
    pool.drain()
    for i, q in enumerate(host[:4]):
        b1, d1, _ = single.check(q, flags=3)
        assert torch.equal(unpack_bits(words[i, :(len(q) + 31) // 32], len(q)), b1) and torch.equal(dists[i, :len(q)], d1)
    # ONE large call cut into passes that run side by side: configurations, edges, gripper poses
    assert pool.engines[0].pass_size == 8192
    big = uniform_configs(robot, 8192 * 5 + 77, 811)
    b1, d1, r1 = single.check(big, flags=15, want_reason=True)
    b, d, r = pool.check(big, flags=15, want_reason=True)
    assert torch.equal(b, b1) and torch.equal(d, d1) and torch.equal(r, r1)
    assert torch.equal(pool.check(big, flags=3, want_dist=False)[0], single.check(big, flags=3, want_dist=False)[0])
    assert pool.check(big[:0], flags=3)[0].numel() == 0
    from jogramop_framework_b200.dist import check_sharded
    from jogramop_framework_b200.engine import unpack_bits as ub
    w, dd = check_sharded(pool, torch.from_numpy(big).cuda(), 15, want_dist=True)       # world size 1: the pool in an engine's place
    assert torch.equal(ub(w, len(big)), b1) and torch.equal(dd, d1)
    qa = uniform_configs(robot, 3001, 812)
    qb = (qa + np.random.default_rng(813).uniform(-0.2, 0.2, qa.shape)).astype(np.float32)
    eb1, ed1 = single.check_edges(qa, qb, substeps=10)
    eb, ed = pool.check_edges(qa, qb, substeps=10)
    assert torch.equal(eb, eb1) and torch.equal(ed, ed1)
    rng = np.random.default_rng(814)
    pq = np.concatenate([rng.uniform(-0.6, 0.6, (20001, 3)) + [0, 0, 0.6], rng.normal(size=(20001, 4))], 1)
    pq[:, 3:] /= np.linalg.norm(pq[:, 3:], axis=1, keepdims=True)
    pq = pq.astype(np.float32)
    links = [robot.end_effector_id - 3, *robot.finger_joint_ids]
    pb1, pd1 = single.check_poses(pq, links=links)
    pb, pd = pool.check_poses(pq, links=links)
    assert torch.equal(pb, pb1) and torch.equal(pd, pd1) and 0.02 < pb1.float().mean().item() < 0.98
    # host buffers: 2 tickets per lane, then refused
    small = [torch.from_numpy(q).pin_memory() for q in host if len(q) <= 8192]
    ht = [pool.submit_host(small[i % len(small)], flags=15, want_reason=True) for i in range(6)]
    with pytest.raises(JogramopError):
        pool.submit_host(small[0], flags=15)
    for i, t in enumerate(ht):
        wb, wd, wr = pool.wait_host(t)
        q = small[i % len(small)]
        b1, d1, r1 = single.check(q.numpy(), flags=15, want_reason=True)
        assert torch.equal(unpack_bits(wb.cuda(), len(q)), b1) and torch.equal(wd.cuda(), d1) and torch.equal(wr.cuda(), r1)
    pool.wait_host(ht[0])                               # waiting twice is harmless
    t = pool.submit_host(small[0], flags=15)
    pool.wait_host(t)
    # against the oracle
    o = Oracle(robot, scene)
    rb, rd = o.check(host[3].astype(np.float64), flags=15)[:2]
    b, d, _ = pool.wait(pool.submit(host[3], flags=15))
    b, d = b.cpu().numpy().astype(bool), d.cpu().numpy()
    outside = np.abs(rd + 0.001) > 1e-3
    assert (b[outside] == rb[outside]).all() and np.abs(d - rd).max() < 1e-4
    pool.close()
    single.close()


def test_engine_lanes_cut_large_calls(robot, scene_loader):
    """jg_engine_set_lanes: a call of several passes runs them on several arenas / internal streams -- same bits, distances
    and reasons as the one-arena engine, counters of the whole call, stream semantics kept (results consumed on the caller's
    stream right after the call), one-pass calls untouched, lanes can be taken away again."""
    from jogramop_framework_b200._lib import JogramopError
    from jogramop_framework_b200.engine import CollisionEngine
    scene = scene_loader(11)
    single = CollisionEngine(robot, scene, chunk=8192)
    laned = CollisionEngine(robot, scene, chunk=8192, lanes=3)
    assert laned._L.jg_engine_lanes(laned._e) == 3 and laned.pass_size == 8192
    big = torch.from_numpy(uniform_configs(robot, 8192 * 7 + 1234, 901)).cuda()
    for fl, wd in ((15, True), (3, True), (3, False)):
        b1, d1, r1 = single.check(big, flags=fl, want_dist=wd, want_reason=True)
        s1 = single.stats()
        b, d, r = laned.check(big, flags=fl, want_dist=wd, want_reason=True)
        nb = int(b.sum().item())                       # consumed on the caller's stream straight away
        assert nb == int(b1.sum().item())
        assert torch.equal(b, b1) and torch.equal(r, r1) and (not wd or torch.equal(d, d1))
        s = laned.stats()
        assert s['configs'] == len(big) == s1['configs'] and s['pairs'] == s1['pairs'] and s['gjk_run'] == s1['gjk_run']
    small = big[:5000]
    assert torch.equal(laned.check(small, flags=15)[0], single.check(small, flags=15)[0])       # one pass: not cut
    assert laned.stats()['configs'] == 5000
    qa = uniform_configs(robot, 5003, 902)
    qb = (qa + np.random.default_rng(903).uniform(-0.2, 0.2, qa.shape)).astype(np.float32)
    eb1, ed1 = single.check_edges(qa, qb, substeps=16)
    eb, ed = laned.check_edges(qa, qb, substeps=16)
    assert torch.equal(eb, eb1) and torch.equal(ed, ed1)
    rng = np.random.default_rng(904)
    pq = np.concatenate([rng.uniform(-0.6, 0.6, (30001, 3)) + [0, 0, 0.6], rng.normal(size=(30001, 4))], 1)
    pq[:, 3:] /= np.linalg.norm(pq[:, 3:], axis=1, keepdims=True)
    pq = pq.astype(np.float32)
    links = [robot.end_effector_id - 3, *robot.finger_joint_ids]
    pb1, pd1 = single.check_poses(pq, links=links)
    pb, pd = laned.check_poses(pq, links=links)
    assert torch.equal(pb, pb1) and torch.equal(pd, pd1)
    x, y, z, w = (pq[:, 3 + k].astype(np.float64) for k in range(4))          # the matrix form has another stride
    T = np.zeros((len(pq), 3, 4), np.float32)
    T[:, 0, :3] = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], 1)
    T[:, 1, :3] = np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], 1)
    T[:, 2, :3] = np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)
    T[:, :, 3] = pq[:, :3]
    pbm1, pdm1 = single.check_poses(T, links=links)
    pbm, pdm = laned.check_poses(T, links=links)
    assert torch.equal(pbm, pbm1) and torch.equal(pdm, pdm1) and torch.allclose(pdm, pd1, atol=1e-4)
    # while profiling nothing is cut (sections of overlapping passes would time each other); results unchanged
    laned.profile(True)
    assert torch.equal(laned.check(big, flags=15)[0], single.check(big, flags=15)[0])
    assert laned.get_profile()['broadphase'][1] == 8
    laned.profile(False)
    laned.set_lanes(1)
    assert torch.equal(laned.check(big, flags=15)[0], single.check(big, flags=15)[0])
    with pytest.raises(JogramopError):
        laned.set_lanes(9)
    laned.close()
    single.close()


def test_non_finite_inputs_are_rejected_on_the_gpu(robot, scene_loader, engines):
    """NaN / inf joint values and poses: bit set, reason 1, distance = query distance, with or without JG_LIMITS -- as the
    oracle does (tests/test_oracle_properties.py); the neighbours of such a row are untouched."""
    e = engines(11)
    o = Oracle(robot, scene_loader(11))
    q = uniform_configs(robot, 4096, 606)
    bad = [5, 64, 1000, 4095]
    q[5, 2] = np.nan
    q[64, 0] = np.inf
    q[1000, 8] = -np.inf
    q[4095, 4] = np.nan
    for fl in (oracle.SCENE | oracle.PLANE, oracle.SCENE | oracle.PLANE | oracle.SELF | oracle.LIMITS):
        rb, rd, rr = o.check(q.astype(np.float64), flags=fl, want_reason=True)
        b, d, r = e.check(q, flags=fl, want_reason=True)
        b, d, r = b.cpu().numpy(), d.cpu().numpy(), r.cpu().numpy()
        assert b[bad].all() and (r[bad] == 1).all() and np.allclose(d[bad], 0.01) and rb[bad].all()
        compare(torch.from_numpy(b), torch.from_numpy(d), rb, rd)
    bv, _, _ = e.check(q, flags=3, want_dist=False)
    assert bv.cpu().numpy()[bad].all()
    eb, ed = e.check_edges(q[:64], q[64:128], substeps=8)
    assert bool(eb[5]) and bool(eb[0] == o.check_edges(q[:1].astype(np.float64), q[64:65].astype(np.float64), substeps=8)[0][0])
    links = [robot.end_effector_id - 3, *robot.finger_joint_ids]
    T = np.tile(np.eye(4, dtype=np.float32)[:3], (64, 1, 1))
    T[:, :, 3] = [2.0, 0.0, 1.0]
    T[7, 1, 1] = np.nan
    pb, pd = e.check_poses(T, links=links)
    assert bool(pb[7]) and not bool(pb[6]) and abs(float(pd[7]) - 0.01) < 1e-7
    # mesh mode: a NaN query must not walk the LBVH (every box test is "not separated" for NaN)
    from jogramop_framework_b200.engine import CollisionEngine
    em = CollisionEngine(robot, scene_loader(11), mesh_mode=True)
    mb, md, _ = em.check(q[:256], flags=oracle.SCENE | oracle.PLANE | oracle.NO_EPA)
    assert bool(mb[5]) and bool(mb[64]) and abs(float(md[5]) - 0.01) < 1e-7
    pm, _ = em.check_poses(T, links=links, flags=oracle.SCENE | oracle.PLANE | oracle.NO_EPA)
    assert bool(pm[7]) and not bool(pm[6])
    em.close()


def test_error_conventions(robot, scene_loader, engines):
    from jogramop_framework_b200._lib import JogramopError
    e = engines(34)
    with pytest.raises(JogramopError):
        e._L.jg_check(e._e, None, 5, 9, 3, None, None, None, None) and None
        from jogramop_framework_b200 import _lib
        _lib.check(e._L.jg_check(e._e, None, 5, 9, 3, None, None, None, None), 'jg_check')
    from jogramop_framework_b200 import _lib
    assert e._L.jg_check(e._e, None, 5, 8, 3, None, None, None, None) == -1      # dof mismatch
    assert b'dof' in e._L.jg_last_error()


def test_full_size_properties(robot, scene_loader, engines):
    """BASELINE config 2 at full size (1M configs, scenario 011): size-independent properties."""
    e = engines(11)
    q = uniform_configs(robot, 1_000_000, 11)
    bits, dist, _ = e.check(q, flags=3)
    assert dist.max().item() <= 0.01 + 1e-9
    assert torch.equal(bits, dist < -0.001)                      # verdict == thresholded distance
    # idempotence / determinism
    bits2, dist2, _ = e.check(q, flags=3)
    assert torch.equal(bits, bits2) and torch.equal(dist, dist2)
    # the clamped mode (JG_NO_EPA) only differs where hull cores overlap: same verdicts, distance clamped at the
    # margin sum (-0.002).  Checked on SCENE only: plane distances are exact signed values in both modes.
    bs, ds, _ = e.check(q, flags=1)
    bs4, ds4, _ = e.check(q, flags=1 | 16)
    assert torch.equal(bs4, bs)
    assert (ds4 - torch.clamp(ds, min=-0.002)).abs().max().item() < 1e-6
    # permutation equivariance: results follow their configurations
    perm = torch.randperm(len(q), generator=torch.Generator().manual_seed(0))
    bits3, dist3, _ = e.check(q[perm.numpy()], flags=3)
    assert torch.equal(bits3.cpu(), bits.cpu()[perm]) and torch.equal(dist3.cpu(), dist.cpu()[perm])
    # monotone in the threshold: collides(thr=-0.001) implies collides(thr=0)
    e0 = engines(11, threshold=0.0)
    bits0, _, _ = e0.check(q, flags=3 | 16)
    assert (bits0 | ~bits).all().item()
    # base translation moves nothing but the platform-relative geometry: q and q shifted outside limits still checks
    rate = bits.float().mean().item()
    assert 0.25 < rate < 0.45          # ~0.31 obstacles + ~0.08 plane
    # 1M-prefix check against the oracle on a 20k sample
    o = Oracle(robot, scene_loader(11))
    idx = np.arange(0, 1_000_000, 50)
    rb, rd = o.check(q[idx].astype(np.float64), flags=3)
    compare(bits[torch.from_numpy(idx)], dist[torch.from_numpy(idx)], rb, rd)


def test_batches_larger_than_one_pass(robot, scene_loader, engines):
    """More configurations than the 2^20 pass capacity: the call splits into passes; results equal the per-slice calls and
    the oracle on a sample that straddles the pass boundary.  Same through the host path (its passes are pipelined)."""
    e = engines(11)
    n = (1 << 20) + (1 << 19) + 77
    q = uniform_configs(robot, n, 123)
    bits, dist, _ = e.check(q, flags=3)
    b0, d0, _ = e.check(q[:1 << 20], flags=3)
    b1, d1, _ = e.check(q[1 << 20:], flags=3)
    assert torch.equal(bits, torch.cat([b0, b1])) and torch.equal(dist, torch.cat([d0, d1]))
    idx = np.concatenate([np.arange((1 << 20) - 3000, (1 << 20) + 3000), np.arange(n - 2000, n)])
    rb, rd = Oracle(robot, scene_loader(11)).check(q[idx].astype(np.float64), flags=3)
    compare(bits[torch.from_numpy(idx)], dist[torch.from_numpy(idx)], rb, rd)
    from jogramop_framework_b200.engine import unpack_bits
    wb, wd, _ = e.check_host(q, flags=3)                 # pageable input, two passes in the pipeline
    assert torch.equal(unpack_bits(wb.cuda(), n), bits) and torch.equal(wd.cuda(), dist)


@pytest.mark.parametrize('sid', [s for s in SCENARIO_IDS if s not in (11, 21, 34, 43, 45)])
def test_remaining_scenarios_parity(sid, robot, scene_loader):
    """The 15 scenarios test_scene_parity does not cover: scene + plane + self + limits on 20 000 configurations each."""
    from jogramop_framework_b200.engine import CollisionEngine
    e = CollisionEngine(robot, scene_loader(sid))
    q = uniform_configs(robot, 20000, 1000 + sid)
    fl = oracle.SCENE | oracle.PLANE | oracle.SELF | oracle.LIMITS
    rb, rd = Oracle(robot, scene_loader(sid)).check(q.astype(np.float64), flags=fl)
    bits, dist, _ = e.check(q, flags=fl)
    compare(bits, dist, rb, rd)
    e.close()
