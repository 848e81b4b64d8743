#!/usr/bin/env python
"""Headline benchmark: configuration collision checks/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (SURVEY.md §8d config 2 = BASELINE.json configs[1]): scenario 011, 1 000 000 uniformly sampled mobile-Panda
configurations per GPU (np.random.default_rng(11 + 1000*rank), float32), flags SCENE+PLANE, outputs collision bit +
signed min distance.  One "step" = one pass of the hot path over that batch.  `--lanes` (4) steps are in flight at a
time: the engine pool (engine.EnginePool) hands the steps round-robin to 4 engines (scratch arenas) on 4 streams, so one
batch's kernels run under the tails of the others'; every step is still a full pass over its own 1 M batch, and all K
finish inside the timed region.  `single_lane` reports the same K steps one batch at a time on one stream.  Weak scaling:
every rank checks its own 1 M shards against a replicated scene; the packed verdict bitmasks of every step are gathered
with one NCCL all-gather inside the timed region.

`value`   : device-resident throughput (inputs already in HBM), CUDA events on the launching stream, max over ranks.
`e2e`     : same metric through the host-buffer entry points (jg_check_host_submit / _wait on the pool's engines; the
            blocking jg_check_host beside it): pinned host q -> H2D -> kernels -> D2H, every step.
`roofline`: narrow-phase kernel (the dominant one) against the measured FP32 FMA peak (the binding roof);
            `roofline_hbm` gives the HBM view of the whole pass.
`cpu_baseline`: the double-precision CPU oracle ("port": pybullet is not installable in this image) on the host cores.
`--impl reference`: the same oracle, all host threads, bounded samples, as the reference arm.

Timing (all on the device, max over ranks): ONE CUDA-event pair around the K timed steps on the launching stream, after a
barrier + synchronize; the steps rotate over 4 different 1 M input batches (144 MB > the 126 MB L2; the pass's own queue
traffic evicts the rest), so no step finds its inputs in L2.  With N > 1 the kernels write their packed verdict words
straight into the rank's slot of the all-gather buffer and the all-gather of step k runs on a high-priority side stream
under the kernels of the following steps; the timed region ends when the last gather has finished.  The kernel split
(`kernel_ms_per_step`) and the roofline durations come from the same K steps run one batch at a time right after the
timed region, with the library's section events on.
`soak`    : the one-batch-at-a-time step looped for 2 s with NVML clock samples every 5 ms (sustained SM clock of THIS workload).
`configs` : BASELINE.json configs 3 / 4 / 5 measured after the timed region (N = 1), each with a parity sample.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_CHECK = 1.5e4         # SURVEY §8d nominal algorithmic work per check, scenario 011
FLOPS_PER_CHECK_NARROW = 1.28e4  # ... of which the narrow phase (W_np)
# DRAM traffic per pass comes from profiles/r02_traffic.json (tools/ncu_summary.py --traffic on an `ncu --set full` capture
# of this very command; the file names the commit it was taken at) -- never a constant in this file.
TRAFFIC_FILE = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
BYTES_PER_CHECK = 40.125        # 36 B in + 4 B distance + 1 bit out
SCENARIO_IDS = [11, 12, 13, 14, 15, 21, 22, 23, 24, 25, 31, 32, 33, 34, 35, 41, 42, 43, 44, 45]   # util.py:10


def load_models(scenario):
    from jogramop_framework_b200.ingest import RobotModel, SceneModel
    data = os.path.join(ROOT, 'jogramop_framework_b200', 'data')
    return (RobotModel.load(os.path.join(data, 'robot_mobile_panda.npz')),
            SceneModel.load(os.path.join(data, f'scene_{scenario:03d}.npz')))


def make_configs(robot, n, seed):
    lim = robot.joint_limits[robot.arm_joint_ids]
    rng = np.random.default_rng(seed)
    return (lim[:, 0] + (lim[:, 1] - lim[:, 0]) * rng.random((n, len(lim)))).astype(np.float32)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML every `period` seconds while a timed region runs."""

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        self.period = period
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = index
            if visible and len(visible.split(',')) > index and visible.split(',')[index].strip().isdigit():
                phys = int(visible.split(',')[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples),
                'sm_mhz_min': float(min(self.samples)) if self.samples else None, 'period_ms': 1e3 * self.period}


def time_oracle(robot, scene, q, flags, n_threads, budget_s):
    """CPU oracle on a bounded prefix of q sized for ~budget_s seconds -> (checks/s, n_used, threads)."""
    import oracle
    o = oracle.Oracle(robot, scene)
    probe = min(len(q), 20000)
    t0 = time.perf_counter()
    o.check(q[:probe].astype(np.float64), flags=flags, n_threads=n_threads)
    rate = probe / (time.perf_counter() - t0)
    n = int(min(len(q), max(probe, rate * budget_s)))
    t0 = time.perf_counter()
    rb, rd = o.check(q[:n].astype(np.float64), flags=flags, n_threads=n_threads)
    dt = time.perf_counter() - t0
    return n / dt, n, (n_threads if n_threads > 0 else os.cpu_count()), rb, rd


def run_reference(args):
    """Reference arm: the CPU implementation of the path on the host cores (oracle port; pybullet not installable)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    robot, scene = load_models(args.scenario)
    import oracle
    flags = oracle.SCENE | oracle.PLANE
    sample = args.ref_sample
    q = make_configs(robot, sample, 11)
    o = oracle.Oracle(robot, scene)
    cores = os.cpu_count()
    for _ in range(args.warmup):
        o.check(q[:max(1000, sample // 10)].astype(np.float64), flags=flags, n_threads=0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.check(q.astype(np.float64), flags=flags, n_threads=0)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        'impl': 'reference', 'metric': 'configuration collision checks/sec', 'value': value, 'unit': 'checks/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'scenario {args.scenario:03d}: uniformly sampled mobile-Panda configurations, '
                               'collision bit + min distance (SCENE+PLANE)',
                   'sample_per_step': sample, 'seed': 11},
        'cpu_baseline': {'value': value, 'unit': 'checks/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{sample} configurations per step (prefix of the 1M workload), '
                                   'CPU oracle restatement - pyBullet not installable in this image'},
        'e2e': {'value': value, 'unit': 'checks/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def random_poses(rng, lo, hi, m):
    pos = (lo + (hi - lo) * rng.random((m, 3))).astype(np.float32)
    quat = rng.normal(size=(m, 4)).astype(np.float32)
    return np.concatenate([pos, quat], 1)


def quat_poses_to_mat(pq):
    p = pq[:, :3].astype(np.float64)
    qn = pq[:, 3:].astype(np.float64)
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    x, y, z, w = qn.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w),
                  1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w),
                  1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
    return np.concatenate([R, p[:, :, None]], 2)


def bench_other_configs(robot, dev, budget_s=1.0):
    """BASELINE.json configs 3, 4 and 5 on ONE GPU, outside the headline's timed region: CUDA events around `reps`
    repetitions after warm-up, plus a parity sample against the CPU oracle for each (the checker, never the timed path)."""
    import torch
    import oracle
    from jogramop_framework_b200.engine import CollisionEngine, SCENE, PLANE
    from jogramop_framework_b200.ingest import SceneModel
    data = os.path.join(ROOT, 'jogramop_framework_b200', 'data')
    LANES = 4
    lim = robot.joint_limits[robot.arm_joint_ids]
    out = {}

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

    def parity(gb, gd, rb, rd):
        gb = gb.cpu().numpy().astype(bool)
        outside = np.abs(rd + 0.001) > 1e-3
        r = {'checked': int(len(rb)), 'verdict_mismatches_outside_1mm_band': int((gb[outside] != rb[outside]).sum()),
             'verdict_mismatches_inside_band': int((gb[~outside] != rb[~outside]).sum())}
        if gd is not None:
            r['max_abs_distance_error_m'] = float(np.abs(gd.cpu().numpy().astype(np.float64) - rd).max())
        return r

    # ---- config 3: every scenario, 1 M configurations each on this GPU (the full-size run is 10 M each over 8 GPUs:
    #      tools/bench_sharded.py; per-GPU work is the same per configuration)
    n3 = 1_000_000
    ms3, rates = 0.0, {}
    for sid in SCENARIO_IDS:
        scene = SceneModel.load(os.path.join(data, f'scene_{sid:03d}.npz'))
        eng = CollisionEngine(robot, scene, device=dev.index)
        q = torch.from_numpy(make_configs(robot, n3, sid)).to(dev)
        ms = timed(lambda: eng.check(q, flags=SCENE | PLANE, packed=True), reps=2, warm=2)
        ms3 += ms
        rates[f'{sid:03d}'] = n3 / ms * 1e3
        eng.close()
        del eng, q
    out['config3_all_20_scenarios'] = {
        'workload': f'{len(SCENARIO_IDS)} scenarios x {n3} uniform configurations (seed = scenario id), bit + min distance, 1 GPU',
        'checks_per_s': len(SCENARIO_IDS) * n3 / ms3 * 1e3, 'ms_total': ms3,
        'slowest_scenario_checks_per_s': min(rates.values()), 'fastest_scenario_checks_per_s': max(rates.values()),
        'parity': 'tests/test_gpu_sweeps.py::test_sweep_all_flags_every_scenario (50 000 per scenario vs the oracle)'}

    # ---- config 4: scenario 034, 1 M edges x 64 substeps
    scene = SceneModel.load(os.path.join(data, 'scene_034.npz'))
    eng = CollisionEngine(robot, scene, device=dev.index)
    m = 1_000_000
    qa = make_configs(robot, m, 34)
    qb = np.clip(qa + np.random.default_rng(3400).uniform(-0.25, 0.25, qa.shape), lim[:, 0], lim[:, 1]).astype(np.float32)
    qa_d, qb_d = torch.from_numpy(qa).to(dev), torch.from_numpy(qb).to(dev)
    ms_d = timed(lambda: eng.check_edges(qa_d, qb_d, substeps=64, packed=True))
    ms_v = timed(lambda: eng.check_edges(qa_d, qb_d, substeps=64, packed=True, want_dist=False))
    k = 2000
    o = oracle.Oracle(robot, scene)
    rb, rd = o.check_edges(qa[:k].astype(np.float64), qb[:k].astype(np.float64), substeps=64)
    gb, gd = eng.check_edges(qa_d[:k], qb_d[:k], substeps=64)
    gbv, _ = eng.check_edges(qa_d[:k], qb_d[:k], substeps=64, want_dist=False)
    bits_all, dist_all = eng.check_edges(qa_d, qb_d, substeps=64)
    eng.close()
    # the same call with its 62 passes side by side on four arenas / internal streams (jg_engine_set_lanes)
    pool = CollisionEngine(robot, scene, device=dev.index, lanes=LANES)
    ms_dp = timed(lambda: pool.check_edges(qa_d, qb_d, substeps=64, packed=True))
    ms_vp = timed(lambda: pool.check_edges(qa_d, qb_d, substeps=64, packed=True, want_dist=False))
    pb, pd = pool.check_edges(qa_d, qb_d, substeps=64)
    pooled4 = {'lanes': LANES, 'api': 'jg_engine_set_lanes(4) + one jg_check_edges call', 'verdict_and_distance_ms': ms_dp, 'verdict_and_distance_substep_checks_per_s': m * 64 / ms_dp * 1e3,
               'verdict_only_ms': ms_vp, 'verdict_only_substep_checks_per_s': m * 64 / ms_vp * 1e3,
               'identical_to_one_engine': bool(torch.equal(pb, bits_all) and torch.equal(pd, dist_all))}
    pool.close()
    del pb, pd, dist_all
    out['config4_edges'] = {
        'workload': 'scenario 034: 1 000 000 edges x 64 interpolated substeps (qa seed 34, qb = clip(qa + U(-0.25, 0.25)) seed 3400); '
                    'checks counted = edges x 64, no early-exit credit',
        'verdict_and_distance': {'ms': ms_d, 'substep_checks_per_s': m * 64 / ms_d * 1e3, 'edges_per_s': m / ms_d * 1e3,
                                 'parity': parity(gb, gd, rb, rd)},
        'verdict_only': {'ms': ms_v, 'substep_checks_per_s': m * 64 / ms_v * 1e3, 'edges_per_s': m / ms_v * 1e3,
                         'parity': parity(gbv, None, rb, rd)},
        'passes_side_by_side': pooled4,
        'edge_collision_rate': float(bits_all.float().mean().item())}
    del qa_d, qb_d

    # ---- config 5: scenario 011, 10 M free-floating gripper poses
    scene = SceneModel.load(os.path.join(data, 'scene_011.npz'))
    eng = CollisionEngine(robot, scene, device=dev.index)
    m = 10_000_000
    tgt = scene.piece(0)
    lo, hi = tgt.min(0) - 0.25, tgt.max(0) + 0.25
    pq = random_poses(np.random.default_rng(5011), lo, hi, m)
    pq_d = torch.from_numpy(pq).to(dev)
    links = [robot.end_effector_id - 3, *robot.finger_joint_ids]
    ms_p = timed(lambda: eng.check_poses(pq_d, links=links, packed=True))
    k = 20000
    T = quat_poses_to_mat(pq[:k])
    o = oracle.Oracle(robot, scene)
    rb, rd = o.check_poses(T, links=links, flags=3)
    gb, gd = eng.check_poses(pq_d[:k], links=links)
    bits_all, dist_all = eng.check_poses(pq_d, links=links)
    eng.close()
    pool = CollisionEngine(robot, scene, device=dev.index, lanes=LANES)
    ms_pp = timed(lambda: pool.check_poses(pq_d, links=links, packed=True))
    pb, pd = pool.check_poses(pq_d, links=links)
    pooled5 = {'lanes': LANES, 'api': 'jg_engine_set_lanes(4) + one jg_check_poses call', 'ms': ms_pp, 'poses_per_s': m / ms_pp * 1e3,
               'identical_to_one_engine': bool(torch.equal(pb, bits_all) and torch.equal(pd, dist_all))}
    pool.close()
    del pb, pd, dist_all
    c5 = {'workload': 'scenario 011: 10 000 000 free-floating gripper poses (position uniform in the target box + 0.25 m, '
                      'uniform rotations, seed 5011) vs target + obstacles + plane',
          'panda_hand_hulls': {'model': 'panda_hand + two finger hulls at 0.04, Bullet margins, convex pieces, bit + signed distance',
                               'ms': ms_p, 'poses_per_s': m / ms_p * 1e3, 'collision_rate': float(bits_all.float().mean().item()),
                               'parity': parity(gb, gd, rb, rd), 'passes_side_by_side': pooled5}}
    # the reference's own filter (create_graspset.py:53-59): burg two-finger gripper vs the visual meshes, boolean
    try:
        from jogramop_framework_b200.scenario import Scenario
        from jogramop_framework_b200.simulation import GraspingSimulator
        from jogramop_framework_b200.graspset import burg_gripper_model
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            scn = Scenario(11, device=dev.index)
        vis = scn.visual_scene_model()
        sim = GraspingSimulator(device=dev.index)
        mb = 2_000_000
        Tw = scn.robot_pose @ np.concatenate([quat_poses_to_mat(pq[:mb]), np.tile([[[0, 0, 0, 1.0]]], (mb, 1, 1))], 1)
        sim.burg_gripper_collision_batch(Tw[:1000], visual_model=vis)        # builds the engine + LBVH
        geng = next(iter(sim._gripper_engines.values()))
        Tr = torch.from_numpy(np.ascontiguousarray(quat_poses_to_mat(pq[:mb]).astype(np.float32))).to(dev)
        ms_b = timed(lambda: geng.check_poses(Tr, links=[0], ee_link=0, want_dist=False, packed=True), reps=2, warm=1)
        hit, _ = geng.check_poses(Tr[:5000], links=[0], ee_link=0, want_dist=False)
        import copy
        tri = copy.copy(vis)
        tri.piece_verts = vis.tri_verts[vis.tris].reshape(-1, 3)
        tri.piece_vert_offset = (np.arange(len(vis.tris) + 1) * 3).astype(np.int32)
        tri.piece_body, tri.piece_margin = vis.tri_body, np.zeros(len(vis.tris))
        og = oracle.Oracle(burg_gripper_model(), tri)
        _, dd = og.check_poses(quat_poses_to_mat(pq[:5000]), links=[0], ee_link=0,
                               flags=oracle.SCENE | oracle.PLANE | oracle.NO_EPA, threshold=1e-9)
        hb, rbg = hit.cpu().numpy().astype(bool), dd <= 1e-9
        touching = (dd > 1e-9) & (dd <= 2e-5)          # clear of every surface by less than 20 um: either verdict is fine
        c5['burg_gripper_visual_meshes'] = {
            'model': 'burg two-finger gripper (3 zero-margin boxes) vs the mesh_fn meshes behind the GPU-built LBVH '
                     f'({len(vis.tris)} triangles) + plane, boolean (the reference filter); {mb} of the 10 M poses',
            'ms': ms_b, 'poses_per_s': mb / ms_b * 1e3, 'collision_rate': float(hb.mean()),
            'parity': {'checked': 5000, 'verdict_mismatches': int((hb != rbg).sum()),
                       'verdict_mismatches_outside_20um_band': int((hb != rbg)[~touching].sum()),
                       'poses_within_20um_of_touching': int(touching.sum())}}
        sim.dismiss()
    except Exception as exc:
        c5['burg_gripper_visual_meshes'] = {'unavailable': repr(exc)[:300]}
    out['config5_gripper_poses'] = c5

    # ---- the headline workload on the simplified geometry profile (mobile_panda_fingersSmallMesh.urdf: boxes around the
    #      links, what the sister planners load; not the pyBullet model of the reference)
    try:
        from jogramop_framework_b200.ingest import RobotModel
        small = RobotModel.load(os.path.join(data, 'robot_mobile_panda_simplified.npz'))
        scene = SceneModel.load(os.path.join(data, 'scene_011.npz'))
        eng = CollisionEngine(small, scene, device=dev.index)
        qh = make_configs(small, 1_000_000, 11)
        q = torch.from_numpy(qh).to(dev)
        ms_s = timed(lambda: eng.check(q, flags=SCENE | PLANE, packed=True), reps=5, warm=3)
        rb, rd = oracle.Oracle(small, scene).check(qh[:20000].astype(np.float64), flags=3)
        gb, gd, _ = eng.check(q[:20000], flags=SCENE | PLANE)
        out['simplified_geometry_profile'] = {
            'workload': 'scenario 011, 1 000 000 uniform configurations, bit + signed min distance, robot = mobile_panda_fingersSmallMesh.urdf '
                        '(13 shapes, 204 hull vertices instead of 1 370)',
            'ms': ms_s, 'checks_per_s': 1e6 / ms_s * 1e3, 'parity': parity(gb, gd, rb, rd)}
        eng.close()
    except Exception as exc:
        out['simplified_geometry_profile'] = {'unavailable': repr(exc)[:300]}
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from jogramop_framework_b200.engine import EnginePool, SCENE, PLANE, NO_EPA, measure_fp32_peak

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device - the engine has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    bound = None
    if not os.environ.get('JG_NO_AFFINITY'):
        from jogramop_framework_b200.dist import bind_to_gpu_cpus
        bound = bind_to_gpu_cpus(local)    # pinned staging buffers of the host path end up on the GPU's NUMA node
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # the collective's kernels get a high-priority stream: with several batches in flight the SMs are never idle, and a
        # pending all-gather block must be the first thing scheduled when a persistent block retires
        # ... and it is a 1 MB collective: two channels move it as fast as NCCL's default channel count, and its spinning
        # CTAs then hold two SMs instead of a dozen while the ranks meet (profiles/r02_scale_probe_8gpu.txt: the gather
        # costs 1.5 % of the 8-GPU step with the default channels, 0.4 % with two)
        os.environ.setdefault('NCCL_MAX_NCHANNELS', '2')
        pg_opts = None
        if not os.environ.get('JG_BENCH_NO_PRIO'):
            pg_opts = dist.ProcessGroupNCCL.Options()
            pg_opts.is_high_priority_stream = True
        dist.init_process_group('nccl', device_id=dev, pg_options=pg_opts)
    robot, scene = load_models(args.scenario)
    flags = SCENE | PLANE | (NO_EPA if args.no_epa else 0)
    # --strong: the SAME total batch (args.n configurations) cut into 32-aligned shards, one per GPU (builder-run record of
    # where the fixed per-pass costs start to show; the driver's scaling run is the weak form)
    n = args.n if not args.strong else ((args.n + world - 1) // world + 31) // 32 * 32
    ROT = 4                                    # rotating input batches: 4 x 36 MB > L2
    q_hosts = [make_configs(robot, n, 11 + 1000 * rank + 7919 * r) for r in range(ROT)]
    q_host = q_hosts[0]
    # args.lanes whole batches in flight: one engine (scratch arena) + stream per lane, steps handed out round-robin
    # (EnginePool); the kernels of one batch fill the tails and idle issue slots of the others
    lanes = max(1, args.lanes)
    pool = EnginePool(robot, scene, device=local, lanes=lanes, chunk=args.chunk)
    eng = pool.engines[0]
    q_devs = [torch.from_numpy(q).to(dev) for q in q_hosts]
    q_dev = q_devs[0]
    words = (n + 31) // 32
    # the receive buffers of the all-gather: a buffer takes the words of G consecutive steps (--gather-every, default 1: one
    # collective per step) and enough buffers rotate that none is rewritten while its gather or a batch in flight uses it;
    # the kernels write straight into this rank's part
    G = max(1, args.gather_every)
    NB = (lanes + G - 1) // G + 1
    gathered = [torch.zeros((world * G * words,), dtype=torch.int32, device=dev) for _ in range(NB)]
    mine = [g[rank * G * words:(rank + 1) * G * words] for g in gathered]
    slots = [m[:words] for m in mine]
    dist_outs = [torch.empty((n,), dtype=torch.float32, device=dev) for _ in range(lanes)]
    dist_out = dist_outs[0]
    side = torch.cuda.Stream(device=dev, priority=0 if os.environ.get('JG_BENCH_NO_PRIO') else -1)
    ev_gather = [torch.cuda.Event() for _ in range(NB)]
    gather_pending = [False] * NB
    group = {'events': [], 'buf': 0}

    fp32_peak = measure_fp32_peak(local) if rank == 0 else None

    def flush():
        b = group['buf']
        if world > 1 and group['events'] and not os.environ.get('JG_BENCH_NO_GATHER'):
            for ev in group['events']:
                side.wait_event(ev)
            with torch.cuda.stream(side):      # the all-gather runs under the kernels of the following steps
                dist.all_gather_into_tensor(gathered[b], mine[b])
                ev_gather[b].record(side)
            gather_pending[b] = True
        group['events'] = []

    def step(i):
        b, j = (i // G) % NB, i % G
        main = torch.cuda.current_stream(dev)
        if j == 0:
            if group['events']:
                flush()                        # (a partial group left by the warm-up)
            group['buf'] = b
            if gather_pending[b]:
                main.wait_event(ev_gather[b])  # the gather that last read this buffer: long finished
        tk = pool.submit(q_devs[i % ROT], flags=flags, packed=True, bits_out=mine[b][j * words:(j + 1) * words],
                         dist_out=dist_outs[pool.next_lane])
        group['events'].append(tk.event)
        if j == G - 1:
            flush()

    def serial_step(i):                         # one batch at a time on the current stream (kernel split, single-lane figure)
        eng.check(q_devs[i % ROT], flags=flags, packed=True, bits_out=slots[0], dist_out=dist_out)

    def drain():
        flush()
        pool.drain()
        torch.cuda.current_stream(dev).wait_stream(side)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_warm = max(args.warmup, 3, lanes)        # every lane's arena has grown to its working size before the timed region
    for i in range(n_warm):
        step(i)
    drain()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.perf_counter()
    ev0.record()
    for s in range(args.steps):
        step(s)
    drain()                                     # the timed region ends when the last all-gather has finished
    ev1.record()
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    stats = eng.stats()
    # single-lane figure: the same K steps, one batch at a time on one stream (no collective: per-GPU kernel time)
    torch.cuda.synchronize()
    sv0, sv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sv0.record()
    for s in range(args.steps):
        serial_step(s)
    sv1.record()
    torch.cuda.synchronize()
    serial_ms = sv0.elapsed_time(sv1) / args.steps
    # kernel split / roofline durations: the SAME K steps again, one batch at a time, with the library's per-section
    # CUDA events on the launching stream.  (The events are not free -- 8 records per pass cost 2.5 % of the step --
    # and sections of overlapping batches would time each other, so the timed region itself carries none.)
    eng.profile(True)
    for s in range(args.steps):
        serial_step(s)
    barrier()
    prof = eng.get_profile()
    eng.profile(False)
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * n * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host-buffer entry point
    q_pins = [torch.from_numpy(q).pin_memory() for q in q_hosts]
    q_pin = q_pins[0]
    out = {'bits': torch.empty((words,), dtype=torch.int32, pin_memory=True),
           'dist': torch.empty((n,), dtype=torch.float32, pin_memory=True), 'reason': None}
    for _ in range(2):
        eng.check_host(q_pin, flags=flags, out=out)
    barrier()
    t0 = time.perf_counter()
    for s_i in range(args.steps):
        eng.check_host(q_pins[s_i % ROT], flags=flags, out=out)      # returns when the results are in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / float(te.item())
    # the same steps with several batches in flight (jg_check_host_submit / _wait on every lane's engine): what a caller
    # streaming batches gets -- every step still uploads its inputs from pinned host memory and downloads its results
    # inside the timed region
    depth = max(2, lanes * args.e2e_per_lane)
    outs = [out] + [{'bits': torch.empty((words,), dtype=torch.int32, pin_memory=True),
                     'dist': torch.empty((n,), dtype=torch.float32, pin_memory=True), 'reason': None} for _ in range(depth - 1)]
    from collections import deque

    def e2e_loop(k_steps):
        tks = deque()
        for s_i in range(k_steps):
            if len(tks) == depth:
                pool.wait_host(tks.popleft())
            tks.append(pool.submit_host(q_pins[s_i % ROT], flags=flags, out=outs[s_i % depth]))
        while tks:
            pool.wait_host(tks.popleft())
    e2e_loop(2 * depth)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    e2e_pipe_s = time.perf_counter() - t0
    tp = torch.tensor([e2e_pipe_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    e2e_pipe_value = world * n * args.steps / float(tp.item())
    launches_per_step = stats['kernel_launches']

    if rank == 0:
        np_ms, np_n = prof['narrowphase']
        bp_ms, bp_n = prof['broadphase']
        fin_ms, fin_n = prof['finalize']
        epa_ms, epa_n = prof['penetration']
        checks_timed = n * args.steps
        # the narrow phase = the GJK stages + the penetration-depth stages (W_np); one "launch" = one pass
        nar_ms = np_ms + epa_ms
        ach_tf = checks_timed * FLOPS_PER_CHECK_NARROW / (nar_ms * 1e-3) / 1e12 if nar_ms > 0 else None
        peaks = {}
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        traffic = None
        try:
            with open(TRAFFIC_FILE) as fh:
                traffic = json.load(fh)
        except Exception:
            pass
        use_traffic = traffic is not None and args.scenario == 11 and not args.no_epa and n == traffic.get('configs_per_pass')
        step_tf = value / world * FLOPS_PER_CHECK / 1e12
        line = {
            'metric': 'configuration collision checks/sec', 'value': value, 'unit': 'checks/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': n_warm, 'ms_per_step': ms_total / args.steps,
            'higher_is_better': True, 'scaling': 'strong' if args.strong else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'scenario {args.scenario:03d}: {n} uniformly sampled mobile-Panda configurations per GPU, '
                                   'collision bit + signed min distance (SCENE+PLANE)',
                       'configs_per_gpu': n, 'seed': '11 + 1000*rank (+ 7919*r for the rotating batches r = 1..3)', 'flags': int(flags),
                       'penetration_depth': 'clamped at -(marginA+marginB) (JG_NO_EPA)' if args.no_epa else 'EPA',
                       'l2': f'inputs larger than L2: the steps rotate over {ROT} different {n}-configuration batches '
                             f'({ROT * n * 36 / 1e6:.0f} MB > 126 MB L2); one CUDA-event pair around all timed steps',
                       'collective': ('ncclAllGather of the packed verdict bits, in place (the kernels write into the rank\'s slot), '
                                      f'one per {G} step(s), NCCL_MAX_NCHANNELS={os.environ.get("NCCL_MAX_NCHANNELS")}, on a high-priority side stream under the following steps\' kernels; '
                                      'the timed region ends after the last gather')
                       if world > 1 else 'none (1 GPU)',
                       'in_flight': f'{lanes} batches in flight: EnginePool, one engine (scratch arena) and CUDA stream per lane, steps '
                                    'handed out round-robin; every step is a full pass over its own batch, all K inside the timed region',
                       'pass_size': args.chunk or 'auto (<= 2^20 configurations per pass)',
                       'cpu_affinity': f'process bound to the {bound} CPUs local to its GPU (NVML)' if bound else 'not set'},
            'clocks': clocks,
            'single_lane': {'value': n / serial_ms * 1e3, 'unit': 'checks/s per GPU', 'ms_per_step': serial_ms,
                            'note': 'one batch at a time on one stream (no overlap between batches, no collective); '
                                    'kernel_ms_per_step and roofline durations are taken in this form'},
            'e2e': {'value': e2e_pipe_value, 'unit': 'checks/s', 'h2d_bytes_per_step': int(n * 36),
                    'd2h_bytes_per_step': int(n * 4 + words * 4),
                    'api': f'jg_check_host_submit / jg_check_host_wait on {lanes} engines, {depth} batches in flight (pinned host buffers)',
                    'synchronous_call_value': e2e_value,
                    'synchronous_api': 'jg_check_host: one blocking call per step (copies overlap only within the batch)'},
            'gpu_launches': int(launches_per_step * args.steps),
            'roofline': {'bound': 'fp32', 'kernel': 'narrow phase (k_narrowphase GJK stages + k_penetration stages)',
                         'achieved': ach_tf, 'peak': fp32_peak,
                         'unit': 'TFLOP/s', 'frac': (ach_tf / fp32_peak) if (ach_tf and fp32_peak) else None,
                         'traffic': traffic.get('whole_pass_bytes') if use_traffic else None,
                         'traffic_detail': ({'source': 'profiles/r02_traffic.json: ncu --set full, dram__bytes_read.sum + '
                                                       'dram__bytes_write.sum summed over ALL kernels of one pass',
                                             'commit': traffic.get('commit'), 'per_kernel_bytes': traffic.get('per_kernel_bytes'),
                                             'narrow_phase_bytes': traffic.get('narrow_phase_bytes'),
                                             'algorithmic_bytes': BYTES_PER_CHECK * n,
                                             'ratio_to_algorithmic': traffic.get('whole_pass_bytes') / (BYTES_PER_CHECK * n)}
                                            if use_traffic else None),
                         'peak_source': 'jg_measure_fp32_peak FMA micro-benchmark on this GPU',
                         'algorithmic_flops_per_check': FLOPS_PER_CHECK_NARROW,
                         'avg_launch_ms': nar_ms / np_n if np_n else None, 'launches': int(np_n),
                         'duration_source': 'per-section CUDA events of the library over the same K steps, run directly after the timed '
                                            'region (the timed region carries no events between kernels: they cost 2.5 % of a step)',
                         'whole_step': {'flops_per_check': FLOPS_PER_CHECK, 'achieved': step_tf,
                                        'frac': step_tf / fp32_peak if fp32_peak else None,
                                        'note': 'SURVEY 8(d): checks/s per GPU x 1.5e4 nominal flops / FP32 peak'},
                         'note': 'algorithmic = nominal work of an unpruned narrow phase; the champion-first pruning '
                                 'executes far fewer flops, so frac is work-avoided + pipe-utilisation combined; '
                                 'counted.executed_frac is the pipe utilisation alone'},
            'roofline_hbm': {'bound': 'hbm', 'achieved': checks_timed * BYTES_PER_CHECK / (ms_total * 1e-3) / 1e9,
                             'peak': hbm_peak, 'unit': 'GB/s',
                             'frac': checks_timed * BYTES_PER_CHECK / (ms_total * 1e-3) / 1e9 / hbm_peak,
                             'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback'},
            'kernel_ms_per_step': {'broadphase': bp_ms / args.steps, 'narrowphase': np_ms / args.steps,
                                   'penetration': epa_ms / args.steps, 'finalize': fin_ms / args.steps},
            'work': {'pairs_per_check': stats['pairs'] / max(stats['configs'], 1),
                     'plane_pairs_per_check': stats['plane_pairs'] / max(stats['configs'], 1),
                     'overlapping_pairs_per_check': stats['epa_pairs'] / max(stats['configs'], 1),
                     'gjk_run_per_check': stats['gjk_run'] / max(stats['configs'], 1),
                     'polytopes_expanded_per_check': stats['epa_run'] / max(stats['configs'], 1)},
            'wall_s_timed_region': wall,
        }
        if world == 1 and not args.no_extras:
            # ---- the pass as ONE CUDA-graph launch (engine.capture_check), same rotating inputs
            try:
                graphs = [eng.capture_check(q, flags=flags) for q in q_devs]
                for g in graphs:
                    g[0]()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for s_i in range(args.steps):
                    graphs[s_i % ROT][0]()
                b.record()
                torch.cuda.synchronize()
                gms = a.elapsed_time(b) / args.steps
                same = torch.equal(graphs[0][1], eng.check(q_devs[0], flags=flags, packed=True)[0])
                line['cuda_graph'] = {'value': n / gms * 1e3, 'ms_per_step': gms, 'api': 'CollisionEngine.capture_check: '
                                      'the pass replayed as one graph launch', 'bits_identical_to_plain_launches': bool(same)}
                del graphs
            except Exception as exc:
                line['cuda_graph'] = {'unavailable': repr(exc)[:200]}
            # ---- soak: the same step looped for 2 s, NVML clock samples every 5 ms
            soak = ClockSampler(local)
            soak.start()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            k = 0
            while time.perf_counter() - t0 < args.soak:
                for _ in range(20):
                    eng.check(q_devs[k % ROT], flags=flags, packed=True, bits_out=slots[0], dist_out=dist_out)
                    k += 1
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            sc = soak.stop()
            line['soak'] = {'seconds': dt, 'steps': k, 'value': k * n / dt, 'unit': 'checks/s (wall clock incl. one sync per 20 steps)',
                            'clocks': sc}
        if world == 1 and not args.no_count:
            # audit of the nominal figure: executed work counted by the instrumented twin of the library, in a
            # subprocess after the timed region (never timed, never the shipped path)
            try:
                import subprocess
                env = dict(os.environ, JG_B200_LIB_VARIANT='count', CUDA_VISIBLE_DEVICES=os.environ.get('CUDA_VISIBLE_DEVICES', str(local)))
                out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'count_work.py'), '--scenario', str(args.scenario),
                                      '--n', str(n), '--seed', '11', '--flags', str(int(flags))],
                                     capture_output=True, text=True, timeout=600, env=env)
                cw = json.loads(out.stdout.strip().splitlines()[-1])
                ex_tf = checks_timed * cw['counted_narrow_flops_per_check'] / (nar_ms * 1e-3) / 1e12
                line['roofline']['counted'] = {
                    'executed_flops_per_check': cw['counted_narrow_flops_per_check'],
                    'support_dots_per_check': cw['support_dots_per_check'],
                    'gjk_iterations_per_check': cw['gjk_iterations_per_check'],
                    'polytope_expansions_per_check': cw['polytope_expansions_per_check'],
                    'executed_tflops': ex_tf, 'executed_frac': ex_tf / fp32_peak if fp32_peak else None,
                    'flops_per_unit': cw['flops_per_unit'],
                    'source': 'libjogramop_b200_count.so (-DJG_COUNT_WORK) on the same inputs, outside the timed region'}
            except Exception as exc:   # the audit is optional evidence; the bench line stands without it
                line['roofline']['counted'] = {'unavailable': repr(exc)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            oflags = oracle.SCENE | oracle.PLANE | (oracle.NO_EPA if args.no_epa else 0)
            v, used, cores, rb, rd = time_oracle(robot, scene, q_host, oflags, 0, args.cpu_budget)
            line['cpu_baseline'] = {'value': v, 'unit': 'checks/s', 'cores': cores, 'kind': 'port',
                                    'sample': f'first {used} of the {n} configurations, all host threads; CPU oracle '
                                              'restatement (double precision) - pyBullet not installable in this image; '
                                              'the oracle itself is pinned by the reference\'s golden artefacts '
                                              '(known-negatives only, see BASELINE.md)'}
            # the oracle's answers for that sample are the checker of the timed path (not timed, not shipped)
            gb, gd, _ = eng.check(q_dev[:used], flags=flags)
            gb, gd = gb.cpu().numpy().astype(bool), gd.cpu().numpy().astype(np.float64)
            outside = np.abs(rd + 0.001) > 1e-3
            line['parity'] = {'checked': int(used), 'against': 'CPU oracle (double precision), same inputs',
                              'verdict_mismatches_outside_1mm_band': int((gb[outside] != rb[outside]).sum()),
                              'verdict_mismatches_inside_band': int((gb[~outside] != rb[~outside]).sum()),
                              'configs_inside_band': int((~outside).sum()),
                              'max_abs_distance_error_m': float(np.abs(gd - rd).max()),
                              'collision_rate': float(rb.mean())}
        if world == 1 and not args.no_extras:
            pool.close()
            del q_devs, q_pins
            torch.cuda.empty_cache()
            try:
                line['configs'] = bench_other_configs(robot, dev)
            except Exception as exc:
                line['configs'] = {'unavailable': repr(exc)[:300]}
        print(json.dumps(line), flush=True)
    pool.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--scenario', type=int, default=11)
    ap.add_argument('--n', type=int, default=1_000_000)
    ap.add_argument('--chunk', type=int, default=0)
    ap.add_argument('--lanes', type=int, default=4, help='batches in flight (engines / streams of the EnginePool)')
    ap.add_argument('--gather-every', type=int, default=1, help='steps whose verdict words share one all-gather (N > 1)')
    ap.add_argument('--e2e-per-lane', type=int, default=1, help='host-path batches in flight per lane (1 or 2)')
    ap.add_argument('--no-epa', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-count', action='store_true', help='skip the counted-work audit (instrumented library)')
    ap.add_argument('--no-extras', action='store_true', help='skip the CUDA-graph line, the soak and the configs 3/4/5 block')
    ap.add_argument('--soak', type=float, default=2.0, help='seconds of the soak loop (sustained clocks)')
    ap.add_argument('--strong', action='store_true', help='strong scaling: --n is the TOTAL batch, split over the GPUs')
    ap.add_argument('--cpu-budget', type=float, default=10.0, help='seconds of CPU work for the cpu_baseline sample')
    ap.add_argument('--ref-sample', type=int, default=200_000, help='configurations per step of the reference arm')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.gpus > 1 and 'WORLD_SIZE' not in os.environ:
        # convenience: re-launch under torchrun
        import subprocess
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', '29511', os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_b200(args)


if __name__ == '__main__':
    main()
