#!/usr/bin/env python
"""Headline benchmark: configuration collision checks/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (SURVEY.md §8d config 2 = BASELINE.json configs[1]): scenario 011, 1 000 000 uniformly sampled mobile-Panda
configurations per GPU (np.random.default_rng(11 + 1000*rank), float32), flags SCENE+PLANE, outputs collision bit +
signed min distance.  One "step" = one pass of the hot path over that batch.  Weak scaling: every rank checks its own
1 M shard against a replicated scene; the packed verdict bitmasks are gathered with one NCCL all-gather inside the
timed region.

`value`   : device-resident throughput (inputs already in HBM), CUDA events on the launching stream, max over ranks.
`e2e`     : same metric through the host-buffer entry point (jg_check_host): pinned host q -> H2D -> kernels -> D2H.
`roofline`: narrow-phase kernel (the dominant one) against the measured FP32 FMA peak (the binding roof);
            `roofline_hbm` gives the HBM view of the whole pass.
`cpu_baseline`: the double-precision CPU oracle ("port": pybullet is not installable in this image) on the host cores.
`--impl reference`: the same oracle, all host threads, bounded samples, as the reference arm.
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
# DRAM bytes (read + write) of the narrow-phase kernels of ONE pass over 1 M configurations of scenario 011, from the
# ncu --set full capture profiles/r01_m_ncu_full_summary.csv (2 x k_narrowphase, 2 x k_penetration, 3 x k_select,
# k_penetration_big): 425.6 MB = 426 B per check, ~10x the algorithmic 40 B (pair queues), 4 % of HBM bandwidth
NARROW_TRAFFIC_BYTES_PER_CHECK_NCU = 395.2
BYTES_PER_CHECK = 40.125        # 36 B in + 4 B distance + 1 bit out


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
    """Samples SM clock / throttle reasons of one GPU every 100 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
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
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


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


def run_b200(args):
    import torch
    import torch.distributed as dist
    from jogramop_framework_b200.engine import CollisionEngine, SCENE, PLANE, NO_EPA, measure_fp32_peak

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
        dist.init_process_group('nccl', device_id=dev)
    robot, scene = load_models(args.scenario)
    flags = SCENE | PLANE | (NO_EPA if args.no_epa else 0)
    n = args.n
    q_host = make_configs(robot, n, 11 + 1000 * rank)
    eng = CollisionEngine(robot, scene, device=local, chunk=args.chunk)
    q_dev = torch.from_numpy(q_host).to(dev)
    words = (n + 31) // 32
    gathered = torch.empty((world * words,), dtype=torch.int32, device=dev)
    my_slot = gathered[rank * words:(rank + 1) * words]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)   # > 126 MB L2

    fp32_peak = measure_fp32_peak(local) if rank == 0 else None

    def step():
        bits, d, _ = eng.check(q_dev, flags=flags, packed=True)
        if world > 1:
            my_slot.copy_(bits)
            dist.all_gather_into_tensor(gathered, my_slot)
        return bits, d

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    eng.profile(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.zero_()                   # evict the 36 MB input from L2 between timed iterations
        ev[s][0].record()
        step()
        ev[s][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    prof = eng.get_profile()
    eng.profile(False)
    stats = eng.stats()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * n * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host-buffer entry point
    q_pin = torch.from_numpy(q_host).pin_memory()
    out = {'bits': torch.empty((words,), dtype=torch.int32, pin_memory=True),
           'dist': torch.empty((n,), dtype=torch.float32, pin_memory=True), 'reason': None}
    for _ in range(2):
        eng.check_host(q_pin, flags=flags, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.check_host(q_pin, flags=flags, out=out)      # returns when the results are in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / float(te.item())
    # the same steps with two batches in flight (jg_check_host_submit / _wait): what a caller streaming batches gets --
    # every step still uploads its inputs from pinned host memory and downloads its results inside the timed region
    out2 = {'bits': torch.empty((words,), dtype=torch.int32, pin_memory=True),
            'dist': torch.empty((n,), dtype=torch.float32, pin_memory=True), 'reason': None}
    outs = [out, out2]
    eng.check_host_wait(eng.check_host_submit(q_pin, flags=flags, out=outs[0]))
    barrier()
    t0 = time.perf_counter()
    tk = eng.check_host_submit(q_pin, flags=flags, out=outs[0])
    for s_i in range(1, args.steps):
        nxt = eng.check_host_submit(q_pin, flags=flags, out=outs[s_i & 1])
        eng.check_host_wait(tk)
        tk = nxt
    eng.check_host_wait(tk)
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
        line = {
            'metric': 'configuration collision checks/sec', 'value': value, 'unit': 'checks/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_total / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'scenario {args.scenario:03d}: {n} uniformly sampled mobile-Panda configurations per GPU, '
                                   'collision bit + signed min distance (SCENE+PLANE)',
                       'configs_per_gpu': n, 'seed': '11 + 1000*rank', 'flags': int(flags),
                       'penetration_depth': 'clamped at -(marginA+marginB) (JG_NO_EPA)' if args.no_epa else 'EPA',
                       'l2': 'flushed between timed iterations (256 MB write)',
                       'collective': 'ncclAllGather of packed verdict bits' if world > 1 else 'none (1 GPU)',
                       'pass_size': args.chunk or 'auto (<= 2^20 configurations per pass)',
                       'cpu_affinity': f'process bound to the {bound} CPUs local to its GPU (NVML)' if bound else 'not set'},
            'clocks': clocks,
            'e2e': {'value': e2e_pipe_value, 'unit': 'checks/s', 'h2d_bytes_per_step': int(n * 36),
                    'd2h_bytes_per_step': int(n * 4 + words * 4),
                    'api': 'jg_check_host_submit / jg_check_host_wait, two 1 M batches in flight (pinned host buffers)',
                    'synchronous_call_value': e2e_value,
                    'synchronous_api': 'jg_check_host: one blocking call per step (copies overlap only within the batch)'},
            'gpu_launches': int(launches_per_step * args.steps),
            'roofline': {'bound': 'fp32', 'kernel': 'narrow phase (k_narrowphase GJK stages + k_penetration stages)',
                         'achieved': ach_tf, 'peak': fp32_peak,
                         'unit': 'TFLOP/s', 'frac': (ach_tf / fp32_peak) if (ach_tf and fp32_peak) else None,
                         'traffic': (NARROW_TRAFFIC_BYTES_PER_CHECK_NCU * n if (args.scenario == 11 and not args.no_epa) else None),
                         'traffic_source': 'ncu dram__bytes_read.sum + dram__bytes_write.sum per pass, profiles/r01_m_ncu_full_summary.csv',
                         'peak_source': 'jg_measure_fp32_peak FMA micro-benchmark on this GPU',
                         'algorithmic_flops_per_check': FLOPS_PER_CHECK_NARROW,
                         'avg_launch_ms': nar_ms / np_n if np_n else None, 'launches': int(np_n),
                         'note': 'algorithmic = nominal work of an unpruned narrow phase; the champion-first pruning '
                                 'executes far fewer flops, so frac is work-avoided + pipe-utilisation combined'},
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
                                              'restatement (double precision) - pyBullet not installable in this image'}
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
        print(json.dumps(line), flush=True)
    eng.close()
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
    ap.add_argument('--no-epa', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-count', action='store_true', help='skip the counted-work audit (instrumented library)')
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
