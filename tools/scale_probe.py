"""Weak-scaling probe of the bench step (4 batches in flight + in-place all-gather of the verdict words) without the rest
of bench.py: what the collective costs at N GPUs.  torchrun ... tools/scale_probe.py [--gather-every G] [--no-gather]
NCCL_MAX_NCHANNELS etc. come from the environment.  Builder tool; numbers of record come from bench.py."""
import argparse, json, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jogramop_framework_b200.engine import EnginePool
from jogramop_framework_b200.ingest import RobotModel, SceneModel
ap = argparse.ArgumentParser()
ap.add_argument('--gather-every', type=int, default=1)
ap.add_argument('--no-gather', action='store_true')
ap.add_argument('--steps', type=int, default=40)
ap.add_argument('--lanes', type=int, default=4)
a = ap.parse_args()
world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    o = dist.ProcessGroupNCCL.Options(); o.is_high_priority_stream = True
    dist.init_process_group('nccl', device_id=dev, pg_options=o)
D = os.path.join(ROOT, 'jogramop_framework_b200', 'data')
robot, scene = RobotModel.load(f'{D}/robot_mobile_panda.npz'), SceneModel.load(f'{D}/scene_011.npz')
lim = robot.joint_limits[robot.arm_joint_ids]
n, words, G, lanes = 1_000_000, 31250, max(1, a.gather_every), a.lanes
qs = [torch.from_numpy((lim[:, 0] + (lim[:, 1] - lim[:, 0]) * np.random.default_rng(11 + 1000 * rank + 7919 * r).random((n, 9))).astype(np.float32)).to(dev) for r in range(4)]
pool = EnginePool(robot, scene, device=local, lanes=lanes)
NB = (lanes + G - 1) // G + 1
gathered = [torch.zeros(world * G * words, dtype=torch.int32, device=dev) for _ in range(NB)]
mine = [g[rank * G * words:(rank + 1) * G * words] for g in gathered]
douts = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(lanes)]
side = torch.cuda.Stream(device=dev, priority=-1)
evg, pend, grp = [torch.cuda.Event() for _ in range(NB)], [False] * NB, {'ev': [], 'b': 0}
def flush():
    b = grp['b']
    if world > 1 and grp['ev'] and not a.no_gather:
        for e in grp['ev']: side.wait_event(e)
        with torch.cuda.stream(side):
            dist.all_gather_into_tensor(gathered[b], mine[b]); evg[b].record(side)
        pend[b] = True
    grp['ev'] = []
def step(i):
    b, j = (i // G) % NB, i % G
    if j == 0:
        if grp['ev']: flush()
        grp['b'] = b
        if pend[b]: torch.cuda.current_stream().wait_event(evg[b])
    t = pool.submit(qs[i % 4], flags=3, packed=True, bits_out=mine[b][j * words:(j + 1) * words], dist_out=douts[pool.next_lane])
    grp['ev'].append(t.event)
    if j == G - 1: flush()
def drain():
    flush(); pool.drain(); torch.cuda.current_stream().wait_stream(side)
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for rep in range(3):
    for i in range(8): step(i)
    drain(); barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps): step(i)
    drain(); e1.record(); barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({'n_gpus': world, 'gather_every': 0 if a.no_gather else G, 'nchannels': os.environ.get('NCCL_MAX_NCHANNELS'), 'rep': rep,
                          'ms_per_step': t.item() / a.steps, 'value': world * n * a.steps / t.item() * 1e3}), flush=True)
pool.close()
if world > 1: dist.destroy_process_group()
