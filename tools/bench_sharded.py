#!/usr/bin/env python
"""BASELINE.json configurations 3, 4 and 5 at FULL size, sharded over the GPUs of one box.

  python tools/bench_sharded.py                       # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
         tools/bench_sharded.py

  config 3  each of the 20 scenarios: 10 M uniform configurations (generated on the device: torch Philox,
            seed = scenario id * 1000 + rank), contiguous 32-aligned shards, bit + signed distance, packed verdict words
            all-gathered (in place) inside the timed region
  config 4  scenario 034: 1 M edges x 64 substeps, qa uniform, qb = clip(qa + U(-0.25, 0.25)^9), sharded by edge
  config 5  scenario 011: 10 M free-floating gripper poses around the target object, sharded by pose

One JSON line per measurement from rank 0: time = max over ranks of the CUDA-event time of the rank's loop.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jogramop_framework_b200.dist import GatherBuffer, gather_bits, shard_bounds  # noqa: E402
from jogramop_framework_b200.engine import PLANE, SCENE, CollisionEngine  # noqa: E402
from jogramop_framework_b200.ingest import RobotModel, SceneModel  # noqa: E402

DATA = os.path.join(ROOT, 'jogramop_framework_b200', 'data')
IDS = [11, 12, 13, 14, 15, 21, 22, 23, 24, 25, 31, 32, 33, 34, 35, 41, 42, 43, 44, 45]
LANES = int(os.environ.get('JG_LANES', 4))   # scratch arenas a multi-pass call uses side by side (jg_engine_set_lanes)


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n_cfg = int(os.environ.get('JG_N', 10_000_000))
    n_edges = int(os.environ.get('JG_EDGES', 1_000_000))
    robot = RobotModel.load(os.path.join(DATA, 'robot_mobile_panda.npz'))
    lim = torch.tensor(robot.joint_limits[robot.arm_joint_ids], dtype=torch.float32, device=dev)
    lo_q, hi_q = lim[:, 0], lim[:, 1]

    def sync_max(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, warm=1):
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        barrier()
        return sync_max(a.elapsed_time(b)), out

    def emit(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    # ---- config 3
    total_ms = 0.0
    for sid in IDS:
        scene = SceneModel.load(os.path.join(DATA, f'scene_{sid:03d}.npz'))
        eng = CollisionEngine(robot, scene, device=local, lanes=LANES)
        lo, hi = shard_bounds(n_cfg, world, rank)
        g = torch.Generator(device=dev)
        g.manual_seed(sid * 1000 + rank)
        q = lo_q + (hi_q - lo_q) * torch.rand((hi - lo, 9), generator=g, device=dev)
        buf = GatherBuffer(n_cfg, dev)        # the kernels write into this rank's slot: in-place all-gather
        dmin = torch.empty(hi - lo, dtype=torch.float32, device=dev)

        def run():   # one call: the library cuts it into passes and runs them side by side on its lanes
            eng.check(q, flags=SCENE | PLANE, packed=True, bits_out=buf.slot, dist_out=dmin)
            return buf.gather()

        ms, allw = timed(run)
        total_ms += ms
        nb = allw.numel() * 32
        rate = (torch.bitwise_and(allw.view(-1, 1) >> torch.arange(32, device=dev, dtype=torch.int32), 1).sum().item()) / n_cfg
        emit(config=3, scenario=sid, n=n_cfg, n_gpus=world, ms=ms, checks_per_s=n_cfg / ms * 1e3, collision_rate=rate,
             pieces=scene.n_pieces, gathered_words=allw.numel(), note='bit + distance, verdict words all-gathered')
        del eng
    emit(config=3, scenario='all 20', n=n_cfg * len(IDS), n_gpus=world, ms=total_ms,
         checks_per_s=n_cfg * len(IDS) / total_ms * 1e3)

    # ---- config 4
    scene = SceneModel.load(os.path.join(DATA, 'scene_034.npz'))
    eng = CollisionEngine(robot, scene, device=local, lanes=LANES)
    lo, hi = shard_bounds(n_edges, world, rank)
    g = torch.Generator(device=dev)
    g.manual_seed(34_000 + rank)
    qa = lo_q + (hi_q - lo_q) * torch.rand((hi - lo, 9), generator=g, device=dev)
    qb = torch.minimum(torch.maximum(qa + (torch.rand((hi - lo, 9), generator=g, device=dev) - 0.5) * 0.5, lo_q), hi_q)

    def run_edges():
        words, _ = eng.check_edges(qa, qb, substeps=64, packed=True)
        return gather_bits(words, n_edges)

    ms, allw = timed(run_edges)
    hit = torch.bitwise_and(allw.view(-1, 1) >> torch.arange(32, device=dev, dtype=torch.int32), 1).sum().item() / n_edges
    emit(config=4, scenario=34, edges=n_edges, substeps=64, n_gpus=world, ms=ms, checks_per_s=n_edges * 64 / ms * 1e3,
         edges_per_s=n_edges / ms * 1e3, edge_collision_rate=hit, outputs='verdict bit + min distance over the substeps')

    def run_edges_bits():
        words, _ = eng.check_edges(qa, qb, substeps=64, packed=True, want_dist=False)
        return gather_bits(words, n_edges)

    ms, allw2 = timed(run_edges_bits)
    emit(config=4, scenario=34, edges=n_edges, substeps=64, n_gpus=world, ms=ms, checks_per_s=n_edges * 64 / ms * 1e3,
         edges_per_s=n_edges / ms * 1e3, outputs='verdict bit only (what an RRT edge validation needs)',
         same_verdicts_as_distance_mode=bool(torch.equal(allw, allw2)))
    del eng

    # ---- config 5
    scene = SceneModel.load(os.path.join(DATA, 'scene_011.npz'))
    eng = CollisionEngine(robot, scene, device=local, lanes=LANES)
    lo, hi = shard_bounds(n_cfg, world, rank)
    tgt = scene.piece(0)
    blo = torch.tensor(tgt.min(0) - 0.25, dtype=torch.float32, device=dev)
    bhi = torch.tensor(tgt.max(0) + 0.25, dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(5011_000 + rank)
    pos = blo + (bhi - blo) * torch.rand((hi - lo, 3), generator=g, device=dev)
    quat = torch.randn((hi - lo, 4), generator=g, device=dev)
    pq = torch.cat([pos, quat], 1).contiguous()

    def run_poses():
        words, _ = eng.check_poses(pq, links=[11, 12, 13], packed=True)
        return gather_bits(words, n_cfg)

    ms, allw = timed(run_poses)
    hit = torch.bitwise_and(allw.view(-1, 1) >> torch.arange(32, device=dev, dtype=torch.int32), 1).sum().item() / n_cfg
    emit(config=5, scenario=11, n=n_cfg, n_gpus=world, ms=ms, checks_per_s=n_cfg / ms * 1e3, collision_rate=hit,
         model='panda_hand + finger hulls (Bullet margins) vs the convex pieces, bit + signed distance')
    del eng
    # the reference's own filter: burg two-finger gripper vs the visual meshes, boolean (create_graspset.py:53-59)
    import contextlib
    import io
    from jogramop_framework_b200.graspset import burg_gripper_model
    from jogramop_framework_b200.scenario import Scenario
    with contextlib.redirect_stdout(io.StringIO()):
        vis = Scenario(11, device=local).visual_scene_model()
    geng = CollisionEngine(burg_gripper_model(), vis, device=local, threshold=1e-6, finger_open=0.0, mesh_mode=True, mesh_margin=0.0)   # (lanes: no gain on the traversal path, measured)

    def run_burg():
        words, _ = geng.check_poses(pq, links=[0], ee_link=0, want_dist=False, packed=True)
        return gather_bits(words, n_cfg)

    ms, allw = timed(run_burg)
    hit = torch.bitwise_and(allw.view(-1, 1) >> torch.arange(32, device=dev, dtype=torch.int32), 1).sum().item() / n_cfg
    emit(config=5, scenario=11, n=n_cfg, n_gpus=world, ms=ms, checks_per_s=n_cfg / ms * 1e3, collision_rate=hit,
         model=f'burg two-finger gripper vs the visual meshes ({len(vis.tris)} triangles, LBVH) + plane, boolean: the reference filter')
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
