"""Multi-rank path on CPU: world_size-2 (and 3) gloo processes shard a batch, each 'checks' its shard with the CPU
oracle standing in for the GPU kernel (test infrastructure), and the all-gathered bitmask must be bit-identical to
the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jogramop_framework_b200 import dist as jd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_align():
    for n in (0, 1, 31, 32, 33, 1000, 1_000_000, 999_999):
        for world in (1, 2, 3, 4, 8):
            per = jd.shard_size(n, world)
            assert per % 32 == 0
            covered = []
            for r in range(world):
                lo, hi = jd.shard_bounds(n, world, r)
                assert lo % 32 == 0 or lo == n
                covered.extend(range(lo, hi)) if n <= 1000 else None
                if r + 1 < world:
                    assert jd.shard_bounds(n, world, r + 1)[0] == hi or hi == n
            if n <= 1000:
                assert covered == list(range(n))
            assert jd.shard_bounds(n, world, world - 1)[1] == n


def test_pack_bits_layout():
    bits = np.zeros(70, bool)
    bits[[0, 31, 32, 69]] = True
    w = jd.pack_bits_numpy(bits).view(np.uint32)
    assert w.tolist() == [0x80000001, 0x00000001, 1 << 5]


def _worker(rank, world, port, n, q, expected_words, expected_dist):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import oracle
        from jogramop_framework_b200.ingest import RobotModel, SceneModel
        data = os.path.join(ROOT, 'jogramop_framework_b200', 'data')
        robot = RobotModel.load(os.path.join(data, 'robot_mobile_panda.npz'))
        scene = SceneModel.load(os.path.join(data, 'scene_034.npz'))
        o = oracle.Oracle(robot, scene)

        class OracleAsEngine:   # same call shape as CollisionEngine.check(packed=True, bits_out=slot)
            device = torch.device('cpu')

            def check(self, qq, flags, want_dist, packed, bits_out):
                b, d = o.check(np.asarray(qq, dtype=np.float64), flags=flags, n_threads=1)
                w = torch.from_numpy(jd.pack_bits_numpy(b).copy())
                bits_out[:len(w)].copy_(w)      # the kernels write straight into the rank's slot of the gather buffer
                return bits_out[:len(w)], torch.from_numpy(d.astype(np.float32)), None

        words, d = jd.check_sharded(OracleAsEngine(), q, flags=3, want_dist=True)
        assert torch.equal(words, torch.from_numpy(expected_words)), f'rank {rank}: gathered bitmask differs'
        assert torch.equal(d, torch.from_numpy(expected_dist)), f'rank {rank}: gathered distances differ'
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n', [(2, 1000), (3, 777), (2, 31)])
def test_gloo_shard_and_gather_matches_single_process(world, n, robot, scene_loader):
    import oracle
    from ref_numpy import uniform_configs
    q = uniform_configs(robot, n, 77)
    b, d = oracle.Oracle(robot, scene_loader(34)).check(q.astype(np.float64), flags=3, n_threads=1)
    expected = jd.pack_bits_numpy(b).copy()
    port = 29600 + world * 7 + n % 50
    mp.spawn(_worker, args=(world, port, n, q, expected, d.astype(np.float32)), nprocs=world, join=True)
