"""The N > 1 path on the CPU: two processes over gloo shard a batch, each checks its block (the oracle stands in for
the CUDA kernel here -- this test is about the sharding and the bitmap all-gather) and every rank must end up with
the full, correctly ordered validity bitmap."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from moveit2_b200 import workloads
    from moveit2_b200.sharding import ShardedStateValidity, shard_bounds
    from util import RES, read
    from oracle.collision_model import OracleEnv
    from oracle.robot_model import RobotModel
    env = OracleEnv(RobotModel(read(RES, "panda_spheres.urdf"), read(RES, "panda.srdf")), "panda_arm", (1, 1, 1),
                    (0, 0, 0.5), 0.02, 0.25)
    env.world.add_points(workloads.make_cloud("C1"))
    lo_b, hi_b = env.joint_bounds()
    q = workloads.random_states(0, n_total, lo_b, hi_b)       # the same global batch on every rank

    def check_local(q_local, bitmap):
        col, _ = env.check(q_local.numpy(), 7, threads=1)
        packed = np.packbits((col == 0).astype(np.uint8), bitorder="little")
        bitmap.zero_()
        bitmap[: len(packed)] = torch.from_numpy(packed)

    sv = ShardedStateValidity(check_local, n_total, torch.device("cpu"))
    lo, hi, per = shard_bounds(n_total, world, rank)
    assert (sv.lo, sv.hi) == (lo, hi) and per % 32 == 0
    sv.check(torch.from_numpy(q[lo:hi]))
    valid = sv.valid_mask()
    ref, _ = env.check(q, 7, threads=1)
    assert np.array_equal(valid, ref == 0), "rank %d: gathered bitmap differs from the unsharded check" % rank
    np.save(os.path.join(out_dir, "valid_%d.npy" % rank), valid)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [1000, 4096 + 17])
def test_two_rank_shard_and_gather(tmp_path, n_total):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "valid_0.npy")
    b = np.load(tmp_path / "valid_1.npy")
    assert np.array_equal(a, b) and 0 < a.sum() < n_total


def test_shard_bounds_cover_everything():
    from moveit2_b200.sharding import shard_bounds
    for n in (0, 1, 31, 32, 33, 1000, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (l0, h0, p0), (l1, h1, p1) in zip(spans, spans[1:]):
                assert h0 == l1 and p0 == p1 and p0 % 32 == 0
