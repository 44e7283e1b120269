"""Multi-rank host logic on CPU: gloo, world_size 2 (the GPU box runs the same code over NCCL)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from monocular_visual_odometry_va4mr_b200 import sharding


def test_shard_range_partitions():
    for total in (0, 1, 7, 64, 65):
        for world in (1, 2, 4, 8):
            blocks = [sharding.shard_range(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(total, world, rank)
    # stand-in for the per-rank tracker: the pose of (sequence, frame) is a pure function of the seed
    poses = torch.stack([torch.stack([torch.full((6,), float(sharding.sequence_seed(s) * 1000 + f), dtype=torch.float64)
                                      for f in range(5)]) for s in range(lo, hi)]) if hi > lo else torch.zeros((0, 5, 6), dtype=torch.float64)
    counts = [sharding.shard_range(total, world, r)[1] - sharding.shard_range(total, world, r)[0] for r in range(world)]
    parts = sharding.gather_trajectories(poses, world, counts)
    full = torch.cat(parts)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_trajectories_gloo_world2(tmp_path, total):
    port = 29600 + total
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert a.shape == (total, 5, 6) and np.array_equal(a, b)          # every rank holds every trajectory
    want = np.array([[sharding.sequence_seed(s) * 1000 + f for f in range(5)] for s in range(total)], np.float64)
    assert np.array_equal(a[:, :, 0], want)                              # ordered by global sequence index
