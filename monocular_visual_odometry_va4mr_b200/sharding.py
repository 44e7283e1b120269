"""Sharding of independent sequences over ranks: the per-frame recursion is
sequential inside a sequence, so the only parallel axis is the sequence; ranks never exchange data
on the hot path, and the trajectories (one pose per sequence and frame) are all-gathered once at
the end.  Backend-agnostic: NCCL on the GPU box, gloo in the CPU tests."""
from __future__ import annotations


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `total` sequences owned by `rank` (sizes differ by at most 1)."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sequence_seed(global_index: int, base_seed: int = 0) -> int:
    """Seed of sequence `global_index` -- independent of how the sequences are sharded."""
    return base_seed * 100003 + global_index


def gather_trajectories(local_poses, world: int, counts=None):
    """all_gather of per-rank pose tensors [n_local, n_frames, 6] -> list ordered by rank.
    `counts` (sequences per rank) is needed when the shards are ragged."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [local_poses]
    if counts is None or len(set(counts)) == 1:
        out = [torch.empty_like(local_poses) for _ in range(world)]
        dist.all_gather(out, local_poses)
        return out
    nmax = max(counts)
    pad = torch.zeros((nmax,) + tuple(local_poses.shape[1:]), dtype=local_poses.dtype, device=local_poses.device)
    pad[: local_poses.shape[0]] = local_poses
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return [o[:c] for o, c in zip(out, counts)]
