"""Batch sharding of the separation path across ranks: contiguous shards, no
data-path collective.  BEST / FORK attend over the batch axis (TDANet_best.py:247-251), so a shard
reproduces the reference run on that shard - exactly what reference DDP computes per rank."""
from typing import Tuple

import torch


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the contiguous shard of `rank`; the first n_items % world_size ranks get one more."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def separate_sharded(model, mixtures: torch.Tensor, world_size: int, rank: int) -> torch.Tensor:
    """Run `model` on this rank's shard of `mixtures` [B, T] (or [B, 1, T]); returns [b_local, n_src, T]."""
    lo, hi = shard_bounds(mixtures.shape[0], world_size, rank)
    if hi == lo:
        return mixtures.new_zeros((0, model.num_sources, mixtures.shape[-1]))
    return model(mixtures[lo:hi])
