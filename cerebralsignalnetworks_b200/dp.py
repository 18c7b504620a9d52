"""Data-parallel plumbing of the train step: batch sharding and the single flat all-reduce.

Reference semantics (LstmDistillation.py:406-414,445): DistributedSampler(drop_last=True) shards the batch,
DDP averages gradients, DINOLoss.update_center all-reduces the teacher column sums and divides by
len(batch) * world (LstmDistillation.py:154-156).  Here both reductions ride in ONE buffer
[gradients | centre sums]; the division by `world` happens inside the fused Adam / centre-EMA kernels."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_range(n_global: int, rank: int, world: int):
    """Even contiguous shard of a global batch; the remainder is dropped (drop_last=True)."""
    per = n_global // world
    return rank * per, (rank + 1) * per


def allreduce_flat_(flat: torch.Tensor) -> torch.Tensor:
    """SUM over ranks, in place (NCCL on GPUs, gloo in the CPU tests)."""
    if world_size() > 1:
        dist.all_reduce(flat)
    return flat


def split_flat(flat: torch.Tensor, n_param: int, world: int, local_batch: int):
    """-> (mean gradients, batch centre) with the reference's normalisation."""
    return flat[:n_param] / world, flat[n_param:] / (local_batch * world)
