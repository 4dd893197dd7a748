"""Egonet-sharded data parallelism: the path shards by QUERY GROUP, replicas + one gradient all-reduce.

The reference has no working multi-device path (torch.nn.DataParallel cannot split a DGL batched graph,
base/base_trainer.py:18-19; every config sets n_gpu = 1).  Here: one process per GPU, each rank holds a full replica and a
contiguous range of queries (a query's 1 positive + `negative_size` negatives stay together so the per-query InfoNCE
soft-max of trainer/trainer.py:52-56 is rank-local), the loss reduction is a sum (model/loss.py:57), so the exact
single-GPU gradient is the SUM over ranks: one all-reduce of the flat gradient buffer per step, no other exchange.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import numpy as np
import torch


def shard_queries(nodes_per_query: np.ndarray, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous query ranges [(begin, end)) per rank, balanced by NODE count (egonet sizes vary 1..57), never
    splitting a query group. Every rank gets at least one query when there are enough queries."""
    nodes_per_query = np.asarray(nodes_per_query, dtype=np.int64)
    q = int(nodes_per_query.shape[0])
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    if q < world_size:
        raise ValueError(f"cannot shard {q} queries over {world_size} ranks")
    csum = np.concatenate([[0], np.cumsum(nodes_per_query)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        cut = int(np.searchsorted(csum, target, side="left"))
        cut = max(cut, bounds[-1] + 1)                 # at least one query per rank
        cut = min(cut, q - (world_size - r))           # leave one query for every later rank
        bounds.append(cut)
    bounds.append(q)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


class FlatGradBucket:
    """All gradients of a module live in ONE flat fp32 buffer (each parameter's .grad is a view into it), so a step needs
    exactly one all-reduce (1 755 303 floats = 7.0 MB for the MAG-CS PGAT+WMR+LBM model)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=dt)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.flat
