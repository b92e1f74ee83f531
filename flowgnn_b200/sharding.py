"""Multi-GPU sharding: graphs are independent, so a batch splits into contiguous graph ranges,
one per rank, with no collective on the data path.  The reference has nothing like
this (one kernel instance, one in-order queue: GIN/config_slr.cfg:2, GIN/src/host.cc:207-209).

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is used only for the throughput tally
and, optionally, to gather the per-graph predictions.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .dataset import Batch, shard_ranges


def shard_of(batch: Batch, rank: int, world_size: int) -> Tuple[Batch, int, int]:
    """Rank ``rank``'s contiguous shard, balanced by node/edge cost; returns (shard, g0, g1)."""
    if world_size <= 1:
        return batch, 0, batch.num_graphs
    bounds = shard_ranges(batch, world_size)
    g0, g1 = int(bounds[rank]), int(bounds[rank + 1])
    return batch.slice(g0, g1), g0, g1


def tally(graphs_done: int, elapsed_ms: float, dist=None, device=None) -> Tuple[int, float]:
    """(sum of graphs over ranks, max elapsed over ranks).  ``dist`` is ``torch.distributed`` or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(graphs_done), float(elapsed_ms)
    import torch
    t_sum = torch.tensor([float(graphs_done)], dtype=torch.float64, device=device)
    t_max = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t_sum, op=dist.ReduceOp.SUM)
    dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    return int(round(t_sum.item())), float(t_max.item())


def gather_predictions(local: np.ndarray, g0: int, total_graphs: int, dist=None, device=None) -> Optional[np.ndarray]:
    """Assemble the per-graph outputs of all ranks (each rank owns out[g0:g0+len(local)]) on every rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    full = torch.zeros(total_graphs, dtype=torch.float32, device=device)
    full[g0:g0 + len(local)] = torch.from_numpy(np.ascontiguousarray(local)).to(full.device)
    # disjoint slices, everything else is zero -> SUM assembles them (NaN stays NaN in its own slot)
    dist.all_reduce(full, op=dist.ReduceOp.SUM)
    return full.cpu().numpy()
