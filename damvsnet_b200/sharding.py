"""Multi-GPU partitioning of the hot path.

Reference views are independent units of work (one meta per reference view,
reference datasets/general_eval.py:24-54), so inference shards with NO data-path
collective: rank r processes views r, r + world, r + 2*world, ...  The only
communication is a max-reduction of the per-rank device time and a sum of the
processed-view counts when throughput is reported.  This replaces the
reference's nn.DataParallel wrapper (test_uni.py:225).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    """Indices of the reference views rank `rank` owns (round robin)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, num_views, world))


def shard_sizes(num_views: int, world: int) -> List[int]:
    return [len(range(r, num_views, world)) for r in range(world)]


def reduce_throughput(views_done: int, elapsed_ms: float, device: torch.device | None = None) -> Tuple[int, float]:
    """(total views over all ranks, max elapsed ms over ranks).  Works on any backend (nccl on GPUs, gloo on CPU)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return views_done, elapsed_ms
    dev = device if device is not None else torch.device("cpu")
    cnt = torch.tensor([float(views_done)], dtype=torch.float64, device=dev)
    ms = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=dev)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return int(round(cnt.item())), ms.item()


def gather_results(local: Sequence[Tuple[int, torch.Tensor]], num_views: int) -> List[torch.Tensor | None]:
    """All-gather small per-view results (e.g. depth-map statistics) into view order on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out: List[torch.Tensor | None] = [None] * num_views
        for i, t in local:
            out[i] = t
        return out
    bucket: List[object] = [None] * dist.get_world_size()
    dist.all_gather_object(bucket, [(i, t.cpu()) for i, t in local])
    out = [None] * num_views
    for part in bucket:
        for i, t in part:
            out[i] = t
    return out
