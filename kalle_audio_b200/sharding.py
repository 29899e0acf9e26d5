"""Batch sharding of independent clips over the GPUs of one node.

Every clip is independent and the convolutions have no cross-batch term, so decode/encode of many clips is
split by batch index over one process per GPU with replicated weights and NO collective on the data path.
The only communication is bookkeeping: the max-over-ranks of the elapsed time, and (optionally) gathering
results onto rank 0 for a caller that wants them in one place.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of rank `rank`: sizes differ by at most one, earlier ranks take the remainder."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world size")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Max of a per-rank scalar (timings are reported as the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_sharded(fn: Callable[[torch.Tensor], torch.Tensor], items: torch.Tensor, gather: bool = False,
                micro_batch: Optional[int] = None) -> Optional[torch.Tensor]:
    """Applies `fn` (e.g. ``ae.decode``) to this rank's slice of `items` [N, ...], optionally in micro-batches.
    Returns the local result, or with ``gather=True`` the concatenated result on rank 0 (None elsewhere)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(items.shape[0], world, rank)
    local = items[lo:hi]
    outs: List[torch.Tensor] = []
    step = micro_batch or max(1, hi - lo)
    for i in range(0, hi - lo, step):
        outs.append(fn(local[i:i + step]))
    out = torch.cat(outs, dim=0) if outs else None
    if not gather or world == 1:
        return out
    sizes = [shard_bounds(items.shape[0], world, r) for r in range(world)]
    if out is None:  # a rank with an empty slice still takes part in the gather
        raise ValueError("gather needs at least one item per rank")
    # gather needs equal sizes on every backend: pad each shard to the largest one, trim on rank 0
    m = max(h - l for l, h in sizes)
    pad = torch.zeros((m,) + tuple(out.shape[1:]), dtype=out.dtype, device=out.device)
    pad[:out.shape[0]] = out
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0)
    if rank != 0:
        return None
    return torch.cat([b[:h - l] for b, (l, h) in zip(bufs, sizes)], dim=0)
