"""Multi-GPU inference: the path shards over clips with no data-path collective.  One process per
GPU (torchrun); rank r enhances the contiguous slice of clips given by ``shard_bounds``; ``gather_waves`` is the
optional host-side collection of the results on rank 0 (what `inference.py`'s single caller would want)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_clips: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first ``num_clips % world_size`` ranks get one extra clip."""
    assert 0 <= rank < world_size
    base, extra = divmod(num_clips, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def enhance_sharded(enhance_fn, noisy: torch.Tensor, world_size: int, rank: int) -> torch.Tensor:
    """Runs ``enhance_fn`` (e.g. ``Inferencer.enhance_batch``) on this rank's clips only."""
    lo, hi = shard_bounds(noisy.shape[0], world_size, rank)
    if hi == lo:
        return noisy.new_empty((0, noisy.shape[1]))
    return enhance_fn(noisy[lo:hi])


def gather_waves(local: torch.Tensor, num_clips: int, group: Optional[dist.ProcessGroup] = None
                 ) -> Optional[torch.Tensor]:
    """Collects the per-rank results on rank 0 in clip order (host-side convenience; not on the timed path)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts: List[Optional[torch.Tensor]] = [None] * world
    dist.all_gather_object(parts, local.cpu(), group=group)
    if rank != 0:
        return None
    out = torch.cat([p for p in parts if p is not None and p.numel()], 0)
    assert out.shape[0] == num_clips
    return out
