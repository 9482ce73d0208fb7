"""Batch-sharded sampling across the GPUs of one node: one process per GPU, weights replicated, each rank samples an
independent contiguous slice of the batch with its own seed, and the only collective is one all-gather of the final
samples.  The reference never shards sampling (it samples on the main process only,
denoising_diffusion.py:1188-1219); this is new, and deliberately has no data-path collective inside the step loop.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) row ranges; the first `batch % world` ranks get one extra row (ragged batches allowed)."""
    base, extra = divmod(batch, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def gather_samples(local: torch.Tensor, bounds: List[Tuple[int, int]], group=None) -> torch.Tensor:
    """All-gather ragged row slices into the full batch on every rank (one collective)."""
    world = len(bounds)
    if world == 1:
        return local
    rows = max(hi - lo for lo, hi in bounds)
    pad = local
    if local.shape[0] < rows:
        pad = torch.cat([local, local.new_zeros((rows - local.shape[0],) + tuple(local.shape[1:]))], dim=0)
    out = local.new_empty((world * rows,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    parts = [out[r * rows: r * rows + (hi - lo)] for r, (lo, hi) in enumerate(bounds)]
    return torch.cat(parts, dim=0)


def sample_sharded(sample_fn: Callable[[int, int], torch.Tensor], batch_size: int, group=None) -> torch.Tensor:
    """`sample_fn(local_batch, rank)` -> local samples.  Returns the gathered [batch_size, ...] tensor on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return sample_fn(batch_size, 0)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if batch_size < world:
        # some ranks would get an empty slice: they cannot run the sampler (B = 0) and have no shape to contribute, so
        # rank 0's block defines the shape: every rank samples ONE row and the first `batch_size` rows are kept
        bounds = [(r, r + 1) for r in range(world)]
        local = sample_fn(1, rank)
        return gather_samples(local, bounds, group)[:batch_size]
    bounds = shard_bounds(batch_size, world)
    lo, hi = bounds[rank]
    local = sample_fn(hi - lo, rank)
    return gather_samples(local, bounds, group)
