"""Sharding a batch of independent views over the GPUs of one box.

Every view is independent, so the batch dimension is split across ranks with NO data-path collective; the
only exchanges are (a) an optional 2-float MIN/MAX all-reduce that makes the composite-depth clamp see the
bounds of the whole batch, exactly like the reference's single-GPU `torch.min/max(depths)`
(ray_marcher.py:50), and (b) one gather of the rendered images to where the eval loop consumes them.
Works with any torch.distributed backend (nccl on the box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_views: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first n % world ranks get one extra view."""
    base, extra = divmod(n_views, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_reduce_depth_bounds(bounds2: torch.Tensor, group=None) -> torch.Tensor:
    """bounds2 = [min_depth, max_depth] of this rank's shard -> of the whole batch (in place)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(bounds2[0:1], op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(bounds2[1:2], op=dist.ReduceOp.MAX, group=group)
    return bounds2


def gather_views(local: torch.Tensor, n_views: int, dst: Optional[int] = None, group=None) -> Optional[torch.Tensor]:
    """Concatenate per-rank shards (dim 0, sizes per `shard_range`) into the full (n_views, ...) tensor.
    dst=None: every rank gets it (all_gather); dst=r: only rank r does (others return None)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_views, world, r) for r in range(world)]
    counts = [b - a for a, b in sizes]
    pad = max(counts)
    buf = local
    if local.shape[0] < pad:                                   # equal-size collective: pad the short shards
        buf = torch.cat([local, local.new_zeros((pad - local.shape[0],) + tuple(local.shape[1:]))])
    buf = buf.contiguous()
    if dst is None or dist.get_backend(group) != 'nccl':       # gloo has no gather into one tensor: all_gather + pick
        out = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(out, buf, group=group)
        if dst is not None and rank != dst:
            return None
        return torch.cat([o[:c] for o, c in zip(out, counts)])
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)]) if rank == dst else None


def render_sharded(renderer, planes, decoder, ray_origins, ray_directions, options, dst: Optional[int] = 0, exact_depth: bool = True,
                   group=None, **flags):
    """Render this rank's slice of a batch that every rank holds the inputs of (e.g. 16 eval views of one subject)
    and gather (rgb, depth, wsum, xyz) to `dst`.  planes with batch size 1 are shared by all views."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = ray_origins.shape[0]
    a, b = shard_range(n, world, rank)
    pl = planes if planes.shape[0] == 1 and n != 1 else planes[a:b]
    if pl.shape[0] == 1 and b - a != 1:
        pl = pl.expand(b - a, -1, -1, -1, -1)
    prev = getattr(renderer, 'depth_bounds_reduce', None)
    if exact_depth and world > 1:
        renderer.depth_bounds_reduce = lambda b2: all_reduce_depth_bounds(b2, group)
    try:
        outs = renderer(pl, decoder, ray_origins[a:b], ray_directions[a:b], options, **flags) if b > a else None
    finally:
        renderer.depth_bounds_reduce = prev
    if outs is None:                                            # more ranks than views: contribute empty shards
        outs = tuple(ray_origins.new_zeros((0, ray_origins.shape[1], c)) for c in (32, 1, 1, 3))
        if exact_depth and world > 1:
            all_reduce_depth_bounds(ray_origins.new_tensor([float('inf'), float('-inf')]), group)
    return tuple(gather_views(o, n, dst, group) for o in outs)
