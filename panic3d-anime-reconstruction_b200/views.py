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


class _RawCuda:
    """__cuda_array_interface__ view of raw device memory (zero-copy torch.as_tensor)."""

    def __init__(self, ptr, shape, typestr='<f4'):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False), 'version': 2}


class PeerGather:
    """Rendered images -> the consumer rank, overlapped with the next render.

    The NCCL gather is an SM kernel; the fused renderer is a persistent kernel that owns every SM (1 CTA/SM, 224 KB of
    shared memory), so a collective launched beside it cannot start before the render ends - the exchange sits on the
    critical path of every step.  Here the consumer rank allocates ONE device buffer (`slots` x world x shard) in the
    library and exports it (`p3d_ipc_alloc`: cudaIpcGetMemHandle); every other rank opens the handle on ITS OWN device
    (`p3d_ipc_open`: the driver enables peer access over NVLink lazily) and from then on delivers its shard with
    `p3d_copy_async` on a side stream: an NVLink DMA by the copy engines, no SM, no NCCL kernel, so it runs under the next
    step's render.  `fence()` makes the consumer's view complete (side streams drained + one barrier); NCCL stays the
    plumbing (handle broadcast, barrier).  Falls back to `gather_views` when the backend is not NCCL / the tensors are not
    CUDA (the CPU tests).  fp32 shards only."""

    def __init__(self, shard_shape, dtype, device, dst: int = 0, group=None, slots: int = 2):
        self.group, self.dst, self.slots = group, dst, slots
        self.shard_shape = tuple(shard_shape)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.p2p = (self.world > 1 and torch.device(device).type == 'cuda' and dist.get_backend(group) == 'nccl'
                    and dtype == torch.float32)
        self.buf = None
        self.last = None
        self._base = None
        if not self.p2p:
            return
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib, C
        L = _lib.lib()
        n = 1
        for d in self.shard_shape:
            n *= int(d)
        self.shard_bytes = n * 4
        total = slots * self.world * self.shard_bytes
        box = [None]
        ptr = C.c_void_p()
        ok, why = 1, ''
        with torch.cuda.device(device):
            # every rank goes through the same collectives whatever happens locally: a rank whose export / import fails (no peer
            # access between two devices, IPC disabled in a container) reports it, and ALL ranks then fall back to the NCCL gather
            try:
                if self.rank == dst:
                    handle = C.create_string_buffer(64)
                    _lib.check(L.p3d_ipc_alloc(total, C.byref(ptr), handle))
                    box[0] = handle.raw
            except Exception as e:                                             # noqa: BLE001 - reported below, collectively
                ok, why = 0, str(e)
            dist.broadcast_object_list(box, src=dst, group=group)
            try:
                if self.rank != dst:
                    if box[0] is None:
                        raise RuntimeError('the consumer rank could not export its buffer')
                    _lib.check(L.p3d_ipc_open(box[0], C.byref(ptr)))
            except Exception as e:                                             # noqa: BLE001
                ok, why = 0, str(e)
            flag = torch.tensor([ok], device=device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                if ptr.value:
                    (L.p3d_ipc_free if self.rank == dst else L.p3d_ipc_close)(ptr.value)
                self.p2p = False
                import warnings
                warnings.warn('PeerGather: peer-to-peer delivery unavailable (%s); using the NCCL gather' % (why or 'another rank failed'))
                return
            self._base = int(ptr.value)
            if self.rank == dst:
                self.buf = torch.as_tensor(_RawCuda(self._base, (slots, self.world) + self.shard_shape), device=device)
            self.stream = torch.cuda.Stream(device=device)
        dist.barrier(group=group)

    def push(self, shard: torch.Tensor, step: int = 0):
        """Deliver this rank's shard of step `step` (asynchronous; ordered after the work queued on the current stream)."""
        if not self.p2p:
            self.last = gather_views(shard, shard.shape[0] * self.world, self.dst, self.group)
            return
        assert shard.is_contiguous() and shard.dtype == torch.float32 and shard.numel() * 4 == self.shard_bytes
        cur = torch.cuda.current_stream(shard.device)
        self.stream.wait_stream(cur)
        dst = self._base + ((step % self.slots) * self.world + self.rank) * self.shard_bytes
        self._lib.check(self._lib.lib().p3d_copy_async(dst, shard.data_ptr(), self.shard_bytes, self.stream.cuda_stream))
        shard.record_stream(self.stream)

    def join_current_stream(self):
        """Make the current stream wait for the copies issued so far (so a CUDA-event bracket includes them)."""
        if self.p2p:
            torch.cuda.current_stream().wait_stream(self.stream)

    def fence(self):
        if self.p2p:
            self.stream.synchronize()
            dist.barrier(group=self.group)

    def result(self, step: int = 0) -> Optional[torch.Tensor]:
        """After fence(): the (world * shard_views, ...) tensor of step `step` on the consumer rank, None elsewhere."""
        if not self.p2p:
            return self.last
        if self.rank != self.dst:
            return None
        b = self.buf[step % self.slots]
        return b.reshape((b.shape[0] * b.shape[1],) + tuple(b.shape[2:]))

    def close(self):
        if self.p2p and self._base:
            self.fence()
            L = self._lib.lib()
            (L.p3d_ipc_free if self.rank == self.dst else L.p3d_ipc_close)(self._base)
            self._base, self.buf = None, None


def render_sharded(renderer, planes, decoder, ray_origins, ray_directions, options, dst: Optional[int] = 0, exact_depth: bool = True,
                   group=None, **flags):
    """Render this rank's slice of a batch that every rank holds the inputs of (e.g. 16 eval views of one subject)
    and gather (rgb, depth, wsum, xyz) to `dst`.  planes with batch size 1 are shared by all views.
    `exact_depth` makes the composite-depth clamp batch-wide.  'auto' ray limits (ray_start = ray_end = 'auto') carry a
    second batch-wide reduction in the reference - rays that miss the box are given the min / max start of the whole
    batch (renderer.py:167-170) - which is NOT exchanged here: it is taken over the shard, so such rays can differ from
    the one-process render; the panic3d configurations use numeric limits."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world > 1 and options.get('ray_start') == 'auto' and options.get('ray_end') == 'auto':
        import warnings
        warnings.warn("render_sharded with 'auto' ray limits: the fill value of rays that miss the box is reduced per shard, not per "
                      "batch (see the docstring)", RuntimeWarning, stacklevel=2)
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
