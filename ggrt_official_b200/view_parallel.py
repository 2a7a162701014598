"""Multi-GPU scaling of the render path: independent target views are sharded across ranks.

Each target view is an independent rasterization of the same Gaussian set (the per-view
loop at /root/reference/ggrt/model/pixelsplat/decoder/cuda_splatting.py:93-127 has no
cross-iteration dependence), so rank r renders views r, r+world, ... of the replicated
Gaussians with no data-path collective in the forward.  The only exchange step is the sum of
the per-view Gaussian gradients: ONE all-reduce over a single contiguous arena
[P, 3 + 6 + 1 + 3K].  One process per GPU, torch.distributed (NCCL over
NVLink on the B200 box; gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment of target views to ranks."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    return list(range(rank, num_views, world_size))


@dataclass
class GradientArena:
    """One flat float32 buffer holding every Gaussian gradient, plus named views into it."""

    flat: torch.Tensor
    views: Dict[str, torch.Tensor]

    @staticmethod
    def allocate(P: int, K: int, device, use_sh: bool = True) -> "GradientArena":
        names = [("dmeans3D", (P, 3)), ("dcov3D", (P, 6)), ("dopacity", (P, 1)),
                 ("dsh", (P, K, 3)) if use_sh else ("dcolors", (P, 3))]
        total = sum(int(torch.Size(s).numel()) for _, s in names)
        flat = torch.empty(total, dtype=torch.float32, device=device)
        views, off = {}, 0
        for name, shape in names:
            n = int(torch.Size(shape).numel())
            views[name] = flat[off: off + n].view(shape)
            off += n
        return GradientArena(flat, views)

    def all_reduce(self, group=None, async_op: bool = False):
        """Sum over ranks, in place (float32: the 1e-3 relative gradient tolerance leaves no room for compression)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def all_reduce_gradients(tensors: Sequence[Optional[torch.Tensor]], group=None) -> None:
    """Sums `.grad` of the given leaf tensors over ranks with one collective (autograd-level path).
    Ranks that rendered no view contribute zeros."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    leaves = [t for t in tensors if t is not None]
    for t in leaves:
        if t.grad is None:
            t.grad = torch.zeros_like(t)
    flat = torch.cat([t.grad.reshape(-1) for t in leaves])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in leaves:
        n = t.grad.numel()
        t.grad.copy_(flat[off: off + n].view_as(t.grad))
        off += n


def render_views_sharded(render_fn, num_views: int, rank: Optional[int] = None, world_size: Optional[int] = None):
    """Calls `render_fn(view_index)` for the views owned by this rank; returns {view_index: result}."""
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    return {v: render_fn(v) for v in shard_views(num_views, rank, world_size)}
