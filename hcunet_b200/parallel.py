"""Data-parallel plumbing for the training path: one process per GPU, patches sharded
across ranks, ONE exchange step per iteration -- the fp32 mean all-reduce of the parameter gradients
(727 009 floats = 2.9 MB for the README model).  BatchNorm uses per-rank batch statistics (plain DDP
semantics); buffers and parameters are broadcast from rank 0 at construction.

The reference has no distributed code at all; this is the build's only collective.
Works on NCCL (GPU) and gloo (CPU tests)."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int):
    """Contiguous balanced partition of ``n_items`` independent units (patches / overlap tiles)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradSync:
    """Flat-bucket gradient all-reduce (mean).  ``allreduce()`` is called after ``loss.backward()``."""

    def __init__(self, module: torch.nn.Module, world: int | None = None, broadcast: bool = True):
        self.module = module
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        self._flat = None
        if self.world > 1 and broadcast:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t, src=0)

    def allreduce(self):
        if self.world <= 1:
            return
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return
        n = sum(g.numel() for g in grads)
        if self._flat is None or self._flat.numel() != n or self._flat.device != grads[0].device:
            self._flat = torch.empty(n, dtype=torch.float32, device=grads[0].device)
        views = []
        o = 0
        for g in grads:
            v = self._flat[o:o + g.numel()].view_as(g)
            views.append(v)
            o += g.numel()
        torch._foreach_copy_(views, grads)
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
        self._flat.mul_(1.0 / self.world)
        torch._foreach_copy_(grads, views)
