"""Data-parallel plumbing for the training path: one process per GPU, patches sharded
across ranks, ONE exchange step per iteration -- the fp32 mean all-reduce of the parameter gradients
(727 009 floats = 2.9 MB for the README model), in place on the engine's flat gradient buffer.  BatchNorm uses per-rank
batch statistics (plain DDP semantics); buffers and parameters are broadcast from rank 0 at construction.

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
    """Gradient all-reduce (mean) of one data-parallel step.  ``allreduce()`` is called after ``loss.backward()``.

    The engine writes every parameter gradient of a step into ONE flat fp32 buffer (parameter order) and hands autograd views
    of it, so the exchange is a single in-place ``all_reduce(AVG)`` on that buffer: no gather / scatter copies, no scaling
    pass, one collective per step (2.9 MB for the README model).  When the gradients are NOT views of one such buffer (a
    foreign module, gradient accumulation into existing ``.grad`` tensors) the generic path packs them into a flat bucket
    first.  gloo (CPU tests) has no AVG: SUM + scale."""

    def __init__(self, module: torch.nn.Module, world: int | None = None, broadcast: bool = True):
        self.module = module
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        self._flat = None
        self.attached = False
        self.in_place_steps = 0     # how many exchanges ran on the engine's own buffer (tests / bench report it)
        if self.world > 1 and broadcast:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t, src=0)

    # ---- all-reduce overlapped with backward ------------------------------------------------------------------------
    def attach(self):
        """Launch the exchange from INSIDE the backward pass: the engine reports when a range of its flat gradient buffer is
        final (`UnetEngine.grad_ready_hook`) -- first the up path + the two deepest levels (~97 % of the README model's
        parameters, ready after ~40 % of the backward's time), then the first levels at the end -- and each range is
        all-reduced (mean) in place on a communication stream while the backward keeps computing.  The backward's stream
        waits for both collectives before it returns, so `optimizer.step()` may follow directly and `allreduce()` becomes a
        no-op.  Capturable: inside `torch.cuda.graph` the collectives become nodes of the step's graph on their own branch.
        Attach BEFORE the first training step (the engine records the bucket split with its step cache)."""
        eng = getattr(self.module, "_engine", None)
        if eng is None:
            raise RuntimeError("GradSync.attach needs a hcunet_b200.Unet_Constructor")
        self._comm = None
        self.overlapped_steps = 0
        self._bucket_calls = 0
        eng.grad_ready_hook = self._on_grads_ready
        self.attached = True
        return self

    def _on_grads_ready(self, flat, lo, hi, events):
        if self.world <= 1 or hi <= lo:
            return None
        if self._comm is None or self._comm.device != flat.device:
            self._comm = torch.cuda.Stream(device=flat.device) if flat.is_cuda else None
        part = flat[lo:hi]
        if self._comm is None:          # CPU tensors (gloo tests): no streams
            self._reduce_mean(part)
            self._bucket_calls += 1
            return None
        for ev in events:
            self._comm.wait_event(ev)
        with torch.cuda.stream(self._comm):
            self._reduce_mean(part)
            done = torch.cuda.Event()
            done.record()
        self._bucket_calls += 1
        if lo == 0:
            self.overlapped_steps += 1
        return done

    def _engine_flat(self):
        """The engine's flat gradient buffer if every ``.grad`` is the view of it the engine returned, else None."""
        eng = getattr(self.module, "_engine", None)
        flat = getattr(eng, "last_grad_flat", None) if eng is not None else None
        if flat is None:
            return None
        off = 0
        base, esz = flat.data_ptr(), flat.element_size()
        for p in self.module.parameters():
            g = p.grad
            if g is None or g.dtype != flat.dtype or not g.is_contiguous() or g.data_ptr() != base + off * esz:
                return None
            off += p.numel()
        return flat if off == flat.numel() else None

    def _reduce_mean(self, flat):
        if dist.get_backend() == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.mul_(1.0 / self.world)

    def allreduce(self):
        if self.world <= 1 or getattr(self, "attached", False):
            return      # attached: the exchange already ran inside backward (see attach())
        flat = self._engine_flat()
        if flat is not None:
            self._reduce_mean(flat)
            self.in_place_steps += 1
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
        self._reduce_mean(self._flat)
        torch._foreach_copy_(grads, views)
