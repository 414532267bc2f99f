"""CUDA-graph capture of a whole training step (launch-bound inner loop -> one graph launch).

A README-model step is ~300 kernel launches of 5 us .. 200 us each; enqueued one by one from Python the step is
bound by the host.  Every kernel of the library is enqueued on the caller's stream without host synchronisation
(include/hcunet_b200.h "Conventions"), so the step -- forward, pixel-weighted loss, backward, optimiser -- can be
captured once and replayed:

    step = GraphedTrainStep(model, optimizer, loss_fn, example=(image, mask, pwl))
    loss = step(image, mask, pwl)          # copies into the static inputs, replays, returns the (static) loss

The reference has nothing comparable (its loop is `tests/r_unet_test.py:24-56`: eager PyTorch, one op at a time).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


class GraphedTrainStep:
    """zero_grad -> model(image) -> loss_fn(logits, mask, pwl) -> backward [-> grad_sync] -> optimizer.step.

    ``grad_sync`` (e.g. ``GradSync.allreduce``) runs eagerly between two graphs (forward/backward and optimiser)
    so that the collective is not captured; without it everything is ONE graph.  The optimizer must be
    capture-safe (``torch.optim.Adam(..., fused=True, capturable=True)``).
    """

    def __init__(self, model, optimizer, loss_fn: Callable, example: Sequence[torch.Tensor],
                 grad_sync: Optional[Callable] = None, warmup: int = 3, input_fn: Optional[Callable] = None, flat=None):
        if warmup < 3:
            # steps 1-2 record the engine's step cache (per-layer packs, a synchronous table upload): not capturable
            raise ValueError("GraphedTrainStep needs warmup >= 3 (the first two steps record the engine's step cache)")
        self.model, self.optimizer, self.loss_fn, self.grad_sync = model, optimizer, loss_fn, grad_sync
        # input_fn(static_in[0]) -> model input, captured with the step (e.g. StackLoader.image on a raw uint8 stack)
        self.input_fn = input_fn
        # flat: hcunet_b200.flat.FlatParameters -- the optimiser holds ONE flat parameter whose gradient is the engine's flat buffer
        self.flat = flat
        self.static_in = [torch.empty_like(t, device=t.device) for t in example]
        for s, t in zip(self.static_in, example):
            s.copy_(t)
        self.loss = None
        # warm-up on a side stream (allocator, lazy attribute setting, optimizer state) as CUDA graphs require
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                if grad_sync is not None:
                    grad_sync()
                optimizer.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.g_main = torch.cuda.CUDAGraph()
        self.g_opt = None
        owner = getattr(grad_sync, "__self__", None)
        if owner is not None and getattr(owner, "attached", False):
            # GradSync.attach(): the collectives are launched inside backward on a communication stream and are captured as a
            # branch of the ONE graph (forward, backward with the all-reduces overlapped, optimiser)
            grad_sync = None
            self.grad_sync = None
        if grad_sync is None:
            with torch.cuda.graph(self.g_main):
                self._fwd_bwd()
                optimizer.step()
        else:
            with torch.cuda.graph(self.g_main):
                self._fwd_bwd()
            self.g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_opt, pool=self.g_main.pool()):
                optimizer.step()

    def _fwd_bwd(self):
        if self.flat is not None:
            self.flat.zero_grad()
        else:
            self.optimizer.zero_grad(set_to_none=True)
        x = self.static_in[0] if self.input_fn is None else self.input_fn(self.static_in[0])
        logits = self.model(x)
        self.loss = self.loss_fn(logits, *self.static_in[1:])
        self.loss.backward()
        if self.flat is not None:
            self.flat.sync_grad()

    def load(self, *tensors, non_blocking=True):
        """Copy one step's inputs (device or pinned-host tensors) into the static input buffers."""
        for s, t in zip(self.static_in, tensors):
            s.copy_(t, non_blocking=non_blocking)

    def run(self):
        self.g_main.replay()
        if self.g_opt is not None:
            self.grad_sync()
            self.g_opt.replay()
        return self.loss

    def __call__(self, *tensors):
        self.load(*tensors)
        return self.run()


class GraphedForward:
    """Inference forward (`model.eval()`, no grad) of a fixed input shape as ONE CUDA-graph launch.

    An eager eval forward of a 1 x 4 x 256 x 256 x 32 stack is ~60 launches of 5 .. 50 us each and is bound by the host
    (1.5 ms); replayed from a graph it is bound by the kernels.  ``out = fwd(x)`` copies ``x`` (device or pinned host) into
    the static input and returns the static output tensor (overwritten by the next call)."""

    def __init__(self, model, example: torch.Tensor, warmup: int = 2):
        if model.training:
            raise RuntimeError("GraphedForward captures an eval-mode forward: call model.eval() first")
        self.model = model
        self.static_in = torch.empty_like(example, device=next(model.parameters()).device)
        self.static_in.copy_(example)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                model(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = model(self.static_in)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out
