"""Adam for `FlatParameters`: the whole optimiser step is ONE kernel launch (`hcu_adam_flat`).

`torch.optim.Adam([fp.flat], fused=True, capturable=True)` + `FlatParameters.guard()` is a fill, a multi-tensor non-finite
check, the fused Adam kernel and its step-counter kernels: ~0.1 ms of launch latency per step for the README model's 727 009
parameters.  `FlatAdam` does the same arithmetic (torch.optim.Adam, `amsgrad=False`; `weight_decay` as the L2 term) and the
skip-step check of fp16-storage training (nothing is touched and the step counter is not advanced when any gradient is
inf / NaN) in one capturable launch.  With data parallelism call it after the gradient all-reduce, like any optimiser.

    fp = FlatParameters(model)
    opt = FlatAdam(fp, lr=1e-3)
    loss.backward(); fp.sync_grad(); opt.step(); opt.zero_grad()
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .flat import FlatParameters


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class FlatAdam:
    def __init__(self, flat: FlatParameters, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if not isinstance(flat, FlatParameters):
            raise TypeError("FlatAdam works on a hcunet_b200.FlatParameters")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("Invalid Adam hyper-parameters")     # torch.optim.Adam's checks
        p = flat.flat
        if not p.is_cuda or p.dtype != torch.float32:
            raise RuntimeError("FlatAdam: CUDA fp32 parameters only (there is no CPU fallback)")
        self.flat = flat
        self.param_groups = [dict(params=[p], lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)]
        self.exp_avg = torch.zeros_like(p)
        self.exp_avg_sq = torch.zeros_like(p)
        self.step_count = torch.zeros((), dtype=torch.int32, device=p.device)
        self._scratch = torch.zeros(2, dtype=torch.float32, device=p.device)   # [skipped flag, grid barrier]

    def step(self):
        p = self.flat.flat
        g = p.grad
        if g is None:
            raise RuntimeError("FlatAdam.step: no gradient (call FlatParameters.sync_grad() after backward)")
        if g.dtype != torch.float32 or g.numel() != p.numel() or not g.is_contiguous():
            raise RuntimeError("FlatAdam.step: the gradient must be the engine's contiguous fp32 flat buffer")
        grp = self.param_groups[0]
        lib = _lib.load()
        st = C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        _lib.check(lib.hcu_adam_flat(_ptr(p.data), _ptr(g), _ptr(self.exp_avg), _ptr(self.exp_avg_sq), p.numel(), grp["lr"],
                                     grp["betas"][0], grp["betas"][1], grp["eps"], grp["weight_decay"], _ptr(self.step_count),
                                     _ptr(self._scratch), st), "adam_flat")

    def zero_grad(self, set_to_none: bool = True):
        self.flat.zero_grad()

    def skipped(self) -> bool:
        """True when the latest step found a non-finite gradient and was skipped (host synchronisation)."""
        return bool(self._scratch[0].item() != 0)

    def state_dict(self):
        return dict(exp_avg=self.exp_avg.clone(), exp_avg_sq=self.exp_avg_sq.clone(), step=int(self.step_count.item()),
                    param_groups=[{k: v for k, v in g.items() if k != "params"} for g in self.param_groups])

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count.fill_(int(sd["step"]))
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
