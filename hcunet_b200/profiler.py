"""CUDA-event profile of the C-ABI calls of a step (bench.py's `roofline` object).

While a ``KernelProfile`` is active every ``hcu_*`` call is bracketed by two CUDA events recorded on the
stream the kernel is launched on (torch's current stream); the engine annotates each call with the layer
it belongs to and its ALGORITHMIC bytes / flops (DESIGN.md "Kernels and rooflines").  Nothing here runs
unless a profile is active."""
from __future__ import annotations

from collections import defaultdict

import torch

from . import _lib


# entry points that launch the same kernel (the *_bn / *_fin forms only add a fused tail)
_ALIAS = {"hcu_conv_tc_fwd_bn": "hcu_conv_tc_fwd", "hcu_conv_tc_fwd_bnbwd": "hcu_conv_tc_fwd", "hcu_bn_bwd_stats_fin": "hcu_bn_bwd_stats"}


class KernelProfile:
    def __init__(self):
        self.records = []

    def __enter__(self):
        _lib.load()
        _lib._ProfState.profiler = self
        return self

    def __exit__(self, *exc):
        _lib._ProfState.profiler = None
        _lib._ProfState.note = None
        torch.cuda.synchronize()
        return False

    def table(self):
        """{kernel name: dict(ms, calls, bytes, flops)} summed over the profiled region."""
        agg = defaultdict(lambda: dict(ms=0.0, calls=0, bytes=0, flops=0))
        for name, note, e0, e1 in self.records:
            a = agg[_ALIAS.get(name, name)]
            a["ms"] += e0.elapsed_time(e1)
            a["calls"] += 1
            if note is not None:
                a["bytes"] += note[1]
                a["flops"] += note[2]
        return dict(agg)

    def by_layer(self):
        agg = defaultdict(lambda: dict(ms=0.0, calls=0, bytes=0, flops=0))
        for name, note, e0, e1 in self.records:
            key = (_ALIAS.get(name, name), note[0] if note else None)
            a = agg[key]
            a["ms"] += e0.elapsed_time(e1)
            a["calls"] += 1
            if note is not None:
                a["bytes"] += note[1]
                a["flops"] += note[2]
        return dict(agg)

    def roofline(self, peaks: dict, region_seconds: float):
        """The `roofline` object for the kernel with the largest summed duration."""
        tab = self.table()
        if not tab:
            return None
        total_ms = sum(v["ms"] for v in tab.values())
        name, top = max(tab.items(), key=lambda kv: kv[1]["ms"])
        hbm = peaks.get("hbm_gbs")
        tf = peaks.get("bf16_tflops_sustained")
        src = "measured (MEASURED_PEAKS.json)"
        if hbm is None:
            hbm, tf, src = 6650.0, 1590.0, "fallback (B200_PROFILING.md)"
        sec = top["ms"] / 1e3
        gbs = top["bytes"] / sec / 1e9 if sec > 0 else 0.0
        tfs = top["flops"] / sec / 1e12 if sec > 0 else 0.0
        t_hbm, t_tc = top["bytes"] / (hbm * 1e9), top["flops"] / (tf * 1e12)
        bound = "tensor" if t_tc > t_hbm else "hbm"
        ach, peak, unit = (tfs, tf, "TFLOP/s") if bound == "tensor" else (gbs, hbm, "GB/s")
        shares = {k: round(v["ms"] / total_ms, 4) for k, v in sorted(tab.items(), key=lambda kv: -kv[1]["ms"])[:8]}
        return {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
                "frac": ach / peak if peak else None, "traffic": None, "peak_source": src,
                "launches": top["calls"], "avg_launch_ms": top["ms"] / max(1, top["calls"]),
                "share_of_kernel_time": top["ms"] / total_ms if total_ms else None,
                "other_bound_frac": (gbs / hbm if bound == "tensor" else tfs / tf),
                "kernel_time_shares": shares,
                "timing": "CUDA events around every C-ABI call on the launching stream (the caller parks the GPU behind a spin "
                          "while the host enqueues, so the pairs bracket kernel time); instrumented steps right after "
                          "the timed region (same step, same inputs)"}
