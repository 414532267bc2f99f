"""CPU oracle for the MIXED path: the reference's algorithm with fp16 STORAGE emulated (TEST INFRASTRUCTURE).

``oracle/unet_oracle.py`` restates the reference in fp32 and is pinned bit-for-bit to reference-minted fixtures.
The benchmarked GPU path stores activations and gradients in fp16 (fp32 accumulate); its deviation from the fp32
reference is dominated by those roundings, amplified by the random-init batch-statistics network.  This file restates
the SAME control flow (`/root/reference/hcat/unet.py:125-143,236-340`, `/root/reference/hcat/loss.py:65-72`) with an
EXPLICIT backward pass and a rounding hook at every point where the engine stores a tensor in fp16
(`hcunet_b200/engine.py`, `hcunet_b200/csrc/conv_tc.cu`, `elementwise.cu`):

    forward : input -> fp16; packed weights -> fp16 (`Up.conv1`: K halves folded in fp32 first, `cat(x, x)`);
              conv / transposed conv: exact fp16 x fp16 products, fp32 accumulate, + bias (fp32) -> y stored fp16;
              BatchNorm statistics from the fp32 accumulators (fp64 reduction), scale = gamma * invstd,
              shift = beta - mean * scale in fp32; the consumer forms relu(fma(y16, scale, shift)) in fp32 and rounds
              it to fp16 (conv operand / pooled value; the arg-max is taken on the fp32 values); logits stay fp32.
    backward: dlogits * S (S = power of two with max|dlogits| * S in (32, 64]) -> fp16; ReLU mask from
              fma(y16, scale, shift) > 0; BN backward sums in fp64, dy = fma(c1, g, fma(c2, y16, c3)) -> fp16;
              weight gradient = sum(operand16 x dy16) in fp32, un-scaled by 1/S; data gradient -> fp16.

With ``emulate=False`` every hook is the identity and the explicit backward must reproduce the reference's autograd
gradients: ``tests/test_oracle_golden.py`` pins that against the reference-minted fixtures (<= 2e-6), so the only
difference between this oracle and the pinned one is WHERE values are rounded to fp16.  The GPU mixed path is gated
against this oracle (``tests/test_gpu_parity.py``): what remains is accumulation order inside fp32 sums.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from .unet_oracle import cross_entropy, normalise_spec

FLOOR_FACTOR = 5.0
"""A tensor's gate is max(stated tolerance, FLOOR_FACTOR x its measured accumulation-order floor): the floor is the largest of
three samples of a heavy-tailed spread, the implementation under test is a fourth draw.  (4.0 until one of six full GPU runs
of one day put the strict-fp32 path's `up_steps.1.batch1.bias` gradient of the full-size cfg1 fixture at 4.0009 x its floor --
5.460e-3 against a gate of 5.459e-3; the fp32 kernels' atomics land in a different order every run.)"""


def gate(stated: float, floor: float) -> float:
    return max(stated, FLOOR_FACTOR * floor)


BN_EPS = 1e-5
BN_MOMENTUM = 0.1
GRAD_SCALE_TARGET = 64.0


def _t(v, dims):
    return (v,) * dims if isinstance(v, int) else tuple(v)


class _Emu:
    """Rounding hooks.  emulate=False: identities (fp32 reference arithmetic)."""

    def __init__(self, emulate: bool):
        self.on = emulate

    def r16(self, t):
        return t.half().float() if self.on else t

    def fma(self, a, b, c):
        """fp32 fused multiply-add (one rounding) of per-channel-broadcast operands."""
        if self.on:
            return (a.double() * b.double() + c.double()).float()
        return a * b + c


def _cshape(v, ndim):
    return v.view((1, -1) + (1,) * (ndim - 2))


ACCUMULATE = "fp32"
"""How the fp32 accumulations are carried out: "fp32" (oneDNN's order, BatchNorm statistics reduced in fp64), "fp64"
(convolutions accumulated in double, rounded once), "split" (two half-K partial sums added at the end) or "stats32"
(BatchNorm sums and sums of squares reduced in fp32, as the conv epilogues' per-CTA partial sums are: ~1e-6 relative on
mean / variance), or ("noise", seed): every fp32 accumulator (conv outputs, BatchNorm forward / backward sums) multiplied
by 1 + eps * N(0, 1) with eps = one fp32 rounding (6e-8; 3e-7 for the long BatchNorm sums) -- a generic model of "the
same sum added in another order", of which any number of independent draws can be taken.  All are legitimate results of
the same fp16-storage arithmetic; the spread between them is the ACCUMULATION-ORDER FLOOR of a case
(``accumulation_floor``): how far two correct implementations may differ on it."""
_NOISE = [None]


def _noisy(t, eps):
    """Multiplicative rounding-level noise when ACCUMULATE is ("noise", seed); identity otherwise."""
    if not (isinstance(ACCUMULATE, tuple) and ACCUMULATE[0] == "noise"):
        return t
    if _NOISE[0] is None or _NOISE[0][0] != ACCUMULATE[1]:
        _NOISE[0] = (ACCUMULATE[1], torch.Generator().manual_seed(7919 * ACCUMULATE[1] + 1))
    return t * (1 + eps * torch.randn(t.shape, generator=_NOISE[0][1], dtype=t.dtype))


def _conv_fn(dims, transposed):
    if transposed:
        f = F.conv_transpose2d if dims == 2 else F.conv_transpose3d
    else:
        f = F.conv2d if dims == 2 else F.conv3d
    if ACCUMULATE == "fp64":
        def f64(a, w, b, **kw):
            return f(a.double(), w.double(), None if b is None else b.double(), **kw).float()
        return f64
    if isinstance(ACCUMULATE, tuple):
        def fnoise(a, w, b, **kw):
            return _noisy(f(a, w, b, **kw), 6e-8)
        return fnoise
    if ACCUMULATE == "split":
        def fsplit(a, w, b, **kw):
            c = a.shape[1]
            if kw.get("groups", 1) != 1 or c < 2:
                return f(a, w, b, **kw)
            h = c // 2
            if transposed:  # weight [Cin, Cout, k]
                return f(a[:, :h], w[:h], b, **kw) + f(a[:, h:], w[h:], None, **kw)
            return f(a[:, :h], w[:, :h], b, **kw) + f(a[:, h:], w[:, h:], None, **kw)
        return fsplit
    return f


def _op_grads(fn, a, w, dy, need_da, **kw):
    """(d a, d w) of out = fn(a, w, **kw) for upstream dy -- autograd of the single primitive."""
    a = a.detach().requires_grad_(need_da)
    w = w.detach().requires_grad_(True)
    with torch.enable_grad():
        out = fn(a, w, None, **kw)
    gs = torch.autograd.grad(out, (a, w) if need_da else (w,), dy)
    return (gs[0], gs[1]) if need_da else (None, gs[0])


def train_step(sd: Dict[str, torch.Tensor], spec: dict, x, mask, pwl, method="pixel", emulate=True, training=True,
               taps: Dict[str, torch.Tensor] = None
               ) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """One training forward + loss + explicit backward.  Returns (loss, logits, {param: grad}, new_buffers).
    ``taps`` (optional dict) receives every tensor the engine stores in HBM during the step, under the tag the engine's
    ``tap`` hook uses for it (NCDHW fp32 holding the fp16 values): teacher-forced layer-by-layer parity."""
    E = _Emu(emulate)
    spec = normalise_spec(spec)
    dims = spec["image_dimensions"]
    feats = list(spec["feature_sizes"])
    nlev = len(feats)
    conv, convT = _conv_fn(dims, False), _conv_fn(dims, True)
    pool_k = _t(spec["max_pool_kernel"], dims)
    pool = F.max_pool2d if dims == 2 else F.max_pool3d
    if taps is None:
        taps = {}
    new_buffers: Dict[str, torch.Tensor] = {}
    tape = []  # backward records, in forward order

    def bn_block(a16, prefix, idx, up_first):
        """conv{idx} -> batch{idx} -> relu (`unet.py:263-266,313-314`).  a16: the operand as the kernel sees it."""
        w = sd[f"{prefix}.conv{idx}.weight"]
        g = spec["groups"][f"conv{idx}"]
        fold = False
        if up_first:
            if g == 1:  # conv1(cat(x, x)): W_eff = W[:, :C] + W[:, C:], folded in fp32 before the fp16 pack
                c = w.shape[1] // 2
                w_eff, fold, g_run = w[:, :c] + w[:, c:], True, 1
            elif g == 2:  # group j reads cat channels [jC, (j+1)C) == all of x: dense conv, weight as stored
                w_eff, g_run = w, 1
            else:
                raise NotImplementedError("groups > 2 on Up.conv1")
        else:
            w_eff, g_run = w, g
        w16 = E.r16(w_eff)
        kw = dict(stride=1, padding=0, dilation=spec["dilation"][f"conv{idx}"], groups=g_run)
        y32 = conv(a16, w16, sd[f"{prefix}.conv{idx}.bias"], **kw)
        bnp = f"{prefix}.batch{idx}"
        gamma, beta = sd[bnp + ".weight"], sd[bnp + ".bias"]
        red = [0] + list(range(2, y32.dim()))
        n = y32.numel() // y32.shape[1]
        if training:
            if ACCUMULATE == "stats32":
                mu = y32.sum(dim=red).double() / n
                var = ((y32 * y32).sum(dim=red).double() / n - mu * mu).clamp_min(0.0)
            else:
                yd = y32.double()
                mu = yd.mean(dim=red)
                ex2 = _noisy((yd * yd).mean(dim=red), 3e-7)
                mu = _noisy(mu, 3e-7)
                var = (ex2 - mu * mu).clamp_min(0.0)
            invstd = (1.0 / torch.sqrt(var + BN_EPS)).float()
            mean = mu.float()
            unbiased = var * n / (n - 1) if n > 1 else var
            new_buffers[bnp + ".running_mean"] = (1 - BN_MOMENTUM) * sd[bnp + ".running_mean"] + BN_MOMENTUM * mean
            new_buffers[bnp + ".running_var"] = (1 - BN_MOMENTUM) * sd[bnp + ".running_var"] + BN_MOMENTUM * unbiased.float()
            new_buffers[bnp + ".num_batches_tracked"] = sd[bnp + ".num_batches_tracked"] + 1
        else:
            mean = sd[bnp + ".running_mean"]
            invstd = torch.rsqrt(sd[bnp + ".running_var"] + BN_EPS)
        scale = gamma * invstd
        shift = beta - mean * scale
        y16 = E.r16(y32)
        z32 = E.fma(y16, _cshape(scale, y16.dim()), _cshape(shift, y16.dim()))  # pre-ReLU, fp32
        taps[f"{prefix}.conv{idx}.y"] = y16
        taps[f"{prefix}.conv{idx}.bn"] = torch.stack([mean, invstd, scale, shift])
        tape.append(dict(kind="conv", prefix=prefix, idx=idx, a16=a16, w16=w16, kw=kw, fold=fold, y16=y16, pos=z32 > 0,
                         scale=scale, mean=mean, invstd=invstd, gamma=gamma, n=n, wshape=w.shape))
        return torch.relu(z32)  # fp32 post-activation; the reader rounds it

    a = E.r16(x)
    taps["input"] = a
    first = True
    for i in range(nlev):
        p = f"down_steps.{i}"
        h = bn_block(a, p, "1", False)
        tape[-1]["first"] = first
        first = False
        h = bn_block(E.r16(h), p, "2", False)
        if i < nlev - 1:  # unet.py:131; arg-max on the fp32 values, pooled tensor stored fp16
            pooled, idx = pool(h, pool_k, return_indices=True)
            tape.append(dict(kind="pool", idx=idx, in_shape=h.shape))
            a = E.r16(pooled)
            taps[f"{p}.conv2.pool"] = a
        else:
            a = E.r16(h)
    for i in range(nlev - 1):  # Up.forward, unet.py:309-315
        p = f"up_steps.{i}"
        w16 = E.r16(sd[p + ".up_conv.weight"])
        kw = dict(stride=spec["upsample_stride"], padding=0)
        up32 = convT(a, w16, sd[p + ".up_conv.bias"], **kw)
        tape.append(dict(kind="up", prefix=p, a16=a, w16=w16, kw=kw))
        a = E.r16(up32)
        taps[p + ".up_conv.out"] = a
        h = bn_block(a, p, "1", True)     # crop(x_up, skip) is a no-op: conv1(cat(x_up, x_up))
        h = bn_block(E.r16(h), p, "2", False)
        a = E.r16(h)
    w16 = E.r16(sd["out_conv.weight"])
    logits = conv(a, w16, sd["out_conv.bias"])
    tape.append(dict(kind="out", a16=a, w16=w16))
    taps["logits"] = logits

    # ---- loss + dlogits (fp32, like the loss kernels) ----------------------------------------------
    lg = logits.detach().requires_grad_(True)
    with torch.enable_grad():
        loss = cross_entropy(lg, mask, pwl, method)
    (dlogits,) = torch.autograd.grad(loss, lg)

    # ---- backward -----------------------------------------------------------------------------------
    S = 1.0
    if emulate:
        amax = float(dlogits.abs().max())
        if amax > 0:
            _, e = math.frexp(GRAD_SCALE_TARGET / amax)
            S = 2.0 ** max(-60, min(60, e - 1))
    inv = 1.0 / S
    grads: Dict[str, torch.Tensor] = {}
    d = E.r16(dlogits * S)
    taps["dlogits"] = d
    taps["grad_scale"] = torch.tensor(S)
    sum_dims = lambda t: [0] + list(range(2, t.dim()))
    for rec in reversed(tape):
        kind = rec["kind"]
        if kind == "out":
            grads["out_conv.bias"] = (d.double().sum(dim=sum_dims(d)) * inv).float()
            da, dw = _op_grads(conv, rec["a16"], rec["w16"], d, True)
            grads["out_conv.weight"] = dw * inv
            d = E.r16(da)
            taps["out_conv.dgrad"] = d
        elif kind == "pool":
            unpool = F.max_unpool2d if dims == 2 else F.max_unpool3d
            d = unpool(d, rec["idx"], pool_k, output_size=rec["in_shape"][2:])
        elif kind == "up":
            p = rec["prefix"]
            grads[p + ".up_conv.bias"] = (d.double().sum(dim=sum_dims(d)) * inv).float()
            da, dw = _op_grads(convT, rec["a16"], rec["w16"], d, True, **rec["kw"])
            grads[p + ".up_conv.weight"] = dw * inv
            d = E.r16(da)
            taps[p + ".up_conv.dgrad"] = d
        else:
            p, idx = rec["prefix"], rec["idx"]
            y16, n = rec["y16"], rec["n"]
            g0 = torch.where(rec["pos"], d, torch.zeros_like(d))
            red = sum_dims(d)
            if ACCUMULATE == "stats32":   # fp32 partial sums, like the per-thread accumulators of the BN-backward kernels
                sg = g0.sum(dim=red).double()
                sgy = (g0 * y16).sum(dim=red).double()
            else:
                sg = _noisy(g0.double().sum(dim=red), 3e-7)
                sgy = _noisy((g0.double() * y16.double()).sum(dim=red), 3e-7)
            mean, invstd = rec["mean"].double(), rec["invstd"].double()
            sgx = invstd * (sgy - mean * sg)
            grads[f"{p}.batch{idx}.weight"] = (sgx * inv).float()
            grads[f"{p}.batch{idx}.bias"] = (sg * inv).float()
            s = rec["gamma"].double() * invstd
            if training:
                mg, mgx = sg / n, sgx / n
                c1, c2, c3 = s, -s * invstd * mgx, s * invstd * mgx * mean - s * mg
                grads[f"{p}.conv{idx}.bias"] = ((c1 * sg + c2 * n * mean + c3 * n) * inv).float()
            else:
                c1, c2, c3 = s, torch.zeros_like(s), torch.zeros_like(s)
                grads[f"{p}.conv{idx}.bias"] = (s * sg * inv).float()
            nd = d.dim()
            c1f, c2f, c3f = (_cshape(c.float(), nd) for c in (c1, c2, c3))
            dy = E.r16(E.fma(c1f, g0, E.fma(c2f, y16, c3f)))
            taps[f"{p}.conv{idx}.dy"] = dy
            need_da = not rec.get("first", False)
            da, dw = _op_grads(conv, rec["a16"], rec["w16"], dy, need_da, **rec["kw"])
            dw = dw * inv
            if rec["fold"]:
                dw = torch.cat([dw, dw], dim=1)
            grads[f"{p}.conv{idx}.weight"] = dw.reshape(rec["wshape"])
            d = E.r16(da) if need_da else None
            if need_da:
                taps[f"{p}.conv{idx}.dgrad"] = d
    return loss.detach(), logits.detach(), grads, new_buffers


def eval_forward(sd: Dict[str, torch.Tensor], spec: dict, x, emulate=True, taps: Dict[str, torch.Tensor] = None) -> torch.Tensor:
    """Inference forward as the engine runs it (`engine.forward`, ``fold_eval``): BatchNorm folded into the conv epilogue,
    out = relu(fma(acc, scale, shift)) with scale = gamma / sqrt(running_var + eps), shift = beta + (bias - running_mean)
    * scale, stored fp16; pooled / up-convolved tensors stored fp16; logits fp32."""
    E = _Emu(emulate)
    spec = normalise_spec(spec)
    dims = spec["image_dimensions"]
    nlev = len(spec["feature_sizes"])
    conv, convT = _conv_fn(dims, False), _conv_fn(dims, True)
    pool = F.max_pool2d if dims == 2 else F.max_pool3d

    def block(a, prefix, idx, up_first):
        w = sd[f"{prefix}.conv{idx}.weight"]
        g = spec["groups"][f"conv{idx}"]
        if up_first and g == 1:
            c = w.shape[1] // 2
            w, g = w[:, :c] + w[:, c:], 1
        elif up_first and g == 2:
            g = 1
        acc = conv(a, E.r16(w), None, stride=1, padding=0, dilation=spec["dilation"][f"conv{idx}"], groups=g)
        bnp = f"{prefix}.batch{idx}"
        isd = 1.0 / torch.sqrt(sd[bnp + ".running_var"] + BN_EPS)
        scale = sd[bnp + ".weight"] * isd
        shift = sd[bnp + ".bias"] + (sd[f"{prefix}.conv{idx}.bias"] - sd[bnp + ".running_mean"]) * scale
        out = E.r16(torch.relu(E.fma(acc, _cshape(scale, acc.dim()), _cshape(shift, acc.dim()))))
        taps[f"{prefix}.conv{idx}.a"] = out
        return out

    if taps is None:
        taps = {}
    a = E.r16(x)
    taps["input"] = a
    for i in range(nlev):
        a = block(block(a, f"down_steps.{i}", "1", False), f"down_steps.{i}", "2", False)
        if i < nlev - 1:
            a = pool(a, spec["max_pool_kernel"])
            taps[f"down_steps.{i}.conv2.pool"] = a
    for i in range(nlev - 1):
        p = f"up_steps.{i}"
        a = E.r16(convT(a, E.r16(sd[p + ".up_conv.weight"]), sd[p + ".up_conv.bias"], stride=spec["upsample_stride"],
                        padding=0))
        taps[p + ".up_conv.out"] = a
        a = block(block(a, p, "1", True), p, "2", False)
    taps["logits"] = conv(a, E.r16(sd["out_conv.weight"]), sd["out_conv.bias"])
    return taps["logits"]


def accumulation_floor(sd, spec, x, mask, pwl, base=None, method="pixel", draws=3):
    """Per-tensor spread of the emulated step under a change of the fp32 accumulation order (see ``ACCUMULATE``).
    Returns (base result, {"logits": rel-L2, "agree": fraction, param name: rel-L2}) -- the largest deviation of the "fp64",
    "split" and "stats32" variants (+ ``draws - 3`` ("noise", seed) draws: small cases, where single rounding flips are rare
    events, need more samples of the spread) from the default one."""
    global ACCUMULATE

    def rel(a, b):
        a, b = a.double(), b.double()
        return float((a - b).norm() / b.norm().clamp_min(1e-30))

    if base is None:
        base = train_step(sd, spec, x, mask, pwl, method)
    floor = {"logits": 0.0, "agree": 1.0}
    saved = ACCUMULATE
    try:
        for mode in ["fp64", "split", "stats32"] + [("noise", i) for i in range(max(0, draws - 3))]:
            ACCUMULATE = mode
            _NOISE[0] = None
            _, lg, gr, _ = train_step(sd, spec, x, mask, pwl, method)
            floor["logits"] = max(floor["logits"], rel(lg, base[1]))
            floor["agree"] = min(floor["agree"], float(((lg > 0) == (base[1] > 0)).float().mean()))
            for k, g in gr.items():
                floor[k] = max(floor.get(k, 0.0), rel(g, base[2][k]))
    finally:
        ACCUMULATE = saved
    return base, floor


def eval_accumulation_floor(sd, spec, x, draws=8):
    """(eval logits of the emulation, their accumulation-order spread as rel-L2, worst thresholded-mask agreement)."""
    global ACCUMULATE
    base = eval_forward(sd, spec, x)
    spread, agree = 0.0, 1.0
    saved = ACCUMULATE
    try:
        for mode in ["fp64", "split"] + [("noise", i) for i in range(max(0, draws - 2))]:
            ACCUMULATE = mode
            _NOISE[0] = None
            lg = eval_forward(sd, spec, x)
            spread = max(spread, float((lg.double() - base.double()).norm() / base.double().norm().clamp_min(1e-30)))
            agree = min(agree, float(((lg > 0) == (base > 0)).float().mean()))
    finally:
        ACCUMULATE = saved
    return base, spread, agree
