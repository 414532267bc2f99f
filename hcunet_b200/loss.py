"""Drop-in mirror of ``hcat/loss.py`` (`loss.py:5-177`) on the B200 kernels.

``cross_entropy`` / ``dice`` / ``L1Loss`` / ``MSELoss`` keep the reference's signatures, argument meaning,
error behaviour and quirks (origin crop of mask / pwl `loss.py:51-56`; ``pwl=None`` means weight 2 because
`loss.py:48` sets ``is_pwl_none = True`` unconditionally; ``'sigmoid'`` applies BCE-with-logits to
``sigmoid(pred)``).  The element-wise BCE and its gradient are one fused memory-bound kernel each
(``hcu_wbce_fwd`` / ``hcu_wbce_bwd``): mask and pwl are read in their storage dtype (fp16 from the
reference dataloader is fine), the crop is folded into the index arithmetic, nothing is materialised.
CUDA only -- no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import HcuLossDesc

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(pred):
    if not isinstance(pred, torch.Tensor):
        raise TypeError(f"pred must be a torch.Tensor, not {type(pred)}")
    if not pred.is_cuda:
        raise RuntimeError("hcunet_b200.loss runs on CUDA (sm_100a) only: there is no CPU fallback")


def _prep_side(t, pred, name):
    """mask / pwl: same device, supported storage dtype, contiguous; shape >= pred's (origin crop)."""
    if t.device != pred.device:
        raise RuntimeError(f"Expected all tensors to be on the same device, but found {name} on {t.device} and pred on "
                           f"{pred.device}")
    if t.dtype not in _DT:
        t = t.float()
    return t.contiguous()


def _desc(pred, mask, pwl, mode=0):
    n_dim = pred.dim()
    if n_dim not in (4, 5):  # loss.py:57-59
        raise IndexError('Unexpected number of predicted mask dimensions. Expected 4 (2D) or 5 (3D) but got' +
                         f' {n_dim} dimensions: {pred.shape}')
    if mask.dim() != n_dim or (pwl is not None and pwl.dim() != n_dim):
        raise IndexError(f"mask / pwl must have {n_dim} dimensions like pred")
    ps = list(pred.shape) + [1] * (5 - n_dim)
    ms = list(mask.shape) + [1] * (5 - n_dim)
    if pwl is not None and list(pwl.shape) != list(mask.shape):
        raise RuntimeError(f"mask {tuple(mask.shape)} and pwl {tuple(pwl.shape)} must have the same shape")
    if ms[0] != ps[0] or ms[1] != ps[1] or any(ms[i] < ps[i] for i in (2, 3, 4)):
        # the reference's slice would silently yield a smaller tensor and then fail to broadcast
        raise RuntimeError(f"The size of tensor a {tuple(pred.shape)} must match the size of tensor b "
                           f"{tuple(mask.shape)} after cropping")
    d = HcuLossDesc()
    d.b, d.c, d.x, d.y, d.z = ps
    d.mx, d.my, d.mz = ms[2], ms[3], ms[4]
    d.dtype_mask = _DT[mask.dtype]
    d.dtype_pwl = _DT[pwl.dtype] if pwl is not None else _lib.F32
    d.mode = mode
    return d


class _WBCE(torch.autograd.Function):
    """mean / worst_z reductions of BCEWithLogits(pred, mask) * (pwl + 1)  (`loss.py:65-80,97-101`)."""

    @staticmethod
    def forward(ctx, pred, mask, pwl, mode, worst_z, weight_mult):
        lib = _lib.load()
        pred = pred.contiguous().float()
        d = _desc(pred, mask, pwl, mode)
        dev = pred.device
        n = pred.numel()
        out = torch.zeros(1, dtype=torch.float64, device=dev)
        zs = torch.zeros(d.z, dtype=torch.float64, device=dev) if worst_z else None
        _lib.check(lib.hcu_wbce_fwd(C.byref(d), _ptr(pred), _ptr(mask), _ptr(pwl), _ptr(out), _ptr(zs), _stream()),
                   "wbce_fwd")
        ctx.d, ctx.mode = d, mode
        ctx.weight_mult = weight_mult
        if worst_z:
            # loss.py:74-80: per-z sums sorted ascending, times linspace(1, 2, Z)^2, / (X * Y), then mean over Z
            Z = d.z
            scaling = (torch.linspace(1, 2, Z) ** 2).to(dev)
            srt, order = torch.sort(zs.float())
            val = ((srt * scaling) / (d.x * d.y)).mean()
            zscale = torch.empty(Z, dtype=torch.float32, device=dev)
            zscale[order] = scaling / (d.x * d.y) / Z
            ctx.mult = 1.0
            ctx.save_for_backward(pred, mask, pwl, zscale)
            return val
        ctx.mult = weight_mult / n
        ctx.save_for_backward(pred, mask, pwl, None)
        return (out * (weight_mult / n)).float().reshape(())

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        pred, mask, pwl, zscale = ctx.saved_tensors
        g = gout.contiguous().float().reshape(1)
        dpred = torch.empty_like(pred)
        _lib.check(lib.hcu_wbce_bwd(C.byref(ctx.d), _ptr(pred), _ptr(mask), _ptr(pwl), _ptr(g), float(ctx.mult),
                                    _ptr(zscale), _ptr(dpred), _stream()), "wbce_bwd")
        return dpred, None, None, None, None, None


class _Pair(torch.autograd.Function):
    """dice / L1 / MSE (`loss.py:104-177`): kind 0 / 1 / 2."""

    @staticmethod
    def forward(ctx, pred, mask, kind):
        lib = _lib.load()
        pred = pred.contiguous().float()
        d = _desc(pred, mask, None)
        sums = torch.zeros(3, dtype=torch.float64, device=pred.device)
        _lib.check(lib.hcu_pair_reduce(C.byref(d), kind, _ptr(pred), _ptr(mask), _ptr(sums), _stream()), "pair_reduce")
        n = pred.numel()
        ctx.d, ctx.kind, ctx.n = d, kind, n
        if kind == 0:
            num = 2 * sums[0] + 1e-10
            den = sums[1] + sums[2] + 1e-10
            ctx.save_for_backward(pred, mask, num, den)
            return (1 - num / den).float()
        ctx.save_for_backward(pred, mask, None, None)
        return (sums[0] / n).float()

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        pred, mask, num, den = ctx.saved_tensors
        g = gout.double().reshape(())
        if ctx.kind == 0:
            coef = torch.stack([-2.0 * g / den, g * num / (den * den)]).float()
        elif ctx.kind == 1:
            coef = torch.stack([g / ctx.n, g * 0]).float()
        else:
            coef = torch.stack([2.0 * g / ctx.n, g * 0]).float()
        dpred = torch.empty_like(pred)
        _lib.check(lib.hcu_pair_bwd(C.byref(ctx.d), ctx.kind, _ptr(pred), _ptr(mask), _ptr(coef), _ptr(dpred),
                                    _stream()), "pair_bwd")
        return dpred, None, None


def cross_entropy(pred: torch.Tensor, mask: torch.Tensor, pwl: torch.Tensor, method='pixel', num_random_pixels=None):
    """`loss.py:5-101`.  ``method`` in {'pixel', 'worst_z', 'random', 'sigmoid'}."""
    _methods = ['pixel', 'worst_z', 'random', 'sigmoid']
    if method not in _methods:
        raise ValueError(f'Viable methods for cross entropy loss are {_methods}, not {method}.')
    if method == 'random':
        if num_random_pixels is None:
            raise ValueError('the number of random pixels to draw is not defined. Please set num_random_pixels to a ' +
                             'value larger than 1.')
        if num_random_pixels <= 1:
            raise ValueError(f'num_random_pixels should be greater than 1 not {num_random_pixels}.')
        if (mask == 0).sum() == 0:
            raise ValueError('There are no background pixels in mask.\n\t(mask==0).sum() == 0 -> True')
    _require_cuda(pred)
    n_dim = pred.dim()
    if n_dim not in (4, 5):
        raise IndexError('Unexpected number of predicted mask dimensions. Expected 4 (2D) or 5 (3D) but got' +
                         f' {n_dim} dimensions: {pred.shape}')
    mask = _prep_side(mask, pred, "mask")
    if pwl is not None:
        pwl = _prep_side(pwl, pred, "pwl")
    if method == 'worst_z' and n_dim != 5:
        raise IndexError("tuple index out of range")  # pred.shape[4] in loss.py:77
    if method == 'random':
        return _random_pixels(pred, mask, num_random_pixels)
    mode = 1 if method == 'sigmoid' else 0
    return _WBCE.apply(pred, mask, pwl, mode, method == 'worst_z', 1.0)


def _random_pixels(pred, mask, num_random_pixels):
    """`loss.py:82-95`: consumes the global CPU RNG (torch.randint) exactly like the reference; the pixel
    selection is index plumbing, the BCE itself runs in the fused kernel (plain BCE == weight-2 BCE / 2)."""
    shape = pred.shape
    if pred.dim() == 5:
        mask = mask[:, :, 0:shape[2], 0:shape[3], 0:shape[4]]
    else:
        mask = mask[:, :, 0:shape[2], 0:shape[3]]
    p = pred.reshape(-1)
    m = mask.reshape(-1).float()
    npos = int((m == 1).sum())
    if npos != 0:
        pos_ind = torch.randint(low=0, high=npos, size=(1, num_random_pixels))[0, :].to(pred.device)
        neg_ind = torch.randint(low=0, high=int((m == 0).sum()), size=(1, num_random_pixels))[0, :].to(pred.device)
        p = torch.cat([p[m == 1][pos_ind], p[m == 0][neg_ind]])
        m = torch.cat([m[m == 1][pos_ind], m[m == 0][neg_ind]])
    n = p.numel()
    return _WBCE.apply(p.reshape(1, 1, 1, 1, n), m.reshape(1, 1, 1, 1, n).contiguous(), None, 0, False, 0.5)


def dice(pred: torch.Tensor, mask: torch.Tensor):
    """`loss.py:104-127`."""
    _require_cuda(pred)
    return _Pair.apply(pred, _prep_side(mask, pred, "mask"), 0)


def L1Loss(pred: torch.Tensor, mask: torch.Tensor):
    """`loss.py:130-152`."""
    _require_cuda(pred)
    return _Pair.apply(pred, _prep_side(mask, pred, "mask"), 1)


def MSELoss(pred: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """`loss.py:155-177`."""
    _require_cuda(pred)
    return _Pair.apply(pred, _prep_side(mask, pred, "mask"), 2)
