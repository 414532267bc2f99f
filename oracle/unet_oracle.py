"""CPU oracle: a functional restatement of the reference U-Net hot path (TEST INFRASTRUCTURE).

This file is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  ``hcunet_b200`` never
does, and fails loudly when its CUDA library is missing.

What it restates
----------------
The reference (`/root/reference/hcat/unet.py`, `/root/reference/hcat/loss.py`) is pure Python that
delegates all arithmetic to the third-party dependency **torch** (ATen + oneDNN on CPU).  The
reference repo pins no version (no requirements.txt / setup.py); this oracle is evaluated with the
image's torch 2.11.0 CPU fp32 and that is the version the golden vectors were minted with
(``oracle/make_golden.py``).  Each function below cites the reference lines it follows and uses
only ``torch.nn.functional`` primitives on explicit weights from a reference-layout ``state_dict``
-- it does not construct ``nn.Module`` s, so it is an independent restatement of the control flow
(the dead skip connection, the origin crop, the unconditional ``is_pwl_none = True`` ...).

Pinning
-------
``tests/test_oracle_golden.py`` checks this oracle bit-for-bit / to 1e-6 against
(a) committed golden vectors produced by executing the UNMODIFIED reference modules
(``tests/golden/*.pt``, minted by ``oracle/make_golden.py``) and (b), when ``/root/reference``
is mounted, the live reference.  The reference's own ``tests/`` hold no golden vectors or asserts
for this path, so (a)/(b) are the pin.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


def _as_dict(v, keys=("conv1", "conv2")):
    """`unet.py:59-64`: tuples / ints are broadcast to {'conv1','conv2'} dicts."""
    if isinstance(v, dict):
        return v
    return {k: v for k in keys}


def normalise_spec(spec: dict) -> dict:
    """Normalise constructor kwargs exactly like `unet.py:59-85`."""
    s = dict(spec)
    s.setdefault("image_dimensions", 2)
    s.setdefault("in_channels", 3)
    s.setdefault("out_channels", 2)
    s.setdefault("feature_sizes", [32, 64, 128, 256, 512, 1024])
    s.setdefault("kernel", (3, 3))
    s.setdefault("upsample_kernel", (2, 2))
    s.setdefault("max_pool_kernel", (2, 2))
    s.setdefault("upsample_stride", 2)
    s.setdefault("dilation", 1)
    s.setdefault("groups", 1)
    if type(s["kernel"]) is tuple:
        s["kernel"] = _as_dict(s["kernel"])
    if type(s["dilation"]) is int or type(s["dilation"]) is tuple:
        s["dilation"] = _as_dict(s["dilation"])
    if type(s["groups"]) is int or type(s["groups"]) is tuple:
        s["groups"] = _as_dict(s["groups"])
    return s


def _bn(x, sd, prefix, training, new_buffers, momentum=0.1, eps=1e-5):
    """`unet.py:259-260,305-306`: nn.BatchNorm{2,3}d defaults (momentum .1, eps 1e-5, affine,
    track_running_stats).  Training: batch statistics (biased var) normalise, running stats get
    the unbiased var; eval: running stats normalise."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training:
        rm2, rv2 = rm.detach().clone(), rv.detach().clone()
        y = F.batch_norm(x, rm2, rv2, w, b, True, momentum, eps)
        new_buffers[prefix + ".running_mean"] = rm2
        new_buffers[prefix + ".running_var"] = rv2
        new_buffers[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
        return y
    return F.batch_norm(x, rm, rv, w, b, False, momentum, eps)


def tf32_round(t: torch.Tensor) -> torch.Tensor:
    """Round fp32 to TF32 (10-bit mantissa, round-half-up in magnitude) -- what cuDNN's default
    ``allow_tf32=True`` path does to conv inputs on the GPU.  Used only to CALIBRATE
    the tolerance of the mixed-precision parity tests: it answers "how far from fp32 is the reference's own
    default GPU arithmetic on this case"."""
    i = t.detach().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class _TF32Conv(torch.autograd.Function):
    """conv / convT whose forward AND backward round their tensor-core inputs to TF32, like cuDNN with
    ``allow_tf32=True`` (fwd: x, w; dgrad: dy, w; wgrad: x, dy).  Calibration only (see tf32_round)."""

    @staticmethod
    def forward(ctx, x, w, b, fn, kw):
        ctx.save_for_backward(x, w)
        ctx.fn, ctx.kw = fn, kw
        return fn(tf32_round(x), tf32_round(w), b, **kw)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        with torch.enable_grad():
            xx, ww = tf32_round(x).requires_grad_(True), tf32_round(w).requires_grad_(True)
            dx, dw = torch.autograd.grad(ctx.fn(xx, ww, None, **ctx.kw), (xx, ww), tf32_round(dy))
        return dx, dw, dy.sum(dim=[0] + list(range(2, dy.dim()))), None, None


def _tconv(fn, tf32, x, w, b, **kw):
    if not tf32:
        return fn(x, w, b, **kw)
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad):
        return _TF32Conv.apply(x, w, b, fn, kw)
    return fn(tf32_round(x), tf32_round(w), b, **kw)


def _block(x, sd, prefix, spec, training, new_buffers, acts, tf32=False):
    """`unet.py:263-266` / `unet.py:313-314`: relu(bn1(conv1(x))); relu(bn2(conv2(x))), padding 0."""
    conv = F.conv2d if spec["image_dimensions"] == 2 else F.conv3d
    for i in ("1", "2"):
        x = _tconv(conv, tf32, x, sd[f"{prefix}.conv{i}.weight"], sd[f"{prefix}.conv{i}.bias"], stride=1, padding=0,
                   dilation=spec["dilation"][f"conv{i}"], groups=spec["groups"][f"conv{i}"])
        if acts is not None:
            acts[f"{prefix}.conv{i}"] = x
        x = F.relu(_bn(x, sd, f"{prefix}.batch{i}", training, new_buffers))
        if acts is not None:
            acts[f"{prefix}.relu{i}"] = x
    return x


def _crop(x, y):
    """`unet.py:318-340`: crop(x, y) slices **x** to y's spatial shape from the origin."""
    assert x.shape[1] == y.shape[1], f"Inputs do not have same number of feature dimmensions: {x.shape} | {y.shape}"
    if x.dim() == 4:
        return x[:, :, 0:y.shape[2], 0:y.shape[3]]
    return x[:, :, 0:y.shape[2], 0:y.shape[3], 0:y.shape[4]]


def unet_forward(sd: Dict[str, torch.Tensor], spec: dict, x: torch.Tensor, training: bool = False,
                 acts: Optional[dict] = None, tf32: bool = False) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """Restates ``Unet_Constructor.forward`` (`unet.py:125-143`) on a reference-layout state_dict.

    Returns (logits, new_buffers); ``new_buffers`` holds the BN running stats a train-mode forward
    would have written in place.  ``acts`` (optional dict) collects every intermediate tensor.
    ``tf32=True`` rounds every conv / convT tensor-core input (forward and backward) to TF32 first.
    """
    spec = normalise_spec(spec)
    dims = spec["image_dimensions"]
    if dims not in (2, 3):
        raise ValueError(f"Does not support {dims} dimensional images")  # unet.py:53
    nlev = len(spec["feature_sizes"])
    pool = F.max_pool2d if dims == 2 else F.max_pool3d
    convT = F.conv_transpose2d if dims == 2 else F.conv_transpose3d
    conv = F.conv2d if dims == 2 else F.conv3d
    new_buffers: Dict[str, torch.Tensor] = {}
    outputs: List[torch.Tensor] = []
    for i in range(nlev - 1):  # unet.py:128-131
        x = _block(x, sd, f"down_steps.{i}", spec, training, new_buffers, acts, tf32)
        outputs.append(x)
        x = pool(x, spec["max_pool_kernel"])  # kernel == stride, floor mode, no padding (unet.py:123)
        if acts is not None:
            acts[f"down_steps.{i}.pool"] = x
    x = _block(x, sd, f"down_steps.{nlev - 1}", spec, training, new_buffers, acts, tf32)  # unet.py:133
    for i in range(nlev - 1):  # unet.py:135-136 -> Up.forward unet.py:309-315
        skip = outputs.pop()
        x = _tconv(convT, tf32, x, sd[f"up_steps.{i}.up_conv.weight"], sd[f"up_steps.{i}.up_conv.bias"],
                   stride=spec["upsample_stride"], padding=0)  # unet.py:294-298: no groups / dilation
        if acts is not None:
            acts[f"up_steps.{i}.up_conv"] = x
        y = _crop(x, skip)  # unet.py:311 -- crops the UPSAMPLED tensor; the skip is only a size donor
        x = torch.cat((x, y), dim=1)  # unet.py:312 -- raises if skip is smaller than x anywhere
        x = _block(x, sd, f"up_steps.{i}", spec, training, new_buffers, acts, tf32)
    x = _tconv(conv, tf32, x, sd["out_conv.weight"], sd["out_conv.bias"])  # unet.py:120,138
    return x, new_buffers


# --------------------------------------------------------------------------------------------
# loss.py
# --------------------------------------------------------------------------------------------

def _crop_to(t, shape):
    """`loss.py:51-59`: origin crop of mask / pwl to pred's spatial shape."""
    if len(shape) == 5:
        return t[:, :, 0:shape[2], 0:shape[3], 0:shape[4]]
    if len(shape) == 4:
        return t[:, :, 0:shape[2], 0:shape[3]]
    raise IndexError("Unexpected number of predicted mask dimensions. Expected 4 (2D) or 5 (3D) but got"
                     f" {len(shape)} dimensions: {shape}")


def cross_entropy(pred, mask, pwl, method="pixel", num_random_pixels=None):
    """Restates `loss.py:5-101`.  NB `loss.py:48` sets ``is_pwl_none = True`` unconditionally, so the
    ``pwl[mask > .5] += 2`` branch (`loss.py:61-63`) never runs; ``pwl=None`` means weight 2."""
    methods = ["pixel", "worst_z", "random", "sigmoid"]
    if method not in methods:
        raise ValueError(f"Viable methods for cross entropy loss are {methods}, not {method}.")
    if method == "random":
        if num_random_pixels is None:
            raise ValueError("the number of random pixels to draw is not defined. Please set num_random_pixels to a "
                             "value larger than 1.")
        if num_random_pixels <= 1:
            raise ValueError(f"num_random_pixels should be greater than 1 not {num_random_pixels}.")
        if (mask == 0).sum() == 0:
            raise ValueError("There are no background pixels in mask.\n\t(mask==0).sum() == 0 -> True")
    if method == "sigmoid":
        pred = torch.sigmoid(pred)  # loss.py:38-40
    shape = pred.shape
    if pwl is None:
        pwl = torch.ones(pred.shape).to(pred.device)  # loss.py:46-47
    mask = _crop_to(mask, shape)
    pwl = _crop_to(pwl, shape)

    def bce(p, m):  # nn.BCEWithLogitsLoss(reduction='none'), loss.py:65
        return F.binary_cross_entropy_with_logits(p, m, reduction="none")

    if method in ("pixel", "sigmoid"):  # loss.py:70-72, 97-99
        loss = bce(pred.float(), mask.float()) * (pwl + 1)
    elif method == "worst_z":  # loss.py:74-80
        loss = bce(pred.float(), mask.float()) * (pwl + 1)
        scaling = torch.linspace(1, 2, pred.shape[4]) ** 2
        loss, _ = torch.sort(loss.sum(dim=[0, 1, 2, 3]))
        loss = loss * scaling.to(loss.device)
        loss = loss / (pred.shape[2] * pred.shape[3])
    else:  # 'random', loss.py:82-95 (consumes the global CPU RNG exactly like the reference)
        pred = pred.reshape(-1)
        mask = mask.reshape(-1)
        if (mask == 1).sum() == 0:
            loss = bce(pred.float(), mask.float())
        else:
            pos_ind = torch.randint(low=0, high=int((mask == 1).sum()), size=(1, num_random_pixels))[0, :]
            neg_ind = torch.randint(low=0, high=int((mask == 0).sum()), size=(1, num_random_pixels))[0, :]
            p = torch.cat([pred[mask == 1][pos_ind], pred[mask == 0][neg_ind]]).unsqueeze(0)
            m = torch.cat([mask[mask == 1][pos_ind], mask[mask == 0][neg_ind]]).unsqueeze(0)
            loss = bce(p.float(), m.float())
    return loss.mean()


def dice(pred, mask):
    """Restates `loss.py:104-127`."""
    mask = _crop_to(mask, pred.shape)
    p = torch.sigmoid(pred)
    return 1 - (2 * (p * mask).sum() + 1e-10) / ((p + mask).sum() + 1e-10)


def L1Loss(pred, mask):
    """Restates `loss.py:130-152`."""
    return F.l1_loss(pred, _crop_to(mask, pred.shape))


def MSELoss(pred, mask):
    """Restates `loss.py:155-177`."""
    return F.mse_loss(pred, _crop_to(mask, pred.shape))


# --------------------------------------------------------------------------------------------
# helpers shared by tests / bench
# --------------------------------------------------------------------------------------------

README_3D = dict(image_dimensions=3, in_channels=4, out_channels=1, feature_sizes=[8, 16, 32, 64, 128],
                 kernel={"conv1": (3, 3, 2), "conv2": (3, 3, 1)}, upsample_kernel=(2, 2, 2),
                 max_pool_kernel=(2, 2, 1), upsample_stride=(2, 2, 1), dilation=1, groups=1)
"""README quickstart kwargs (`README.md:17-26`) with the real kwarg name (SURVEY.md section 8d)."""


def golden_inputs(kwargs: dict, xshape, seed: int):
    """Seeded synthetic (image, mask, pwl) triple used by ``oracle/make_golden.py`` and the parity
    tests (CPU generator => identical on every box with the same torch build; the fixtures carry an
    ``input_checksum`` to prove it)."""
    g = torch.Generator().manual_seed(seed + 2000)
    x = torch.randn(tuple(xshape), generator=g)
    mshape = list(xshape)
    mshape[1] = kwargs["out_channels"]
    mask = (torch.rand(mshape, generator=g) > 0.7).float()
    pwl = torch.rand(mshape, generator=g) * 3
    return x, mask, pwl


def train_step_grads(sd: Dict[str, torch.Tensor], spec: dict, x, mask, pwl, method="pixel", tf32=False):
    """One reference-semantics training forward+backward on CPU.  Returns
    (loss, logits, {param: grad}, new_buffers)."""
    leaf = {}
    for k, v in sd.items():
        if v.is_floating_point() and not (k.endswith("running_mean") or k.endswith("running_var")):
            leaf[k] = v.detach().clone().requires_grad_(True)
        else:
            leaf[k] = v.detach().clone()
    logits, new_buffers = unet_forward(leaf, spec, x, training=True, tf32=tf32)
    loss = cross_entropy(logits, mask, pwl, method)
    loss.backward()
    grads = {k: v.grad for k, v in leaf.items() if v.requires_grad}
    return loss.detach(), logits.detach(), grads, new_buffers


def seeded_state_dict(ctor, kwargs: dict, seed: int) -> Dict[str, torch.Tensor]:
    """The state_dict every fixture starts from: ``torch.manual_seed(seed); ctor(**kwargs)`` (the reference's default
    initialisation -- ``ctor`` is the reference class in ``oracle/make_golden.py`` and ``hcunet_b200.Unet_Constructor``,
    whose parameter containers consume the RNG identically, in the GPU tests), then non-trivial BatchNorm affine
    parameters / running statistics from a second seeded generator so that parity exercises them."""
    torch.manual_seed(seed)
    model = ctor(**kwargs)
    g = torch.Generator().manual_seed(seed + 1000)
    with torch.no_grad():
        for k, v in model.state_dict().items():
            if "batch" in k and k.endswith("weight"):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            elif "batch" in k and k.endswith("bias"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.2)
            elif k.endswith("running_mean"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
            elif k.endswith("running_var"):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
    return model, {k: v.detach().clone() for k, v in model.state_dict().items()}


def state_checksum(sd: Dict[str, torch.Tensor]) -> float:
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


SAMPLE = 8192


def sample_tensor(t: torch.Tensor, key: str) -> torch.Tensor:
    """Full tensor when small, else SAMPLE seeded entries (the seed is a hash of the parameter name): the big fixtures
    store a gradient as (norm, sample) instead of 124 MB of floats."""
    flat = t.detach().reshape(-1)
    if flat.numel() <= SAMPLE:
        return flat.clone()
    seed = sum((i + 1) * ord(c) for i, c in enumerate(key)) % (2 ** 31)
    idx = torch.randint(flat.numel(), (SAMPLE,), generator=torch.Generator().manual_seed(seed))
    return flat.cpu()[idx].clone()


def batch_statistics_buffers(sd_before, buffers_after, momentum=0.1):
    """Running statistics := the batch statistics of the step that produced ``buffers_after`` (inverting the momentum
    update).  One train-mode forward leaves running stats 90 % at their random initial values, which kills every ReLU of
    some fixtures in eval mode (constant logits: a vacuous eval check); the batch statistics keep the network alive."""
    out = {}
    for k, v in buffers_after.items():
        if k.endswith("running_mean"):
            out[k] = (v - (1 - momentum) * sd_before[k]) / momentum
        elif k.endswith("running_var"):
            out[k] = ((v - (1 - momentum) * sd_before[k]) / momentum).clamp_min(1e-4)
        else:
            out[k] = v
    return out
