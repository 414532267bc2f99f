"""Mint golden vectors by executing the UNMODIFIED reference modules (TEST INFRASTRUCTURE).

Run in the build container only (needs ``/root/reference``):

    python oracle/make_golden.py            # rewrites tests/golden/*.pt

The reference ships no golden vectors / known-answer tests for this path, so
these fixtures -- reference ``hcat/unet.py`` + ``hcat/loss.py`` run on CPU fp32 with torch 2.11.0
under fixed seeds -- are what pins both the oracle restatement (``oracle/unet_oracle.py``) and the
CUDA path.  Each fixture stores: constructor kwargs, the reference-initialised ``state_dict``, the
inputs, train-mode logits, the pixel-weighted loss, every parameter gradient, the BN buffers after
that one train-mode forward, and the eval-mode logits computed with those updated buffers.
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import build_reference_unet, load_reference_loss  # noqa: E402
from oracle.unet_oracle import README_3D, golden_inputs  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

K3 = {"conv1": (3, 3, 2), "conv2": (3, 3, 1)}

CASES = {
    # name: (ctor kwargs, input shape, seed)
    "g3d_small": (dict(image_dimensions=3, in_channels=4, out_channels=1, feature_sizes=[4, 8, 16], kernel=K3,
                       upsample_kernel=(2, 2, 2), max_pool_kernel=(2, 2, 1), upsample_stride=(2, 2, 1),
                       dilation=1, groups=1), (2, 4, 46, 44, 6), 11),
    "g3d_readme": (README_3D, (1, 4, 192, 192, 8), 0),
    "g2d_small": (dict(image_dimensions=2, in_channels=3, out_channels=2, feature_sizes=[4, 8, 16], kernel=(3, 3),
                       upsample_kernel=(2, 2), max_pool_kernel=(2, 2), upsample_stride=2, dilation=1, groups=1),
                  (2, 3, 46, 44), 12),
    "g3d_prod": (dict(image_dimensions=3, in_channels=4, out_channels=1, feature_sizes=[4, 8, 16], kernel=K3,
                      upsample_kernel=(8, 8, 2), max_pool_kernel=(2, 2, 1), upsample_stride=(2, 2, 1),
                      dilation=1, groups=2), (1, 4, 60, 60, 6), 13),
    "g3d_dil": (dict(image_dimensions=3, in_channels=2, out_channels=3, feature_sizes=[4, 8], kernel=K3,
                     upsample_kernel=(2, 2, 2), max_pool_kernel=(2, 2, 1), upsample_stride=(2, 2, 1),
                     dilation={"conv1": (2, 2, 1), "conv2": 1}, groups=1), (2, 2, 30, 30, 5), 14),
}


def mint_model_case(name, kwargs, xshape, seed):
    loss_mod = load_reference_loss()
    torch.manual_seed(seed)
    model = build_reference_unet(**kwargs)
    # make BN affine params / buffers non-trivial so parity exercises them
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
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x, mask, pwl = golden_inputs(kwargs, xshape, seed)
    model.train()
    logits = model(x)
    loss = loss_mod.cross_entropy(logits, mask, pwl, "pixel")
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    sd1 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.eval()
    with torch.no_grad():
        logits_eval = model(x)
    # only what the forward changed (BN buffers) is kept from the post-step state_dict
    sd1 = {k: v for k, v in sd1.items() if "running_" in k or "num_batches" in k}
    fx = dict(kwargs=kwargs, seed=seed, xshape=tuple(xshape), state_dict=sd0, logits_train=logits.detach(),
              loss=loss.detach(), grads=grads, buffers_after=sd1, logits_eval=logits_eval,
              torch_version=torch.__version__,
              input_checksum=float(x.double().sum() + mask.double().sum() + pwl.double().sum()))
    if x.numel() < 200_000:  # big inputs are regenerated from the seed (oracle.unet_oracle.golden_inputs)
        fx.update(x=x, mask=mask, pwl=pwl)
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(f"{name}: logits {tuple(logits.shape)} loss {loss.item():.6f} "
          f"bytes {os.path.getsize(os.path.join(OUT, name + '.pt'))}")


def mint_loss_cases():
    loss_mod = load_reference_loss()
    g = torch.Generator().manual_seed(77)
    out = {"torch_version": torch.__version__, "cases": []}
    for shape, big in (((2, 1, 9, 7, 5), (2, 1, 12, 9, 8)), ((3, 2, 10, 6), (3, 2, 13, 9))):
        pred = torch.randn(shape, generator=g) * 3
        mask = (torch.rand(big, generator=g) > 0.6).float()
        pwl = torch.rand(big, generator=g) * 3
        for method in ("pixel", "sigmoid", "worst_z"):
            if method == "worst_z" and len(shape) != 5:
                continue
            for variant in ("fp32", "none", "fp16"):
                p = pred.clone().requires_grad_(True)
                m, w = mask, pwl
                if variant == "none":
                    w = None
                if variant == "fp16":
                    m, w = mask.half(), pwl.half()
                val = loss_mod.cross_entropy(p, m, w, method)
                val.backward()
                out["cases"].append(dict(fn="cross_entropy", method=method, variant=variant, pred=pred, mask=m,
                                         pwl=w, value=val.detach(), grad=p.grad.detach().clone()))
        for fn in ("dice", "L1Loss", "MSELoss"):
            p = pred.clone().requires_grad_(True)
            val = getattr(loss_mod, fn)(p, mask)
            val.backward()
            out["cases"].append(dict(fn=fn, pred=pred, mask=mask, value=val.detach(), grad=p.grad.detach().clone()))
    # 'random' consumes the global CPU RNG (loss.py:88-89)
    pred = torch.randn((1, 1, 8, 8, 4), generator=g)
    mask = (torch.rand((1, 1, 8, 8, 4), generator=g) > 0.5).float()
    torch.manual_seed(5)
    p = pred.clone().requires_grad_(True)
    val = loss_mod.cross_entropy(p, mask, None, "random", num_random_pixels=16)
    val.backward()
    out["cases"].append(dict(fn="cross_entropy", method="random", variant="seed5", pred=pred, mask=mask, pwl=None,
                             num_random_pixels=16, value=val.detach(), grad=p.grad.detach().clone()))
    torch.save(out, os.path.join(OUT, "loss_cases.pt"))
    print("loss_cases:", len(out["cases"]))


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (kwargs, xshape, seed) in CASES.items():
        mint_model_case(name, kwargs, xshape, seed)
    mint_loss_cases()


if __name__ == "__main__":
    main()
