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


FULL_CASES = {
    # BASELINE.json configs at their full sizes.  name: (ctor kwargs, input shape, seed)
    "full_cfg1": (README_3D, (1, 4, 256, 256, 32), 0),          # configs[0] with the 256 x 256 correction
    "full_cfg2": (README_3D, (4, 4, 256, 256, 32), 0),          # configs[1]: the benchmarked train step, batch 4
    "full_cfg3": (dict(image_dimensions=2, in_channels=3, out_channels=2, feature_sizes=[32, 64, 128, 256, 512, 1024],
                       kernel=(3, 3), upsample_kernel=(2, 2), max_pool_kernel=(2, 2), upsample_stride=2, dilation=1,
                       groups=1), (2, 3, 572, 572), 0),          # configs[2] at batch 2 (batch 16 needs ~100 GB on the CPU)
}


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def mint_full_case(name, kwargs, xshape, seed):
    """Full-size fixture: the UNMODIFIED reference's logits / loss / gradients (sampled where large) + its own fp64
    reproducibility floor, and the fp16-storage emulation's (oracle/mixed_oracle.py) with its accumulation-order floor --
    computed here so that the GPU box only loads them."""
    import time

    from oracle import mixed_oracle as M
    from oracle import unet_oracle as O

    t0 = time.time()
    loss_mod = load_reference_loss()
    model, sd0 = O.seeded_state_dict(build_reference_unet, kwargs, seed)
    x, mask, pwl = golden_inputs(kwargs, xshape, seed)
    model.train()
    logits = model(x)
    loss = loss_mod.cross_entropy(logits, mask, pwl, "pixel")
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    buf1 = {k: v.detach().clone() for k, v in model.state_dict().items() if "running_" in k or "num_batches" in k}
    print(f"{name}: reference step {time.time() - t0:.1f}s", flush=True)
    # eval-mode logits with running statistics := this batch's statistics (keeps the network alive)
    bstats = O.batch_statistics_buffers(sd0, buf1)
    model.load_state_dict({**sd0, **bstats})
    model.eval()
    with torch.no_grad():
        logits_eval = model(x)
    # fp32 reproducibility floor: the reference in float64
    model64, _ = O.seeded_state_dict(build_reference_unet, kwargs, seed)
    model64 = model64.double().train()
    l64 = model64(x.double())
    loss_mod.cross_entropy(l64, mask.double(), pwl.double(), "pixel").backward()
    floor32 = {k: _rel(grads[k], p.grad) for k, p in model64.named_parameters()}
    floor32["logits"] = _rel(logits.detach(), l64.detach())
    del model64, l64
    print(f"{name}: fp64 floor {time.time() - t0:.1f}s (logits {floor32['logits']:.1e})", flush=True)
    # fp16-storage emulation + its accumulation-order floor
    (eloss, elogits, egrads, ebuf), floor16 = M.accumulation_floor(sd0, kwargs, x, mask, pwl)
    elogits_eval, eval_spread, eval_agree = M.eval_accumulation_floor({**sd0, **bstats}, kwargs, x)
    print(f"{name}: emulation {time.time() - t0:.1f}s (logits floor {floor16['logits']:.1e})", flush=True)
    fx = dict(kwargs=kwargs, seed=seed, xshape=tuple(xshape), torch_version=torch.__version__,
              state_checksum=O.state_checksum(sd0),
              input_checksum=float(x.double().sum() + mask.double().sum() + pwl.double().sum()),
              logits_train=logits.detach(), loss=loss.detach(), logits_eval=logits_eval, eval_buffers=bstats,
              buffers_after=buf1,
              grad_norm={k: float(g.double().norm()) for k, g in grads.items()},
              grad_sample={k: O.sample_tensor(g, k) for k, g in grads.items()},
              floor32=floor32,
              emu=dict(logits_train=elogits, loss=eloss, logits_eval=elogits_eval, eval_floor=eval_spread,
                       eval_agree_floor=eval_agree, buffers_after={k: v for k, v in ebuf.items()},
                       grad_norm={k: float(g.double().norm()) for k, g in egrads.items()},
                       grad_sample={k: O.sample_tensor(g, k) for k, g in egrads.items()}, floor=floor16,
                       vs_fp32=dict(logits=_rel(elogits, logits.detach()),
                                    agree_eval=float(((elogits_eval > 0) == (logits_eval > 0)).float().mean()))))
    if name == "full_cfg2":  # same patch as full_cfg1, whose eval logits are stored: keep this fixture under 5 MB
        for d in (fx, fx["emu"]):
            d.pop("logits_eval")
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(f"{name}: logits {tuple(logits.shape)} loss {loss.item():.6f} emu-vs-fp32 logits {fx['emu']['vs_fp32']['logits']:.2e} "
          f"eval agreement {fx['emu']['vs_fp32']['agree_eval']:.5f} bytes {os.path.getsize(os.path.join(OUT, name + '.pt'))}",
          flush=True)


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


def mint_loader_cases():
    """Input path: seeded raw stacks through the UNMODIFIED reference transforms (`transforms.py`: to_float -> reshape ->
    normalize -> to_tensor; mask / pwl: to_float -> reshape -> to_tensor), as `Stack.__getitem__` composes them
    (`dataloader.py:68-92`, `tests/transforms_test.py:22-52`)."""
    import numpy as np

    from oracle.ref_loader import load_reference_transforms

    t = load_reference_transforms()
    rng = np.random.default_rng(123)
    cases = []
    specs = [  # (shape [Z, Y, X, C] or [Y, X, C], dtype, mean, std)
        ((6, 10, 12, 4), np.uint8, None, None),
        ((35, 3, 70, 4), np.uint8, None, None),                      # crosses the 32 x 32 transpose tiles, ragged
        ((5, 9, 33, 4), np.uint16, [0.4, 0.5, 0.45, 0.55], [0.2, 0.25, 0.3, 0.22]),
        ((4, 6, 34, 2), np.uint16, [0.1, 0.9], [0.7, 0.3]),          # channel count != 4: scalar loads
        ((40, 37, 3), np.uint8, [0.5, 0.4, 0.3], [0.5, 0.25, 0.125]),  # 2D
    ]
    for shape, dt, mean, std in specs:
        hi = 256 if dt == np.uint8 else 65536
        img = rng.integers(0, hi, size=shape, dtype=dt)
        mask = (rng.random(shape[:-1]) > 0.5).astype(np.uint8) * 255
        pwl = rng.integers(0, 65536, size=shape[:-1], dtype=np.uint16)
        image, m, w = t.to_float()([img.copy(), np.expand_dims(mask, mask.ndim), np.expand_dims(pwl, pwl.ndim)])
        image, m, w = t.reshape()([image, m, w])
        image = t.normalize(mean, std)(image)
        image, m, w = t.to_tensor()([image, m, w])
        cases.append(dict(raw=torch.from_numpy(img.view(np.int16) if dt == np.uint16 else img), raw_dtype=str(np.dtype(dt)),
                          mask_raw=torch.from_numpy(mask), pwl_raw=torch.from_numpy(pwl.view(np.int16)), mean=mean, std=std,
                          image=image.contiguous(), mask=m.contiguous(), pwl=w.contiguous()))
    torch.save({"torch_version": torch.__version__, "numpy_version": np.__version__, "cases": cases},
               os.path.join(OUT, "loader_cases.pt"))
    print("loader_cases:", len(cases), os.path.getsize(os.path.join(OUT, "loader_cases.pt")), "bytes")


TILER_KW = dict(image_dimensions=3, in_channels=4, out_channels=1, feature_sizes=[4, 8, 16, 32], kernel=K3,
                upsample_kernel=(8, 8, 2), max_pool_kernel=(2, 2, 1), upsample_stride=(2, 2, 1), dilation=1, groups=2)
TILER_CUDA_MEM = 4.2e9     # hcat.__CUDA_MEM__ of a 4 GB GPU: EVAL [128, 128, 6], PAD (128, 128, 10) (segment.py:48-57)


def mint_tiler_case():
    """`predict_segmentation_mask` (hcat/segment.py:21-136) of the UNMODIFIED reference on CPU: the deployed architecture of
    `main.py:46-55` (`groups=2`, `ConvTranspose3d k=(8,8,2)`) at a quarter of its width, a seeded [1, 4, 140, 150, 12] stack with a
    NaN, a +inf and a -inf voxel, the 4 GB row of the tile table -> 2 x 2 x 2 overlapping tiles of 383 x 383 x 25.  BatchNorm
    running statistics come from 25 train-mode forwards (random affine parameters); the output bias is then centred on the median logit of the first
    tile, so that the mask is about half ones with an intricate boundary (a random-init model otherwise says 1 everywhere).
    The reference is called with a NUMPY image: with torch 2.11 its NaN scrub (`image[np.isnan(image)] = 0`, segment.py:66)
    raises on a torch tensor (np.isnan returns a uint8 tensor), as a numpy array it runs unmodified."""
    from oracle.ref_loader import load_reference_tiler

    torch.manual_seed(5)
    model = build_reference_unet(**TILER_KW)
    g = torch.Generator().manual_seed(6)
    img = torch.randn((1, 4, 140, 150, 12), generator=g)
    img[0, 1, 3, 4, 5] = float("nan")
    img[0, 2, 100, 7, 1] = float("inf")
    img[0, 0, 50, 60, 7] = float("-inf")
    utils, seg = load_reference_tiler(cuda_mem=TILER_CUDA_MEM)
    with torch.no_grad():
        for k, v in model.state_dict().items():   # non-trivial BatchNorm affine parameters, like the other fixtures
            if "batch" in k and k.endswith("weight"):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            elif "batch" in k and k.endswith("bias"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.3 + 0.2)
        model.train()
        for _ in range(25):                       # running statistics close to the batch statistics (momentum 0.1)
            model(torch.randn((1, 4, 140, 140, 12), generator=g))
        model.eval()
        clean = img.clone()
        clean[torch.isnan(clean)] = 0
        clean[torch.isinf(clean)] = 1
        tile = utils.pad_image_with_reflections(clean, pad_size=(128, 128, 10))[:, :, 0:383, 0:383, 0:25]
        model.out_conv.bias -= model(tile)[:, :, 128:256, 128:256, 10:16].median()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    mask = seg.predict_segmentation_mask(model, img.clone().numpy(), "cpu")
    prob = seg.predict_segmentation_mask(model, img.clone().numpy(), "cpu", use_probability_map=True)
    print()
    assert mask.dtype == torch.uint8 and prob.dtype == torch.float32
    frac = float(mask.float().mean())
    assert 0.2 < frac < 0.8, frac
    import numpy as np
    torch.save({"torch_version": torch.__version__, "kwargs": TILER_KW, "cuda_mem": TILER_CUDA_MEM, "state_dict": sd, "image": img,
                "mask_bits": torch.from_numpy(np.packbits(mask.numpy().reshape(-1))), "mask_shape": list(mask.shape),
                "prob": prob.half(), "ones_fraction": frac,
                "x_ind": utils.calculate_indexes(128, 128, 140, 396), "z_ind": utils.calculate_indexes(10, 6, 12, 32)},
               os.path.join(OUT, "tiler_prod.pt"))
    print("tiler_prod: ones", frac, os.path.getsize(os.path.join(OUT, "tiler_prod.pt")), "bytes")


def main():
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["small", "full"]
    if "small" in which:
        for name, (kwargs, xshape, seed) in CASES.items():
            mint_model_case(name, kwargs, xshape, seed)
        mint_loss_cases()
    if "small" in which or "loader" in which:
        mint_loader_cases()
    if "small" in which or "tiler" in which:
        mint_tiler_case()
    for name, (kwargs, xshape, seed) in FULL_CASES.items():
        if "full" in which or name in which:
            mint_full_case(name, kwargs, xshape, seed)


if __name__ == "__main__":
    main()
