"""Input path: ``hcunet_b200.StackLoader`` against the UNMODIFIED reference transforms.

tests/golden/loader_cases.pt holds seeded raw stacks ([Z, Y, X, C] uint8 / uint16 as skimage.io.imread yields them) and
what `to_float -> reshape -> normalize -> to_tensor` of `/root/reference/hcat/transforms.py` make of them (minted by
oracle/make_golden.py).  Integer in, fp16 out: the bar is bit-exact."""
import os

import pytest
import torch

from conftest import GOLDEN
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu


def cases():
    return torch.load(os.path.join(GOLDEN, "loader_cases.pt"), weights_only=False)["cases"]


def tiny_model(dims, cin):
    import hcunet_b200 as H

    if dims == 3:
        kw = dict(image_dimensions=3, in_channels=cin, out_channels=1, feature_sizes=[8, 16], kernel=(1, 1, 1),
                  upsample_kernel=(2, 2, 1), max_pool_kernel=(2, 2, 1), upsample_stride=(2, 2, 1))
    else:
        kw = dict(image_dimensions=2, in_channels=cin, out_channels=1, feature_sizes=[8, 16], kernel=(1, 1),
                  upsample_kernel=(2, 2), max_pool_kernel=(2, 2), upsample_stride=2)
    m = H.Unet_Constructor(**kw)
    m.precision = "mixed"
    return m.cuda()


@pytest.mark.parametrize("i", range(5))
def test_loader_matches_reference_transforms_bit_for_bit(i):
    import hcunet_b200 as H

    c = cases()[i]
    want = c["image"]                       # [1, C, X, Y, Z] / [1, C, X, Y] fp16
    dims = want.dim() - 2
    model = tiny_model(dims, want.shape[1])
    loader = H.StackLoader(model, c["mean"], c["std"])
    x = loader.image(c["raw"])
    assert x.shape == want.shape and x.dtype == torch.float16 and x.is_cuda
    assert torch.equal(x.cpu(), want), float((x.cpu().float() - want.float()).abs().max())
    # the storage behind the view: channels-last, pitch 8, zero padding
    store = x.as_strided((x.shape[0], want[0, 0].numel(), 8), (want[0, 0].numel() * 8, 8, 1))
    assert float(store[:, :, want.shape[1]:].float().abs().max()) == 0.0
    # labels: the full extent reproduces the reference's to_float -> reshape -> to_tensor; a crop is its origin crop
    full = tuple(want.shape[2:])
    for raw, ref in ((c["mask_raw"], c["mask"]), (c["pwl_raw"], c["pwl"])):
        got = loader.labels(raw, full)
        assert torch.equal(got.cpu(), ref)
        ext = tuple(max(1, s - 3) for s in full)
        crop = loader.labels(raw.pin_memory(), ext)       # read in place from pinned host memory
        sl = (slice(None), slice(None)) + tuple(slice(0, e) for e in ext)
        assert torch.equal(crop.cpu(), ref[sl])
    # a batch of two stacks == the two stacks alone
    both = loader.image(torch.stack([c["raw"], c["raw"].flip(0)]))
    assert torch.equal(both[0].cpu(), want[0]) and torch.equal(both[1].cpu(), loader.image(c["raw"].flip(0))[0].cpu())


def test_loader_rejects_what_to_float_rejects():
    import hcunet_b200 as H

    loader = H.StackLoader(tiny_model(3, 4))
    with pytest.raises(TypeError):
        loader.image(torch.zeros((4, 8, 8, 4), dtype=torch.float32))     # to_float: TypeError for anything but uint8/16
    with pytest.raises(RuntimeError):
        loader.image(torch.zeros((4, 8, 8, 3), dtype=torch.uint8))       # channel count


def test_model_reads_the_loader_layout_without_a_layout_pass():
    """README model: loader output (channels-last view) and the reference-style fp16 NCDHW tensor give the same logits bit
    for bit, the former without the NCDHW -> NDHWC kernel; gradients agree; the fp32 path accepts the view as well."""
    import hcunet_b200 as H
    from hcunet_b200 import _lib

    torch.manual_seed(0)
    m = H.Unet_Constructor(**O.README_3D)
    m.precision = "mixed"
    m = m.cuda().train()
    g = torch.Generator().manual_seed(3)
    raw = torch.randint(0, 256, (2, 9, 190, 192, 4), generator=g, dtype=torch.uint8)       # [B, Z, Y, X, C]
    mask = (torch.rand((2, 9, 190, 192), generator=g) > 0.7).to(torch.uint8) * 255
    pwl = torch.rand((2, 9, 190, 192), generator=g).half()
    loader = H.StackLoader(m)
    x, mk, w = loader(raw.pin_memory(), mask.pin_memory(), pwl.pin_memory())
    assert x.shape == (2, 4, 192, 190, 9) and mk.shape == w.shape
    # reference-style tensors of the same data (to_float -> reshape -> normalize -> to_tensor in torch on the host)
    xr = ((raw.double() / 256 - 0.5) / 0.5).float().half().permute(0, 4, 3, 2, 1).contiguous()
    mr = (mask.double() / 256).float().half().permute(0, 3, 2, 1).unsqueeze(1).contiguous()
    wr = pwl.permute(0, 3, 2, 1).unsqueeze(1).contiguous()
    assert torch.equal(x.cpu(), xr)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    for _ in range(2):   # the first two steps record the engine's step cache (more launches): compare steady-state steps
        H.cross_entropy(m(xr.cuda()), mr.cuda(), wr.cuda(), "pixel").backward()
    outs = []
    for xin, mi, wi in ((x, mk, w), (xr.cuda(), mr.cuda(), wr.cuda())):
        m.load_state_dict(sd)
        m.zero_grad(set_to_none=True)
        n0 = _lib.launch_count()
        logits = m(xin)
        assert mk.shape[2:] == logits.shape[2:], "the loader's label crop is the prediction's extent"
        loss = H.cross_entropy(logits, mi, wi, "pixel")
        loss.backward()
        torch.cuda.synchronize()
        outs.append((logits.detach().clone(), float(loss), {k: p.grad.clone() for k, p in m.named_parameters()},
                     _lib.launch_count() - n0))
    assert torch.equal(outs[0][0], outs[1][0])
    assert outs[0][1] == outs[1][1]
    assert outs[0][3] == outs[1][3] - 1, "the loader layout should save exactly the NCDHW -> NDHWC launch"
    for k, gr in outs[0][2].items():
        if not (k.endswith(".bias") and (".conv" in k or ".up_conv" in k)):
            assert float((gr - outs[1][2][k]).double().norm() / outs[1][2][k].double().norm().clamp_min(1e-30)) <= 1e-4, k
    # fp32 path: the view goes through the same layout pass as the NCDHW tensor.  eval(): BN folded, no cross-CTA
    # reductions -> bit-equal; train(): the FFMA kernel's fp32 statistics atomics are order-dependent (~1e-5 on this
    # 5-level model with a 4x4 bottom level), so the two calls agree to the fp32 path's stated tolerance
    m.precision = "fp32"
    m.load_state_dict(sd)
    with torch.no_grad():
        a, b = m(x), m(xr.cuda())
        assert float((a - b).double().norm() / b.double().norm()) <= 1e-4
        m.eval()
        a, b = m(x), m(xr.cuda())
    assert torch.equal(a, b)
