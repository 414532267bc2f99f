"""Pins the oracle restatement (oracle/unet_oracle.py) to the reference.

(a) against the committed golden vectors minted by executing the unmodified reference
    (tests/golden/*.pt, oracle/make_golden.py), and
(b) when /root/reference is mounted (build container only), against the live reference modules.
"""
import pytest
import torch

from conftest import MODEL_CASES, load_golden
from oracle import unet_oracle as O
from oracle.ref_loader import build_reference_unet, load_reference_loss, reference_available


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("name", MODEL_CASES)
def test_oracle_matches_golden_model(name):
    fx = load_golden(name)
    loss, logits, grads, newbuf = O.train_step_grads(fx["state_dict"], fx["kwargs"], fx["x"], fx["mask"], fx["pwl"])
    # same torch build, same primitive ops in the same order => bit-exact
    assert torch.equal(logits, fx["logits_train"])
    assert torch.equal(loss, fx["loss"])
    for k, g in fx["grads"].items():
        assert torch.equal(grads[k], g), k
    for k, v in fx["buffers_after"].items():
        assert torch.equal(newbuf[k], v), k
    sd = dict(fx["state_dict"])
    sd.update(newbuf)
    with torch.no_grad():
        ev, _ = O.unet_forward(sd, fx["kwargs"], fx["x"], training=False)
    assert torch.equal(ev, fx["logits_eval"])


@pytest.mark.parametrize("name", MODEL_CASES)
def test_mixed_oracle_with_rounding_off_matches_golden(name):
    """oracle/mixed_oracle.py restates the same step with an EXPLICIT backward and fp16 rounding hooks; with the hooks
    off it must reproduce the reference's autograd gradients -- so the only thing it adds is WHERE values are rounded."""
    from oracle import mixed_oracle as M

    fx = load_golden(name)
    loss, logits, grads, newbuf = M.train_step(fx["state_dict"], fx["kwargs"], fx["x"], fx["mask"], fx["pwl"], emulate=False)
    floor = 1.5e-5 if name == "g3d_readme" else 5e-6   # the 5-level fixture is reproducible to 1.1e-5 in fp32 (fp64 floor)
    assert rel_l2(logits, fx["logits_train"]) <= floor
    assert abs(float(loss) - float(fx["loss"])) <= 1e-6 * abs(float(fx["loss"]))
    gmax = max(float(g.abs().max()) for g in fx["grads"].values())
    for k, g in fx["grads"].items():
        if k.endswith(".bias") and (".conv1." in k or ".conv2." in k or ".up_conv." in k):
            assert float((grads[k] - g).abs().max()) <= 1e-5 * gmax, k   # analytically zero
        else:
            assert rel_l2(grads[k], g) <= 4 * floor, (k, rel_l2(grads[k], g))
    for k, v in fx["buffers_after"].items():
        assert rel_l2(newbuf[k].float(), v.float()) <= 5e-6, k
    # and with the hooks ON the emulation differs from fp32 by what fp16 storage costs on this case -- a sanity band
    _, l16, _, _ = M.train_step(fx["state_dict"], fx["kwargs"], fx["x"], fx["mask"], fx["pwl"], emulate=True)
    assert 1e-4 < rel_l2(l16, fx["logits_train"]) < 5e-2


def test_full_size_fixture_is_reproducible_from_its_seed():
    """tests/golden/full_cfg1.pt (minted from the unmodified reference at 1 x 4 x 256 x 256 x 32) regenerates: same
    seeded state_dict / inputs, and the oracle restatement reproduces its logits and sampled gradients bit for bit."""
    import hcunet_b200 as H

    fx = load_golden("full_cfg1")
    _, sd = O.seeded_state_dict(H.Unet_Constructor, fx["kwargs"], fx["seed"])
    assert abs(O.state_checksum(sd) - fx["state_checksum"]) <= 1e-9 * fx["state_checksum"]
    loss, logits, grads, _ = O.train_step_grads(sd, fx["kwargs"], fx["x"], fx["mask"], fx["pwl"])
    assert torch.equal(logits, fx["logits_train"]) and torch.equal(loss, fx["loss"])
    for k, smp in fx["grad_sample"].items():
        assert torch.equal(O.sample_tensor(grads[k], k), smp), k


def test_oracle_matches_golden_losses():
    fx = torch.load(__import__("os").path.join(__import__("conftest").GOLDEN, "loss_cases.pt"), weights_only=False)
    assert len(fx["cases"]) >= 20
    for c in fx["cases"]:
        p = c["pred"].clone().requires_grad_(True)
        if c["fn"] == "cross_entropy":
            if c["method"] == "random":
                torch.manual_seed(5)
                v = O.cross_entropy(p, c["mask"], c["pwl"], c["method"], c["num_random_pixels"])
            else:
                v = O.cross_entropy(p, c["mask"], c["pwl"], c["method"])
        else:
            v = getattr(O, c["fn"])(p, c["mask"])
        v.backward()
        assert torch.equal(v.detach(), c["value"]), (c["fn"], c.get("method"), c.get("variant"))
        assert torch.equal(p.grad, c["grad"]), (c["fn"], c.get("method"), c.get("variant"))


def test_loss_error_behaviour():
    p = torch.zeros(1, 1, 4, 4, 2)
    with pytest.raises(ValueError):
        O.cross_entropy(p, p, p, method="nope")
    with pytest.raises(ValueError):
        O.cross_entropy(p, p, p, method="random")
    with pytest.raises(ValueError):
        O.cross_entropy(p, p, p, method="random", num_random_pixels=1)
    with pytest.raises(IndexError):
        O.cross_entropy(torch.zeros(4, 4, 4), torch.zeros(4, 4, 4), None)
    # pwl=None => weight 2 (loss.py:46-48)
    a = O.cross_entropy(p + 0.3, torch.ones_like(p), None)
    b = O.cross_entropy(p + 0.3, torch.ones_like(p), torch.ones_like(p))
    assert torch.equal(a, b)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted")
def test_oracle_matches_live_reference_and_quirks():
    torch.manual_seed(3)
    kwargs = dict(O.README_3D, feature_sizes=[4, 8, 16])
    ref = build_reference_unet(**kwargs)
    ref.train()
    x = torch.randn(1, 4, 44, 46, 7)
    out_ref = ref(x)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    # state_dict was read AFTER the train forward: rewind buffers is not possible, so compare eval
    ref.eval()
    with torch.no_grad():
        ev_ref = ref(x)
        ev, _ = O.unet_forward(sd, kwargs, x, training=False)
    assert torch.equal(ev, ev_ref)
    assert out_ref.shape == ev.shape
    # the skip connection is dead -- Up(x, skip) ignores skip's values
    up = ref.up_steps[0]
    xx = torch.randn(1, 16, 5, 5, 3)
    up.eval()
    with torch.no_grad():
        a = up(xx, torch.randn(1, 8, 12, 12, 6))
        b = up(xx, torch.zeros(1, 8, 12, 12, 6))
    assert torch.equal(a, b)
    # 128x128 is too small for the 5-level README model
    big = build_reference_unet(**O.README_3D)
    with pytest.raises(RuntimeError):
        big(torch.randn(1, 4, 128, 128, 32))
    sd5 = {k: v.detach() for k, v in big.state_dict().items()}
    with pytest.raises(RuntimeError):
        O.unet_forward(sd5, O.README_3D, torch.randn(1, 4, 128, 128, 32))
    # the reference cannot construct 2D models as shipped
    from oracle.ref_loader import load_reference_unet
    with pytest.raises(RuntimeError):
        load_reference_unet().Unet_Constructor()
    # README spelling raises TypeError
    with pytest.raises(TypeError):
        load_reference_unet().Unet_Constructor(image_dimmensions=3)
    # loss module parity on a fresh random case
    L = load_reference_loss()
    p = torch.randn(2, 1, 5, 6, 3)
    m = (torch.rand(2, 1, 7, 7, 4) > 0.5).float()
    w = torch.rand(2, 1, 7, 7, 4)
    for method in ("pixel", "sigmoid", "worst_z"):
        assert torch.equal(L.cross_entropy(p, m, w, method), O.cross_entropy(p, m, w, method))
    assert torch.equal(L.dice(p, m), O.dice(p, m))


# ---- overlap-tile driver (hcat/segment.py:21-136, hcat/utils.py:33-124) --------------------------------------------------

def _tiler_fixture():
    import numpy as np

    fx = torch.load(__import__("os").path.join(__import__("conftest").GOLDEN, "tiler_prod.pt"), weights_only=False)
    n = 1
    for v in fx["mask_shape"]:
        n *= v
    fx["mask"] = torch.from_numpy(np.unpackbits(fx["mask_bits"].numpy())[:n].reshape(fx["mask_shape"]))
    return fx


def test_tiler_oracle_matches_golden_mask():
    """oracle/tiler_oracle.py (restated tile loop) around oracle/unet_oracle.unet_forward (restated network, eval mode) ==
    the mask the unmodified reference produced, voxel for voxel; probabilities to fp16 storage of the fixture."""
    from oracle import tiler_oracle as T

    fx = _tiler_fixture()

    def forward(t):
        return O.unet_forward(fx["state_dict"], fx["kwargs"], t, training=False)[0]

    mask, skipped = T.predict_segmentation_mask(forward, fx["image"].clone(), cuda_mem=fx["cuda_mem"])
    assert skipped == 0 and mask.dtype == torch.uint8 and list(mask.shape) == fx["mask_shape"]
    assert torch.equal(mask, fx["mask"])
    assert 0.3 < float(mask.float().mean()) < 0.7      # a non-trivial mask (about half ones)
    prob, _ = T.predict_segmentation_mask(forward, fx["image"].clone(), use_probability_map=True, cuda_mem=fx["cuda_mem"])
    assert torch.equal(prob.half(), fx["prob"])


def test_tile_index_arithmetic_matches_the_reference():
    """`calculate_indexes` of the product (pure integers), of the oracle and -- when mounted -- of the reference agree,
    including its quirks (tiles eval + 2 pad - 1 wide, the fallback pair for short extents, [[0, n]] for eval > n)."""
    from hcunet_b200 import segment as S
    from oracle import tiler_oracle as T

    fx = _tiler_fixture()
    assert S.calculate_indexes(128, 128, 140, 396) == fx["x_ind"] and S.calculate_indexes(10, 6, 12, 32) == fx["z_ind"]
    cases = [(128, 128, 140, 396), (128, 300, 330, 586), (10, 6, 8, 28), (10, 15, 18, 38), (128, 300, 2048, 2304),
             (10, 15, 128, 148), (128, 300, 200, 456), (128, 300, 300, 556), (30, 50, 49, 109), (6, 7, 100, 112)]
    ref = None
    from oracle.ref_loader import reference_is_live
    if reference_is_live():
        from oracle.ref_loader import load_reference_tiler
        ref = load_reference_tiler()[0]
    for c in cases:
        assert S.calculate_indexes(*c) == T.calculate_indexes(*c), c
        if ref is not None:
            assert S.calculate_indexes(*c) == ref.calculate_indexes(*c), c
    # tile order: z outermost, then x, then y (segment.py:78-80)
    tl = S.tile_list((140, 150, 12), (128, 128, 10), (128, 128, 6))
    assert len(tl) == 8 and tl[0] == ([0, 383], [0, 383], [0, 25]) and tl[1][1] == [22, 405] and tl[4][2] == [6, 31]


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["x"]).reference_is_live(), reason="reference not mounted")
def test_tiler_oracle_matches_live_reference_padding():
    """utils.py:33-74 on a seeded stack: the oracle's three flips + cats == the reference's numpy flips, bit for bit."""
    from oracle import tiler_oracle as T
    from oracle.ref_loader import load_reference_tiler

    utils, _ = load_reference_tiler()
    g = torch.Generator().manual_seed(3)
    x = torch.randn((2, 3, 9, 11, 8), generator=g)
    # (pad 0 is not compared: the reference's `image[pad-1::-1]` then mirrors the WHOLE extent, utils.py:50; the tiler always
    # pads by (128, 128, 10), segment.py:53-56)
    for pad in ((2, 4, 6), (8, 10, 8), (4, 2, 2)):
        assert torch.equal(T.pad_image_with_reflections(x, pad), utils.pad_image_with_reflections(x, pad))
    with pytest.raises(ValueError):
        T.pad_image_with_reflections(x, (3, 2, 2))
