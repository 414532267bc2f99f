"""Overlap-tile inference driver on the GPU (`hcunet_b200.segment`, `csrc/tiler.cu`) against the mask the UNMODIFIED reference
`predict_segmentation_mask` (hcat/segment.py:21-136) produced on CPU (tests/golden/tiler_prod.pt, oracle/make_golden.py):
the deployed architecture (`groups=2`, `ConvTranspose3d k=(8,8,2)`, main.py:46-55) at a quarter width, a seeded stack with
NaN / +-inf voxels, 2 x 2 x 2 overlapping tiles.  Bar (BASELINE.json north_star): >= 99.9 % thresholded-mask voxel agreement."""
import ctypes as C

import pytest
import torch

from test_oracle_golden import _tiler_fixture

pytestmark = pytest.mark.gpu


def _model(fx, precision):
    import hcunet_b200 as H

    m = H.Unet_Constructor(**fx["kwargs"])
    m.load_state_dict(fx["state_dict"])
    m.precision = precision
    return m.cuda().eval()


@pytest.mark.parametrize("precision", ["fp32", "mixed"])
def test_predict_segmentation_mask_matches_the_reference(precision):
    from hcunet_b200 import segment as S

    fx = _tiler_fixture()
    m = _model(fx, precision)
    mask = S.predict_segmentation_mask(m, fx["image"].clone(), "cuda", cuda_mem=fx["cuda_mem"])
    assert mask.dtype == torch.uint8 and mask.device.type == "cpu" and list(mask.shape) == fx["mask_shape"]
    agree = float((mask == fx["mask"]).float().mean())
    prob = S.predict_segmentation_mask(m, fx["image"].clone(), "cuda", use_probability_map=True, cuda_mem=fx["cuda_mem"])
    assert prob.dtype == torch.float32
    dp = float((prob - fx["prob"].float()).abs().max())
    print(f"tiler/{precision}: mask agreement {agree:.6f} (ones {float(mask.float().mean()):.3f}), max |dp| {dp:.2e}")
    assert agree >= 0.999, agree
    assert dp <= (2e-3 if precision == "fp32" else 2e-2)      # the fixture stores fp16 probabilities (5e-4)


def test_ranks_reproduce_the_single_process_mask():
    """Tiles sharded over 3 emulated ranks (contiguous ranges of the reference's tile order, no communication): merging the
    ranks' blocks in rank order == the single-process mask, bit for bit (later tiles overwrite earlier ones)."""
    from hcunet_b200 import segment as S

    fx = _tiler_fixture()
    m = _model(fx, "mixed")
    img = fx["image"].pin_memory()
    whole = S.predict_segmentation_mask(m, img, "cuda", cuda_mem=fx["cuda_mem"])
    parts = [S.predict_segmentation_mask(m, img, "cuda", cuda_mem=fx["cuda_mem"], world=3, rank=r, return_written=True)
             for r in range(3)]
    assert sum(int(w.sum()) for _, w in parts) >= whole.numel()
    assert torch.equal(S.merge_rank_masks(parts), whole)


def test_reflection_padding_scrub_and_skip_kernels():
    """`hcu_tile_gather` == scrub (NaN -> 0, +-inf -> 1) + the reference's reflection padding + slice, in both layouts;
    `pad_image_with_reflections` == the oracle's, no scrub; `hcu_tile_flags` == 0 exactly for an all -1 tile."""
    from hcunet_b200 import _lib, segment as S
    from oracle import tiler_oracle as T

    lib = _lib.load()
    g = torch.Generator().manual_seed(9)
    x = torch.randn((1, 3, 20, 14, 10), generator=g)
    x[0, 0, 0, 0, 0] = float("nan"); x[0, 1, 19, 13, 9] = float("inf"); x[0, 2, 5, 6, 7] = float("-inf")
    pad = (8, 6, 4)
    want_raw = T.pad_image_with_reflections(x.clone(), pad)
    got_raw = S.pad_image_with_reflections(x.cuda(), pad).cpu()
    assert torch.equal(torch.nan_to_num(got_raw, 7.0, 8.0, 9.0), torch.nan_to_num(want_raw, 7.0, 8.0, 9.0))
    clean = x.clone()
    clean[torch.isnan(clean)] = 0
    clean[torch.isinf(clean)] = 1
    want = T.pad_image_with_reflections(clean, pad)
    xs = x.cuda()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for org, ext in (((0, 0, 0), (36, 26, 18)), ((3, 5, 1), (17, 9, 12)), ((20, 12, 8), (16, 14, 10))):
        geom = S._geom(3, (20, 14, 10), pad, org, ext)
        ref = want[0, :, org[0]:org[0] + ext[0], org[1]:org[1] + ext[1], org[2]:org[2] + ext[2]]
        nc = torch.empty((3,) + ext, dtype=torch.float32, device="cuda")
        _lib.check(lib.hcu_tile_gather(C.byref(geom), C.c_void_p(xs.data_ptr()), _lib.F32, C.c_void_p(nc.data_ptr()), _lib.F32, 0, 3, st))
        assert torch.equal(nc.cpu(), ref)
        cl = torch.empty(ext + (8,), dtype=torch.float16, device="cuda")
        _lib.check(lib.hcu_tile_gather(C.byref(geom), C.c_void_p(xs.data_ptr()), _lib.F32, C.c_void_p(cl.data_ptr()), _lib.F16, 1, 8, st))
        assert torch.equal(cl.cpu()[..., :3].permute(3, 0, 1, 2), ref.half()) and float(cl[..., 3:].abs().max()) == 0.0
    # resident sub-stack: only the original voxels the tile reads
    geom = S._geom(3, (20, 14, 10), pad, (3, 5, 1), (17, 9, 12), sorg=(0, 0, 0), ssize=(12, 8, 9))
    sub = xs[0, :, 0:12, 0:8, 0:9].contiguous()
    nc = torch.empty((3, 17, 9, 12), dtype=torch.float32, device="cuda")
    _lib.check(lib.hcu_tile_gather(C.byref(geom), C.c_void_p(sub.data_ptr()), _lib.F32, C.c_void_p(nc.data_ptr()), _lib.F32, 0, 3, st))
    assert torch.equal(nc.cpu(), want[0, :, 3:20, 5:14, 1:13])
    assert S._orig_range(3, 20, 8, 20) == (0, 12) and S._orig_range(5, 14, 6, 14) == (0, 8) and S._orig_range(1, 13, 4, 10) == (0, 9)
    # skip test
    y = torch.full((1, 2, 12, 12, 6), -1.0)
    y[0, 1, 11, 11, 5] = 0.5
    ys = y.cuda()
    flags = torch.zeros(2, dtype=torch.int32, device="cuda")
    g0 = S._geom(2, (12, 12, 6), (2, 2, 2), (0, 0, 0), (8, 8, 6))      # never reaches voxel (11, 11, 5) or its mirrors
    g1 = S._geom(2, (12, 12, 6), (2, 2, 2), (8, 8, 4), (8, 8, 6))
    _lib.check(lib.hcu_tile_flags(C.byref(g0), C.c_void_p(ys.data_ptr()), _lib.F32, C.c_void_p(flags.data_ptr()), st))
    _lib.check(lib.hcu_tile_flags(C.byref(g1), C.c_void_p(ys.data_ptr()), _lib.F32, C.c_void_p(flags.data_ptr() + 4), st))
    f = flags.cpu().tolist()
    assert f[0] == 0 and f[1] > 0


def test_all_minus_one_tiles_are_skipped_and_errors_match():
    """A stack of -1 everywhere: every tile is skipped, the mask stays the float32 zeros the reference starts from
    (segment.py:59); a model whose output is smaller than pad + eval raises the reference's RuntimeError."""
    import hcunet_b200 as H
    from hcunet_b200 import segment as S

    fx = _tiler_fixture()
    m = _model(fx, "mixed")
    img = torch.full((1, 4, 140, 150, 12), -1.0)
    n0 = H._lib.launch_count()
    mask = S.predict_segmentation_mask(m, img, "cuda", cuda_mem=fx["cuda_mem"])
    assert mask.dtype == torch.float32 and float(mask.abs().max()) == 0.0
    assert H._lib.launch_count() - n0 == 8, "8 skip tests, no network launch"
    with pytest.raises(RuntimeError, match="Amount of padding is not sufficient"):
        S.predict_segmentation_mask(m, fx["image"].clone(), "cuda", eval_image_size=[128, 128, 6], pad_size=[16, 16, 2])
    with pytest.raises(KeyError):
        S.predict_segmentation_mask(m, fx["image"].clone(), "cuda", cuda_mem=180e9)     # segment.py:54: no such row
    m.train()
    with pytest.raises(RuntimeError):
        S.predict_segmentation_mask(m, fx["image"].clone(), "cuda", cuda_mem=fx["cuda_mem"])


def test_evaluate_is_as_unfinished_as_the_reference():
    """`unet.py:198-233`: checks, eval(), padding, forward over the slices, result dropped -> None."""
    import hcunet_b200 as H
    from oracle import unet_oracle as O

    m = H.Unet_Constructor(**dict(O.README_3D, feature_sizes=[4, 8])).cuda().train()
    with pytest.raises(ValueError):
        m.evaluate([1, 2, 3])
    with pytest.raises(ImportError):
        m.evaluate(torch.zeros((1, 3, 120, 120, 10)))
    n0 = H._lib.launch_count()
    assert m.evaluate(torch.randn((1, 4, 120, 104, 10))) is None
    assert not m.training and H._lib.launch_count() > n0
