"""CPU-side tests (no GPU): the drop-in surface, the planner's geometry / error behaviour, and that the
C-ABI library loads and exports every symbol ``include/hcunet_b200.h`` declares (no compute calls)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import MODEL_CASES, ROOT, load_golden
from oracle import unet_oracle as O


def test_cabi_library_exports_every_declared_symbol():
    from hcunet_b200 import _lib

    header = open(os.path.join(ROOT, "include", "hcunet_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(hcu_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    assert os.path.exists(_lib.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    loaded = _lib.load()
    assert loaded.hcu_abi_version() == _lib.ABI_VERSION
    m = re.search(r"#define HCU_ABI_VERSION (\d+)", header)
    assert int(m.group(1)) == _lib.ABI_VERSION
    # struct sizes agree with the header's layout (all int32 / int64 fields, natural alignment)
    assert ctypes.sizeof(_lib.HcuConvDesc) == 4 * (3 + 3 + 4 + 3 + 3 + 4 + 3 * 6 + 2 + 4)
    assert ctypes.sizeof(_lib.HcuLossDesc) == 4 * 14


@pytest.mark.parametrize("name", MODEL_CASES)
def test_surface_matches_reference_state_dict_and_plan(name):
    import hcunet_b200 as H
    from hcunet_b200.engine import plan_unet

    fx = load_golden(name)
    torch.manual_seed(fx["seed"])
    m = H.Unet_Constructor(**fx["kwargs"])
    sd = m.state_dict()
    assert list(sd.keys()) == list(fx["state_dict"].keys())
    for k, v in fx["state_dict"].items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k
        # same construction order => same RNG stream => same initial conv weights as the reference
        if "conv" in k:
            assert torch.equal(sd[k], v), k
    m.load_state_dict(fx["state_dict"])
    plan = plan_unet(m.model_specification, fx["xshape"])
    dims = fx["kwargs"]["image_dimensions"]
    assert tuple(plan.out_sz[:dims]) == tuple(fx["logits_train"].shape[2:])
    assert m.model_specification == O.normalise_spec(fx["kwargs"])


def test_constructor_errors_and_spelling():
    import hcunet_b200 as H

    with pytest.raises(ValueError):
        H.Unet_Constructor(image_dimensions=4)
    with pytest.raises(ValueError):
        H.Unet_Constructor(feature_sizes=[8])
    with pytest.raises(AssertionError):
        H.Unet_Constructor(feature_sizes=[8, 12])
    with pytest.raises(TypeError):
        H.Unet_Constructor(image_dimmensions=3)  # README.md:17 spelling raises in the reference too
    m = H.Unet_Constructor(**O.README_3D)
    assert sum(p.numel() for p in m.parameters()) == 727009 and len(m.state_dict()) == 136


def test_planner_raises_like_the_reference():
    from hcunet_b200.engine import plan_unet

    spec = O.normalise_spec(O.README_3D)
    with pytest.raises(RuntimeError):
        plan_unet(spec, (1, 4, 128, 128, 32))
    with pytest.raises(RuntimeError):
        plan_unet(spec, (1, 3, 256, 256, 32))   # channels
    with pytest.raises(RuntimeError):
        plan_unet(spec, (1, 4, 256, 256))       # rank
    p = plan_unet(spec, (1, 4, 256, 256, 32))
    assert p.out_sz == (68, 68, 27)
    p = plan_unet(spec, (1, 4, 256, 256, 64))
    assert p.out_sz == (68, 68, 59)
    # geometry table of SURVEY 8a (cfg4)
    convs = {g.name: g for g in p.steps if hasattr(g, "taps")}
    assert convs["down_steps.0.conv1"].out_sz == (254, 254, 63)
    assert convs["down_steps.4.conv2"].out_sz == (8, 8, 59)
    assert convs["up_steps.3.conv2"].out_sz == (68, 68, 59)
    assert convs["up_steps.0.conv1"].fold and convs["up_steps.0.conv1"].cin_g == 64
    # 2D classic: 572 -> 388 in the textbook; the reference's valid convs + dead skips give 196
    spec2 = O.normalise_spec({})
    assert plan_unet(spec2, (16, 3, 572, 572)).out_sz[:2] == (196, 196)
    # a skip smaller than the upsampled tensor makes torch.cat raise in the reference (unet.py:312)
    spec3 = O.normalise_spec(dict(O.README_3D, feature_sizes=[4, 8], upsample_kernel=(12, 12, 2)))
    with pytest.raises(RuntimeError):
        plan_unet(spec3, (1, 4, 20, 20, 4))


def test_no_cpu_fallback_and_shim():
    import hcat
    import hcunet_b200 as H

    assert hcat.unet.Unet_Constructor is H.Unet_Constructor and hcat.loss.cross_entropy is H.cross_entropy
    m = H.Unet_Constructor(**dict(O.README_3D, feature_sizes=[4, 8]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 20, 20, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.cross_entropy(torch.zeros(1, 1, 4, 4, 2), torch.zeros(1, 1, 4, 4, 2), None)
    # argument validation happens before the device check, like the reference's ordering (loss.py:25-36)
    with pytest.raises(ValueError):
        H.cross_entropy(torch.zeros(1, 1, 4, 4, 2), torch.zeros(1, 1, 4, 4, 2), None, method="nope")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hcunet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|import_module\(.oracle|oracle/", src, flags=re.M), \
                    f"{f} reaches into oracle/"


def test_every_conv_of_the_baseline_configs_gets_a_tensor_core_kernel():
    """Host-only routing check (no GPU): every convolution / data gradient / weight gradient of BASELINE.json configs[1]
    (README 3D model) and configs[2] (classic 2D U-Net, 572x572x3, batch 16) is taken by a tensor-core kernel -- the
    64..1024-channel levels of the 2D model by the K-streamed one -- and no layer falls back to the FFMA kernels."""
    import ctypes as C

    from hcunet_b200 import _lib
    from hcunet_b200.engine import ConvGeom, conv_desc, plan_unet
    import hcunet_b200 as H

    lib = _lib.load()

    def describe(d):
        buf = C.create_string_buffer(256)
        assert lib.hcu_conv_tc_describe(C.byref(d), buf, 256) == 0
        return buf.value.decode()

    readme = dict(image_dimensions=3, in_channels=4, out_channels=1, feature_sizes=[8, 16, 32, 64, 128],
                  kernel={"conv1": (3, 3, 2), "conv2": (3, 3, 1)}, upsample_kernel=(2, 2, 2), max_pool_kernel=(2, 2, 1),
                  upsample_stride=(2, 2, 1), dilation=1, groups=1)
    cases = [(H.Unet_Constructor(**readme).model_specification, (4, 4, 256, 256, 32)),
             (H.Unet_Constructor().model_specification, (16, 3, 572, 572))]
    kinds = set()
    for spec, shape in cases:
        plan = plan_unet(spec, shape)
        for g in plan.steps:
            if not isinstance(g, ConvGeom) or g.name == "out_conv":   # the 1x1 logits conv has its own fp32-output descriptors
                continue
            cpi, cpo = max(8, g.cin_t), max(8, g.cout_t)
            d = conv_desc(_lib.F16, _lib.F16, plan.batch, g.in_sz, cpi, 0, g.cin_g, g.cin_g, g.out_sz, g.out_sz, cpo, 0,
                          g.cout_g, g.groups, g.taps, g.dil)
            pad = tuple((g.taps[i] - 1) * g.dil[i] for i in range(3))
            dd = conv_desc(_lib.F16, _lib.F16, plan.batch, g.out_sz, cpo, 0, g.cout_g, g.cout_g, g.in_sz, g.in_sz, cpi, 0,
                           g.cin_g, g.groups, g.taps, g.dil, pad=pad)
            for what, desc in (("fwd", d), ("dgrad", dd)):
                text = describe(desc)
                assert text.startswith(("classic ", "ks ")), (g.name, what, text)
                kinds.add(text.split()[0])
                if g.cin_t >= 256 and what == "fwd":
                    assert text.startswith("ks "), (g.name, text)     # neither the weights nor an x-plane fit otherwise
            wg = lib.hcu_conv_wgrad_tc5_supported(C.byref(d)) or lib.hcu_conv_wgrad_ws_supported(C.byref(d)) or \
                lib.hcu_conv_wgrad_tc_supported(C.byref(d))
            assert wg, g.name
            # the channel-poor levels of the 3D model go to the TMA-fed row-stacked tcgen05 kernel, the 2D model has no z rows for it
            rows = bool(lib.hcu_conv_wgrad_rows_supported(C.byref(d)))
            assert rows == (spec["image_dimensions"] == 3 and cpi <= 32 and cpo <= 64), (g.name, rows)
    assert kinds == {"classic", "ks"}
    # the kernel hint and the forced tile are honoured
    d = conv_desc(_lib.F16, _lib.F16, 16, (66, 66, 1), 256, 0, 256, 256, (64, 64, 1), (64, 64, 1), 256, 0, 256, 1, (3, 3, 1))
    d.reserved[0], d.reserved[1] = 2, 2 | (256 << 8) | (4 << 20)
    assert describe(d).startswith("ks M=256 Nc=256 nsplit=1 PC=4")
    d.reserved[0], d.reserved[1] = 1, 0
    assert describe(d).startswith(("classic ", "unsupported"))


def test_checkpoints_interchange_with_the_reference_both_ways(tmp_path):
    """`unet.py:145-196`: a file written by the UNMODIFIED reference's save() loads into this implementation (same
    state_dict bit for bit, same model_specification, eval mode, hyperparameters returned) and vice versa."""
    from oracle import ref_loader as R

    if not R.reference_available():
        pytest.skip("reference modules not available (neither /root/reference nor oracle/_ref)")
    import hcunet_b200 as H

    kwargs = dict(O.README_3D, feature_sizes=[4, 8])
    cwd = os.getcwd()
    os.chdir(tmp_path)   # save() snapshots ./**/*.py (unet.py:150-160): keep that small
    try:
        torch.manual_seed(3)
        ref = R.build_reference_unet(**kwargs)
        assert ref.save("ref.unet", hyperparameters={"lr": 3e-4}) is None
        mine = H.Unet_Constructor(image_dimensions=3, in_channels=1, out_channels=1, feature_sizes=[2, 4], kernel=(3, 3, 1),
                                  upsample_kernel=(2, 2, 1), max_pool_kernel=(2, 2, 1), upsample_stride=(2, 2, 1))
        assert mine.load("ref.unet", to_cuda=False) == {"lr": 3e-4}
        assert not mine.training and mine.model_specification == ref.model_specification
        for k, v in ref.state_dict().items():
            assert torch.equal(mine.state_dict()[k], v), k
        # and back: this implementation's file into the reference
        torch.manual_seed(4)
        src = H.Unet_Constructor(**kwargs)
        src.save("mine.unet", hyperparameters=[1, 2])
        back = R.build_reference_unet(**dict(kwargs, feature_sizes=[2, 4]))
        assert back.load("mine.unet", to_cuda=False) == [1, 2]
        for k, v in src.state_dict().items():
            assert torch.equal(back.state_dict()[k], v), k
        blob = torch.load("mine.unet", weights_only=False)
        assert set(blob) == {"state_dict", "model_specifications", "hyperparameters", "python_files", "tree_structure"}
    finally:
        os.chdir(cwd)
