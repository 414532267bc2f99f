import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    import torch

    from oracle.unet_oracle import golden_inputs

    fx = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    if "x" not in fx:
        x, mask, pwl = golden_inputs(fx["kwargs"], fx["xshape"], fx["seed"])
        chk = float(x.double().sum() + mask.double().sum() + pwl.double().sum())
        assert abs(chk - fx["input_checksum"]) < 1e-6 * max(1.0, abs(chk)), "seeded inputs differ from the fixture's"
        fx.update(x=x, mask=mask, pwl=pwl)
    return fx


MODEL_CASES = ["g3d_small", "g2d_small", "g3d_prod", "g3d_dil", "g3d_readme"]
