"""Loader for the UNMODIFIED reference modules (test infrastructure only).

Only tests/, ``__graft_entry__.smoke()``, ``oracle/make_golden.py`` and bench.py's
``cpu_baseline`` / ``--impl reference`` leg may import this file.  The product package
``hcunet_b200`` never does.

The reference (`/root/reference/hcat/unet.py`, `/root/reference/hcat/loss.py`) cannot be
imported with a plain ``import hcat`` in this image: ``hcat/__init__.py:1-5`` eagerly imports
``segment``/``main`` (skimage, GPy, matplotlib ... absent) and ``hcat/unet.py:6-9`` imports
``hcat.utils.pad_image_with_reflections``.  We register an empty ``hcat`` package and a stub
``hcat.utils`` in ``sys.modules`` and then execute the two reference source files unmodified
from where they lie.  ``/root/reference`` only exists in the build
container, never on the GPU box, so everything here is gated on ``reference_available()``.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
"""``oracle/_ref/hcat/{unet,loss}.py``: byte-for-byte copies of the two reference modules of the hot path, staged by
``stage_reference()`` (called from ``__graft_entry__.build()`` in the build container).  The directory is git-ignored
(no reference source enters the history) but travels with the repo snapshot to the GPU box, where ``/root/reference``
does not exist -- so ``bench.py --impl reference`` can time the UNMODIFIED reference there."""


def _pick_root() -> str:
    env = os.environ.get("HCUNET_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile(os.path.join("/root/reference", "hcat", "unet.py")):
        return "/root/reference"
    return STAGED_ROOT


REFERENCE_ROOT = _pick_root()


def stage_reference() -> bool:
    """Copy `hcat/unet.py` and `hcat/loss.py` from the mounted reference into ``oracle/_ref/hcat/`` (unmodified; a
    sha256 manifest is written next to them).  Returns False when the reference is not mounted."""
    import hashlib
    import shutil

    src = "/root/reference/hcat"
    if not os.path.isfile(os.path.join(src, "unet.py")):
        return False
    dst = os.path.join(STAGED_ROOT, "hcat")
    os.makedirs(dst, exist_ok=True)
    lines = []
    for f in ("unet.py", "loss.py"):
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
        with open(os.path.join(dst, f), "rb") as fh:
            lines.append(f"{hashlib.sha256(fh.read()).hexdigest()}  hcat/{f}")
    with open(os.path.join(STAGED_ROOT, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return True


def reference_is_live() -> bool:
    """True when the modules come from the mounted reference tree (build container), not the staged copies."""
    return REFERENCE_ROOT != STAGED_ROOT and reference_available()

_cache = {}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "hcat", "unet.py"))


def _load(name: str, filename: str):
    key = (name, filename)
    if key in _cache:
        return _cache[key]
    # private names so a real ``hcat`` shim package (repo root) is never shadowed
    pkg_name = "_hcat_reference"
    if pkg_name not in sys.modules:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = []  # mark as package
        sys.modules[pkg_name] = pkg
    saved = {k: sys.modules.get(k) for k in ("hcat", "hcat.utils")}
    stub_pkg = types.ModuleType("hcat")
    stub_pkg.__path__ = []
    stub_utils = types.ModuleType("hcat.utils")

    def pad_image_with_reflections(*a, **k):  # only used by the unfinished ``evaluate``
        raise NotImplementedError("stub: hcat.utils is outside the hot path")

    stub_utils.pad_image_with_reflections = pad_image_with_reflections
    sys.modules["hcat"] = stub_pkg
    sys.modules["hcat.utils"] = stub_utils
    try:
        spec = importlib.util.spec_from_file_location(f"{pkg_name}.{name}", filename)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache[key] = mod
    return mod


def load_reference_unet():
    """Return the reference ``hcat.unet`` module (unmodified source)."""
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    return _load("unet", os.path.join(REFERENCE_ROOT, "hcat", "unet.py"))


def load_reference_loss():
    """Return the reference ``hcat.loss`` module (unmodified source)."""
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    return _load("loss", os.path.join(REFERENCE_ROOT, "hcat", "loss.py"))


def load_reference_transforms():
    """The reference ``hcat/transforms.py`` (unmodified source, build container only -- it is not staged).  Its imports of
    skimage / elasticdeform (absent from this image) are satisfied by empty stub modules: the deterministic transforms
    this project mirrors (``to_float``, ``reshape``, ``normalize``, ``to_tensor``) are numpy / torch only."""
    path = os.path.join("/root/reference", "hcat", "transforms.py")
    if not os.path.isfile(path):
        raise FileNotFoundError("reference transforms not mounted")
    key = ("transforms", path)
    if key in _cache:
        return _cache[key]
    names = ["skimage", "skimage.exposure", "skimage.transform", "skimage.io", "skimage.morphology", "elasticdeform", "cv2"]
    saved = {k: sys.modules.get(k) for k in names}
    try:
        for k in names:
            try:
                __import__(k)
            except Exception:
                m = types.ModuleType(k)
                m.__path__ = []
                sys.modules[k] = m
        for k in names:   # `import skimage.io as io` needs the attribute on the parent stub
            if "." in k:
                parent, child = k.rsplit(".", 1)
                if not hasattr(sys.modules[parent], child):
                    setattr(sys.modules[parent], child, sys.modules[k])
        import numpy as np

        had_float = "float" in np.__dict__
        if not had_float:
            np.float = float   # the reference's era of numpy still had this alias (annotations at transforms.py:190 ...)
        try:
            spec = importlib.util.spec_from_file_location("_hcat_reference.transforms", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        finally:
            if not had_float:
                del np.float
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache[key] = mod
    return mod


def build_reference_unet(**kwargs):
    """Construct the reference Unet_Constructor.

    2D models cannot be constructed by the reference as shipped (`hcat/unet.py:293-303`
    raises).  For ``image_dimensions == 2`` we alias ``torch.nn.modules.conv.ConvTranspose3d``
    to ``ConvTranspose2d`` DURING CONSTRUCTION ONLY so the gate at `unet.py:293` passes
    (SURVEY.md section 8c "2D oracle").  This deviation is disclosed wherever 2D parity is
    reported.
    """
    import torch

    ref = load_reference_unet()
    if kwargs.get("image_dimensions", 2) == 2:
        conv_mod = torch.nn.modules.conv
        saved = conv_mod.ConvTranspose3d
        conv_mod.ConvTranspose3d = torch.nn.ConvTranspose2d
        try:
            return ref.Unet_Constructor(**kwargs)
        finally:
            conv_mod.ConvTranspose3d = saved
    return ref.Unet_Constructor(**kwargs)


_TILER_STUBS = ["skimage", "skimage.exposure", "skimage.filters", "skimage.morphology", "skimage.feature", "skimage.segmentation",
                "skimage.transform", "skimage.io", "GPy", "matplotlib", "matplotlib.pyplot", "torchvision", "torchvision.ops"]


def load_reference_tiler(cuda_mem=None):
    """The reference's overlap-tile driver, UNMODIFIED source (build container only; not staged):
    ``hcat/utils.py`` (``pad_image_with_reflections``, ``calculate_indexes``) and ``hcat/segment.py``
    (``predict_segmentation_mask``).  Their module-level imports of skimage / GPy / matplotlib (absent from this image;
    none of them is touched by the three functions) are satisfied by empty stub modules, ``hcat.haircell`` by a stub class,
    and the package attribute ``hcat.__CUDA_MEM__`` (`hcat/__init__.py`: the GPU's memory in bytes, or None) is ``cuda_mem``:
    it selects the tile size table of `segment.py:48-57`.  Returns (utils module, segment module)."""
    root = "/root/reference/hcat"
    if not os.path.isfile(os.path.join(root, "segment.py")):
        raise FileNotFoundError("reference tiler not mounted")
    key = ("tiler", cuda_mem)
    if key in _cache:
        return _cache[key]
    names = list(_TILER_STUBS) + ["hcat", "hcat.utils", "hcat.haircell"]
    saved = {k: sys.modules.get(k) for k in names}
    try:
        for k in _TILER_STUBS:
            try:
                __import__(k)
            except Exception:
                m = types.ModuleType(k)
                m.__path__ = []
                sys.modules[k] = m
        for k in _TILER_STUBS:
            if "." in k:
                parent, child = k.rsplit(".", 1)
                if not hasattr(sys.modules[parent], child):
                    setattr(sys.modules[parent], child, sys.modules[k])
        pkg = types.ModuleType("hcat")
        pkg.__path__ = []
        pkg.__CUDA_MEM__ = cuda_mem
        sys.modules["hcat"] = pkg
        hc = types.ModuleType("hcat.haircell")
        hc.HairCell = type("HairCell", (), {})
        sys.modules["hcat.haircell"] = hc
        spec = importlib.util.spec_from_file_location("hcat.utils", os.path.join(root, "utils.py"))
        utils = importlib.util.module_from_spec(spec)
        sys.modules["hcat.utils"] = utils
        spec.loader.exec_module(utils)
        pkg.utils = utils
        spec = importlib.util.spec_from_file_location("_hcat_reference.segment", os.path.join(root, "segment.py"))
        seg = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(seg)
        seg.hcat = pkg     # keep the stub package (with this cuda_mem) bound to the module after sys.modules is restored
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache[key] = (utils, seg)
    return utils, seg
