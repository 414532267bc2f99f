"""Drop-in shim: ``from hcat.unet import Unet_Constructor`` / ``from hcat.loss import cross_entropy`` resolve to
the B200 implementation without the reference's heavy imports (`hcat/__init__.py:1-5` pulls skimage, GPy ...).
Only the hot path is provided: ``hcat.unet`` (also importable as ``hcat.unet`` via `hcat/__init__.py:2`) and
``hcat.loss``."""
from . import loss, unet  # noqa: F401
