"""Drop-in shim: ``from hcat.unet import Unet_Constructor`` / ``from hcat.loss import cross_entropy`` resolve to
the B200 implementation without the reference's heavy imports (`hcat/__init__.py:1-5` pulls skimage, GPy ...).
Only the hot path and its direct caller are provided: ``hcat.unet``, ``hcat.loss``, and the overlap-tile driver
``hcat.segment.predict_segmentation_mask`` with ``hcat.utils.pad_image_with_reflections`` / ``calculate_indexes``."""
from . import loss, segment, unet, utils  # noqa: F401
