"""hcunet_b200 -- B200-native (sm_100a) implementation of the HcUnet U-Net hot path.

Public surface (mirrors ``hcat.unet`` / ``hcat.loss`` of wisamreid/HcUnet):

    from hcunet_b200 import Unet_Constructor, cross_entropy, dice, L1Loss, MSELoss

Everything numeric runs in ``libhcunet_b200.so`` (hand-written CUDA, C ABI in ``include/hcunet_b200.h``).
Importing the package is cheap and does not need the library; the first compute call loads it and
raises if it is missing -- there is no CPU / PyTorch fallback.
"""
from .flat import FlatParameters  # noqa: F401
from .optim import FlatAdam  # noqa: F401
from .loader import StackLoader  # noqa: F401
from .loss import L1Loss, MSELoss, cross_entropy, dice  # noqa: F401
from .segment import calculate_indexes, pad_image_with_reflections, predict_segmentation_mask  # noqa: F401
from .unet import Down, Unet_Constructor, Up  # noqa: F401

__all__ = ["Unet_Constructor", "Down", "Up", "cross_entropy", "dice", "L1Loss", "MSELoss", "StackLoader", "FlatParameters", "FlatAdam",
           "predict_segmentation_mask", "pad_image_with_reflections", "calculate_indexes"]
__version__ = "0.1.0"
