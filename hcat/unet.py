"""`hcat/unet.py` drop-in: the same public names, implemented by hcunet_b200."""
from hcunet_b200.unet import Down, Unet_Constructor, Up  # noqa: F401
