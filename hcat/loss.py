"""`hcat/loss.py` drop-in: the same public names, implemented by hcunet_b200."""
from hcunet_b200.loss import L1Loss, MSELoss, cross_entropy, dice  # noqa: F401
