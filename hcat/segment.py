"""`from hcat.segment import predict_segmentation_mask` -> the B200 implementation (hcunet_b200.segment)."""
from hcunet_b200.segment import predict_segmentation_mask  # noqa: F401
