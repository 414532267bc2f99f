"""`from hcat.utils import pad_image_with_reflections, calculate_indexes` -> the B200 implementation (hcunet_b200.segment)."""
from hcunet_b200.segment import calculate_indexes, pad_image_with_reflections  # noqa: F401
