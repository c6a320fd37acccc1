"""Host-side image pre-processing entry point of the reference (model/accessory/data/transform.py)."""
from .transform import CLIP_MEAN, CLIP_STD, get_transform  # noqa: F401
