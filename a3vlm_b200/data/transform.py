"""`get_transform(transform_type, size)` with the reference's three pipelines (model/accessory/data/
transform.py:37-79): PIL image -> float tensor (3, size, size), CLIP-normalised -- what
`MetaModel.forward / generate` expect as `images`.  CPU work on PIL / torchvision, not on the accelerated
path; inside the reference tree its own module is used unchanged.

  padded_resize        pad to a square with the CLIP mean colour (centred), bicubic resize      (eval + fine-tune default)
  resized_center_crop  bicubic resize of the short side, centre crop
  random_resized_crop  RandomResizedCrop(scale 0.9-1.0, ratio 3/4-4/3), bicubic
"""
from __future__ import annotations

from typing import Sequence

from PIL import Image

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class PadToSquare:
    """Centre the image on a square canvas of the longer side, filled with `background_color` (RGB in [0, 1],
    truncated to 8 bits like the reference, transform.py:13-36)."""

    def __init__(self, background_color: Sequence[float] = CLIP_MEAN):
        self.fill = tuple(int(c * 255) for c in background_color)

    def __call__(self, img: Image.Image) -> Image.Image:
        w, h = img.size
        if w == h:
            return img
        side = max(w, h)
        canvas = Image.new(img.mode, (side, side), self.fill)
        canvas.paste(img, ((side - w) // 2, (side - h) // 2))
        return canvas

    def __repr__(self) -> str:
        return f"PadToSquare(fill={self.fill})"


def get_transform(transform_type: str, size: int = 224):
    import torchvision.transforms as T

    bicubic = T.InterpolationMode.BICUBIC
    geometry = {
        "padded_resize": lambda: [PadToSquare(CLIP_MEAN), T.Resize(size, interpolation=bicubic)],
        "resized_center_crop": lambda: [T.Resize(size, interpolation=bicubic), T.CenterCrop(size)],
        "random_resized_crop": lambda: [T.RandomResizedCrop(size=(size, size), scale=(0.9, 1.0), ratio=(0.75, 1.3333),
                                                            interpolation=bicubic, antialias=None)],
    }
    if transform_type not in geometry:
        raise ValueError(f"unknown transform type: {transform_type}")
    return T.Compose(geometry[transform_type]() + [T.ToTensor(), T.Normalize(mean=list(CLIP_MEAN), std=list(CLIP_STD))])
