"""Image ingest for the export loops: ``datasets/Coco.py:165-179 _read_image``.

The JPEG / PNG decode stays ``cv2.imread`` on the host (a library call in the reference as well); everything behind
it -- INTER_AREA resize, the RGB2GRAY conversion on imread's BGR memory order, ``float32 / 255`` -- runs on the GPU,
bit-exact against OpenCV (``csrc/ingest.cu``), so only the uint8 pixels cross PCIe and the image arrives where the
HA pass wants it."""
from __future__ import annotations

import torch

from . import ops

__all__ = ["read_image_b200"]


def read_image_b200(path, resize, device=None) -> torch.Tensor:
    """``_read_image(path)`` of datasets/Coco.py:165-179 with ``self.sizer = resize`` -> cuda fp32 [H,W] in [0,1]."""
    import cv2
    img = cv2.imread(str(path))
    if img is None:  # (the reference would fail inside cv2.resize with an opaque assertion)
        raise FileNotFoundError(f"cv2.imread could not read {path}")
    return ops.ingest_resize_gray(img, (int(resize[0]), int(resize[1])), device=device)
