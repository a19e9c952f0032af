"""Sub-pixel key-point refinement: ``models/model_wrap.py:200-234`` + ``utils/losses.py:48-129``.

The soft-argmax itself lives in ``torchgeometry`` (unpinned in the reference's requirements.txt:19 and absent
from this image), so this step is restated from its published definition and is PARITY-UNPINNED.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def soft_argmax_points(heat, pts, patch_size: int = 5) -> np.ndarray:
    """heat [H,W] (cuda tensor or numpy), pts float64 [3,K] -> float64 [3,K] with sub-pixel x, y."""
    if not torch.cuda.is_available():
        raise RuntimeError("soft_argmax_points needs a CUDA device; there is no CPU fallback")
    h = torch.as_tensor(heat) if not torch.is_tensor(heat) else heat
    dev = h.device if h.is_cuda else torch.device("cuda", torch.cuda.current_device())  # stay on the heat-map's GPU
    h = h.to(dev, torch.float32).squeeze().contiguous()
    H, W = h.shape
    K = pts.shape[1]
    out = np.array(pts, dtype=np.float64, copy=True)
    if K == 0:
        return out
    p = torch.as_tensor(np.ascontiguousarray(np.asarray(pts, np.float64).T), dtype=torch.float32).to(h.device)
    dxdy = torch.empty((K, 2), device=h.device, dtype=torch.float32)
    with torch.cuda.device(h.device):
        _lib.check(_lib.lib().spb_soft_argmax_points(h.data_ptr(), H, W, p.contiguous().data_ptr(), None, K,
                                                     patch_size, dxdy.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream), "soft_argmax_points")
    out[:2, :] = out[:2, :] + dxdy.cpu().numpy().T - patch_size // 2
    return out
