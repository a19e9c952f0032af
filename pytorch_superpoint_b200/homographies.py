"""Homography sampling for homography adaptation (host side, numpy; the step before the hot path).

Restates ``utils/homographies.py:12-141`` (``sample_homography_np``) and the HA branch of
``datasets/Coco.py:262-276``.  The RNG call sequence (scipy ``truncnorm.rvs`` x3, ``rvs(n_scales)``,
``randint``, ``uniform`` x2, ``randint``) is kept so that ``np.random.seed(s)`` reproduces the
reference's draws; pinned by ``tests/golden/homographies.npz``.
"""
from __future__ import annotations

from math import pi

import numpy as np

__all__ = ["sample_homography", "make_ha_homographies", "sample_ha_homographies_device", "EXPORT_PARAMS"]

# configs/magicpoint_coco_export.yaml:19-28
EXPORT_PARAMS = dict(translation=True, rotation=True, scaling=True, perspective=True,
                     scaling_amplitude=0.2, perspective_amplitude_x=0.2, perspective_amplitude_y=0.2,
                     allow_artifacts=True, patch_ratio=0.85)


def _perspective_transform(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """3x3 H with dst ~ H.src for 4 point pairs (what cv2.getPerspectiveTransform solves)."""
    try:
        import cv2
        return cv2.getPerspectiveTransform(np.float32(src), np.float32(dst))
    except ImportError:  # same 8x8 system, float64
        s, d = np.float32(src).astype(np.float64), np.float32(dst).astype(np.float64)
        A, b = np.zeros((8, 8)), np.zeros(8)
        for i in range(4):
            A[i] = [s[i, 0], s[i, 1], 1, 0, 0, 0, -s[i, 0] * d[i, 0], -s[i, 1] * d[i, 0]]
            A[i + 4] = [0, 0, 0, s[i, 0], s[i, 1], 1, -s[i, 0] * d[i, 1], -s[i, 1] * d[i, 1]]
            b[i], b[i + 4] = d[i, 0], d[i, 1]
        return np.append(np.linalg.solve(A, b), 1.0).reshape(3, 3)


def sample_homography(shape, shift=0, perspective=True, scaling=True, rotation=True, translation=True,
                      n_scales=5, n_angles=25, scaling_amplitude=0.1, perspective_amplitude_x=0.1,
                      perspective_amplitude_y=0.1, patch_ratio=0.5, max_angle=pi / 2,
                      allow_artifacts=False, translation_overflow=0.0) -> np.ndarray:
    """Random homography between a random patch of the unit square and the full frame."""
    from scipy.stats import truncnorm

    corners = np.array([[0.0, 0.0], [0.0, 1.0], [1.0, 1.0], [1.0, 0.0]])
    margin = (1 - patch_ratio) / 2
    patch = margin + patch_ratio * corners
    trunc = 2

    if perspective:
        if not allow_artifacts:
            perspective_amplitude_x = min(perspective_amplitude_x, margin)
            perspective_amplitude_y = min(perspective_amplitude_y, margin)
        dy = truncnorm(-trunc, trunc, loc=0, scale=perspective_amplitude_y / 2).rvs(1)[0]
        dxl = truncnorm(-trunc, trunc, loc=0, scale=perspective_amplitude_x / 2).rvs(1)[0]
        dxr = truncnorm(-trunc, trunc, loc=0, scale=perspective_amplitude_x / 2).rvs(1)[0]
        patch = patch + np.array([[dxl, dy], [dxl, -dy], [dxr, dy], [dxr, -dy]])

    if scaling:
        s = truncnorm(-trunc, trunc, loc=1, scale=scaling_amplitude / 2).rvs(n_scales)
        s = np.concatenate(([1.0], s))
        c = patch.mean(axis=0, keepdims=True)
        cand = (patch - c)[None] * s[:, None, None] + c
        if allow_artifacts:
            ok = np.arange(n_scales)
        else:
            ok = np.where(((cand >= 0.0) & (cand < 1.0)).all(axis=(1, 2)))[0]
        patch = cand[int(ok[np.random.randint(ok.shape[0], size=1)].squeeze())]

    if translation:
        lo, hi = patch.min(axis=0), (1 - patch).min(axis=0)
        if allow_artifacts:
            lo, hi = lo + translation_overflow, hi + translation_overflow
        patch = patch + np.array([np.random.uniform(-lo[0], hi[0], 1), np.random.uniform(-lo[1], hi[1], 1)]).T

    if rotation:
        ang = np.concatenate((np.linspace(-max_angle, max_angle, num=n_angles), [0.0]))
        c = patch.mean(axis=0, keepdims=True)
        R = np.stack([np.cos(ang), -np.sin(ang), np.sin(ang), np.cos(ang)], axis=1).reshape(-1, 2, 2)
        cand = np.matmul((patch - c)[None], R) + c
        if allow_artifacts:
            ok = np.arange(n_angles)
        else:
            ok = np.where(((cand >= 0.0) & (cand < 1.0)).all(axis=(1, 2)))[0]
        patch = cand[int(ok[np.random.randint(ok.shape[0], size=1)].squeeze())]

    wh = np.asarray(shape, dtype=np.float64)[::-1]
    return _perspective_transform(corners * wh[None] + shift, patch * wh[None] + shift)


def make_ha_homographies(num: int, seed: int | None = None, params: dict | None = None):
    """datasets/Coco.py:265-276 -> (homographies, inv_homographies), each float32 [num,3,3].

    ``homographies[i] = inv(sampled_i)`` with ``homographies[0] = I`` is what export.py un-warps
    with; ``inv_homographies = inverse(homographies)`` (fp32) is what the image and the valid
    mask are warped with.
    """
    import torch

    if seed is not None:
        np.random.seed(seed)
    p = dict(EXPORT_PARAMS if params is None else params)
    hs = np.stack([np.linalg.inv(sample_homography(np.array([2, 2]), shift=-1, **p)) for _ in range(num)])
    hs[0] = np.identity(3)
    hs = torch.tensor(hs, dtype=torch.float32)
    inv = torch.stack([torch.inverse(hs[i]) for i in range(num)])
    return hs.numpy(), inv.numpy()


def sample_ha_homographies_device(num: int, images: int = 1, seed: int = 0, params: dict | None = None, device=None,
                                  return_raw: bool = False):
    """GPU form of datasets/Coco.py:262-276 for a whole batch: ``images`` x ``num`` homographies drawn by one kernel
    launch (libspb200 ``spb_sample_ha_homographies``; Philox counters, so a (seed, index) pair is reproducible but the
    stream differs from numpy's -- distributional parity with ``utils/homographies.py:12-141``).

    Returns cuda float32 ``(homographies, inv_homographies)`` of shape ``[images, num, 3, 3]`` with
    ``homographies[:, 0] = I`` (and the raw ``sample_homography_np`` matrices when ``return_raw``).  The host sampler
    above needs ~20 ms per homography (a scipy ``truncnorm`` object per draw, like the reference); this one is what
    ``export_detector_homoAdapt_b200`` uses when the data loader does not provide the matrices."""
    import torch

    from . import _lib
    if not torch.cuda.is_available():
        raise RuntimeError("sample_ha_homographies_device needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    p = dict(perspective=True, scaling=True, rotation=True, translation=True, n_scales=5, n_angles=25,
             scaling_amplitude=0.1, perspective_amplitude_x=0.1, perspective_amplitude_y=0.1, patch_ratio=0.5,
             max_angle=pi / 2, allow_artifacts=False, translation_overflow=0.0)  # defaults of sample_homography_np
    p.update(EXPORT_PARAMS if params is None else params)
    total = int(images) * int(num)
    h = torch.empty((images, num, 3, 3), dtype=torch.float32, device=dev)
    inv = torch.empty_like(h)
    raw = torch.empty_like(h) if return_raw else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().spb_sample_ha_homographies(
            int(seed) & 0xFFFFFFFFFFFFFFFF, total, int(num), int(bool(p["perspective"])), int(bool(p["scaling"])),
            int(bool(p["rotation"])), int(bool(p["translation"])), int(p["n_scales"]), int(p["n_angles"]),
            float(p["scaling_amplitude"]), float(p["perspective_amplitude_x"]), float(p["perspective_amplitude_y"]),
            float(p["patch_ratio"]), float(p["max_angle"]), int(bool(p["allow_artifacts"])),
            float(p["translation_overflow"]), h.data_ptr(), inv.data_ptr(), raw.data_ptr() if return_raw else None,
            torch.cuda.current_stream().cuda_stream), "sample_ha_homographies")
    return (h, inv, raw) if return_raw else (h, inv)
