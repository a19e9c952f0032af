"""CPU restatement (numpy / torch-fp32) of the reference's HA-export / extraction path.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Every function cites the
reference ``file:line`` it restates (paths relative to ``/root/reference``).

Float discipline: everything that decides an integer / 0-1 result (coordinates,
nearest-neighbour masks, thresholds) is computed in explicit IEEE float32 with one
rounding per elementary operation and NO fused multiply-add, in the operation
order written here.  The CUDA kernels use ``__fmul_rn/__fadd_rn/__fdiv_rn`` in the
same order, so kernel-vs-oracle coordinates and masks are bit-identical; the
oracle-vs-reference delta (the reference's matmul / vectorised grid_sample order
is library-defined) is pinned by the golden fixtures.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

f32 = np.float32

__all__ = [
    "linspace_f32", "warp_grid", "inv_warp_image_batch", "compute_valid_mask",
    "flatten_detection", "combine_heatmap", "nms_fast", "get_pts_from_heatmap",
    "sample_desc_from_points", "nn_match_two_way", "canonical_params", "net_forward",
    "fold_bn", "ha_export_one", "extract_one", "soft_argmax_points",
    "LAYERS", "synthetic_state_dict",
]


# --------------------------------------------------------------------------------------
# coordinates: utils/utils.py:341-344 (linspace grid), 286-314 (warp_points),
#              torch grid_sample(align_corners=True) un-normalisation
# --------------------------------------------------------------------------------------
def linspace_f32(steps: int) -> np.ndarray:
    """``torch.linspace(-1, 1, steps)`` in float32 (utils/utils.py:341).

    torch fills the first half as ``fma(step, i, start)`` and the second half as
    ``fma(-step, steps-1-i, end)`` with ``step = fp32(2/(steps-1))`` (verified bit-equal in
    tests/test_oracle_golden.py).  The fused multiply-add is emulated exactly in float64: the
    product of an fp32 step and a small integer plus/minus 1 is exact there, so the single final
    rounding to fp32 is the FMA's rounding.
    """
    if steps == 1:
        return np.array([-1.0], dtype=f32)
    step = np.float64(f32(f32(2.0) / f32(steps - 1)))
    i = np.arange(steps)
    lo = (-1.0 + step * i).astype(f32)
    hi = (1.0 - step * (steps - 1 - i)).astype(f32)
    return np.where(i < steps // 2, lo, hi).astype(f32)


def warp_grid(mats: np.ndarray, H: int, W: int):
    """Source pixel coordinates (ix, iy), each [N,H,W] fp32, for ``out(p) = img(M.p)``.

    utils/utils.py:286-314 (``[u,v,1] . M^T`` then divide by w) followed by grid_sample's
    ``((g + 1) / 2) * (size - 1)``.
    """
    m = np.asarray(mats, dtype=f32).reshape(-1, 3, 3)
    u = linspace_f32(W)[None, None, :]
    v = linspace_f32(H)[None, :, None]

    def row(r):
        a = m[:, r, 0][:, None, None] * u
        b = m[:, r, 1][:, None, None] * v
        return (a + b) + m[:, r, 2][:, None, None]

    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        X, Y, Z = row(0), row(1), row(2)
        gx = X / Z
        gy = Y / Z
        ix = ((gx + f32(1.0)) / f32(2.0)) * f32(W - 1)
        iy = ((gy + f32(1.0)) / f32(2.0)) * f32(H - 1)
    return ix.astype(f32), iy.astype(f32)


def _gather(img: np.ndarray, xf: np.ndarray, yf: np.ndarray):
    """img [N or 1, (C,) H, W]; returns (values, in_bounds) with zero padding semantics."""
    H, W = img.shape[-2:]
    with np.errstate(invalid="ignore"):
        ok = (xf >= 0) & (xf <= W - 1) & (yf >= 0) & (yf <= H - 1)
    xi = np.where(ok, xf, 0).astype(np.int64)
    yi = np.where(ok, yf, 0).astype(np.int64)
    n = np.arange(xf.shape[0])[:, None, None] if img.shape[0] == xf.shape[0] else np.zeros((1, 1, 1), np.int64)
    return img[n, yi, xi], ok


def _bilinear(img: np.ndarray, ix: np.ndarray, iy: np.ndarray) -> np.ndarray:
    """4-tap zero-padded bilinear, accumulation order nw, ne, sw, se (ATen grid_sampler)."""
    with np.errstate(invalid="ignore", over="ignore"):
        x0 = np.floor(ix)
        y0 = np.floor(iy)
        x1 = x0 + f32(1)
        y1 = y0 + f32(1)
        w_nw = (x1 - ix) * (y1 - iy)
        w_ne = (ix - x0) * (y1 - iy)
        w_sw = (x1 - ix) * (iy - y0)
        w_se = (ix - x0) * (iy - y0)
        acc = np.zeros(ix.shape, dtype=f32)
        for xf, yf, w in ((x0, y0, w_nw), (x1, y0, w_ne), (x0, y1, w_sw), (x1, y1, w_se)):
            val, ok = _gather(img, xf, yf)
            acc = acc + np.where(ok, val * w, f32(0))
    return acc.astype(f32)


def inv_warp_image_batch(img, mats, mode: str = "bilinear") -> np.ndarray:
    """utils/utils.py:318-351.  img [N,H,W] (or [H,W] / [1,H,W] broadcast), mats [N,3,3] -> [N,H,W]."""
    img = np.asarray(img, dtype=f32)
    if img.ndim == 2:
        img = img[None]
    if img.ndim == 4:
        img = img[:, 0]
    mats = np.asarray(mats, dtype=f32).reshape(-1, 3, 3)
    H, W = img.shape[-2:]
    ix, iy = warp_grid(mats, H, W)
    if mode == "bilinear":
        return _bilinear(img, ix, iy)
    if mode == "nearest":
        with np.errstate(invalid="ignore"):
            val, ok = _gather(img, np.rint(ix), np.rint(iy))  # nearbyint: half-to-even
        return np.where(ok, val, f32(0)).astype(f32)
    raise ValueError(mode)


def compute_valid_mask(image_shape, inv_homography, erosion_radius: int = 0) -> np.ndarray:
    """utils/utils.py:677-704: nearest-warp of an all-ones image (+ optional cv2 ellipse erosion)."""
    H, W = int(image_shape[0]), int(image_shape[1])
    mats = np.asarray(inv_homography, dtype=f32).reshape(-1, 3, 3)
    mask = inv_warp_image_batch(np.ones((1, H, W), f32), mats, mode="nearest")
    if erosion_radius > 0:
        import cv2  # same library call as the reference (utils/utils.py:699-702)
        k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (erosion_radius * 2,) * 2)
        mask = np.stack([cv2.erode(m, k, iterations=1) for m in mask])
    return mask.astype(f32)


# --------------------------------------------------------------------------------------
# heat-map: utils/utils.py:480-525 (flattenDetection), utils/d2s.py:8-25, export.py:49-63
# --------------------------------------------------------------------------------------
def flatten_detection(semi: np.ndarray) -> np.ndarray:
    """semi [B,65,Hc,Wc] -> heat [B,8Hc,8Wc]: softmax over 65, drop dustbin, depth-to-space(8).

    ``heat[y,x] = p[8*(y%8) + (x%8), y//8, x//8]`` (DepthToSpace(8) == pixel_shuffle).
    """
    s = np.asarray(semi, dtype=f32)
    B, C, Hc, Wc = s.shape
    assert C == 65
    e = np.exp(s - s.max(axis=1, keepdims=True))
    p = (e / e.sum(axis=1, keepdims=True, dtype=f32)).astype(f32)[:, :64]
    return p.reshape(B, 8, 8, Hc, Wc).transpose(0, 3, 1, 4, 2).reshape(B, Hc * 8, Wc * 8)


def combine_heatmap(heat: np.ndarray, mats_unwarp: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """export.py:49-63.  heat, mask [N,H,W]; returns [H,W] = sum unwarp(heat*mask) / sum unwarp(mask)."""
    hm = (np.asarray(heat, f32) * np.asarray(mask, f32)).astype(f32)
    num = inv_warp_image_batch(hm, mats_unwarp, "bilinear").sum(axis=0, dtype=f32)
    den = inv_warp_image_batch(mask, mats_unwarp, "bilinear").sum(axis=0, dtype=f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (num / den).astype(f32)


# --------------------------------------------------------------------------------------
# key-points: models/model_wrap.py:117-180 (nms_fast), 252-279 (getPtsFromHeatmap)
# --------------------------------------------------------------------------------------
def nms_fast(in_corners: np.ndarray, H: int, W: int, dist_thresh: int):
    """Greedy grid NMS in confidence order (ties: input order, i.e. stable sort).

    Returns (corners[3,K] sorted by -conf, indices[K] into the input columns).
    """
    n = in_corners.shape[1]
    if n == 0:  # model_wrap.py:150-151
        return np.zeros((3, 0)).astype(int), np.zeros(0).astype(int)
    order = np.argsort(-in_corners[2, :], kind="stable")
    xy = in_corners[:2, order].round().astype(int)
    if n == 1:  # model_wrap.py:152-154
        return np.vstack((xy, in_corners[2])).reshape(3, 1), np.zeros(1).astype(int)
    d = int(dist_thresh)
    blocked = np.zeros((H + 2 * d, W + 2 * d), dtype=bool)
    # model_wrap.py:156-158: when several corners round to one pixel the grid remembers the LAST
    # one in sorted order (the lowest confidence); that is the column the reference returns.
    owner = np.zeros((H, W), dtype=int)
    owner[xy[1], xy[0]] = np.arange(n)  # numpy assigns in order: the last duplicate wins
    kept = []
    for j in range(n):
        x, y = xy[0, j] + d, xy[1, j] + d
        if not blocked[y, x]:
            blocked[y - d:y + d + 1, x - d:x + d + 1] = True
            kept.append(owner[y - d, x - d])
    kept = np.asarray(kept, dtype=int)
    # reference collects survivors in row-major grid order, then re-sorts by -conf (170-179)
    rm = np.lexsort((xy[0, kept], xy[1, kept]))
    kept = kept[rm]
    out = in_corners[:, order[kept]]
    again = np.argsort(-out[2, :], kind="stable")
    return out[:, again], order[kept][again]


def get_pts_from_heatmap(heat: np.ndarray, conf_thresh: float = 0.015, nms_dist: int = 4,
                         border_remove: int = 4) -> np.ndarray:
    """models/model_wrap.py:252-279 -> float64 [3,K] rows (x, y, conf), conf descending."""
    heat = np.asarray(heat).squeeze()
    H, W = heat.shape
    ys, xs = np.where(heat >= conf_thresh)  # row-major; compare happens in heat.dtype
    if len(ys) == 0:
        return np.zeros((3, 0))
    pts = np.zeros((3, len(ys)))
    pts[0], pts[1], pts[2] = xs, ys, heat[ys, xs]
    pts, _ = nms_fast(pts, H, W, nms_dist)
    # model_wrap.py:271-272 re-sorts ascending then reverses; with a stable sort that flips
    # the tie order to (larger row-major index first).  Keep that behaviour.
    pts = pts[:, np.argsort(pts[2, :], kind="stable")[::-1]]
    b = border_remove
    drop = (pts[0] < b) | (pts[0] >= W - b) | (pts[1] < b) | (pts[1] >= H - b)
    return pts[:, ~drop]


def sample_desc_from_points(coarse_desc: np.ndarray, pts: np.ndarray, cell: int = 8) -> np.ndarray:
    """models/model_wrap.py:281-299.  coarse [1,D,Hc,Wc] fp32, pts [3,K] -> [D,K] fp32 unit columns."""
    c = np.asarray(coarse_desc, dtype=f32)
    D, Hc, Wc = c.shape[1:]
    H, W = Hc * cell, Wc * cell
    K = pts.shape[1]
    if K == 0:
        return np.zeros((D, 0))
    gx = (pts[0, :].astype(np.float64) / (float(W) / 2.0) - 1.0).astype(f32)
    gy = (pts[1, :].astype(np.float64) / (float(H) / 2.0) - 1.0).astype(f32)
    ix = ((gx + f32(1)) / f32(2)) * f32(Wc - 1)
    iy = ((gy + f32(1)) / f32(2)) * f32(Hc - 1)
    img = c[0]  # [D,Hc,Wc] treated as a batch of D single-channel images sharing the grid
    ixb = np.broadcast_to(ix[None, None, :], (D, 1, K)).astype(f32)
    iyb = np.broadcast_to(iy[None, None, :], (D, 1, K)).astype(f32)
    desc = _bilinear(img, ixb, iyb)[:, 0, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        desc = desc / np.linalg.norm(desc, axis=0)[None, :]
    return desc.astype(f32)


def nn_match_two_way(desc1: np.ndarray, desc2: np.ndarray, nn_thresh: float) -> np.ndarray:
    """models/model_wrap.py:437-480 -> [3,L] (idx1, idx2, L2 distance)."""
    assert desc1.shape[0] == desc2.shape[0]
    if desc1.shape[1] == 0 or desc2.shape[1] == 0:
        return np.zeros((3, 0))
    if nn_thresh < 0.0:
        raise ValueError("'nn_thresh' should be non-negative")
    d = np.sqrt(2 - 2 * np.clip(desc1.T @ desc2, -1, 1))
    fwd = d.argmin(axis=1)
    sc = d[np.arange(d.shape[0]), fwd]
    bwd = d.argmin(axis=0)
    keep = (sc < nn_thresh) & (bwd[fwd] == np.arange(len(fwd)))
    out = np.zeros((3, int(keep.sum())))
    out[0], out[1], out[2] = np.arange(desc1.shape[1])[keep], fwd[keep], sc[keep]
    return out


def soft_argmax_points(heat: np.ndarray, pts: np.ndarray, patch_size: int = 5) -> np.ndarray:
    """models/model_wrap.py:200-234 + utils/losses.py:48-129 (sub-pixel refinement).

    PARITY UNPINNED: the spatial soft-argmax lives in ``torchgeometry`` (unpinned in
    requirements.txt:19, not installed); restated from its published definition:
    softmax over the log-normalised 5x5 patch, expectation of pixel coordinates.
    """
    heat = np.asarray(heat, f32).squeeze()
    r = patch_size // 2
    pad = np.pad(heat, r, mode="constant")
    out = pts.copy()
    for k in range(pts.shape[1]):
        x, y = int(round(pts[0, k])), int(round(pts[1, k]))
        p = pad[y:y + patch_size, x:x + patch_size].astype(f32)
        p = p / (p.sum() + f32(1e-6))
        p = np.where(p < 0, f32(1e-6), p)  # utils/losses.py:126-129: negatives only; log(0) = -inf
        with np.errstate(divide='ignore'):
            lg = np.log(p)
        e = np.exp(lg - lg.max())
        den = e.sum() + f32(1e-6)
        gy, gx = np.mgrid[0:patch_size, 0:patch_size].astype(f32)
        out[0, k] = pts[0, k] + float((gx * e).sum() / den) - r
        out[1, k] = pts[1, k] + float((gy * e).sum() / den) - r
    return out


# --------------------------------------------------------------------------------------
# network: models/SuperPointNet_pretrained.py:46-76, SuperPointNet_gauss2.py:43-70 (+unet_parts.py:10-48),
#          SuperPointNet.py:154-224
# --------------------------------------------------------------------------------------
# (name, cin, cout, ksize, relu, pool_after)
LAYERS = [
    ("conv1a", 1, 64, 3, True, False), ("conv1b", 64, 64, 3, True, True),
    ("conv2a", 64, 64, 3, True, False), ("conv2b", 64, 64, 3, True, True),
    ("conv3a", 64, 128, 3, True, False), ("conv3b", 128, 128, 3, True, True),
    ("conv4a", 128, 128, 3, True, False), ("conv4b", 128, 128, 3, True, False),
    ("convPa", 128, 256, 3, True, False), ("convPb", 256, 65, 1, False, False),
    ("convDa", 128, 256, 3, True, False), ("convDb", 256, 256, 1, False, False),
]

_GAUSS2 = {
    "conv1a": ("inc.conv.conv.0", "inc.conv.conv.1"), "conv1b": ("inc.conv.conv.3", "inc.conv.conv.4"),
    "conv2a": ("down1.mpconv.1.conv.0", "down1.mpconv.1.conv.1"), "conv2b": ("down1.mpconv.1.conv.3", "down1.mpconv.1.conv.4"),
    "conv3a": ("down2.mpconv.1.conv.0", "down2.mpconv.1.conv.1"), "conv3b": ("down2.mpconv.1.conv.3", "down2.mpconv.1.conv.4"),
    "conv4a": ("down3.mpconv.1.conv.0", "down3.mpconv.1.conv.1"), "conv4b": ("down3.mpconv.1.conv.3", "down3.mpconv.1.conv.4"),
    "convPa": ("convPa", "bnPa"), "convPb": ("convPb", "bnPb"), "convDa": ("convDa", "bnDa"), "convDb": ("convDb", "bnDb"),
}


def canonical_params(state_dict) -> dict:
    """Map any of the three reference state-dict layouts to
    ``{layer: {'w','b', optional 'bn': (gamma,beta,mean,var)}}`` (fp32 torch tensors)."""
    sd = {k.replace("module.", "", 1) if k.startswith("module.") else k: v for k, v in state_dict.items()}
    out = {}
    gauss2 = "inc.conv.conv.0.weight" in sd
    for name, *_ in LAYERS:
        cname, bname = _GAUSS2[name] if gauss2 else (name, "bn" + name[4:])
        ent = {"w": sd[cname + ".weight"].detach().float(), "b": sd[cname + ".bias"].detach().float()}
        if bname + ".running_mean" in sd:
            ent["bn"] = tuple(sd[bname + s].detach().float() for s in (".weight", ".bias", ".running_mean", ".running_var"))
        out[name] = ent
    return out


def fold_bn(params: dict, eps: float = 1e-5) -> dict:
    """Eval-mode BatchNorm folded into conv weight/bias: w' = w*g/sqrt(v+eps), b' = (b-m)*g/sqrt(v+eps)+beta."""
    out = {}
    for name, ent in params.items():
        w, b = ent["w"], ent["b"]
        if "bn" in ent:
            g, beta, m, v = ent["bn"]
            s = g / torch.sqrt(v + eps)
            w = w * s[:, None, None, None]
            b = (b - m) * s + beta
        out[name] = {"w": w.contiguous(), "b": b.contiguous()}
    return out


def net_forward(params: dict, x: torch.Tensor, bn_mode: str = "eval", want_desc: bool = True):
    """x [B,1,H,W] fp32 -> (semi [B,65,H/8,W/8], desc [B,256,H/8,W/8] unit-norm or None).

    ``bn_mode='eval'`` (running statistics; the parity target) or ``'batch'`` (train-mode batch
    statistics = the as-written reference, which never calls ``.eval()``: model_wrap.py:111).
    """
    def block(name, t, relu):
        ent = params[name]
        t = F.conv2d(t, ent["w"], ent["b"], padding=ent["w"].shape[-1] // 2)
        if "bn" in ent:
            g, beta, m, v = ent["bn"]
            if bn_mode == "batch":
                t = F.batch_norm(t, None, None, g, beta, training=True, eps=1e-5)
            else:
                t = F.batch_norm(t, m, v, g, beta, training=False, eps=1e-5)
        return F.relu(t) if relu else t

    with torch.no_grad():
        t = x.float()
        for name, _, _, _, relu, pool in LAYERS[:8]:
            t = block(name, t, relu)
            if pool:
                t = F.max_pool2d(t, 2, 2)
        semi = block("convPb", block("convPa", t, True), False)
        desc = None
        if want_desc:
            d = block("convDb", block("convDa", t, True), False)
            desc = d / torch.norm(d, p=2, dim=1, keepdim=True)
    return semi, desc


def synthetic_state_dict(layout: str = "gauss2", seed: int = 0) -> dict:
    """Deterministic synthetic weights (no checkpoints ship with the reference: .MISSING_LARGE_BLOBS).

    He-scaled normal conv weights, small biases, non-trivial BN affine + running statistics.
    Same generator on both sides of every parity test.  ``layout`` in {'v1','bnflat','gauss2'}.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, cin, cout, k, _, _ in LAYERS:
        cname, bname = _GAUSS2[name] if layout == "gauss2" else (name, "bn" + name[4:])
        std = (2.0 / (cin * k * k)) ** 0.5
        sd[cname + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * std
        sd[cname + ".bias"] = torch.randn(cout, generator=g) * 0.05
        if layout != "v1":
            sd[bname + ".weight"] = 1.0 + 0.1 * torch.randn(cout, generator=g)
            sd[bname + ".bias"] = 0.05 * torch.randn(cout, generator=g)
            sd[bname + ".running_mean"] = 0.05 * torch.randn(cout, generator=g)
            sd[bname + ".running_var"] = 1.0 + 0.2 * torch.rand(cout, generator=g)
            sd[bname + ".num_batches_tracked"] = torch.tensor(1)
    return sd


# --------------------------------------------------------------------------------------
# whole-path drivers
# --------------------------------------------------------------------------------------
def ha_export_one(img: np.ndarray, mats_unwarp: np.ndarray, mats_fwd: np.ndarray, params: dict,
                  conf_thresh=0.015, nms_dist=4, top_k=600, bn_mode="eval", border_remove=4,
                  return_heat=False):
    """One image through datasets/Coco.py:262-284 (warp + mask) and export.py:283-328.

    mats_unwarp = sample['homographies'] (inverse of the sampled homography, [0]=I);
    mats_fwd    = sample['inv_homographies'] = inverse(mats_unwarp).
    Returns pts float64 [K,3] (x, y, conf), K <= top_k (and the aggregated heat-map if asked).
    """
    img = np.asarray(img, f32)
    H, W = img.shape
    warped = inv_warp_image_batch(img[None], mats_fwd, "bilinear")          # Coco.py:279
    mask = compute_valid_mask((H, W), mats_fwd, 0)                            # Coco.py:282
    semi, _ = net_forward(params, torch.from_numpy(warped)[:, None], bn_mode, want_desc=False)
    heat = flatten_detection(semi.numpy())                                    # model_wrap.py:358
    agg = combine_heatmap(heat, mats_unwarp, mask)                            # export.py:310
    pts = get_pts_from_heatmap(agg, conf_thresh, nms_dist, border_remove).T  # export.py:311,322
    if top_k and pts.shape[0] > top_k:
        pts = pts[:top_k]
    return (pts, agg) if return_heat else pts


def extract_one(img: np.ndarray, params: dict, conf_thresh=0.015, nms_dist=4, bn_mode="eval"):
    """Single-frame extraction (export.py:127-145 via Val_model_heatmap.py:88-155):
    returns (pts [3,K] f64, desc [256,K] f32, heat [H,W] f32)."""
    x = torch.from_numpy(np.asarray(img, f32))[None, None]
    semi, desc = net_forward(params, x, bn_mode, want_desc=True)
    heat = flatten_detection(semi.numpy())[0]
    pts = get_pts_from_heatmap(heat, conf_thresh, nms_dist)
    return pts, sample_desc_from_points(desc.numpy(), pts), heat


# --------------------------------------------------------------------------------------
# image ingest: datasets/Coco.py:165-179 (_read_image behind cv2.imread)
# --------------------------------------------------------------------------------------
def _area_tab(ssize: int, dsize: int, scale: float):
    """OpenCV's computeResizeAreaTab (imgproc/resize.cpp): (dst index, src index, weight) entries."""
    import math
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, f32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, f32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, f32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def _resize_area_linear_u8(src: np.ndarray, H: int, W: int) -> np.ndarray:
    """OpenCV's INTER_AREA when an axis is enlarged: the 8-bit linear resize with "area" coefficients (imgproc/resize.cpp:
    s = floor(d*scale), f = (float)((d+1) - (s+1)*inv_scale) reduced to its fractional part, 11-bit fixed-point
    weights, HResizeLinear into int32, VResizeLinear (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2)."""
    import math
    Hs, Ws, C = src.shape

    def axis(dsz, ssz):
        inv = dsz / ssz
        scale = 1.0 / inv
        ofs, w, dmax = np.zeros(dsz, np.int64), np.zeros((dsz, 2), np.int64), dsz
        for d in range(dsz):
            s = math.floor(d * scale)
            f = f32((d + 1) - (s + 1) * inv)
            f = f32(0.0) if f <= 0 else f32(f - math.floor(f))
            if s < 0:
                f, s = f32(0.0), 0
            if s + 1 >= ssz:
                dmax = min(dmax, d)
                if s >= ssz - 1:
                    f, s = f32(0.0), ssz - 1
            ofs[d] = s
            w[d, 0] = int(np.rint(f32((f32(1.0) - f) * f32(2048.0))))
            w[d, 1] = int(np.rint(f32(f * f32(2048.0))))
        return ofs, w, dmax

    xo, xa, xmax = axis(W, Ws)
    yo, ya, _ = axis(H, Hs)
    S = src.astype(np.int64)
    rows = S[:, xo, :] * xa[None, :, 0, None] + S[:, np.minimum(xo + 1, Ws - 1), :] * xa[None, :, 1, None]
    if xmax < W:
        rows[:, xmax:, :] = S[:, xo[xmax:], :] * 2048
    r0, r1 = rows[yo], rows[np.minimum(yo + 1, Hs - 1)]
    v = (((ya[:, 0, None, None] * (r0 >> 4)) >> 16) + ((ya[:, 1, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(v, 0, 255).astype(np.uint8)


def resize_area_u8(src: np.ndarray, H: int, W: int) -> np.ndarray:
    """``cv2.resize(src, (W, H), interpolation=cv2.INTER_AREA)`` for uint8 [Hs,Ws,C]: OpenCV's linear variant when an axis
    is enlarged, else resizeAreaFast_ (integer factors: box sums, (s+2)>>2 for 2x2, round-half-even of s * (1.f/area) otherwise) and
    resizeArea_ (per-axis cell tables, fp32 products and sums in table order) restated."""
    Hs, Ws, C = src.shape
    sx, sy = 1.0 / (W / Ws), 1.0 / (H / Hs)
    if sx < 1.0 or sy < 1.0:
        return _resize_area_linear_u8(src, H, W)
    ix, iy = int(round(sx)), int(round(sy))
    if abs(sx - ix) < np.finfo(np.float64).eps and abs(sy - iy) < np.finfo(np.float64).eps:
        s = src[:H * iy, :W * ix].astype(np.int64).reshape(H, iy, W, ix, C).sum(axis=(1, 3))
        if ix == 2 and iy == 2:
            return ((s + 2) >> 2).astype(np.uint8)
        return np.clip(np.rint(s.astype(f32) * (f32(1.0) / f32(ix * iy))), 0, 255).astype(np.uint8)
    xtab, ytab = _area_tab(Ws, W, sx), _area_tab(Hs, H, sy)
    xd = np.array([t[0] for t in xtab])
    xs = np.array([t[1] for t in xtab])
    xa = np.array([t[2] for t in xtab], f32)
    dst = np.zeros((H, W, C), np.uint8)
    acc = np.zeros((W, C), f32)
    prev = ytab[0][0]
    for dy, sy_i, beta in ytab:
        row = src[sy_i].astype(f32)
        buf = np.zeros((W, C), f32)
        for k in range(len(xtab)):  # (entries of one destination cell are consecutive: sequential fp32 adds)
            buf[xd[k]] = buf[xd[k]] + row[xs[k]] * xa[k]
        if dy != prev:
            dst[prev] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
            acc = (beta * buf).astype(f32)
            prev = dy
        else:
            acc = (acc + beta * buf).astype(f32)
    dst[prev] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return dst


def ingest_resize_gray(img_bgr: np.ndarray, H: int, W: int) -> np.ndarray:
    """datasets/Coco.py:170-178: INTER_AREA resize, ``cv2.COLOR_RGB2GRAY`` on imread's BGR memory order (OpenCV's 15-bit
    fixed point: 9798, 19235, 3735), ``astype('float32') / 255``."""
    r = resize_area_u8(np.asarray(img_bgr, np.uint8), H, W).astype(np.int64)
    g = (r[..., 0] * 9798 + r[..., 1] * 19235 + r[..., 2] * 3735 + (1 << 14)) >> 15
    return g.astype(f32) / f32(255.0)
