"""Host-side mirror of the reference's free functions on the hot path, backed by libspb200.so.

Same names, argument meaning and return conventions as the reference
(``utils/utils.py``: inv_warp_image_batch 318-351, compute_valid_mask 677-704, flattenDetection
480-525; ``export.py``: combine_heatmap 49-63; ``models/model_wrap.py``: getPtsFromHeatmap 252-279,
sample_desc_from_points 281-299).  PyTorch supplies device memory and the stream only; every
computation is a hand-written sm_100a kernel reached through the C ABI.  No CPU fallback: a missing
library or a non-CUDA build raises.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

__all__ = ["inv_warp_image_batch", "compute_valid_mask", "ellipse_spans", "flattenDetection", "combine_heatmap",
           "getPtsFromHeatmap", "get_pts_from_heatmap_device", "sample_desc_from_points",
           "sample_desc_device", "sample_desc_batch_device", "nn_match_pairs_device", "ha_aggregate", "ha_finalize",
           "ha_warp_views", "ingest_resize_gray"]


def _dev(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("pytorch_superpoint_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    d = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if d.type != "cuda":
        d = torch.device("cuda", torch.cuda.current_device())
    return d


def _cu(t, dev) -> torch.Tensor:
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.asarray(t))
    return t.to(device=dev, dtype=torch.float32).contiguous()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _back(t: torch.Tensor, device) -> torch.Tensor:
    """The reference returns results on ``device``; honour a CPU request with a D2H copy."""
    if device is not None and torch.device(device).type == "cpu":
        return t.cpu()
    return t


def inv_warp_image_batch(img, mat_homo_inv, device=None, mode="bilinear"):
    """utils/utils.py:318-351.  img [B,1,H,W] (or [H,W]); mat_homo_inv [B,3,3] -> [B,1,H,W].

    A single image with B matrices is broadcast (the HA case) without materialising the copies."""
    dev = _dev(device if device is not None else (img.device if torch.is_tensor(img) and img.is_cuda else None))
    x = _cu(img, dev)
    if x.dim() in (2, 3):
        x = x.reshape(1, 1, x.shape[-2], x.shape[-1])
    m = _cu(mat_homo_inv, dev).reshape(-1, 3, 3)
    B, ch, H, W = x.shape
    if ch != 1:
        raise ValueError("inv_warp_image_batch: single-channel images only (as in the reference's use)")
    N = m.shape[0]
    if B not in (1, N):
        raise ValueError(f"batch mismatch: {B} images, {N} homographies")
    out = torch.empty((N, 1, H, W), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().spb_inv_warp_image_batch(
            x.data_ptr(), 0 if B == 1 else H * W, m.data_ptr(), out.data_ptr(), N, H, W,
            {"bilinear": _lib.MODE_BILINEAR, "nearest": _lib.MODE_NEAREST}[mode], _stream()), "inv_warp_image_batch")
    return _back(out, device)


def compute_valid_mask(image_shape, inv_homography, device=None, erosion_radius=0):
    """utils/utils.py:677-704 -> [B,H,W] fp32 in {0,1}."""
    dev = _dev(device)
    m = _cu(inv_homography, dev).reshape(-1, 3, 3)
    H, W = int(image_shape[0]), int(image_shape[1])
    out = torch.empty((m.shape[0], H, W), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().spb_compute_valid_mask(m.data_ptr(), out.data_ptr(), m.shape[0], H, W, _stream()),
                   "compute_valid_mask")
    if erosion_radius > 0:  # utils/utils.py:699-702, on the device (spb_erode)
        spans = ellipse_spans(erosion_radius * 2)
        eroded = torch.empty_like(out)
        sp = torch.as_tensor(spans, dtype=torch.int32).to(dev).contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().spb_erode(out.data_ptr(), eroded.data_ptr(), m.shape[0], H, W, sp.data_ptr(),
                                            len(spans), erosion_radius, erosion_radius, _stream()), "erode")
        out = eroded
    return _back(out, device)


def ellipse_spans(ksize: int):
    """Rows of ``cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (ksize, ksize))`` as column spans [j1, j2) -- OpenCV's
    formula restated (half-axes r = c = ksize // 2, dx = round_half_even(c * sqrt(1 - dy^2 / r^2)), row i covers
    [max(c - dx, 0), min(c + dx + 1, ksize))); the anchor cv2.erode uses by default is (ksize // 2, ksize // 2)."""
    r = c = ksize // 2
    inv_r2 = 1.0 / (float(r) * r) if r else 0.0
    spans = []
    for i in range(ksize):
        dy = i - r
        if abs(dy) <= r:
            dx = int(np.rint(c * np.sqrt((r * r - dy * dy) * inv_r2)))
            spans.append([max(c - dx, 0), min(c + dx + 1, ksize)])
        else:
            spans.append([0, 0])
    return spans


def ha_warp_views(imgs, mats_fwd, view_image=None):
    """The image side of the HA branch (datasets/Coco.py:279-282) with the kernels of the fused pass: imgs [nb,H,W]
    (or [H,W]); mats_fwd [V,3,3]; view_image int [V] (default: all views belong to image 0, or V/nb views per image
    in order).  Returns (views [V,H,W] fp32, bits [V,H,W/8] uint8): the views within ~1e-3 of
    ``inv_warp_image_batch``, the packed valid masks EXACTLY those of ``compute_valid_mask``."""
    dev = _dev(imgs.device if torch.is_tensor(imgs) and imgs.is_cuda else None)
    x = _cu(imgs, dev)
    if x.dim() == 2:
        x = x[None]
    nb, H, W = x.shape
    if H < 8 or W < 8 or W % 8:
        raise ValueError(f"ha_warp_views: needs H >= 8 and W a multiple of 8 (packed mask bytes), got {H}x{W}")
    m = _cu(mats_fwd, dev).reshape(-1, 3, 3)
    V = m.shape[0]
    if view_image is None:
        if V % nb:
            raise ValueError(f"{V} views do not divide over {nb} images: pass view_image")
        view_image = torch.arange(V, device=dev, dtype=torch.int32) // (V // nb)
    vi = torch.as_tensor(view_image).to(dev, torch.int32).contiguous()
    views = torch.empty((V, H, W), device=dev, dtype=torch.float32)
    bits = torch.empty((V, H, W // 8), device=dev, dtype=torch.uint8)
    L = _lib.lib()
    ws = torch.empty(int(L.spb_ha_warp_views_workspace_bytes(nb, V, H, W)), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        _lib.check(L.spb_ha_warp_views(x.data_ptr(), m.data_ptr(), vi.data_ptr(), None, nb, V, H, W, views.data_ptr(),
                                       bits.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "ha_warp_views")
    return views, bits


def ingest_resize_gray(img_bgr, size, device=None):
    """datasets/Coco.py:165-179 behind ``cv2.imread``: uint8 [Hs,Ws,3] (numpy or tensor, imread's BGR order) ->
    INTER_AREA resize to ``size`` = (H, W), OpenCV's RGB2GRAY on that memory order, float32 / 255 -> cuda fp32 [H,W].
    Bit-exact against OpenCV (down- and up-scaling); the uint8 image (3 bytes per source pixel) is what crosses to the GPU."""
    dev = _dev(device if device is not None else (img_bgr.device if torch.is_tensor(img_bgr) and img_bgr.is_cuda else None))
    x = img_bgr if torch.is_tensor(img_bgr) else torch.from_numpy(np.ascontiguousarray(img_bgr))
    if x.dtype != torch.uint8 or x.dim() != 3 or x.shape[2] != 3:
        raise ValueError(f"expected a uint8 [H,W,3] image, got {x.dtype} {tuple(x.shape)}")
    x = x.to(dev).contiguous()
    H, W = int(size[0]), int(size[1])
    out = torch.empty((H, W), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().spb_ingest_resize_gray(x.data_ptr(), x.shape[0], x.shape[1], out.data_ptr(), H, W,
                                                     _stream()), "ingest_resize_gray")
    return out


def flattenDetection(semi, tensor=False):
    """utils/utils.py:480-525: [B,65,Hc,Wc] -> [B,1,H,W]; [65,Hc,Wc] -> [1,H,W]."""
    dev = _dev(semi.device if torch.is_tensor(semi) and semi.is_cuda else None)
    s = _cu(semi, dev)
    batch = s.dim() == 4
    if not batch:
        s = s.unsqueeze(0)
    B, ch, Hc, Wc = s.shape
    if ch != 65:
        raise ValueError("flattenDetection expects 65 channels")
    heat = torch.empty((B, 1, Hc * 8, Wc * 8), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().spb_flatten_detection(s.data_ptr(), 0, heat.data_ptr(), B, Hc, Wc, _stream()),
                   "flatten_detection")
    out = heat if batch else heat[0]
    return out if (not torch.is_tensor(semi) or semi.is_cuda) else out.cpu()


def combine_heatmap(heatmap, inv_homographies, mask_2D, device=None):
    """export.py:49-63.  heatmap, mask_2D [N,1,H,W]; inv_homographies [1,N,3,3] -> [1,H,W]."""
    dev = _dev(device if device is not None else (heatmap.device if heatmap.is_cuda else None))
    h = _cu(heatmap, dev)
    k = _cu(mask_2D, dev)
    m = _cu(inv_homographies, dev).reshape(-1, 3, 3)
    N, H, W = h.shape[0], h.shape[-2], h.shape[-1]
    out = torch.empty((1, H, W), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().spb_combine_heatmap(h.data_ptr(), k.data_ptr(), m.data_ptr(), out.data_ptr(), N, H, W,
                                                  _stream()), "combine_heatmap")
    return _back(out, device)


def ha_aggregate(semi_nhwc, mats_unwarp, mats_fwd, H, W, sums=None, accumulate=False):
    """Fused softmax/d2s/mask/un-warp/aggregate (spb_ha_aggregate).  semi_nhwc [N,Hc,Wc,C>=65] fp32 cuda.
    Returns sums [2,H,W] = (sum heat*mask, sum mask)."""
    dev = semi_nhwc.device
    N, Hc, Wc, cs = semi_nhwc.shape
    assert Hc * 8 == H and Wc * 8 == W and semi_nhwc.is_contiguous() and semi_nhwc.dtype == torch.float32
    mu, mf = _cu(mats_unwarp, dev).reshape(-1, 3, 3), _cu(mats_fwd, dev).reshape(-1, 3, 3)
    if sums is None:
        sums = torch.empty((2, H, W), device=dev, dtype=torch.float32)
        accumulate = False
    L = _lib.lib()
    nbytes = L.spb_ha_aggregate_workspace_bytes(N, H, W)
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        _lib.check(L.spb_ha_aggregate(semi_nhwc.data_ptr(), cs, mu.data_ptr(), mf.data_ptr(), N, H, W, sums.data_ptr(),
                                      int(bool(accumulate)), ws.data_ptr(), nbytes, _stream()), "ha_aggregate")
    return sums


def ha_finalize(sums):
    """export.py:63: heat = sums[0] / sums[1]."""
    H, W = sums.shape[-2:]
    heat = torch.empty((H, W), device=sums.device, dtype=torch.float32)
    with torch.cuda.device(sums.device):
        _lib.check(_lib.lib().spb_ha_finalize(sums.data_ptr(), heat.data_ptr(), H, W, _stream()), "ha_finalize")
    return heat


def get_pts_from_heatmap_device(heat, conf_thresh, nms_dist, border_remove=4, top_k=0, capacity=None,
                                background=False):
    """Device-resident form: heat [B,H,W] cuda fp32 -> (pts [B,capacity,3] fp32, counts [B] i32, total [B] i32).
    ``background``: the launch runs beside other kernels (a side stream under the next batch's convolutions): keep
    the NMS kernel at one 1024-thread CTA per SM instead of the batch-dependent size."""
    assert heat.is_cuda and heat.dtype == torch.float32 and heat.dim() == 3
    heat = heat.contiguous()
    B, H, W = heat.shape
    L = _lib.lib()
    if capacity is None:
        d1 = nms_dist + 1
        capacity = top_k if top_k and top_k > 0 else -(-H // d1) * -(-W // d1)
    pts = torch.empty((B, capacity, 3), device=heat.device, dtype=torch.float32)
    counts = torch.empty((B,), device=heat.device, dtype=torch.int32)
    total = torch.empty((B,), device=heat.device, dtype=torch.int32)
    nbytes = L.spb_nms_workspace_bytes(B, H, W, nms_dist)
    ws = torch.empty(nbytes, device=heat.device, dtype=torch.uint8)
    with torch.cuda.device(heat.device):
        _lib.check(L.spb_get_pts_from_heatmap_ex(heat.data_ptr(), B, H, W, float(conf_thresh), int(nms_dist),
                                                 int(border_remove), int(top_k or 0), capacity, pts.data_ptr(),
                                                 counts.data_ptr(), total.data_ptr(), ws.data_ptr(), nbytes,
                                                 1024 if background else 0, _stream()), "get_pts_from_heatmap")
    return pts, counts, total


def getPtsFromHeatmap(heatmap, conf_thresh, nms_dist, border_remove=4, top_k=0):
    """models/model_wrap.py:252-279: heat [H,W] (numpy or tensor) -> float64 [3,K] rows (x, y, conf)."""
    dev = _dev(heatmap.device if torch.is_tensor(heatmap) and heatmap.is_cuda else None)
    h = _cu(heatmap, dev).squeeze()
    pts, counts, _ = get_pts_from_heatmap_device(h[None], conf_thresh, nms_dist, border_remove, top_k)
    k = int(counts[0].item())
    if k == 0:
        return np.zeros((3, 0))
    return pts[0, :k].cpu().numpy().astype(np.float64).T.copy()


def sample_desc_device(coarse, pts, count=None, nhwc=False, cell=8):
    """coarse [D,Hc,Wc] (NCHW slice) or [Hc,Wc,D] (nhwc=True) cuda fp32; pts [K,3] cuda fp32 -> [K,D]."""
    coarse = coarse.contiguous()
    if nhwc:
        Hc, Wc, D = coarse.shape
    else:
        D, Hc, Wc = coarse.shape
    K = pts.shape[0]
    out = torch.empty((K, D), device=coarse.device, dtype=torch.float32)
    if K == 0:
        return out
    with torch.cuda.device(coarse.device):
        _lib.check(_lib.lib().spb_sample_desc_from_points(
            coarse.data_ptr(), int(nhwc), pts.data_ptr(), count.data_ptr() if count is not None else None, K, Hc, Wc,
            D, cell, out.data_ptr(), _stream()), "sample_desc_from_points")
    return out


def sample_desc_batch_device(coarse_nhwc, pts, counts, cell=8, normalize=True):
    """ONE launch for a batch: coarse [B,Hc,Wc,D] cuda fp32 (the network's NHWC output), pts [B,K,3], counts [B] int32
    -> [B,K,D] unit-norm descriptors, the rows behind counts[b] zero filled.  No host synchronisation."""
    coarse = coarse_nhwc.contiguous()
    B, Hc, Wc, D = coarse.shape
    K = pts.shape[1]
    out = torch.empty((B, K, D), device=coarse.device, dtype=torch.float32)
    if K == 0 or B == 0:
        return out
    with torch.cuda.device(coarse.device):
        _lib.check(_lib.lib().spb_sample_desc_batch(coarse.data_ptr(), pts.contiguous().data_ptr(), counts.data_ptr(), B,
                                                    K, Hc, Wc, D, cell, int(bool(normalize)), out.data_ptr(), _stream()),
                   "sample_desc_batch")
    return out


def nn_match_pairs_device(desc1, desc2, nn_thresh):
    """Two-way matching of P image pairs in one launch (models/model_wrap.py:437-480 per pair): desc1 / desc2 [P,K,D]
    point-major cuda fp32, zero rows = padding -> (idx1 [P,K] int32, idx2, score [P,K] fp32, count [P] int32) on the
    device; pair p's matches are the first count[p] entries, ascending idx1."""
    if nn_thresh < 0.0:
        raise ValueError("'nn_thresh' should be non-negative")
    d1, d2 = desc1.contiguous(), desc2.contiguous()
    P, K, D = d1.shape
    if tuple(d2.shape) != (P, K, D):
        raise ValueError("nn_match_pairs_device: both descriptor sets must be [P,K,D]")
    dev = d1.device
    idx1 = torch.empty((P, K), dtype=torch.int32, device=dev)
    idx2 = torch.empty((P, K), dtype=torch.int32, device=dev)
    score = torch.empty((P, K), dtype=torch.float32, device=dev)
    count = torch.empty((P,), dtype=torch.int32, device=dev)
    L = _lib.lib()
    ws = torch.empty(int(L.spb_nn_match_pairs_workspace_bytes(P, K)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.spb_nn_match_pairs(d1.data_ptr(), d2.data_ptr(), P, K, D, float(nn_thresh), idx1.data_ptr(),
                                        idx2.data_ptr(), score.data_ptr(), count.data_ptr(), ws.data_ptr(), ws.numel(),
                                        _stream()), "nn_match_pairs")
    return idx1, idx2, score, count


def sample_desc_from_points(coarse_desc, pts, cell=8):
    """models/model_wrap.py:281-299: coarse [1,D,Hc,Wc], pts float64 [3,K] -> float32 [D,K]."""
    dev = _dev(coarse_desc.device if torch.is_tensor(coarse_desc) and coarse_desc.is_cuda else None)
    c = _cu(coarse_desc, dev)
    D = c.shape[1]
    if pts.shape[1] == 0:
        return np.zeros((D, 0))
    p = torch.as_tensor(np.ascontiguousarray(np.asarray(pts, dtype=np.float64).T), dtype=torch.float32).to(dev)
    if p.shape[1] == 2:
        p = torch.cat([p, torch.zeros_like(p[:, :1])], dim=1)
    out = sample_desc_device(c[0], p.contiguous(), None, nhwc=False, cell=cell)
    return out.cpu().numpy().T.copy()


def nn_match_two_way(desc1, desc2, nn_thresh):
    """models/model_wrap.py:437-480 (PointTracker.nn_match_two_way): desc1 [D,K1], desc2 [D,K2] -> float64 [3,L]
    rows (index in desc1, index in desc2, L2 distance), mutual nearest neighbours below ``nn_thresh``."""
    d1 = np.asarray(desc1) if not torch.is_tensor(desc1) else desc1
    d2 = np.asarray(desc2) if not torch.is_tensor(desc2) else desc2
    assert d1.shape[0] == d2.shape[0]
    if d1.shape[1] == 0 or d2.shape[1] == 0:
        return np.zeros((3, 0))
    if nn_thresh < 0.0:
        raise ValueError("'nn_thresh' should be non-negative")
    dev = _dev(d1.device if torch.is_tensor(d1) and d1.is_cuda else None)
    a = torch.as_tensor(d1, dtype=torch.float32).to(dev).contiguous()
    b = torch.as_tensor(d2, dtype=torch.float32).to(dev).contiguous()
    D, K1, K2 = a.shape[0], a.shape[1], b.shape[1]
    L = _lib.lib()
    n = min(K1, K2)
    idx1 = torch.empty(n, dtype=torch.int32, device=dev)
    idx2 = torch.empty(n, dtype=torch.int32, device=dev)
    score = torch.empty(n, dtype=torch.float32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(int(L.spb_nn_match_workspace_bytes(K1, K2)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.spb_nn_match_two_way(a.data_ptr(), K1, b.data_ptr(), K2, D, float(nn_thresh), idx1.data_ptr(),
                                          idx2.data_ptr(), score.data_ptr(), count.data_ptr(), ws.data_ptr(), ws.numel(),
                                          torch.cuda.current_stream().cuda_stream), "nn_match_two_way")
    k = int(count.item())
    out = np.zeros((3, k))
    out[0], out[1], out[2] = idx1[:k].cpu().numpy(), idx2[:k].cpu().numpy(), score[:k].cpu().numpy()
    return out
