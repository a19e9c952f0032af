"""``SuperPointFrontend_b200`` -- same surface as ``models/model_wrap.py:36-408 SuperPointFrontend_torch``.

Constructor, ``loadModel``, ``net_parallel``, ``run``, ``getPtsFromHeatmap``, ``nms_fast``,
``sample_desc_from_points``, ``heatmap`` / ``pts`` attributes keep the reference's names, argument
meaning, return types (numpy float64 ``[3,K]`` points, float32 ``[256,K]`` descriptors) and error
behaviour (plain Python exceptions).  The work is done by libspb200.so kernels.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .net import SuperPointNet_b200


class SuperPointFrontend_b200(object):
    """Wrapper around the B200 net for pre/post processing (mirror of SuperPointFrontend_torch)."""

    def __init__(self, config, weights_path, nms_dist, conf_thresh, nn_thresh,
                 cuda=False, trained=False, device="cuda", grad=False, load=True):
        self.config = config
        self.name = "SuperPoint"
        self.cuda = cuda
        self.nms_dist = nms_dist
        self.conf_thresh = conf_thresh
        self.nn_thresh = nn_thresh
        self.cell = 8
        self.border_remove = 4
        self.sparsemap = None
        self._heatmap = None
        self.pts = None
        self.pts_subpixel = None
        self.patches = None
        self.device = device
        self.subpixel = bool(self.config["model"]["subpixel"]["enable"])
        self.net = None
        if load:
            self.loadModel(weights_path)

    # models/model_wrap.py:72-111.  Both branches of the reference are honoured: '*.tar' checkpoints carry
    # {'model_state_dict': ...}; anything else is a bare state_dict (MagicLeap superpoint_v1.pth layout).
    def loadModel(self, weights_path):
        ckpt = torch.load(weights_path, map_location=lambda storage, loc: storage)
        sd = ckpt["model_state_dict"] if weights_path[-4:] == ".tar" else ckpt
        self.net = SuperPointNet_b200()
        self.net.load_state_dict(sd)

    def net_parallel(self):
        """models/model_wrap.py:113-115 wraps the net in nn.DataParallel.  Here multi-GPU is homography
        sharding + one NCCL all-reduce (pytorch_superpoint_b200.sharding); nothing to wrap."""
        print("=== SuperPointFrontend_b200: homography sharding over", torch.cuda.device_count(), "GPUs (no DataParallel)")

    @property
    def points(self):
        return self.pts

    @property
    def heatmap(self):
        return self._heatmap

    @heatmap.setter
    def heatmap(self, heatmap):
        self._heatmap = heatmap

    def getSparsemap(self):
        return self.sparsemap

    # ---- models/model_wrap.py:117-180 --------------------------------------------------------------------
    def nms_fast(self, in_corners, H, W, dist_thresh):
        """Greedy grid NMS of arbitrary corners [3,N] (x, y, conf) -> (corners[3,K], indices[K]).

        The confidence order (a stable argsort, host side) is turned into exact fp32 priorities so that
        float64 confidences keep their order on the device; duplicates that round to one pixel follow
        the reference: the first (most confident) decides suppression, the last one is reported."""
        n = in_corners.shape[1]
        if n == 0:
            return np.zeros((3, 0)).astype(int), np.zeros(0).astype(int)
        inds1 = np.argsort(-in_corners[2, :], kind="stable")
        corners = in_corners[:, inds1]
        rc = corners[:2, :].round().astype(int)
        if n == 1:
            return np.vstack((rc, in_corners[2])).reshape(3, 1), np.zeros((1)).astype(int)
        if n >= (1 << 24):
            raise ValueError("nms_fast: too many corners")
        prio = np.zeros((H, W), np.float32)
        order = np.arange(n)
        prio[rc[1, ::-1], rc[0, ::-1]] = (n - order[::-1]).astype(np.float32)  # first duplicate wins
        owner = np.zeros((H, W), int)
        owner[rc[1], rc[0]] = order                                              # last duplicate wins
        kept = ops.getPtsFromHeatmap(prio, 0.5, dist_thresh, border_remove=0)
        ky, kx = kept[1].astype(int), kept[0].astype(int)
        rm = np.lexsort((kx, ky))                                                 # np.where order
        inds_keep = owner[ky[rm], kx[rm]]
        out = corners[:, inds_keep]
        inds2 = np.argsort(-out[-1, :], kind="stable")
        return out[:, inds2], inds1[inds_keep[inds2]]

    # ---- models/model_wrap.py:252-279 --------------------------------------------------------------------
    def getPtsFromHeatmap(self, heatmap):
        if torch.is_tensor(heatmap):
            h = heatmap.squeeze()
            self.sparsemap = (h >= self.conf_thresh)
        else:
            h = np.asarray(heatmap).squeeze()
            self.sparsemap = (h >= self.conf_thresh)
        return ops.getPtsFromHeatmap(h, self.conf_thresh, self.nms_dist, self.border_remove)

    # ---- models/model_wrap.py:281-299 --------------------------------------------------------------------
    def sample_desc_from_points(self, coarse_desc, pts):
        return ops.sample_desc_from_points(coarse_desc, pts, cell=self.cell)

    # ---- models/model_wrap.py:200-234 (sub-pixel refinement; torchgeometry soft-argmax, parity unpinned) --
    def soft_argmax_points(self, pts, patch_size=5):
        from .subpixel import soft_argmax_points
        heat = self.heatmap
        out = soft_argmax_points(heat, pts[0], patch_size)
        self.pts_subpixel = [out]
        return [out.copy()]

    # ---- models/model_wrap.py:322-408 --------------------------------------------------------------------
    def run(self, inp, onlyHeatmap=False, train=True):
        """inp [B,1,H,W] float32 in [0,1].  onlyHeatmap -> heat [B,1,H,W] (cuda tensor);
        else (pts list of [3,K] f64, pts_desc list of [K,256], dense_desc [B,256,H,W], heatmap)."""
        if self.net is None:
            raise RuntimeError("no network loaded")
        dev = torch.device(self.device) if torch.device(self.device).type == "cuda" else torch.device("cuda")
        inp = inp.to(dev)
        B, H, W = inp.shape[0], inp.shape[2], inp.shape[3]
        with torch.no_grad():
            semi, desc = self.net.forward_nhwc(inp, want_desc=not onlyHeatmap)
        from . import _lib
        heat = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().spb_flatten_detection(semi.data_ptr(), 65, heat.data_ptr(), B, H // 8, W // 8,
                                                        torch.cuda.current_stream().cuda_stream), "flatten_detection")
        self.heatmap = heat
        if onlyHeatmap:
            return heat
        p_dev, c_dev, _ = ops.get_pts_from_heatmap_device(heat[:, 0], self.conf_thresh, self.nms_dist, self.border_remove)
        # models/model_wrap.py:393-404: dense = normalise(interpolate(coarse, x8)) in one fused kernel; the key-point
        # columns of all images are gathered in one launch, and only they (not the 78 MB/image dense tensor) travel
        D = desc.shape[-1]
        dense = torch.empty((B, D, H, W), device=dev, dtype=torch.float32)
        cols = torch.empty((B, p_dev.shape[1], D), device=dev, dtype=torch.float32)
        L = _lib.lib()
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(L.spb_dense_desc(desc.data_ptr(), B, H // self.cell, W // self.cell, D, self.cell,
                                        dense.data_ptr(), st), "dense_desc")
            _lib.check(L.spb_gather_dense_desc(dense.data_ptr(), p_dev.data_ptr(), c_dev.data_ptr(), B, p_dev.shape[1], D,
                                               H, W, cols.data_ptr(), st), "gather_dense_desc")
        counts = c_dev.cpu().numpy()
        p = p_dev.cpu().numpy().astype(np.float64)
        cols = cols.cpu().numpy()
        pts = [p[i, :counts[i]].T.copy() for i in range(B)]
        self.pts = pts
        # (numpy's dense_cpu[i, :, ys, xs].transpose() in the reference comes out as [D, K]; same here)
        pts_desc = [cols[i, :counts[i]].T.copy() for i in range(B)]
        if self.subpixel:
            raise NotImplementedError("residual sub-pixel head (SubpixelNet) is outside the hot path")
        return pts, pts_desc, dense, heat


def _extract_batch(self, inp, top_k=0):
    """Batched single-frame extraction (export.py:127-145 per image, BASELINE configs[0]/[2]) with everything left on
    the device: inp [B,1,H,W] -> (pts [B,cap,3] fp32 rows (x, y, conf) sorted by confidence, counts [B] int32,
    desc [B,cap,256] fp32 unit-norm descriptors sampled at the key-points).  No host synchronisation."""
    if self.net is None:
        raise RuntimeError("no network loaded")
    dev = torch.device(self.device) if torch.device(self.device).type == "cuda" else torch.device("cuda")
    inp = inp.to(dev)
    B, H, W = inp.shape[0], inp.shape[2], inp.shape[3]
    from . import _lib
    with torch.no_grad():
        semi, desc = self.net.forward_nhwc(inp, want_desc=True)
    heat = torch.empty((B, H, W), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().spb_flatten_detection(semi.data_ptr(), 65, heat.data_ptr(), B, H // 8, W // 8,
                                                    torch.cuda.current_stream().cuda_stream), "flatten_detection")
    pts, counts, _ = ops.get_pts_from_heatmap_device(heat, self.conf_thresh, self.nms_dist, self.border_remove, top_k)
    out = ops.sample_desc_batch_device(desc, pts, counts, cell=self.cell)  # one launch for the batch
    self.heatmap = heat[:, None]
    return pts, counts, out


SuperPointFrontend_b200.extract_batch = _extract_batch


from .tracker import PointTracker_b200  # noqa: E402,F401  (models/model_wrap.py:411-583; kept importable from here)
