"""``Val_model_heatmap_b200`` and ``export_descriptor_b200`` -- the single-frame extraction drop-ins.

``Val_model_heatmap_b200`` mirrors ``Val_model_heatmap.py:33-155`` (constructor from ``config['model']``,
``loadModel()``, ``run(images) -> np [B,1,H,W]``, ``heatmap_to_pts()``, ``soft_argmax_points()``,
``desc_to_sparseDesc()``, attributes ``heatmap`` / ``pts_nms_batch`` / ``desc_sparse_batch`` / ``outs``); it is picked
by ``front_end_model: 'Val_model_heatmap_b200'`` through ``utils/loader.py:147-153`` (INTEGRATION.md shows the one-line
top-level module the reference side needs).  Behind the API the batch stays on the device: network (both heads),
soft-max + depth-to-space, NMS / top-k of all images and the descriptor sampling of all images are one launch each;
only what the caller asks for crosses to the host.

``export_descriptor_b200`` mirrors ``export.py:70-192``: image pairs -> key-points, descriptors and the two-way
matches of ``PointTracker.update`` -> ``<output_dir>/predictions/<i>.npz`` with the reference's keys.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib, ops
from .frontend import SuperPointFrontend_b200
from .net import SuperPointNet_b200
from .tracker import PointTracker_b200

__all__ = ["Val_model_heatmap_b200", "export_descriptor_b200"]


class Val_model_heatmap_b200(SuperPointFrontend_b200):
    def __init__(self, config, device="cuda", verbose=False):
        # Val_model_heatmap.py:33-58 reads these keys of config['model'] (KeyError like the reference when absent)
        self.config = config
        self.model = config["name"]
        self.params = config["params"]
        self.weights_path = config["pretrained"]
        self.device = device
        self.nms_dist = config["nms"]
        self.conf_thresh = config["detection_threshold"]
        self.nn_thresh = config["nn_thresh"]
        self.cell = 8
        self.cell_size = 8
        self.border_remove = 4
        self.sparsemap = None
        self._heatmap = None
        self.pts = None
        self.pts_subpixel = None
        self.pts_nms_batch = None
        self.desc_sparse_batch = None
        self.patches = None
        self.subpixel = False
        self.name = "SuperPoint"
        self.net = None
        self.outs = None
        self._heat_dev = self._pts_dev = self._cnt_dev = None

    def _cuda(self) -> torch.device:
        d = torch.device(self.device)
        if d.type != "cuda":
            if not torch.cuda.is_available():
                raise RuntimeError("Val_model_heatmap_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            d = torch.device("cuda", torch.cuda.current_device())
        return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())

    # Val_model_heatmap.py:61-73: '*.tar' checkpoints carry {'model_state_dict': ...}
    def loadModel(self):
        ckpt = torch.load(self.weights_path, map_location=lambda storage, loc: storage)
        self.net = SuperPointNet_b200(**{k: v for k, v in (self.params or {}).items() if k == "subpixel_channel"},
                                      bn_mode=self.config.get("b200_bn_mode", "eval"))
        self.net.load_state_dict(ckpt["model_state_dict"])

    # Val_model_heatmap.py:88-111
    def run(self, images):
        """images [B,1,H,W] -> heat-map numpy [B,1,H,W]; ``self.outs`` = {'semi','desc'} (device tensors)."""
        if self.net is None:
            raise RuntimeError("no network loaded")
        dev = self._cuda()
        x = images.to(dev)
        B, H, W = x.shape[0], x.shape[2], x.shape[3]
        with torch.no_grad():
            semi, desc = self.net.forward_nhwc(x, want_desc=True)
        heat = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().spb_flatten_detection(semi.data_ptr(), 65, heat.data_ptr(), B, H // 8, W // 8,
                                                        torch.cuda.current_stream().cuda_stream), "flatten_detection")
        self.outs = {"semi": semi.permute(0, 3, 1, 2), "desc": desc.permute(0, 3, 1, 2)}
        self._desc_nhwc, self._heat_dev = desc, heat
        self._pts_dev = self._cnt_dev = None
        self.heatmap = heat.cpu().numpy()
        return self.heatmap

    # Val_model_heatmap.py:114-119: [getPtsFromHeatmap(h) for h in heatmap] -- all images in one launch
    def heatmap_to_pts(self):
        if self._heat_dev is None:
            raise RuntimeError("run() first")
        pts, counts, _ = ops.get_pts_from_heatmap_device(self._heat_dev[:, 0], self.conf_thresh, self.nms_dist,
                                                         self.border_remove)
        self._pts_dev, self._cnt_dev = pts, counts
        c = counts.cpu().numpy()
        p = pts.cpu().numpy().astype(np.float64)
        self.pts_nms_batch = [p[i, :c[i]].T.copy() for i in range(p.shape[0])]
        return self.pts_nms_batch

    # models/model_wrap.py:200-234 (export.py:134-135): refines the FIRST image's points, like the reference
    def soft_argmax_points(self, pts, patch_size=5):
        from .subpixel import soft_argmax_points
        out = soft_argmax_points(self._heat_dev[0, 0], pts[0], patch_size)
        self.pts_subpixel = [out]
        return [out.copy()]

    # Val_model_heatmap.py:151-155 (valid for B = 1 upstream; batched semantics = the per-image loop), one launch
    def desc_to_sparseDesc(self):
        if self._pts_dev is None:
            raise RuntimeError("heatmap_to_pts() first")
        d = ops.sample_desc_batch_device(self._desc_nhwc, self._pts_dev, self._cnt_dev, cell=self.cell)
        c = self._cnt_dev.cpu().numpy()
        d = d.cpu().numpy()
        self.desc_sparse_batch = [d[i, :c[i]].T.copy() if c[i] else np.zeros((d.shape[2], 0)) for i in range(d.shape[0])]
        return self.desc_sparse_batch


def export_descriptor_b200(config, output_dir, args=None, loader=None, net: SuperPointNet_b200 = None):
    """Drop-in for ``export.py:70-192``.  Reads ``config['model']`` (``name, params, pretrained, nms,
    detection_threshold, nn_thresh, subpixel.{enable,patch_size}``), iterates samples with ``image`` /
    ``warped_image`` ``[1,1,H,W]`` and ``homography`` and writes ``<output_dir>/predictions/<i>.npz`` with the keys
    ``image, prob, desc, warped_image, warped_prob, warped_desc, homography, matches`` (shapes and dtypes of the
    reference: prob [K,3] f64, desc [K,256] f32, matches [L,4] f64).  Returns the number of pairs written.
    ``loader`` None imports the reference's ``utils.loader.dataLoader_test`` (running inside its repo); ``net``
    overrides the checkpoint named by ``config['model']['pretrained']``."""
    save_output = os.path.join(str(output_dir), "predictions")
    os.makedirs(save_output, exist_ok=True)
    sub = config["model"].get("subpixel", {}) or {}
    subpixel, patch_size = bool(sub.get("enable", False)), int(sub.get("patch_size", 5))
    if loader is None:
        from utils.loader import dataLoader_test  # reference module (only when run inside its repo)
        loader = dataLoader_test(config, dataset=config["data"]["dataset"])["test_loader"]
    agent = Val_model_heatmap_b200(config["model"], device="cuda")
    if net is not None:
        agent.net = net
    else:
        agent.loadModel()
    tracker = PointTracker_b200(max_length=2, nn_thresh=agent.nn_thresh)

    def pts_desc(img):  # export.py:127-145
        agent.run(img)
        pts = agent.heatmap_to_pts()
        if subpixel:
            pts = agent.soft_argmax_points(pts, patch_size=patch_size)
        desc = agent.desc_to_sparseDesc()  # (sampled at the INTEGER points, Val_model_heatmap.py:153)
        return pts[0], desc[0]

    def squeeze_np(t):
        return (t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)).squeeze()

    count = 0
    for sample in loader:
        img_0, img_1 = torch.as_tensor(sample["image"]), torch.as_tensor(sample["warped_image"])
        pts, desc = pts_desc(img_0)
        tracker.update(pts, desc)
        pred = {"image": squeeze_np(img_0), "prob": pts.transpose(), "desc": desc.transpose()}
        pts, desc = pts_desc(img_1)
        tracker.update(pts, desc)
        pred.update({"warped_image": squeeze_np(img_1), "warped_prob": pts.transpose(),
                     "warped_desc": desc.transpose(), "homography": squeeze_np(sample["homography"]),
                     "matches": tracker.get_matches().transpose()})
        tracker.clear_desc()  # export.py:180: the next pair starts without a previous frame's descriptors
        np.savez_compressed(os.path.join(save_output, "{}.npz".format(count)), **pred)
        count += 1
    return count
