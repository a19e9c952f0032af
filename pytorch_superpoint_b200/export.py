"""Homography-adaptation export (``export.py:196-357 export_detector_homoAdapt_gpu``) on the B200 path.

``HAExporter`` is the per-image engine: host image in -> ``pts`` float64 ``[K,3]`` (x, y, conf) out, doing on the
GPU what the reference splits between the Dataset worker (N image warps + valid masks on the CPU,
datasets/Coco.py:262-284), the network, ``combine_heatmap`` (export.py:49-63) and the NumPy NMS
(models/model_wrap.py:252-279).  ``export_detector_homoAdapt_b200`` keeps the reference's function signature,
config keys and output files (``<output_dir>/predictions/<export_folder>/<name>.npz`` with key ``pts``).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .homographies import make_ha_homographies
from .net import SuperPointNet_b200
from .sharding import owner_of_image, shard_bounds, world_info

__all__ = ["HAExporter", "export_detector_homoAdapt_b200"]


class HAExporter:
    """One process per GPU.  With ``torch.distributed`` initialised (NCCL) the homographies of every image are
    sharded over the ranks and combined by a single all-reduce per batch; otherwise everything runs locally."""

    def __init__(self, net: SuperPointNet_b200, conf_thresh=0.015, nms_dist=4, top_k=600, border_remove=4,
                 device: Optional[torch.device] = None, group=None, chunk: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("HAExporter needs a CUDA device (sm_100a); there is no CPU fallback")
        self.net = net
        self.conf_thresh, self.nms_dist, self.top_k, self.border_remove = conf_thresh, nms_dist, top_k, border_remove
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.group = group
        self.chunk = chunk
        self._bufs = {}

    def _buffers(self, B, H, W):
        key = (B, H, W)
        if key not in self._bufs:
            d1 = self.nms_dist + 1
            cap = self.top_k if self.top_k else -(-H // d1) * -(-W // d1)
            dev = self.device
            self._bufs = {key: dict(
                img=torch.empty((B, H, W), device=dev, dtype=torch.float32),
                sums=torch.empty((B, 2, H, W), device=dev, dtype=torch.float32),
                heat=torch.empty((B, H, W), device=dev, dtype=torch.float32),
                pts_host=torch.empty((B, cap, 3), dtype=torch.float32).pin_memory(),
                cnt_host=torch.empty((B,), dtype=torch.int32).pin_memory(), cap=cap)}
        return self._bufs[key]

    def heatmaps(self, imgs_dev: torch.Tensor, mats_unwarp: torch.Tensor, mats_fwd: torch.Tensor) -> torch.Tensor:
        """imgs_dev [B,H,W] cuda; mats [N,3,3] (shared by the batch) or [B,N,3,3] -> aggregated heat [B,H,W]."""
        B, H, W = imgs_dev.shape
        buf = self._buffers(B, H, W)
        rank, world = world_info(self.group)
        per_image = mats_unwarp.dim() == 4
        N = mats_unwarp.shape[-3]
        lo, hi = shard_bounds(N, rank, world)
        sums = buf["sums"]
        for b in range(B):
            mu = mats_unwarp[b] if per_image else mats_unwarp
            mf = mats_fwd[b] if per_image else mats_fwd
            if hi > lo:
                self.net.ha_heatmap_sums(imgs_dev[b], mu[lo:hi], mf[lo:hi], sums=sums[b], chunk=self.chunk)
            else:
                sums[b].zero_()
        if world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        from . import _lib
        L = _lib.lib()
        st = torch.cuda.current_stream().cuda_stream
        for b in range(B):
            _lib.check(L.spb_ha_finalize(sums[b].data_ptr(), buf["heat"][b].data_ptr(), H, W, st), "ha_finalize")
        return buf["heat"]

    def export_batch(self, imgs_host: Sequence[torch.Tensor], mats_unwarp: torch.Tensor, mats_fwd: torch.Tensor,
                     sync: bool = True):
        """imgs_host: B pinned (or pageable) CPU tensors [H,W] fp32.  Returns a list of float64 [K,3] arrays for
        the images this rank owns (``b % world == rank``; None for the others).  H2D, compute and D2H are all
        enqueued on the current stream; one synchronisation at the end."""
        B = len(imgs_host)
        H, W = imgs_host[0].shape[-2:]
        buf = self._buffers(B, H, W)
        rank, world = world_info(self.group)
        with torch.cuda.device(self.device):
            for b in range(B):
                buf["img"][b].copy_(imgs_host[b].reshape(H, W), non_blocking=True)
            mu = mats_unwarp.to(self.device, torch.float32, non_blocking=True)
            mf = mats_fwd.to(self.device, torch.float32, non_blocking=True)
            heat = self.heatmaps(buf["img"], mu, mf)
            mine = [b for b in range(B) if owner_of_image(b, world) == rank]
            if mine:
                sel = heat[mine] if len(mine) < B else heat
                pts, counts, _ = ops.get_pts_from_heatmap_device(sel.contiguous(), self.conf_thresh, self.nms_dist,
                                                                 self.border_remove, self.top_k, buf["cap"])
                buf["pts_host"][:len(mine)].copy_(pts, non_blocking=True)
                buf["cnt_host"][:len(mine)].copy_(counts, non_blocking=True)
            if sync:
                torch.cuda.current_stream().synchronize()
        self._last = (mine, buf)
        return self.collect() if sync else None

    def collect(self) -> List[Optional[np.ndarray]]:
        mine, buf = self._last
        B = buf["img"].shape[0]
        out: List[Optional[np.ndarray]] = [None] * B
        for j, b in enumerate(mine):
            k = int(buf["cnt_host"][j])
            out[b] = buf["pts_host"][j, :k].numpy().astype(np.float64)
        return out

    def export_image(self, img_host: torch.Tensor, mats_unwarp: torch.Tensor, mats_fwd: torch.Tensor) -> np.ndarray:
        """One image -> float64 [K,3]; with several ranks only rank 0 returns the points (others None)."""
        return self.export_batch([img_host], mats_unwarp, mats_fwd)[0]


def export_detector_homoAdapt_b200(config, output_dir, args=None, loader=None, net: SuperPointNet_b200 = None):
    """Drop-in for ``export.py:196-357``.  Reads the same config keys (model.nms, model.top_k,
    model.detection_threshold, data.homography_adaptation.{num,homographies.params}, data.export_folder,
    pretrained) and writes the same ``.npz`` files.  ``loader`` yields the reference's sample dicts
    (``image_2D`` [1,1,H,W] or ``image``, ``name``, optionally ``homographies`` / ``inv_homographies``); when it is
    None the reference's own ``utils.loader.dataLoader_test`` is imported (running inside the reference repo)."""
    nms_dist = config["model"]["nms"]
    top_k = config["model"]["top_k"]
    conf_thresh = config["model"]["detection_threshold"]
    ha = config["data"]["homography_adaptation"]
    num = ha["num"]
    export_task = config["data"]["export_folder"]
    save_output = os.path.join(str(output_dir), "predictions", export_task)
    os.makedirs(save_output, exist_ok=True)
    if loader is None:
        from utils.loader import dataLoader_test  # reference module (only when run inside its repo)
        loader = dataLoader_test(config, dataset=config["data"]["dataset"], export_task=export_task)["test_loader"]
    if net is None:
        path = config["pretrained"]
        ckpt = torch.load(path, map_location=lambda storage, loc: storage)
        net = SuperPointNet_b200()
        net.load_state_dict(ckpt["model_state_dict"] if path[-4:] == ".tar" else ckpt)
    if config["model"].get("subpixel", {}).get("enable", False):
        from .subpixel import soft_argmax_points
    exporter = HAExporter(net, conf_thresh, nms_dist, top_k)
    rank, world = world_info()
    count = 0
    for sample in loader:
        name = sample["name"][0] if not isinstance(sample["name"], str) else sample["name"]
        target = os.path.join(save_output, "{}.npz".format(name))
        if os.path.exists(target):  # export.py:302-306 resume behaviour
            continue
        img = sample["image_2D"] if "image_2D" in sample else sample["image"]
        img = torch.as_tensor(np.asarray(img), dtype=torch.float32).squeeze()
        if "homographies" in sample:
            mu = torch.as_tensor(np.asarray(sample["homographies"]), dtype=torch.float32).reshape(-1, 3, 3)
            mf = torch.as_tensor(np.asarray(sample["inv_homographies"]), dtype=torch.float32).reshape(-1, 3, 3)
        else:
            h, inv = make_ha_homographies(num, params=ha["homographies"]["params"])
            mu, mf = torch.from_numpy(h), torch.from_numpy(inv)
        pts = exporter.export_image(img, mu, mf)
        if pts is None:
            continue
        if config["model"].get("subpixel", {}).get("enable", False):
            heat = exporter._bufs[next(iter(exporter._bufs))]["heat"][0]
            pts = soft_argmax_points(heat, pts.T, 5).T
        if "scene_name" in sample:
            os.makedirs(os.path.join(save_output, sample["scene_name"][0]), exist_ok=True)
        np.savez_compressed(target, pts=pts)
        count += 1
    return count
