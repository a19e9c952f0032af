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
from .homographies import make_ha_homographies, sample_ha_homographies_device  # noqa: F401
from .net import SuperPointNet_b200
from .sharding import balanced_view_plan, owner_of_image, shard_bounds, world_info

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
        self._plan_key, self._plan = None, None

    def _buffers(self, B, H, W):
        key = (B, H, W)
        if key not in self._bufs:
            d1 = self.nms_dist + 1
            cap = self.top_k if self.top_k else -(-H // d1) * -(-W // d1)
            dev = self.device
            slots = [dict(heat=torch.empty((B, H, W), device=dev, dtype=torch.float32),
                          pts_host=torch.empty((B, cap, 3), dtype=torch.float32).pin_memory(),
                          cnt_host=torch.empty((B,), dtype=torch.int32).pin_memory(),
                          done=None, mine=[]) for _ in range(2)]
            self._bufs = {key: dict(img=[torch.empty((B, H, W), device=dev, dtype=torch.float32) for _ in range(2)],
                                    img_free=[None, None], kin=0,
                                    sums=torch.empty((B, 2, H, W), device=dev, dtype=torch.float32),
                                    slots=slots, cap=cap, k=0)}
            self._side = torch.cuda.Stream(device=dev)
            self._copy = torch.cuda.Stream(device=dev)
        return self._bufs[key]

    def heatmaps(self, imgs_dev: torch.Tensor, mats_unwarp: torch.Tensor, mats_fwd: torch.Tensor,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """imgs_dev [B,H,W] cuda; mats [N,3,3] (shared by the batch) or [B,N,3,3] -> aggregated heat [B,H,W]."""
        B, H, W = imgs_dev.shape
        buf = self._buffers(B, H, W)
        rank, world = world_info(self.group)
        per_image = mats_unwarp.dim() == 4
        N = mats_unwarp.shape[-3]
        lo, hi = shard_bounds(N, rank, world)
        sums = buf["sums"]
        if world > 1 and N % world and B >= world and -(-N // world) * B <= 800:
            # uneven shares (100 views over 8 ranks = 13 or 12 per image): rotate them over the images of the batch so
            # that every rank processes the same number of views per step (sharding.balanced_view_plan)
            key = (B, N, rank, world)
            if self._plan_key != key:
                sel, vi, vo, most = balanced_view_plan(B, N, rank, world)
                self._plan_key, self._plan = key, (sel.to(self.device), vi.to(self.device), vo.to(self.device), most)
            sel, vi, vo, most = self._plan
            mats_unwarp = mats_unwarp.to(self.device, torch.float32)
            mats_fwd = mats_fwd.to(self.device, torch.float32)
            if per_image:
                mu, mf = mats_unwarp.reshape(B * N, 3, 3)[sel], mats_fwd.reshape(B * N, 3, 3)[sel]
            else:
                mu, mf = mats_unwarp[sel % N], mats_fwd[sel % N]
            self.net.ha_heatmap_sums_ragged(imgs_dev, mu, mf, vi, vo, most, sums=sums)
        elif hi > lo:
            mu = mats_unwarp[:, lo:hi] if per_image else mats_unwarp[lo:hi]
            mf = mats_fwd[:, lo:hi] if per_image else mats_fwd[lo:hi]
            self.net.ha_heatmap_sums_batch(imgs_dev, mu, mf, sums=sums, chunk=self.chunk)
        else:
            sums.zero_()
        if world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        # export.py:63 -- [B,2,H,W] sums -> [B,H,W] heat: one launch over the batch (sums[b,0]/sums[b,1])
        heat = out if out is not None else buf["slots"][0]["heat"]
        from . import _lib
        _lib.check(_lib.lib().spb_ha_finalize_batch(sums.data_ptr(), heat.data_ptr(), B, H, W,
                                                    torch.cuda.current_stream().cuda_stream), "ha_finalize")
        return heat

    def submit_resident(self, imgs_dev: torch.Tensor, mats_unwarp: torch.Tensor, mats_fwd: torch.Tensor):
        """Device-resident inputs -> enqueue heat-maps on the current stream and NMS/top-k + the D2H of the
        key-points on a side stream, so they overlap the next batch's convolutions.  Returns a ticket for collect()."""
        B, H, W = imgs_dev.shape
        buf = self._buffers(B, H, W)
        rank, world = world_info(self.group)
        slot = buf["slots"][buf["k"] & 1]
        buf["k"] += 1
        main = torch.cuda.current_stream()
        if slot["done"] is not None:
            main.wait_event(slot["done"])  # the side stream still reads this slot's heat-map from two batches ago
        heat = self.heatmaps(imgs_dev, mats_unwarp, mats_fwd, out=slot["heat"])
        ready = torch.cuda.Event()
        ready.record(main)
        mine = [b for b in range(B) if owner_of_image(b, world) == rank]
        slot["mine"] = mine
        side = self._side  # (on the main stream instead the step is 3 % slower: measured)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            if mine:
                sel = heat[mine].contiguous() if len(mine) < B else heat
                pts, counts, _ = ops.get_pts_from_heatmap_device(sel, self.conf_thresh, self.nms_dist,
                                                                 self.border_remove, self.top_k, buf["cap"])
                slot["pts_host"][:len(mine)].copy_(pts, non_blocking=True)
                slot["cnt_host"][:len(mine)].copy_(counts, non_blocking=True)
            slot["done"] = torch.cuda.Event()
            slot["done"].record(side)
        return slot

    def submit(self, imgs_host: Sequence[torch.Tensor], mats_unwarp: torch.Tensor, mats_fwd: torch.Tensor):
        """Host images in (pinned for true async copies): H2D + submit_resident.  Returns a ticket for collect().
        The copies run on a copy stream into one of two device image buffers, so the H2D of batch i+1 overlaps the
        kernels of batch i (the main stream only waits for its own batch's copy event)."""
        collated = torch.is_tensor(imgs_host)  # a DataLoader-style [B,H,W] / [B,1,H,W] batch: ONE copy
        B = imgs_host.shape[0] if collated else len(imgs_host)
        H, W = imgs_host.shape[-2:] if collated else imgs_host[0].shape[-2:]
        buf = self._buffers(B, H, W)
        with torch.cuda.device(self.device):
            j = buf["kin"] & 1
            buf["kin"] += 1
            img = buf["img"][j]
            main = torch.cuda.current_stream()
            if buf["img_free"][j] is not None:
                self._copy.wait_event(buf["img_free"][j])  # the batch that last read this buffer has consumed it
            with torch.cuda.stream(self._copy):
                if collated:
                    img.copy_(imgs_host.reshape(B, H, W), non_blocking=True)
                else:
                    for b in range(B):
                        img[b].copy_(imgs_host[b].reshape(H, W), non_blocking=True)
                mu = mats_unwarp.to(self.device, torch.float32, non_blocking=True)
                mf = mats_fwd.to(self.device, torch.float32, non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(self._copy)
            main.wait_event(copied)
            mu.record_stream(main)
            mf.record_stream(main)
            ticket = self.submit_resident(img, mu, mf)
            buf["img_free"][j] = torch.cuda.Event()
            buf["img_free"][j].record(main)  # (after the heat-map kernels of this batch, the only readers of img)
            return ticket

    def collect(self, ticket) -> List[Optional[np.ndarray]]:
        """Wait for a ticket; float64 [K,3] (x, y, conf) per image this rank owns (b % world == rank), else None."""
        ticket["done"].synchronize()
        B = ticket["heat"].shape[0]
        out: List[Optional[np.ndarray]] = [None] * B
        for j, b in enumerate(ticket["mine"]):
            k = int(ticket["cnt_host"][j])
            out[b] = ticket["pts_host"][j, :k].numpy().astype(np.float64)
        return out

    def export_batch(self, imgs_host: Sequence[torch.Tensor], mats_unwarp: torch.Tensor, mats_fwd: torch.Tensor):
        """Synchronous form: B host images -> list of float64 [K,3] arrays (None for images another rank owns)."""
        return self.collect(self.submit(imgs_host, mats_unwarp, mats_fwd))

    def export_image(self, img_host: torch.Tensor, mats_unwarp: torch.Tensor, mats_fwd: torch.Tensor) -> np.ndarray:
        """One image -> float64 [K,3]; with several ranks only rank 0 returns the points (others None)."""
        return self.export_batch([img_host], mats_unwarp, mats_fwd)[0]

    def last_heat(self, B, H, W) -> torch.Tensor:
        buf = self._buffers(B, H, W)
        return buf["slots"][(buf["k"] - 1) & 1]["heat"]


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
    from .subpixel import soft_argmax_points
    exporter = HAExporter(net, conf_thresh, nms_dist, top_k)
    rank, world = world_info()
    subpixel = bool(config["model"].get("subpixel", {}).get("enable", False))
    batch_images = max(1, int(config["model"].get("b200_batch_images", 8)))  # images per GPU step (ours, optional)
    seed0 = int(config.get("seed", 0)) * 1000003  # every rank draws the same matrices for image k (shards must agree)
    from concurrent.futures import ThreadPoolExecutor
    writers = ThreadPoolExecutor(max_workers=2)  # np.savez_compressed (zlib) off the submitting thread
    writes, pending, batch = [], [], []
    count = 0
    seen = 0

    def collect_oldest():
        nonlocal count
        ticket, metas = pending.pop(0)
        for b, pts in enumerate(exporter.collect(ticket)):
            if pts is None:  # another rank owns this image
                continue
            target, scene = metas[b]
            if subpixel:
                pts = soft_argmax_points(ticket["heat"][b], pts.T, 5).T
            if scene is not None:
                os.makedirs(os.path.join(save_output, scene), exist_ok=True)
            writes.append(writers.submit(np.savez_compressed, target, pts=pts))
            count += 1

    def flush():
        """Submit the accumulated images as ONE batch (H2D on the copy stream, kernels, NMS on the side stream) and
        collect the batch before it: the host work of batch i overlaps the GPU work of batch i+1."""
        nonlocal seen
        if not batch:
            return
        imgs = [m[0] for m in batch]
        if all(m[1] is not None for m in batch):
            mu, mf = torch.stack([m[1] for m in batch]), torch.stack([m[2] for m in batch])
        else:  # drawn on the GPU, one launch per batch (the host sampler costs ~2 s per 100, like the reference's)
            mu, mf = sample_ha_homographies_device(num, len(batch), seed=seed0 + seen,
                                                   params=ha["homographies"]["params"])
        seen += len(batch)
        pending.append((exporter.submit(imgs, mu, mf), [(m[3], m[4]) for m in batch]))
        batch.clear()
        while len(pending) > 1:
            collect_oldest()

    for sample in loader:
        name = sample["name"][0] if not isinstance(sample["name"], str) else sample["name"]
        target = os.path.join(save_output, "{}.npz".format(name))
        if os.path.exists(target):  # export.py:302-306 resume behaviour
            continue
        img = sample["image_2D"] if "image_2D" in sample else sample["image"]
        img = torch.as_tensor(np.asarray(img), dtype=torch.float32).squeeze()
        if img.dim() != 2:
            raise ValueError(f"expected one gray image per sample, got {tuple(img.shape)}")
        if batch and batch[0][0].shape != img.shape:
            flush()
        mu = mf = None
        if "homographies" in sample:
            mu = torch.as_tensor(np.asarray(sample["homographies"]), dtype=torch.float32).reshape(-1, 3, 3)
            mf = torch.as_tensor(np.asarray(sample["inv_homographies"]), dtype=torch.float32).reshape(-1, 3, 3)
        scene = sample["scene_name"][0] if "scene_name" in sample else None
        batch.append((img.pin_memory() if not img.is_pinned() else img, mu, mf, target, scene))
        if len(batch) == batch_images:
            flush()
    flush()
    while pending:
        collect_oldest()
    for w in writes:
        w.result()  # re-raise I/O errors like the reference's synchronous savez would
    writers.shutdown()
    if world > 1:  # count over all ranks (each image is written by exactly one)
        t = torch.tensor([count], device=exporter.device)
        dist.all_reduce(t, group=exporter.group)
        count = int(t.item())
    return count
