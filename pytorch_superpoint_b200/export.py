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
from .sharding import PeerExchange, PeerUnavailable, owned_images, plan_batch, same_host, sharded_heatmap_sums, world_info

__all__ = ["HAExporter", "export_detector_homoAdapt_b200"]


class HAExporter:
    """One process per GPU.  With ``torch.distributed`` initialised (NCCL) the homographies of every image are
    sharded over the ranks (``sharding.plan_batch``: rotated view blocks) and combined by ONE exchange per batch that
    runs, with the division / NMS / top-k of batch i, under the warps and convolutions of batch i+1.
    ``exchange``: ``'peer'`` -- the owner of an image adds the partial planes straight out of its peers' memory over
    NVLink inside the kernel that divides (``sharding.PeerExchange``, CUDA IPC, one box); ``'nccl'`` -- an asynchronous
    ``reduce_scatter`` over the image dimension (``sharding.exchange_sums``); ``'auto'`` (default; env
    ``SPB200_EXCHANGE`` overrides): peer memory when every rank runs on this host, NCCL otherwise.  Without
    ``torch.distributed`` everything runs locally.  ``subpixel_patch`` > 0 adds the soft-argmax refinement of
    export.py:314-319 on the device."""

    def __init__(self, net: SuperPointNet_b200, conf_thresh=0.015, nms_dist=4, top_k=600, border_remove=4,
                 device: Optional[torch.device] = None, group=None, chunk: int = 0, subpixel_patch: int = 0,
                 sharded: bool = True, exchange: str = "auto"):
        if not torch.cuda.is_available():
            raise RuntimeError("HAExporter needs a CUDA device (sm_100a); there is no CPU fallback")
        self.net = net
        self.conf_thresh, self.nms_dist, self.top_k, self.border_remove = conf_thresh, nms_dist, top_k, border_remove
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.group = group
        self.chunk = chunk
        self.subpixel_patch = int(subpixel_patch)
        self.sharded = bool(sharded)  # False: ignore torch.distributed (a purely local exporter, e.g. a cross-check)
        exchange = os.environ.get("SPB200_EXCHANGE", exchange)
        if exchange not in ("auto", "peer", "nccl"):
            raise ValueError(f"exchange must be 'auto', 'peer' or 'nccl', got {exchange!r}")
        self.exchange = exchange  # resolved to 'peer' / 'nccl' when the first multi-rank buffers are built (collective)
        self._bufs = {}
        self._plans = {}
        self._side = self._copy = None

    def _world(self):
        return world_info(self.group) if self.sharded else (0, 1)

    # ---- buffers: everything a batch shape needs is allocated once -----------------------------------------
    def _buffers(self, B, H, W):
        key = (B, H, W)
        if key not in self._bufs:
            d1 = self.nms_dist + 1
            cap = self.top_k if self.top_k else -(-H // d1) * -(-W // d1)
            dev = self.device
            rank, world = self._world()
            lo, hi = owned_images(B, rank, world)
            nown = max(hi - lo, 1)
            f32 = dict(device=dev, dtype=torch.float32)
            for old in self._bufs.values():  # one batch shape at a time: release the previous shape's peer mappings
                if old.get("peer") is not None:
                    old["peer"].close()
            peer = None
            if world > 1:
                auto = self.exchange == "auto"
                if auto:
                    self.exchange = "peer" if same_host(self.group) else "nccl"
                if self.exchange == "peer":
                    try:
                        peer = PeerExchange(B, H, W, dev, self.group)  # collective: all ranks build the same shapes
                    except PeerUnavailable:
                        if not auto:
                            raise
                        self.exchange = "nccl"  # (every rank takes this branch: the constructor agrees on it)
            slots = [dict(sums=peer.sums[i] if peer is not None else torch.empty((B, 2, H, W), **f32),  # partial planes
                          owned=None if peer is not None else torch.empty((nown, 2, H, W), **f32),  # reduce-scatter out
                          heat=torch.empty((nown, H, W), **f32),
                          dxdy=torch.zeros((nown, cap, 2), **f32),
                          pts_host=torch.empty((nown, cap, 3), dtype=torch.float32).pin_memory(),
                          dxdy_host=torch.empty((nown, cap, 2), dtype=torch.float32).pin_memory(),
                          cnt_host=torch.empty((nown,), dtype=torch.int32).pin_memory(),
                          done=None, own=(lo, hi), B=B) for i in range(2)]
            self._bufs = {key: dict(img=[torch.empty((B, H, W), **f32) for _ in range(2)], img_free=[None, None], kin=0,
                                    slots=slots, cap=cap, k=0, peer=peer)}
            if self._side is None:
                self._side = torch.cuda.Stream(device=dev)
                self._copy = torch.cuda.Stream(device=dev)
        return self._bufs[key]

    def plan(self, B, N, H, W, shared: bool):
        """The cached ``sharding.ShardPlan`` of this rank for a batch shape, its pass index arrays on the device."""
        rank, world = self._world()
        cap = self.chunk if self.chunk and self.chunk > 0 else self.net.ha_pass_views(H, W, self.device)
        key = (B, N, H, W, bool(shared), rank, world, cap)
        if key not in self._plans:
            plan = plan_batch(B, N, rank, world, cap)
            if getattr(self.net, "bn_mode", "eval") == "batch" and (world > 1 or any(p.accumulate for p in plan.passes)):
                # batch statistics = the statistics of ALL views of an image (the reference's batch): they cannot be
                # split over ranks or passes without changing the result
                raise NotImplementedError(
                    "bn_mode='batch' needs all homographies of an image in one pass on one GPU "
                    f"(got world={world}, {N} views, {cap} views per pass): use HAExporter(sharded=False) and give "
                    "every rank its own images")
            dev_passes = []
            for p in plan.passes:
                vm = p.view_mat % N if shared else p.view_mat
                dev_passes.append((p, p.view_image.to(self.device), vm.to(self.device), p.view_offsets.to(self.device)))
            self._plans[key] = (plan, dev_passes)
        return self._plans[key]

    def _partial_fn(self, imgs_dev, mats_unwarp, mats_fwd, plan_dev):
        """The ``partial_fn`` of ``sharding.sharded_heatmap_sums``: one pass of this rank's views (warp -> detector
        network -> aggregate) written / added into the partial (sum heat*mask, sum mask) planes of its images."""
        mu = mats_unwarp.to(self.device, torch.float32).reshape(-1, 3, 3)
        mf = mats_fwd.to(self.device, torch.float32).reshape(-1, 3, 3)
        by_pass = {id(p): t for p, *t in plan_dev}

        def partial_fn(p, out):
            vi, vm, vo = by_pass[id(p)]
            self.net.ha_heatmap_sums_ragged(imgs_dev[p.b0:p.b0 + p.nb], mu, mf, vi, vo, p.max_views,
                                            sums=out[p.b0:p.b0 + p.nb], accumulate=p.accumulate, view_mat=vm)
        return partial_fn

    def submit_resident(self, imgs_dev: torch.Tensor, mats_unwarp: torch.Tensor, mats_fwd: torch.Tensor):
        """Device-resident inputs (imgs [B,H,W]; mats [N,3,3] shared by the batch or [B,N,3,3]) -> enqueue the
        partial sums on the current stream, the exchange on NCCL's stream and the division + NMS / top-k + the D2H of
        the key-points on a side stream, so they overlap the next batch's kernels.  Returns a ticket for collect()."""
        B, H, W = imgs_dev.shape
        buf = self._buffers(B, H, W)
        shared = mats_unwarp.dim() == 3
        N = mats_unwarp.shape[-3]
        plan, plan_dev = self.plan(B, N, H, W, shared)
        step = buf["k"]
        slot = buf["slots"][step & 1]
        buf["k"] += 1
        peer = buf["peer"]
        main = torch.cuda.current_stream()
        if slot["done"] is not None:
            main.wait_event(slot["done"])  # the exchange / side stream of two batches ago still reads this slot
        partial_fn = self._partial_fn(imgs_dev, mats_unwarp, mats_fwd, plan_dev)
        if peer is not None:
            peer.wait_reusable(step, main.cuda_stream)  # every owner has read this slot's planes of two batches ago
            owned, work, _ = sharded_heatmap_sums(partial_fn, B, N, slot["sums"], group=self.group, plan=plan,
                                                  exchange=lambda sums, pl: (None, None))
            peer.publish(step, main.cuda_stream)        # this rank's partial planes of the batch are complete
        else:
            owned, work, _ = sharded_heatmap_sums(partial_fn, B, N, slot["sums"], group=self.group, plan=plan,
                                                  out=slot["owned"], async_op=True)
        ready = torch.cuda.Event()
        ready.record(main)
        lo, hi = plan.own
        side = self._side
        side.wait_event(ready)
        from . import _lib
        L = _lib.lib()
        with torch.cuda.stream(side):
            if work is not None:
                work.wait()  # the side stream (not the main one) waits for the collective
            if peer is not None:  # reduce over the ranks + export.py:63 in one kernel, straight from peer memory
                peer.finalize(step, lo, hi, slot["heat"], side.cuda_stream)
            if hi > lo:
                heat = slot["heat"][:hi - lo]
                if peer is None:
                    _lib.check(L.spb_ha_finalize_batch(owned.data_ptr(), heat.data_ptr(), hi - lo, H, W,
                                                       side.cuda_stream), "ha_finalize")  # export.py:63
                pts, counts, _ = ops.get_pts_from_heatmap_device(heat, self.conf_thresh, self.nms_dist,
                                                                 self.border_remove, self.top_k, buf["cap"],
                                                                 background=True)  # beside the next batch's kernels
                if self.subpixel_patch:  # model_wrap.py:200-234: all owned images in ONE launch, no host round trip
                    _lib.check(L.spb_soft_argmax_points_batch(heat.data_ptr(), hi - lo, H, W, pts.data_ptr(),
                                                              counts.data_ptr(), buf["cap"], self.subpixel_patch,
                                                              slot["dxdy"].data_ptr(), side.cuda_stream), "soft_argmax")
                    slot["dxdy_host"][:hi - lo].copy_(slot["dxdy"][:hi - lo], non_blocking=True)
                slot["pts_host"][:hi - lo].copy_(pts, non_blocking=True)
                slot["cnt_host"][:hi - lo].copy_(counts, non_blocking=True)
                slot["pts_dev"], slot["cnt_dev"] = pts, counts  # device-resident consumers (descriptors, matching)
            slot["done"] = torch.cuda.Event()
            slot["done"].record(side)
        return slot

    def heatmaps(self, imgs_dev: torch.Tensor, mats_unwarp: torch.Tensor, mats_fwd: torch.Tensor) -> torch.Tensor:
        """Aggregated heat-maps [own_hi - own_lo, H, W] of the images this rank owns (all B on a single GPU)."""
        slot = self.submit_resident(imgs_dev, mats_unwarp, mats_fwd)
        torch.cuda.current_stream().wait_event(slot["done"])
        lo, hi = slot["own"]
        return slot["heat"][:hi - lo]

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
        """Wait for a ticket; float64 [K,3] (x, y, conf) per image this rank owns (``sharding.owned_images``), None
        for the images other ranks finalise."""
        ticket["done"].synchronize()
        lo, hi = ticket["own"]
        out: List[Optional[np.ndarray]] = [None] * ticket["B"]
        for j in range(hi - lo):
            k = int(ticket["cnt_host"][j])
            pts = ticket["pts_host"][j, :k].numpy().astype(np.float64)
            if self.subpixel_patch:  # model_wrap.py:229-231: refined = point + soft-argmax offset - patch // 2
                pts[:, :2] += ticket["dxdy_host"][j, :k].numpy().astype(np.float64) - self.subpixel_patch // 2
            out[lo + j] = pts
        return out

    def export_batch(self, imgs_host: Sequence[torch.Tensor], mats_unwarp: torch.Tensor, mats_fwd: torch.Tensor):
        """Synchronous form: B host images -> list of float64 [K,3] arrays (None for images another rank owns)."""
        return self.collect(self.submit(imgs_host, mats_unwarp, mats_fwd))

    def export_image(self, img_host: torch.Tensor, mats_unwarp: torch.Tensor, mats_fwd: torch.Tensor) -> np.ndarray:
        """One image -> float64 [K,3]; with several ranks only rank 0 returns the points (others None)."""
        return self.export_batch([img_host], mats_unwarp, mats_fwd)[0]

    def close(self):
        """Collective when the peer exchange is active: unmap / free the shared buffers (all ranks call it)."""
        for buf in self._bufs.values():
            if buf.get("peer") is not None:
                buf["peer"].close()
                buf["peer"] = None
        self._bufs = {}

    def last_heat(self, B, H, W) -> torch.Tensor:
        """Heat-maps [owned,H,W] of the most recent batch of this shape (valid after its collect())."""
        buf = self._buffers(B, H, W)
        slot = buf["slots"][(buf["k"] - 1) & 1]
        lo, hi = slot["own"]
        return slot["heat"][:hi - lo]


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
        # model.b200_bn_mode: 'eval' (default, the product path) or 'batch' (the as-written train-mode statistics)
        net = SuperPointNet_b200(bn_mode=config["model"].get("b200_bn_mode", "eval"))
        net.load_state_dict(ckpt["model_state_dict"] if path[-4:] == ".tar" else ckpt)
    sub = config["model"].get("subpixel", {}) or {}
    subpixel = bool(sub.get("enable", False))
    # export.py:314-319 refines with patch_size = config['model']['subpixel']['patch_size'] (5 in the shipped YAMLs)
    exporter = HAExporter(net, conf_thresh, nms_dist, top_k, subpixel_patch=int(sub.get("patch_size", 5)) if subpixel else 0)
    rank, world = world_info()
    # datasets/Coco.py:281-283 erodes the valid masks of the views with this margin (0 in every shipped export YAML);
    # the fused path keeps the masks as bits inside one kernel chain and has no erosion stage -> refuse, do not ignore
    margin = (((config["data"].get("augmentation") or {}).get("homographic") or {}).get("valid_border_margin", 0)) or 0
    if margin:
        raise NotImplementedError("data.augmentation.homographic.valid_border_margin != 0 is not supported by the fused "
                                  "export path (use ops.compute_valid_mask(erosion_radius=...) + ops.combine_heatmap)")
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
        have = [m[1] is not None for m in batch]
        if all(have):
            mu, mf = torch.stack([m[1] for m in batch]), torch.stack([m[2] for m in batch])
        elif any(have):
            raise ValueError("a batch mixes samples with and without 'homographies': refusing to discard the provided "
                             "matrices (flush the loader per kind, or provide them for every sample)")
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
        if batch and (batch[0][0].shape != img.shape or (batch[0][1] is not None) != ("homographies" in sample)):
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
    exporter.close()  # (collective with the peer-memory exchange: unmap / free the shared buffers)
    return count
