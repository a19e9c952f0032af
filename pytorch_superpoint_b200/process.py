"""``SuperPointNet_process_b200`` -- the batched, tensor-in / tensor-out post-processing that deepFEPE-style callers
use (``models/model_utils.py:9-207 SuperPointNet_process`` + ``SuperPointNet_gauss2.process_output``, 72-99):
heat-map -> NMS mask -> soft-argmax residuals -> per image (points, residuals, descriptors at the refined points)
cropped or padded to ``out_num_points``.  Same method names, arguments and return conventions; the work is done by
libspb200 kernels (exact-greedy NMS, roi-pooled soft-argmax, bilinear descriptor sampling)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops


def crop_or_pad_choice(in_num_points, out_num_points, shuffle=False):
    """utils/utils.py:904-926 -- same numpy RNG calls in the same order, so a seeded run selects the same rows."""
    choice = np.random.permutation(in_num_points) if shuffle else np.arange(in_num_points)
    assert out_num_points > 0, "out_num_points = %d must be positive int!" % out_num_points
    if in_num_points >= out_num_points:
        return choice[:out_num_points]
    pad = np.random.choice(choice, out_num_points - in_num_points, replace=True)
    return np.concatenate([choice, pad])


class SuperPointNet_process_b200(object):
    def __init__(self, **config):
        self.out_num_points = config.get("out_num_points", 500)
        self.patch_size = config.get("patch_size", 5)
        self.device = config.get("device", "cuda:0")
        self.nms_dist = config.get("nms_dist", 4)
        self.conf_thresh = config.get("conf_thresh", 0.015)
        self.heatmap = None
        self.heatmap_nms_batch = None
        if self.patch_size != 5:
            raise ValueError("pred_soft_argmax is built for the reference's patch_size = 5")

    def _dev(self):
        if not torch.cuda.is_available():
            raise RuntimeError("SuperPointNet_process_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        d = torch.device(self.device)
        return d if d.type == "cuda" else torch.device("cuda", torch.cuda.current_device())

    # models/model_utils.py:24-60
    def pred_soft_argmax(self, labels_2D, heatmap):
        """labels_2D, heatmap [B,1,H,W] -> {'pred': [N,2] (dx, dy) residuals of the N labelled pixels in
        ``labels_2D.nonzero()`` order, 'patches': [N,1,5,5]}."""
        dev = self._dev()
        heat = torch.as_tensor(heatmap).to(dev, torch.float32).contiguous()
        B, _, H, W = heat.shape
        idx = torch.as_tensor(labels_2D).to(dev).nonzero().contiguous()  # [N,4] (batch, 0, y, x), row-major
        N = idx.shape[0]
        dxdy = torch.empty((N, 2), device=dev, dtype=torch.float32)
        patches = torch.empty((N, 1, 5, 5), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().spb_roi_soft_argmax(heat.data_ptr(), B, H, W, idx.data_ptr(), N, dxdy.data_ptr(),
                                                      patches.data_ptr(), torch.cuda.current_stream().cuda_stream),
                       "roi_soft_argmax")
        return {"pred": dxdy - self.patch_size // 2, "patches": patches}

    # models/model_utils.py:62-90
    @staticmethod
    def sample_desc_from_points(coarse_desc, pts, cell_size=8):
        """coarse_desc [1,D,Hc,Wc], pts [N,2] (x, y) -> [1,N,D] raw bilinear samples (no re-normalisation)."""
        D = coarse_desc.shape[1]
        if pts.shape[0] == 0:  # (the reference tests shape[1]; an empty set gives ones((1,1,D)) there as well)
            return torch.ones((1, 1, D), device=coarse_desc.device)
        c = coarse_desc.to(torch.float32).contiguous()
        p = pts.to(c.device, torch.float32).contiguous()
        out = torch.empty((p.shape[0], D), device=c.device, dtype=torch.float32)
        with torch.cuda.device(c.device):
            _lib.check(_lib.lib().spb_sample_desc_from_points_ex(c.data_ptr(), 0, p.data_ptr(), 2, None, p.shape[0],
                                                                 c.shape[2], c.shape[3], D, cell_size, 0,
                                                                 out.data_ptr(),
                                                                 torch.cuda.current_stream().cuda_stream),
                       "sample_desc_from_points_ex")
        return out.unsqueeze(0)

    # models/model_utils.py:93-105
    @staticmethod
    def ext_from_points(labels_res, points):
        labels_res = labels_res.transpose(1, 2).transpose(2, 3).unsqueeze(1)
        return labels_res[points[:, 0], points[:, 1], points[:, 2], points[:, 3], :]

    # models/model_utils.py:125-153 (+ 156-175 heatmap_nms)
    def heatmap_to_nms(self, heatmap, tensor=False, boxnms=False):
        """heatmap [B,1,H,W] -> 0/1 mask of the NMS survivors [B,1,H,W] (numpy, or a cuda tensor with tensor=True)."""
        if boxnms:
            raise NotImplementedError("box_nms (torchvision IoU NMS) is not on the path")
        dev = self._dev()
        heat = torch.as_tensor(heatmap).to(dev, torch.float32)
        B, _, H, W = heat.shape
        pts, counts, _ = ops.get_pts_from_heatmap_device(heat[:, 0].contiguous(), self.conf_thresh, self.nms_dist, 4)
        mask = torch.zeros((B, 1, H, W), device=dev, dtype=torch.float32)
        k = torch.arange(pts.shape[1], device=dev)[None, :] < counts[:, None]
        b = torch.arange(B, device=dev)[:, None].expand_as(k)[k]
        mask[b, 0, pts[..., 1][k].long(), pts[..., 0][k].long()] = 1.0
        self.heatmap = heatmap
        self.heatmap_nms_batch = mask if tensor else mask.cpu().numpy()
        return self.heatmap_nms_batch

    # models/model_utils.py:178-207
    def batch_extract_features(self, desc, heatmap_nms_batch, residual):
        """desc [B,D,Hc,Wc], mask [B,1,H,W], residual [N,2] in mask.nonzero() order -> dict of
        'pts_int' [B,M,2] (x, y), 'pts_offset' [B,M,2], 'pts_desc' [B,M,D] with M = out_num_points."""
        mask = torch.as_tensor(heatmap_nms_batch)
        B = mask.shape[0]
        idx = mask.nonzero()                                   # [N,4] (b, 0, y, x), sorted by image then row-major
        per_image = torch.bincount(idx[:, 0], minlength=B).tolist()
        xy = idx[:, [3, 2]].float()                            # integer key-points as (x, y)
        pts_int, pts_offset, pts_desc = [], [], []
        start = 0
        for b, n_b in enumerate(per_image):
            xy_b, res_b = xy[start:start + n_b], residual[start:start + n_b]
            start += n_b
            d_b = self.sample_desc_from_points(desc[b:b + 1], xy_b + res_b)[0]      # descriptors at the refined points
            pick = torch.as_tensor(crop_or_pad_choice(n_b, self.out_num_points, shuffle=True), device=xy_b.device)
            pts_int.append(xy_b[pick])
            pts_offset.append(res_b[pick])
            pts_desc.append(d_b[pick])
        return {"pts_int": torch.stack(pts_int, dim=0), "pts_offset": torch.stack(pts_offset, dim=0),
                "pts_desc": torch.stack(pts_desc, dim=0)}
