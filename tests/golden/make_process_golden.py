#!/usr/bin/env python
"""Golden vectors for SURVEY 8 f-3 (models/model_utils.py SuperPointNet_process), produced by the UNMODIFIED reference.
Build container only (needs /root/reference):  python tests/golden/make_process_golden.py
Writes tests/golden/process.npz: a seeded heat-map batch, the reference's NMS mask, its roi-pooled patches, its raw
descriptor samples and its batch_extract_features output for a seeded numpy RNG.  The soft-argmax itself lives in
torchgeometry (absent): the residuals fed to batch_extract_features are seeded random numbers."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
refshim.load()
np.int = int  # models/model_utils.py:167 uses the alias numpy 2 removed (harness-side shim, no reference file edited)
from models.model_utils import SuperPointNet_process  # noqa: E402  (the reference)
from utils.losses import extract_patches  # noqa: E402

rng = np.random.RandomState(7)
B, H, W, D = 2, 64, 96, 256
heat = (rng.rand(B, 1, H, W) ** 6 * 0.6).astype(np.float32)      # sparse peaks above the 0.015 threshold
proc = SuperPointNet_process(out_num_points=60, patch_size=5, device="cpu", nms_dist=4, conf_thresh=0.015)
mask = proc.heatmap_to_nms(torch.from_numpy(heat), tensor=True)   # [B,1,H,W] float tensor on cpu
idx = mask.nonzero()
patches = extract_patches(idx, torch.from_numpy(heat), patch_size=5)
desc = torch.from_numpy(rng.randn(B, D, H // 8, W // 8).astype(np.float32))
desc = desc / desc.norm(dim=1, keepdim=True)
pts = torch.from_numpy(np.stack([rng.uniform(0, W - 1, 40), rng.uniform(0, H - 1, 40)], 1).astype(np.float32))
samp = SuperPointNet_process.sample_desc_from_points(desc[:1], pts.clone())
residual = torch.from_numpy(rng.uniform(-0.5, 0.5, (idx.shape[0], 2)).astype(np.float32))
np.random.seed(99)
feat = proc.batch_extract_features(desc, mask, residual)
np.savez_compressed(os.path.join(HERE, "process.npz"), heat=heat, mask=mask.numpy(), idx=idx.numpy(),
                    patches=patches.numpy(), desc=desc.numpy(), pts=pts.numpy(), samp=samp.numpy(),
                    residual=residual.numpy(), pts_int=feat["pts_int"].numpy(), pts_offset=feat["pts_offset"].numpy(),
                    pts_desc=feat["pts_desc"].numpy())
print({k: tuple(v.shape) for k, v in feat.items()}, "N =", idx.shape[0], "patches", tuple(patches.shape))
