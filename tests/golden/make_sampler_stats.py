#!/usr/bin/env python
"""Draws from the UNMODIFIED reference sampler (utils/homographies.py:12-141 sample_homography_np) for the
distributional parity test of the GPU sampler.  Build container only (needs /root/reference):

    python tests/golden/make_sampler_stats.py

Writes tests/golden/sampler_draws.npz: for each parameter set, K sampled 3x3 homographies (float32) in the
[-1,1] frame used by datasets/Coco.py:265-267 (shape (2,2), shift -1).  No reference source is copied."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
refshim.load()  # puts /root/reference on sys.path with the import shims
from utils.homographies import sample_homography_np  # noqa: E402  (the reference)
from pytorch_superpoint_b200.homographies import EXPORT_PARAMS  # noqa: E402

K = 3000
SETS = {
    "export": dict(EXPORT_PARAMS),                                            # configs/magicpoint_coco_export.yaml
    "noartifacts": dict(EXPORT_PARAMS, allow_artifacts=False, patch_ratio=0.5,  # the border-collision branches
                        scaling_amplitude=0.4, perspective_amplitude_x=0.3, perspective_amplitude_y=0.3),
}
out = {}
for name, p in SETS.items():
    np.random.seed(1234)
    out[name] = np.stack([sample_homography_np(np.array([2, 2]), shift=-1, **p) for _ in range(K)]).astype(np.float32)
    print(name, out[name].shape)
np.savez_compressed(os.path.join(HERE, "sampler_draws.npz"), **out)
