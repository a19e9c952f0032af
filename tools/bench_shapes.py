#!/usr/bin/env python
"""One-off throughput of the HA export at the other BASELINE shapes (configs[3] KITTI 384x1248, configs[4] 480x640):
device-resident images, per-image homographies drawn on the GPU, CUDA events.  Not the bench contract (bench.py is)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_superpoint_b200 import SuperPointNet_b200, sample_ha_homographies_device  # noqa: E402
from pytorch_superpoint_b200.export import HAExporter  # noqa: E402
from pytorch_superpoint_b200.synthetic import synthetic_state_dict  # noqa: E402

net = SuperPointNet_b200()
net.load_state_dict(synthetic_state_dict("gauss2", seed=1))
CONV_GFLOP = {(240, 320): 12.161, (384, 1248): 75.884, (480, 640): 48.643}
for (H, W), B in (((240, 320), 32), ((384, 1248), 8), ((480, 640), 8)):
    exp = HAExporter(net, 0.015, 4, 600)
    rng = np.random.RandomState(0)
    imgs = torch.from_numpy(rng.rand(B, H, W).astype(np.float32)).cuda()
    mu, mf = sample_ha_homographies_device(100, B, seed=3)
    for _ in range(2):
        exp.submit_resident(imgs, mu, mf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 6
    e0.record()
    for _ in range(steps):
        t = exp.submit_resident(imgs, mu, mf)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ips = B / ms * 1e3
    print(f"{H}x{W}: {ips:8.1f} images/s  ({ms:.2f} ms per {B}-image step, "
          f"{CONV_GFLOP[(H, W)] * 100 * ips / 1e3:.0f} TFLOP/s of detector convs)")
