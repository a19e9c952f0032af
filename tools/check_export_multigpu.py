#!/usr/bin/env python
"""Multi-GPU check of the .npz entry point (run under torchrun, NCCL): every rank walks the same loader, the
homographies are sharded over the ranks, each image is written by exactly one rank, and the key-points agree with a
single-GPU export of the same images.   torchrun --nproc-per-node 2 tools/check_export_multigpu.py <outdir>"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_superpoint_b200 import SuperPointNet_b200  # noqa: E402
from pytorch_superpoint_b200.export import export_detector_homoAdapt_b200  # noqa: E402
from pytorch_superpoint_b200.homographies import sample_ha_homographies_device  # noqa: E402
from pytorch_superpoint_b200.synthetic import synthetic_state_dict  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = sys.argv[1]
net = SuperPointNet_b200()
net.load_state_dict(synthetic_state_dict("gauss2", seed=1))
rng = np.random.RandomState(0)
N, H, W = 20, 120, 160
samples = []
for i in range(11):
    img = np.full((H, W), 0.3, np.float32)
    for _ in range(10):
        x0, y0 = rng.randint(0, W - 12), rng.randint(0, H - 12)
        img[y0:y0 + rng.randint(6, 40), x0:x0 + rng.randint(6, 50)] = rng.uniform(0, 1)
    samples.append({"image_2D": torch.from_numpy(img)[None, None], "name": [f"g{i}"]})
params = {"translation": True, "rotation": True, "scaling": True, "perspective": True, "scaling_amplitude": 0.2,
          "perspective_amplitude_x": 0.2, "perspective_amplitude_y": 0.2, "allow_artifacts": True, "patch_ratio": 0.85}
cfg = {"model": {"nms": 4, "top_k": 300, "detection_threshold": 0.015, "subpixel": {"enable": False},
                 "b200_batch_images": 4},
       "data": {"export_folder": "train", "dataset": "Coco",
                "homography_adaptation": {"num": N, "homographies": {"params": params}}}}
n = export_detector_homoAdapt_b200(cfg, out, loader=samples, net=net)
dist.barrier()
if rank == 0:
    assert n == len(samples), n
    # the same images and the same (counter-based) homographies on one GPU
    worst = 1.0
    seen = 0
    for b0 in range(0, len(samples), 4):
        chunk = samples[b0:b0 + 4]
        mu, mf = sample_ha_homographies_device(N, len(chunk), seed=seen, params=params)
        seen += len(chunk)
        for j, s in enumerate(chunk):
            sums = net.ha_heatmap_sums(s["image_2D"][0, 0].cuda(), mu[j], mf[j])
            from pytorch_superpoint_b200.ops import ha_finalize, getPtsFromHeatmap
            ref = getPtsFromHeatmap(ha_finalize(sums), 0.015, 4).T[:300]
            got = np.load(os.path.join(out, "predictions", "train", s["name"][0] + ".npz"))["pts"]
            a = {(int(x), int(y)) for x, y, _ in got}
            b = {(int(x), int(y)) for x, y, _ in ref}
            worst = min(worst, len(a & b) / max(len(a | b), 1))
    print(f"multi-GPU export ok: {n} files from {world} ranks, worst key-point Jaccard vs single-GPU {worst:.3f}")
    assert worst > 0.9
dist.barrier()
dist.destroy_process_group()
