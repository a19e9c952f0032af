"""Heat-map error of the fused HA pass against the fp32 CPU oracle on the headline configuration (one 240x320 image,
100 homographies): max / mean abs error and the key-point overlap.   python tools/check_heat_error.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from pytorch_superpoint_b200 import SuperPointNet_b200, _lib, make_ha_homographies
from pytorch_superpoint_b200.export import HAExporter
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import structured_images
H, W, N = 240, 320, 100
img = structured_images(1, H, W, seed=0)[0]
mu, mf = make_ha_homographies(N, seed=0)
sd = O.synthetic_state_dict("gauss2", seed=1)
net = SuperPointNet_b200(); net.load_state_dict(sd)
ref_pts, ref_heat = O.ha_export_one(img, mu, mf, O.canonical_params(sd), return_heat=True)
for _ in (0,):
    exp = HAExporter(net, 0.015, 4, 600)
    pts = exp.export_batch([torch.from_numpy(img)], torch.from_numpy(mu), torch.from_numpy(mf))[0]
    heat = exp.last_heat(1, H, W)[0].cpu().numpy()
    e = np.abs(heat - ref_heat)
    a = {(int(x), int(y)) for x, y, _ in pts}; b = {(int(x), int(y)) for x, y, _ in ref_pts}
    print(f"heat max-abs {e.max():.3e} mean-abs {e.mean():.3e} (heat mean {ref_heat.mean():.4f}), key-point Jaccard vs fp32 oracle {len(a & b) / len(a | b):.4f}")
