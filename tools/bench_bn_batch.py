"""Timing of the opt-in train-mode BatchNorm path (SuperPointNet_b200(bn_mode='batch'), csrc/bn.cu) next to the product
path on the export's batch (all N views of one 240x320 image):   python tools/bench_bn_batch.py [N=100] [H=240] [W=320]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_superpoint_b200 import SuperPointNet_b200, make_ha_homographies
from pytorch_superpoint_b200.ops import ha_finalize
from pytorch_superpoint_b200.synthetic import synthetic_state_dict
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
H = int(sys.argv[2]) if len(sys.argv) > 2 else 240
W = int(sys.argv[3]) if len(sys.argv) > 3 else 320
sd = synthetic_state_dict("gauss2", 1)
img = torch.rand(H, W, device="cuda")
mu, mf = (torch.from_numpy(m).cuda() for m in make_ha_homographies(N, seed=1))
heat = {}
for mode in ("eval", "batch"):
    net = SuperPointNet_b200(bn_mode=mode); net.load_state_dict(sd)
    for _ in range(2):
        sums = net.ha_heatmap_sums(img, mu, mf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        sums = net.ha_heatmap_sums(img, mu, mf)
    e1.record(); torch.cuda.synchronize()
    heat[mode] = ha_finalize(sums)
    print(f"bn_mode={mode}: {e0.elapsed_time(e1) / 3:.2f} ms per image ({N} views of {H}x{W}), peak memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
print(f"heat-map difference between the two BN modes on these synthetic weights: max {float((heat['eval'] - heat['batch']).abs().max()):.3e}")
