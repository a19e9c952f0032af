"""A/B of first-block variants on the whole detector forward: spb_conv_tc_set_fuse1(1) vs (mode; 0 = two kernels), outputs
compared and both timed with CUDA events.   python tools/ab_fuse1.py [mode=0] [views=800] [H=240] [W=320]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_superpoint_b200 import SuperPointNet_b200, _lib
from pytorch_superpoint_b200.synthetic import synthetic_state_dict
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
views = int(sys.argv[2]) if len(sys.argv) > 2 else 800
H = int(sys.argv[3]) if len(sys.argv) > 3 else 240
W = int(sys.argv[4]) if len(sys.argv) > 4 else 320
L = _lib.lib()
net = SuperPointNet_b200(); net.load_state_dict(synthetic_state_dict("gauss2", 1))
x = torch.rand(views, 1, H, W, device="cuda")
out = {}
for m in (1, mode):
    L.spb_conv_tc_set_fuse1(m)
    for _ in range(2):
        s, _d = net.forward_nhwc(x, want_desc=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        s, _d = net.forward_nhwc(x, want_desc=False)
    e1.record(); torch.cuda.synchronize()
    out[m] = s.clone()
    print(f"fuse1={m}: {e0.elapsed_time(e1) / 4:.4f} ms per {views} views of {H}x{W}")
L.spb_conv_tc_set_fuse1(-1)
d = (out[1] - out[mode]).abs()
print(f"max |diff| {d.max().item():.3e}  mean {d.mean().item():.3e}  scale {out[1].abs().max().item():.3e}  finite {bool(torch.isfinite(out[mode]).all())}")
