#!/usr/bin/env python
"""Micro-benchmark of one network layer through spb_net_layer (CUDA events), for ncu captures and tuning.
    python tools/bench_layer.py --layer 1 --views 20 --h 240 --w 320 [--halo 0|1] [--iters 5]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_superpoint_b200 import SuperPointNet_b200, _lib  # noqa: E402
from pytorch_superpoint_b200.synthetic import synthetic_state_dict  # noqa: E402

SHAPES = [(1, 64, 3), (64, 64, 3), (64, 64, 3), (64, 64, 3), (64, 128, 3), (128, 128, 3), (128, 128, 3),
          (128, 128, 3), (128, 256, 3), (256, 65, 1), (128, 256, 3), (256, 256, 1)]
POOL = [0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0]

ap = argparse.ArgumentParser()
ap.add_argument("--layer", type=int, default=1)
ap.add_argument("--views", type=int, default=20)
ap.add_argument("--h", type=int, default=240)
ap.add_argument("--w", type=int, default=320)
ap.add_argument("--halo", type=int, default=1)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--impl", default="tcgen05")
ap.add_argument("--forward", action="store_true", help="run the whole detector forward instead of one layer")
ap.add_argument("--fuse1", type=int, default=1)
a = ap.parse_args()
L = _lib.lib()
net = SuperPointNet_b200(impl=a.impl)
net.load_state_dict(synthetic_state_dict("gauss2", 1))
h = net.handle(torch.device("cuda", 0))
L.spb_conv_tc_set_halo(a.halo)
L.spb_conv_tc_set_fuse1(a.fuse1)
if a.forward:
    x = torch.rand(a.views, 1, a.h, a.w, device="cuda")
    for _ in range(2):
        net.forward_nhwc(x, want_desc=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        net.forward_nhwc(x, want_desc=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print(f"forward views {a.views} {a.h}x{a.w} fuse1={a.fuse1}: {ms:.4f} ms  {12.161e-3 * a.views / ms * (a.h * a.w) / 76800:.1f} TFLOP/s")
    sys.exit(0)
cin, cout, ks = SHAPES[a.layer]
if a.layer == 0:
    x = torch.rand(a.views, a.h, a.w, device="cuda")
else:
    x = (torch.randn(a.views, a.h, a.w, cin, device="cuda") * 0.5).to(torch.bfloat16)
oh, ow = (a.h // 2, a.w // 2) if POOL[a.layer] else (a.h, a.w)
out = torch.empty(a.views, oh, ow, cout, device="cuda", dtype=torch.float32 if a.layer in (9, 11) else torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(L.spb_net_layer(h, a.layer, x.data_ptr(), out.data_ptr(), a.views, a.h, a.w, st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    _lib.check(L.spb_net_layer(h, a.layer, x.data_ptr(), out.data_ptr(), a.views, a.h, a.w, st))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
flops = 2.0 * a.views * a.h * a.w * cin * cout * ks * ks
tiles = a.views * ((a.h + 15) // 16) * ((a.w + 7) // 8)
print(f"layer {a.layer} views {a.views} {a.h}x{a.w} halo={a.halo}: {ms:.4f} ms  {flops / ms / 1e9:.1f} TFLOP/s  "
      f"{ms * 1e-3 * 1.965e9 * 148 / tiles:.0f} cyc/tile@1.965GHz")
