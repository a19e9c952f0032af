"""Post-network kernels in isolation (CUDA events; also the ncu target for nms / descriptor sampling / matcher /
homography sampler): prints ms per call and the achieved algorithmic GB/s of each.

    python tools/bench_post.py [--H 240 --W 320 --B 32 --K 1000]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_superpoint_b200 import ops  # noqa: E402
from pytorch_superpoint_b200.homographies import sample_ha_homographies_device  # noqa: E402


ONCE = False


def timed(fn, n=20):
    if ONCE:  # ncu target: every kernel exactly once
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--H", type=int, default=240)
    ap.add_argument("--W", type=int, default=320)
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--K", type=int, default=1000)
    ap.add_argument("--once", action="store_true", help="run every operator once (the ncu capture target)")
    a = ap.parse_args()
    global ONCE
    ONCE = a.once
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(0)
    H, W, B, K = a.H, a.W, a.B, a.K
    # sparse, realistic heat-maps: smooth blobs + noise floor below the threshold
    heat = torch.rand((B, H, W), generator=g) * 0.01
    idx = torch.randint(0, H * W, (B, 4000), generator=g)
    heat.view(B, -1).scatter_(1, idx, torch.rand((B, 4000), generator=g))
    heat = heat.to(dev)
    ms = timed(lambda: ops.get_pts_from_heatmap_device(heat, 0.015, 4, 4, K))
    print(f"nms+topk  B={B} {H}x{W}: {ms:.4f} ms  -> {B * H * W * 4 / ms / 1e6:.1f} GB/s algorithmic (heat read once)")
    pts, counts, _ = ops.get_pts_from_heatmap_device(heat, 0.015, 4, 4, K)
    print("  mean key-points", float(counts.float().mean()))
    D, Hc, Wc = 256, H // 8, W // 8
    coarse = torch.nn.functional.normalize(torch.randn((B, Hc, Wc, D), generator=g), dim=-1).to(dev)
    ms = timed(lambda: ops.sample_desc_batch_device(coarse, pts, counts))
    kk = float(counts.sum())
    print(f"sample_desc (batched) B={B} K<={K}: {ms:.4f} ms -> {kk * D * 4 * 5 / ms / 1e6:.1f} GB/s (4 taps read + 1 write)")
    dd = ops.sample_desc_batch_device(coarse, pts, counts)
    ms = timed(lambda: ops.nn_match_pairs_device(dd[0::2], dd[1::2], 0.7))
    print(f"nn_match_pairs {B // 2} pairs x {K}x{K}, no host round trip: {ms:.4f} ms -> {B // 2 * 2 * K * K * D / ms / 1e9:.3f} TFLOP/s")
    ms = timed(lambda: ops.sample_desc_device(coarse[0], pts[0], counts[0:1], nhwc=True))
    k0 = int(counts[0])
    print(f"sample_desc (one image, K={k0}): {ms:.4f} ms -> {k0 * D * 4 * 5 / ms / 1e6:.1f} GB/s")
    d1 = torch.nn.functional.normalize(torch.randn((D, K), generator=g), dim=0).numpy()
    d2 = torch.nn.functional.normalize(torch.randn((D, K), generator=g), dim=0).numpy()
    a1, a2 = torch.from_numpy(d1).to(dev), torch.from_numpy(d2).to(dev)
    ms = timed(lambda: ops.nn_match_two_way(a1, a2, 0.7), n=10)
    print(f"nn_match_two_way {K}x{K} (incl. host round trip): {ms:.4f} ms -> {2 * K * K * D / ms / 1e9:.3f} TFLOP/s")
    ms = timed(lambda: sample_ha_homographies_device(100, B, seed=1, device=dev))
    print(f"sample_ha_homographies {B}x100: {ms:.4f} ms")


if __name__ == "__main__":
    main()
