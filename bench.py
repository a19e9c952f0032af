#!/usr/bin/env python
"""bench.py -- HA-export throughput (BASELINE.json metric) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: IMAGES 240x320 images, each
homography-adapted with 100 homographies (configs[1]: magicpoint_coco_export.yaml shape): image warp x100 ->
detector network -> fused softmax/d2s/un-warp/aggregate -> (NCCL all-reduce when the homographies are sharded
over ranks) -> divide -> threshold + greedy NMS + top-600.  Rank 0 prints ONE JSON line.

  value      images/s with inputs resident in HBM (CUDA events, barrier + synchronize both sides, max over ranks)
  e2e        same metric through HAExporter.export_batch with pinned HOST images in and HOST key-points out
  roofline   the dominant kernel (conv1_fused_kernel: conv1a+conv1b+pool, 44 % of the conv FLOPs) against the
             measured bf16 tensor peak
  cpu_baseline  the CPU oracle port of the reference path on this box's host cores (rank 0, N=1 only)
  --impl reference   times that CPU arm as the main line (bounded sample per step)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, NVIEWS, TOPK, THR, NMS = 240, 320, 100, 600, 0.015, 4
METRIC = "ha_export_images_per_s"
UNIT = "images/s"
WORKLOAD = ("HA export: 100 homographies x 240x320 (configs[1], magicpoint_coco_export.yaml shape); "
            "SuperPointNet_gauss2 weights, seeded random init, eval-mode BN folded; thr 0.015, nms 4, top_k 600")
CONV_GFLOP_DET = 12.161  # conv FLOPs per 240x320 view, detector path only
# dram__bytes_read.sum + dram__bytes_write.sum of conv1_fused_kernel per view: 0.2461 GB + 1.9083 GB over one 800-view
# launch (ncu --set full, profiles/r01_ncu_top_kernels.txt); algorithmic = 0.307 MB fp32 in + 2.458 MB pooled bf16 out
TRAFFIC_BYTES_PER_VIEW = (0.246116e9 + 1.908269e9) / 800


def structured_images(n, seed=0):
    """Seeded synthetic images with edges/corners (random rectangles + noise), fp32 in [0,1]."""
    rng = np.random.RandomState(seed)
    out = np.empty((n, H, W), np.float32)
    for i in range(n):
        img = np.full((H, W), 0.35, np.float32) + 0.03 * rng.randn(H, W).astype(np.float32)
        for _ in range(14):
            x0, y0 = rng.randint(0, W - 10), rng.randint(0, H - 10)
            x1, y1 = x0 + rng.randint(8, W // 3), y0 + rng.randint(8, H // 3)
            img[y0:y1, x0:x1] = rng.uniform(0, 1)
        out[i] = np.clip(img, 0, 1)
    return out


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, sustained)"
    except Exception:
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        """Keep the samples taken inside [t0, t1] (the timed region); nvidia-smi needs ~100 ms to start, so the
        sampler is started before the warm-up steps.  If the region was shorter than one sampling period, keep the
        samples closest to it (taken under the same load: the warm-up steps run the same kernels)."""
        time.sleep(0.06)
        inside = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.05]
        if not inside and self.rows:
            near = sorted(self.rows, key=lambda tr: min(abs(tr[0] - t0), abs(tr[0] - t1)))[:2]
            inside = [r for _, r in near]
        self.rows = [(0.0, r) for r in inside]

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        rows = [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) >= 6 and r[2 + j].startswith("Active") for r in rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path (oracle/ is test infrastructure; here it is the thing TIMED
# as the CPU baseline, never part of the GPU path)
# ------------------------------------------------------------------------------------------------------------
def cpu_step(params, img, mu, mf, views):
    from oracle import oracle as O
    return O.ha_export_one(img, mu[:views], mf[:views], params, THR, NMS, TOPK)


def run_reference(args):
    from oracle import oracle as O
    from pytorch_superpoint_b200.homographies import make_ha_homographies
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = O.canonical_params(O.synthetic_state_dict("gauss2", seed=1))
    imgs = structured_images(2, seed=0)
    mu, mf = make_ha_homographies(NVIEWS, seed=0)
    views = args.ref_views
    for i in range(args.warmup):
        cpu_step(params, imgs[i % 2], mu, mf, views)
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_step(params, imgs[i % 2], mu, mf, views)
    dt = (time.perf_counter() - t0) / args.steps
    value = 1.0 / (dt * NVIEWS / views)
    sample = (f"{views} of the {NVIEWS} homographies of one 240x320 image per step (warp, mask, net fp32, flatten, "
              f"combine, NMS), time scaled x{NVIEWS / views:g} to a full image")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "views_per_step": views},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from pytorch_superpoint_b200 import SuperPointNet_b200, _lib
    from pytorch_superpoint_b200.synthetic import synthetic_state_dict
    from pytorch_superpoint_b200.export import HAExporter
    from pytorch_superpoint_b200.homographies import make_ha_homographies

    os.environ.pop("NCCL_DEBUG", None)  # keep stdout to the single JSON line (NCCL prints its version banner there)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    net = SuperPointNet_b200()
    net.load_state_dict(synthetic_state_dict("gauss2", seed=1))
    exporter = HAExporter(net, THR, NMS, TOPK, device=dev, chunk=args.chunk)
    B = args.images
    imgs_np = structured_images(B, seed=0)
    # every image gets its OWN 100 homographies, like datasets/Coco.py:262-276.  Resident arm: pre-sampled outside
    # the timed region.  End-to-end arm: drawn afresh on the GPU inside every step (sampler.cu; all
    # ranks use the same counter-based seed, so the shards agree on the matrices).
    from pytorch_superpoint_b200.homographies import sample_ha_homographies_device
    imgs_host = torch.from_numpy(imgs_np).pin_memory()  # one collated, pinned [B,H,W] batch (DataLoader style)
    imgs_dev = torch.from_numpy(imgs_np).to(dev)
    mu_dev, mf_dev = sample_ha_homographies_device(NVIEWS, B, seed=0, device=dev)
    from pytorch_superpoint_b200 import ops

    def step_resident():
        # heat-maps on the current stream; NMS/top-k of this batch on the exporter's side stream (overlaps the
        # next batch's convolutions).  The closing barrier + device synchronize covers both streams.
        return exporter.submit_resident(imgs_dev, mu_dev, mf_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with ClockSampler(local) as clk:
        for _ in range(max(args.warmup, 3)):
            step_resident()
        launches0 = L.spb_launch_count()
        t_region0 = time.monotonic()
        ms = timed(step_resident, args.steps)
        clk.window(t_region0, time.monotonic())
    launches = int(L.spb_launch_count() - launches0)
    value = B * args.steps / (ms / 1e3)

    # ---- end to end through the public API: pinned host images in, host key-points out ------------------
    pending = []
    e2e_seen = []

    def step_e2e():
        # pinned host images in -> host key-points out; batch i's NMS + D2H overlap batch i+1, results of batch
        # i-1 are collected (host numpy arrays) while batch i runs
        mu_step, mf_step = sample_ha_homographies_device(NVIEWS, B, seed=1 + len(e2e_seen), device=dev)
        e2e_seen.append(0)
        pending.append(exporter.submit(imgs_host, mu_step, mf_step))
        if len(pending) > 1:
            exporter.collect(pending.pop(0))

    def drain():
        while pending:
            exporter.collect(pending.pop(0))

    for _ in range(2):
        step_e2e()
    drain()

    def timed_e2e(steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            step_e2e()
        drain()  # every batch's key-points are on the host before the clock stops
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([max(e0.elapsed_time(e1), wall)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    ms_e2e = timed_e2e(args.steps)
    e2e_value = B * args.steps / (ms_e2e / 1e3)
    h2d = B * H * W * 4  # the homographies are drawn on the device
    d2h = -(-B // world) * (TOPK * 3 * 4 + 4)

    # ---- roofline of the dominant kernel: a profiled pass (CUDA events around every stage launch) -------
    h = net.handle(dev)
    _lib.check(L.spb_net_profile(h, 1))
    prof_steps = min(args.steps, 10)
    ms_prof = timed(step_resident, prof_steps)
    import ctypes as C
    st_ms, st_n = (C.c_float * 16)(), (C.c_int * 16)()
    _lib.check(L.spb_net_profile_read(h, st_ms, st_n))
    _lib.check(L.spb_net_profile(h, 0))
    names = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b", "convPa", "convPb",
             "convDa", "convDb", "warp", "aggregate"]
    stage_ms = {names[i]: round(st_ms[i] / prof_steps, 4) for i in range(14) if st_n[i]}
    from pytorch_superpoint_b200.sharding import balanced_view_plan, shard_bounds
    lo, hi = shard_bounds(NVIEWS, rank, world)
    views_local = hi - lo  # views per image on this rank ...
    if world > 1 and NVIEWS % world and B >= world and -(-NVIEWS // world) * B <= 800:
        views_local = len(balanced_view_plan(B, NVIEWS, rank, world)[0]) / B  # ... averaged over the rotated shares
    peak_tf, peak_hbm, peak_src = peaks()
    # stage 1 = conv1_fused_kernel: conv1a (1->64) + conv1b (64->64) + pool in one launch
    k_launches = st_n[1]
    k_ms = st_ms[1] / max(k_launches, 1)
    views_per_launch = B * views_local * prof_steps / max(k_launches, 1)
    flop_per_launch = 2.0 * H * W * 64 * 9 * (64 + 1) * views_per_launch
    achieved = flop_per_launch / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    net_ms = sum(st_ms[i] for i in range(12)) / prof_steps
    # DRAM traffic of that kernel from the ncu --set full capture in profiles/ (bytes per view, measured at
    # 100 views per launch), scaled to this run's views per launch
    traffic = TRAFFIC_BYTES_PER_VIEW * views_per_launch if TRAFFIC_BYTES_PER_VIEW else None
    roofline = {"bound": "tensor", "kernel": "conv1_fused_kernel (conv1a+conv1b+pool: 44.2% of the detector's conv FLOPs)",
                "achieved": round(achieved, 2), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(achieved / peak_tf, 4),
                "traffic": traffic, "algorithmic_bytes": (H * W * 4 + (H // 2) * (W // 2) * 64 * 2) * views_per_launch,
                "peak_source": peak_src, "avg_launch_ms": round(k_ms, 4), "launches": int(k_launches),
                "views_per_launch": views_per_launch,
                "whole_net_tflops": round(CONV_GFLOP_DET * 1e-3 * B * views_local / (net_ms * 1e-3), 2) if net_ms else None}

    line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_step": B, "homographies": NVIEWS,
                       "homography_sets": "one set of 100 per image (datasets/Coco.py:262-276); resident arm: pre-sampled; "
                                          "e2e arm: drawn on the GPU inside every step",
                       "parallelism": f"homography-sharded x{world} (uneven shares rotated over the images of the batch: "
                                      f"{B * NVIEWS // world} views per rank and step), 1 NCCL all-reduce of [B,2,H,W] per step",
                       "l2": "per-image activation working set ~1.3 GB streams through the 126 MB L2 every step "
                             "(inputs larger than L2; no explicit flush)", "ha_chunk": args.chunk},
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(ms_e2e / args.steps, 4)},
            "gpu_launches": launches, "roofline": roofline, "stage_ms_per_step": stage_ms,
            "profiled_ms_per_step": round(ms_prof / prof_steps, 4), "clocks": clk.summary()}

    # ---- second half of the BASELINE metric: single-frame extraction ms/image (configs[2]: batch 64, top-k 1000,
    # detector + descriptor heads, NMS, sparse descriptors) next to the CPU oracle on one image (configs[0]) -----
    if rank == 0 and world == 1:
        from pytorch_superpoint_b200.frontend import SuperPointFrontend_b200
        fe = SuperPointFrontend_b200({"model": {"subpixel": {"enable": False}}}, "", NMS, THR, 0.7, load=False,
                                     device=str(dev))
        fe.net = net
        xb = torch.from_numpy(structured_images(64, seed=5)[:, None]).to(dev)
        for _ in range(3):
            fe.extract_batch(xb, top_k=1000)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            p_, c_, d_ = fe.extract_batch(xb, top_k=1000)
        e1.record()
        torch.cuda.synchronize()
        line["extract"] = {"ms_per_image": round(e0.elapsed_time(e1) / 10 / 64, 5), "batch": 64, "top_k": 1000,
                           "workload": "configs[2]: key-point + descriptor extraction, batch 64 at 240x320, device resident",
                           "mean_keypoints": round(float(c_.float().mean().item()), 1)}

    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O  # the CPU baseline leg: the one place this arm touches oracle/
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        params = O.canonical_params(synthetic_state_dict("gauss2", seed=1))
        mu_np, mf_np = mu_dev[0].cpu().numpy(), mf_dev[0].cpu().numpy()  # image 0's matrices, same on both sides
        cpu_step(params, imgs_np[0], mu_np, mf_np, 5)  # warm-up
        t0 = time.perf_counter()
        pts_cpu = cpu_step(params, imgs_np[0], mu_np, mf_np, NVIEWS)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": round(1.0 / dt, 4), "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "one full 240x320 image with all 100 homographies through oracle/ "
                                          "(numpy warp/mask/combine/NMS + torch fp32 conv on all host threads)"}
        # the same image through the GPU path: key-point overlap with the fp32 CPU oracle (informational)
        pts_gpu = exporter.export_image(imgs_host[0], mu_dev[0], mf_dev[0])
        a = {(int(x), int(y)) for x, y, _ in pts_gpu}
        b = {(int(x), int(y)) for x, y, _ in pts_cpu}
        line["keypoint_jaccard_vs_cpu_fp32"] = round(len(a & b) / max(len(a | b), 1), 4)
        # configs[0]: the same single-frame extraction on the host cores (CPU oracle, one image)
        x1 = torch.from_numpy(structured_images(1, seed=5)[:, None])
        t0 = time.perf_counter()
        semi_c, desc_c = O.net_forward(params, x1)
        pts_c = O.get_pts_from_heatmap(O.flatten_detection(semi_c.numpy())[0], THR, NMS)[:, :1000]
        O.sample_desc_from_points(desc_c.numpy(), pts_c)
        line["extract"]["cpu_ms_per_image"] = round((time.perf_counter() - t0) * 1e3, 2)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=32, help="images per step (one batch)")
    ap.add_argument("--chunk", type=int, default=0, help="views per network pass (0 = all of the rank's views)")
    ap.add_argument("--ref-views", type=int, default=20, help="--impl reference: homographies per step (bounded sample)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
