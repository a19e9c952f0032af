#!/usr/bin/env python
"""bench.py -- HA-export throughput (BASELINE.json metric) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4|c5|c1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs; the default c2 is the configuration the metric is quoted on):
  c2  HA export, 100 homographies x 240x320 (configs[1], magicpoint_coco_export.yaml shape), top-600
  c4  HA export, 100 homographies x 384x1248 (configs[3], magicpoint_kitti_export.yaml shape)
  c5  HA export, 100 homographies x 480x640 + descriptors at the key-points + two-way matching of image pairs
      (configs[4]: HPatches shape, homographies sharded over the GPUs with the NCCL exchange of the summed planes)
  c3  single-frame extraction, batch 64 at 240x320, top-k 1000 (configs[2]);  c1: the same with ONE image (configs[0])

A "step" is one pass of the hot path over one batch of synthetic input: IMAGES x n_gpus images (weak scaling: every
GPU works on the same number of views per step at every N), each homography-adapted: image warp x100 -> detector
network -> fused softmax/d2s/un-warp/aggregate -> (when the homographies are sharded over ranks: the owner of an image
adds the partial planes out of its peers' memory over NVLink inside the dividing kernel; SPB200_EXCHANGE=nccl: a
reduce-scatter over the image dimension) -> divide -> threshold + greedy NMS + top-k.  Rank 0 prints ONE JSON line.

  value      images/s with inputs resident in HBM (CUDA events, barrier + synchronize both sides, max over ranks)
  e2e        same metric through HAExporter.submit/collect with pinned HOST images in and HOST key-points out
  roofline   the dominant kernel (conv1_fused_kernel: conv1a+conv1b+pool, 44 % of the conv FLOPs) against the
             measured bf16 tensor peak (sustained when the clock record shows the power cap, burst otherwise)
  cpu_baseline  the reference's own CPU path on this box's host cores (rank 0, N=1 only): the UNMODIFIED reference
             functions from baseline/_ref when that copy travelled (kind "reference"), else the oracle port; with the
             wall time of every stage (stages_ms)
  train_mode_bn (c2, N=1) the as-written reference (BN layers left in train mode) on the host cores next to
             SuperPointNet_b200(bn_mode='batch'), one image each, with the key-point overlap of the two (informational)
  shard_check (N>1) a few images recomputed unsharded on rank 0 outside the timed region vs the sharded result
  image_sharded (N>1) comparison line: the same batch with images instead of homographies sharded (no collective)
  --impl reference   times that CPU arm as the main line (one full image, all 100 homographies, per step)
"""
from __future__ import annotations

import argparse
import glob
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

THR, NMS = 0.015, 4
METRIC = "ha_export_images_per_s"
UNIT = "images/s"
# conv FLOPs per view, detector path only / with the descriptor head
WORKLOADS = {
    "c2": dict(kind="ha", H=240, W=320, N=100, topk=600, images=32, gflop_det=12.161, gflop_full=13.026,
               name="HA export: 100 homographies x 240x320 (configs[1], magicpoint_coco_export.yaml shape)"),
    "c4": dict(kind="ha", H=384, W=1248, N=100, topk=600, images=4, gflop_det=75.884, gflop_full=81.282,
               name="HA export: 100 homographies x 384x1248 (configs[3], magicpoint_kitti_export.yaml shape)"),
    "c5": dict(kind="ha", H=480, W=640, N=100, topk=1000, images=8, gflop_det=48.643, gflop_full=52.104, match=True,
               name="HA export: 100 homographies x 480x640 + descriptors + two-way matching of image pairs "
                    "(configs[4], HPatches shape)"),
    "c3": dict(kind="extract", H=240, W=320, N=1, topk=1000, images=64, gflop_det=12.161, gflop_full=13.026,
               name="key-point + descriptor extraction, batch 64 at 240x320, top-k 1000 (configs[2])"),
    "c1": dict(kind="extract", H=240, W=320, N=1, topk=0, images=1, gflop_det=12.161, gflop_full=13.026,
               name="single 240x320 image: heat-map, nms_fast (dist 4, conf 0.015), descriptors (configs[0])"),
}
NET_NOTE = "SuperPointNet_gauss2 weights, seeded random init, eval-mode BN folded; thr 0.015, nms 4"


def workload_name(key):
    wl = WORKLOADS[key]
    return f"{key}: {wl['name']}; {NET_NOTE}, top_k {wl['topk']}"


def structured_images(n, H, W, seed=0):
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
        return (float(p["bf16_tflops_sustained"]), float(p["bf16_tflops"]), float(p["hbm_gbs"]),
                "measured (MEASURED_PEAKS.json)")
    except Exception:
        return 1400.0, 1650.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per view of the dominant kernel from the newest ncu --set full summary committed under profiles/
    (profiles/r*_ncu_traffic.json: {"kernel": ..., "dram_bytes_per_view": ..., "views_per_launch": ..., "source": ...})."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))
    for f in reversed(files):
        try:
            with open(f) as fh:
                d = json.load(fh)
            return float(d["dram_bytes_per_view"]), os.path.basename(f)
        except Exception:
            continue
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""

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
# CPU arm.  oracle/ and baseline/_ref are test / baseline infrastructure; here they are the thing TIMED as the CPU
# baseline, never part of the GPU path.
# ------------------------------------------------------------------------------------------------------------
class CpuArm:
    """One HA-exported image on the host cores: the UNMODIFIED reference functions (baseline/_ref, copied by
    baseline/install_ref.py in the build container) when present, else the oracle port of the same path."""

    def __init__(self, wl):
        from oracle import oracle as O
        self.O, self.wl = O, wl
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.sd = O.synthetic_state_dict("gauss2", seed=1)
        self.ref = None
        ref_dir = os.path.join(ROOT, "baseline", "_ref")
        if os.path.isfile(os.path.join(ref_dir, "export.py")):
            try:
                from oracle import refshim
                refshim.REF = ref_dir
                self.ref = refshim.load()
                net = self.ref.nets["gauss2"]()
                net.load_state_dict(self.sd)
                net.eval()  # the parity target (eval-mode BN); no reference file is edited
                self.fe = self.ref.frontend(conf_thresh=THR, nms_dist=NMS)
                self.fe.net = net
            except Exception as e:  # noqa: BLE001 -- any import problem: fall back to the port, say so
                self.ref, self.why = None, f"baseline/_ref import failed: {type(e).__name__}: {e}"
        self.kind = "reference" if self.ref is not None else "port"
        self.params = None if self.ref is not None else O.canonical_params(self.sd)

    def ha_image(self, img, mu, mf, train_bn=False):
        """img [H,W] numpy, mu / mf [N,3,3] numpy -> float64 [K,3] key-points (export.py:283-328 for one image).
        ``train_bn``: leave the network in train mode = the as-written reference (model_wrap.py:111 never calls
        .eval()), batch statistics over the N views.  ``self.stages`` receives the wall time of every stage in ms."""
        wl = self.wl
        if self.ref is None:
            t0 = time.perf_counter()
            pts = self.O.ha_export_one(img, mu, mf, self.params, THR, NMS, wl["topk"], bn_mode="batch" if train_bn else "eval")
            self.stages = {"all": round((time.perf_counter() - t0) * 1e3, 1)}
            return pts
        ref, fe = self.ref, self.fe
        H, W = img.shape
        x = torch.from_numpy(img)[None, None]
        inv = torch.from_numpy(mf)
        tm = [time.perf_counter()]
        fe.net.train(bool(train_bn))
        try:
            with torch.no_grad():
                warped = ref.inv_warp_image_batch(x.repeat(len(mf), 1, 1, 1), inv, mode="bilinear")     # Coco.py:279
                tm.append(time.perf_counter())
                mask = ref.compute_valid_mask(torch.tensor([H, W]), inv_homography=inv, erosion_radius=0)[:, None]
                tm.append(time.perf_counter())
                heat = fe.run(warped, onlyHeatmap=True, train=False)                                      # export.py:309
                tm.append(time.perf_counter())
                agg = ref.combine_heatmap(heat, torch.from_numpy(mu)[None], mask, device="cpu")           # export.py:310
                tm.append(time.perf_counter())
        finally:
            fe.net.eval()
        pts = fe.getPtsFromHeatmap(agg.detach().cpu().squeeze().numpy()).transpose()                # export.py:311
        tm.append(time.perf_counter())
        self.stages = {k: round((b - a) * 1e3, 1) for k, a, b in
                       zip(("warp", "valid_mask", "net_and_flatten", "combine_heatmap", "nms_topk"), tm[:-1], tm[1:])}
        return pts[:wl["topk"]] if wl["topk"] else pts

    def extract_image(self, img):
        """Single-frame extraction (export.py:127-145): net -> flatten -> getPtsFromHeatmap -> descriptors."""
        wl, O = self.wl, self.O
        x = torch.from_numpy(img)[None, None]
        if self.ref is None:
            semi, desc = O.net_forward(self.params, x)
            pts = O.get_pts_from_heatmap(O.flatten_detection(semi.numpy())[0], THR, NMS)
            pts = pts[:, :wl["topk"]] if wl["topk"] else pts
            return pts, O.sample_desc_from_points(desc.numpy(), pts)
        fe = self.fe
        with torch.no_grad():
            out = fe.net(x)
            heat = self.ref.flattenDetection(out["semi"], tensor=True)
        pts = fe.getPtsFromHeatmap(heat.squeeze().numpy())
        pts = pts[:, :wl["topk"]] if wl["topk"] else pts
        return pts, fe.sample_desc_from_points(out["desc"], pts)

    def sample(self):
        wl = self.wl
        if wl["kind"] == "ha":
            what = (f"one full {wl['H']}x{wl['W']} image with all {wl['N']} homographies per step (warp, mask, net fp32, "
                    "flatten, combine, NMS, top-k)")
        else:
            what = f"one {wl['H']}x{wl['W']} image per step (net fp32, flatten, NMS, descriptor sampling)"
        src = ("the UNMODIFIED reference functions (baseline/_ref: utils/utils.py, models/model_wrap.py, export.py, "
               "SuperPointNet_gauss2; net.eval())") if self.ref is not None else "oracle/ (numpy + torch fp32 port)"
        return f"{what} through {src} on all host threads"


def run_reference(args):
    from pytorch_superpoint_b200.homographies import make_ha_homographies
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    arm = CpuArm(wl)
    imgs = structured_images(2, wl["H"], wl["W"], seed=0)
    mu, mf = make_ha_homographies(wl["N"], seed=0)

    def step(i):
        if wl["kind"] == "ha":
            arm.ha_image(imgs[i % 2], mu, mf)
        else:
            arm.extract_image(imgs[i % 2])

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = 1.0 / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "images_per_step": 1, "homographies": wl["N"],
                       "sample": "one image of the workload per step (a bounded sample: the GPU arm's step is a batch)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": arm.sample()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import ctypes as C
    import torch.distributed as dist
    from pytorch_superpoint_b200 import SuperPointNet_b200, _lib, ops
    from pytorch_superpoint_b200.export import HAExporter
    from pytorch_superpoint_b200.frontend import SuperPointFrontend_b200
    from pytorch_superpoint_b200.homographies import sample_ha_homographies_device
    from pytorch_superpoint_b200.sharding import plan_batch
    from pytorch_superpoint_b200.synthetic import synthetic_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NCCL_DEBUG is left as the caller set it (the communicator lines "... rank r nranks N ..." are the evidence of
    # the rank count).  NCCL writes them to file descriptor 1, also at process exit, behind anything this script
    # prints: to keep stdout to the ONE JSON line, fd 1 is pointed at stderr for the life of the process and the
    # JSON line goes to a duplicate of the original stdout.
    json_fd = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    wl = WORKLOADS[args.workload]
    H, W, NV, TOPK = wl["H"], wl["W"], wl["N"], wl["topk"]
    ha = wl["kind"] == "ha"
    per_gpu = args.images if args.images > 0 else wl["images"]
    B = per_gpu * world if ha else per_gpu  # weak scaling: the same number of views per GPU and step at every N

    net = SuperPointNet_b200()
    net.load_state_dict(synthetic_state_dict("gauss2", seed=1))
    exporter = HAExporter(net, THR, NMS, TOPK, device=dev, chunk=args.chunk)
    fe = SuperPointFrontend_b200({"model": {"subpixel": {"enable": False}}}, "", NMS, THR, 0.7, load=False, device=str(dev))
    fe.net = net
    imgs_np = structured_images(B, H, W, seed=0)
    imgs_host = torch.from_numpy(imgs_np).pin_memory()  # one collated, pinned [B,H,W] batch (DataLoader style)
    imgs_dev = torch.from_numpy(imgs_np).to(dev)
    # every image gets its OWN homographies, like datasets/Coco.py:262-276.  Resident arm: pre-sampled outside the
    # timed region.  End-to-end arm: drawn afresh on the GPU inside every step (sampler.cu; all ranks
    # use the same counter-based seed, so the shards agree on the matrices).
    mu_dev = mf_dev = None
    if ha:
        mu_dev, mf_dev = sample_ha_homographies_device(NV, B, seed=0, device=dev)
    own_lo, own_hi = exporter._buffers(B, H, W)["slots"][0]["own"] if ha else (0, B)
    match_state = {}

    def match_owned(ticket):
        """c5: descriptors of the owned images at their HA key-points + two-way matching of the pairs (2k, 2k+1), all
        on the exporter's side stream behind the ticket's NMS (no host synchronisation)."""
        lo, hi = ticket["own"]
        if hi - lo < 2:
            return
        side = exporter._side
        x = imgs_dev[lo:hi, None]
        main = torch.cuda.current_stream()
        _, desc = net.forward_nhwc(x, want_desc=True)  # main stream: shares the network workspace with the HA passes
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            d = ops.sample_desc_batch_device(desc, ticket["pts_dev"], ticket["cnt_dev"])  # [nown,cap,256], zero padded
            res = ops.nn_match_pairs_device(d[0::2], d[1::2], 0.7)
            match_state["last"] = res
            ticket["done"] = torch.cuda.Event()
            ticket["done"].record(side)
        desc.record_stream(side)

    def step_resident():
        if ha:
            t = exporter.submit_resident(imgs_dev, mu_dev, mf_dev)
            if wl.get("match"):
                match_owned(t)
            return t
        return fe.extract_batch(imgs_dev[:, None], top_k=TOPK)

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
    clocks = clk.summary()

    # ---- end to end through the public API: pinned host images in, host key-points out ------------------
    pending, e2e_seen, host_out = [], [], {}

    def step_e2e():
        if not ha:  # extraction: H2D of the batch, D2H of key-points + sparse descriptors into pinned buffers
            p_, c_, d_ = fe.extract_batch(imgs_host.to(dev, non_blocking=True)[:, None], top_k=TOPK)
            if "p" not in host_out:
                host_out.update(p=torch.empty(p_.shape, dtype=p_.dtype).pin_memory(),
                                c=torch.empty(c_.shape, dtype=c_.dtype).pin_memory(),
                                d=torch.empty(d_.shape, dtype=d_.dtype).pin_memory())
            host_out["p"].copy_(p_, non_blocking=True)
            host_out["c"].copy_(c_, non_blocking=True)
            host_out["d"].copy_(d_, non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the results are on the host before the next step starts
            return host_out
        # pinned host images in -> host key-points out; batch i's exchange + NMS + D2H overlap batch i+1, results of
        # batch i-1 are collected (host numpy arrays) while batch i runs
        mu_step, mf_step = sample_ha_homographies_device(NV, B, seed=1 + len(e2e_seen), device=dev)
        e2e_seen.append(0)
        t = exporter.submit(imgs_host, mu_step, mf_step)
        if wl.get("match"):
            match_owned(t)
        pending.append(t)
        if len(pending) > 1:
            exporter.collect(pending.pop(0))

    def drain():
        while pending:
            exporter.collect(pending.pop(0))
        if match_state.get("last") is not None:
            match_state["last"][3].cpu()  # the match counts of the last batch reach the host before the clock stops

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
    h2d = B * H * W * 4  # every rank reads the whole batch (it holds views of every image); homographies: on device
    cap = TOPK if TOPK else -(-H // 5) * -(-W // 5)
    d2h = (own_hi - own_lo) * (cap * 3 * 4 + 4) + (0 if ha else B * cap * 256 * 4)

    # ---- roofline of the dominant kernel: a profiled pass (CUDA events around every stage launch) -------
    h = net.handle(dev)
    net.set_profile(True, dev)
    prof_steps = min(args.steps, 10)
    ms_prof = timed(step_resident, prof_steps)
    st_ms, st_n = (C.c_float * 16)(), (C.c_int * 16)()
    _lib.check(L.spb_net_profile_read(h, st_ms, st_n))
    net.set_profile(False, dev)
    names = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b", "convPa", "convPb",
             "convDa", "convDb", "warp", "aggregate"]
    stage_ms = {names[i]: round(st_ms[i] / prof_steps, 4) for i in range(14) if st_n[i]}
    if ha:
        ha_views = plan_batch(B, NV, rank, world, net.ha_pass_views(H, W, dev)).local_views
        views_step = ha_views + ((own_hi - own_lo) if wl.get("match") else 0)
    else:
        ha_views, views_step = 0, B
    peak_sus, peak_burst, peak_hbm, peak_src = peaks()
    capped = "sw_power_cap" in (clocks.get("reasons") or [])
    peak_tf = peak_sus if capped else peak_burst  # a kernel timed in an uncapped region is measured against burst
    # stage 1 = conv1_fused_kernel: conv1a (1->64) + conv1b (64->64) + pool in one launch
    k_launches = st_n[1]
    k_ms = st_ms[1] / max(k_launches, 1)
    views_per_launch = views_step * prof_steps / max(k_launches, 1)
    flop_per_launch = 2.0 * H * W * 64 * 9 * (64 + 1) * views_per_launch
    achieved = flop_per_launch / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    net_ms = sum(st_ms[i] for i in range(12)) / prof_steps
    per_view, traffic_src = ncu_traffic()
    traffic = per_view * views_per_launch * (H * W) / (240 * 320) if per_view else None
    det_flop = wl["gflop_det"] * 1e-3 * ha_views + wl["gflop_full"] * 1e-3 * (views_step - ha_views)
    roofline = {"bound": "tensor", "kernel": "conv1_fused_kernel (conv1a+conv1b+pool: 44.2% of the detector's conv FLOPs)",
                "achieved": round(achieved, 2), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(achieved / peak_tf, 4),
                "frac_of_sustained": round(achieved / peak_sus, 4), "frac_of_burst": round(achieved / peak_burst, 4),
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes": (H * W * 4 + (H // 2) * (W // 2) * 64 * 2) * views_per_launch,
                "peak_source": peak_src + (": sustained (sw_power_cap active in the timed region)" if capped else
                                           ": burst (no cap in the clock record)"),
                "avg_launch_ms": round(k_ms, 4), "launches": int(k_launches), "views_per_launch": views_per_launch,
                "whole_net_tflops": round(det_flop / (net_ms * 1e-3), 2) if net_ms else None,
                "whole_path_frac": round(value / world * wl["gflop_det"] * 1e-3 * NV / peak_tf, 4) if ha else None}

    line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args.workload),
                       "images_per_step": B, "images_per_gpu_and_step": per_gpu, "homographies": NV,
                       "homography_sets": "one set per image (datasets/Coco.py:262-276); resident arm: pre-sampled; "
                                          "e2e arm: drawn on the GPU inside every step",
                       "parallelism": (f"homography-sharded x{world}: every rank warps / runs the detector on its rotated "
                                       f"block of the views of EVERY image ({views_step} views per rank and step); ONE "
                                       "exchange of the [B,2,H,W] planes per step, overlapped with the next step: " +
                                       ("the owner of an image adds the partial planes straight out of its peers' memory "
                                        "over NVLink (CUDA IPC) inside the kernel that divides (csrc/peer.cu)"
                                        if getattr(exporter, "exchange", "") == "peer" else
                                        "an asynchronous NCCL reduce-scatter over the image dimension") +
                                       "; each rank runs NMS on its B/N images") if ha else
                                      "single GPU batch",
                       "exchange": getattr(exporter, "exchange", None) if (ha and world > 1) else None,
                       "l2": "per-pass activation working set (~2 GB) streams through the 126 MB L2 every step "
                             "(inputs larger than L2; no explicit flush)", "ha_chunk": args.chunk},
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(ms_e2e / args.steps, 4)},
            "gpu_launches": launches, "roofline": roofline, "stage_ms_per_step": stage_ms,
            "stage_sum_over_step": round(sum(stage_ms.values()) / (ms_prof / prof_steps), 4),
            "profiled_ms_per_step": round(ms_prof / prof_steps, 4), "clocks": clocks}
    if wl.get("match") and match_state.get("last") is not None:
        line["matches_per_pair"] = round(float(match_state["last"][3].float().mean().item()), 1)

    # ---- N > 1: the sharded result against an unsharded recomputation on rank 0 (outside the timed region) ------
    if ha and world > 1:
        t = exporter.submit_resident(imgs_dev, mu_dev, mf_dev)
        pts_sh = exporter.collect(t)
        heat_sh = exporter.last_heat(B, H, W).clone()
        if rank == 0:
            k = min(own_hi - own_lo, 4)
            local = HAExporter(net, THR, NMS, TOPK, device=dev, chunk=args.chunk, sharded=False)
            pts_lo = local.collect(local.submit_resident(imgs_dev[:k].contiguous(), mu_dev[:k].contiguous(),
                                                         mf_dev[:k].contiguous()))
            heat_lo = local.last_heat(k, H, W)
            jac = []
            for b in range(k):
                a = {(int(x), int(y)) for x, y, _ in pts_sh[b]}
                c = {(int(x), int(y)) for x, y, _ in pts_lo[b]}
                jac.append(len(a & c) / max(len(a | c), 1))
            line["shard_check"] = {"images": k, "heat_max_abs": float((heat_sh[:k] - heat_lo).abs().max().item()),
                                   "keypoint_jaccard": round(min(jac), 4),
                                   "note": "sharded (this run) vs unsharded on rank 0: fp32 summation order only"}

    # ---- N > 1: SURVEY 8(e) "alternative noted" -- the same batch with IMAGES sharded instead of homographies (every
    # rank: its own B / world images with all their views, no collective), as a comparison line outside the timed region
    if ha and world > 1:
        img_ex = HAExporter(net, THR, NMS, TOPK, device=dev, chunk=args.chunk, sharded=False)
        lo_i, hi_i = rank * per_gpu, (rank + 1) * per_gpu
        im_l, mu_l, mf_l = (t[lo_i:hi_i].contiguous() for t in (imgs_dev, mu_dev, mf_dev))

        def step_img():
            return img_ex.submit_resident(im_l, mu_l, mf_l)

        for _ in range(3):
            step_img()
        n_img = min(args.steps, 10)
        ms_img = timed(step_img, n_img)
        line["image_sharded"] = {"value": round(B * n_img / (ms_img / 1e3), 3), "unit": UNIT, "steps": n_img,
                                 "note": "comparison only: images (not homographies) sharded over the ranks, no "
                                         "collective; north_star mandates homography sharding + one exchange (the "
                                         "value above), which also shortens the latency of a single image by world x"}

    # ---- second half of the BASELINE metric: single-frame extraction ms/image (configs[2]) next to the CPU arm -----
    if rank == 0 and world == 1 and ha and args.workload == "c2":
        xb = torch.from_numpy(structured_images(64, 240, 320, seed=5)[:, None]).to(dev)
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
        arm = CpuArm(wl)  # the CPU baseline leg: the one place this arm touches oracle/ and baseline/_ref
        if ha:
            mu_np, mf_np = mu_dev[0].cpu().numpy(), mf_dev[0].cpu().numpy()  # image 0's matrices, same on both sides
            small = min(NV, 5)
            arm.ha_image(imgs_np[0], mu_np[:small], mf_np[:small])  # warm-up
            t0 = time.perf_counter()
            pts_cpu = arm.ha_image(imgs_np[0], mu_np, mf_np)
            dt = time.perf_counter() - t0
            # the same image through the GPU path: key-point overlap with the fp32 CPU path (informational)
            pts_gpu = exporter.export_image(imgs_host[0], mu_dev[0], mf_dev[0])
            a = {(int(x), int(y)) for x, y, _ in pts_gpu}
            b = {(int(x), int(y)) for x, y, _ in pts_cpu}
            line["keypoint_jaccard_vs_cpu_fp32"] = round(len(a & b) / max(len(a | b), 1), 4)
        else:
            arm.extract_image(imgs_np[0])
            t0 = time.perf_counter()
            for i in range(5):
                arm.extract_image(imgs_np[i % B])
            dt = (time.perf_counter() - t0) / 5
        line["cpu_baseline"] = {"value": round(1.0 / dt, 4), "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                                "sample": arm.sample().replace("per step", "(timed once)")}
        if ha:
            line["cpu_baseline"]["stages_ms"] = dict(arm.stages)
        if ha and args.workload == "c2":
            # SURVEY 8(d) "report both BN modes": the as-written reference leaves the BN layers in train mode
            # (model_wrap.py:111); the same image once more that way, next to our opt-in fp32 path for it
            t0 = time.perf_counter()
            pts_train = arm.ha_image(imgs_np[0], mu_np, mf_np, train_bn=True)
            dt_train = time.perf_counter() - t0
            from pytorch_superpoint_b200 import SuperPointNet_b200 as _Net
            net_bn = _Net(bn_mode="batch")
            net_bn.load_state_dict(net.state_dict())
            ex_bn = HAExporter(net_bn, THR, NMS, TOPK, device=dev, sharded=False)
            ex_bn.export_image(imgs_host[0], mu_dev[0], mf_dev[0])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pts_bn = ex_bn.export_image(imgs_host[0], mu_dev[0], mf_dev[0])
            dt_bn = time.perf_counter() - t0
            a = {(int(x), int(y)) for x, y, _ in pts_bn}
            b = {(int(x), int(y)) for x, y, _ in pts_train}
            line["train_mode_bn"] = {
                "cpu_reference_images_per_s": round(1.0 / dt_train, 4),
                "b200_bn_mode_batch_images_per_s": round(1.0 / dt_bn, 2),
                "keypoint_jaccard": round(len(a & b) / max(len(a | b), 1), 4),
                "note": "the as-written reference (BN layers left in train mode: statistics of the 100 views) on the host "
                        "cores vs SuperPointNet_b200(bn_mode='batch') (fp32 CUDA-core path, csrc/bn.cu), one image each, "
                        "single calls; informational -- the metric above is the eval-mode product path"}
        if "extract" in line:  # configs[0]: the same single-frame extraction on the host cores, one image
            x1 = structured_images(1, 240, 320, seed=5)[0]
            arm1 = CpuArm(WORKLOADS["c3"])
            arm1.extract_image(x1)
            t0 = time.perf_counter()
            arm1.extract_image(x1)
            line["extract"]["cpu_ms_per_image"] = round((time.perf_counter() - t0) * 1e3, 2)
            line["extract"]["cpu_kind"] = arm1.kind
    if world > 1:
        try:
            if ha:
                exporter.close()  # collective: unmap / free the peer-memory exchange buffers
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass
    if rank == 0:
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--images", type=int, default=0, help="images per GPU and step (0 = the workload's default)")
    ap.add_argument("--chunk", type=int, default=0, help="views per network pass (0 = library default: bounded by bytes)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
