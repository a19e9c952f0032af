#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Everything written is small seeded input/output data (no reference source).  The tests replay
these against ``oracle/`` on CPU and against the CUDA path on the GPU box, where
``/root/reference`` does not exist.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from oracle.oracle import synthetic_state_dict  # noqa: E402
from pytorch_superpoint_b200.homographies import EXPORT_PARAMS  # noqa: E402

REF = refshim.REF


def structured_image(H, W, seed):
    """Synthetic image with corners/edges: random rectangles and lines, lightly blurred."""
    import cv2
    rng = np.random.RandomState(seed)
    img = np.full((H, W), 0.3, np.float32)
    for _ in range(12):
        x0, y0 = rng.randint(0, W - 8), rng.randint(0, H - 8)
        x1, y1 = x0 + rng.randint(6, max(W // 3, 8)), y0 + rng.randint(6, max(H // 3, 8))
        cv2.rectangle(img, (x0, y0), (x1, y1), float(rng.uniform(0, 1)), -1)
    for _ in range(6):
        p = tuple(int(v) for v in (rng.randint(0, W), rng.randint(0, H)))
        q = tuple(int(v) for v in (rng.randint(0, W), rng.randint(0, H)))
        cv2.line(img, p, q, float(rng.uniform(0, 1)), 1)
    return cv2.GaussianBlur(img, (3, 3), 0).astype(np.float32)


def ha_mats(ref, n, seed):
    """datasets/Coco.py:265-276 with the reference's own sampler."""
    np.random.seed(seed)
    hs = np.stack([np.linalg.inv(ref.sample_homography_np(np.array([2, 2]), shift=-1, **EXPORT_PARAMS))
                   for _ in range(n)])
    hs[0] = np.identity(3)
    hs = torch.tensor(hs, dtype=torch.float32)
    inv = torch.stack([torch.inverse(hs[i]) for i in range(n)])
    return hs, inv


def main():
    ref = refshim.load()
    torch.manual_seed(0)
    out = {}

    # ---- homography sampler (utils/homographies.py:12-141) -------------------------------------
    np.random.seed(7)
    raw = np.stack([ref.sample_homography_np(np.array([2, 2]), shift=-1, **EXPORT_PARAMS) for _ in range(8)])
    np.random.seed(11)
    raw_noart = np.stack([ref.sample_homography_np(np.array([2, 2]), shift=-1, perspective=True, scaling=True,
                                                   rotation=True, translation=True, allow_artifacts=False,
                                                   patch_ratio=0.5) for _ in range(8)])
    hs, inv = ha_mats(ref, 6, 3)
    np.savez_compressed(os.path.join(HERE, "homographies.npz"), raw_seed7=raw, raw_noart_seed11=raw_noart,
                        ha_seed3_h=hs.numpy(), ha_seed3_inv=inv.numpy())

    # ---- warp + valid mask (utils/utils.py:318-351, 677-704) -----------------------------------
    H, W = 48, 64
    img = torch.from_numpy(structured_image(H, W, 1))
    hs, inv = ha_mats(ref, 6, 3)
    warped = ref.inv_warp_image_batch(img.repeat(6, 1, 1, 1), inv, mode="bilinear")[:, 0]
    nearest = ref.inv_warp_image_batch(img.repeat(6, 1, 1, 1), inv, mode="nearest")[:, 0]
    mask = ref.compute_valid_mask(torch.tensor([H, W]), inv_homography=inv, erosion_radius=0)
    mask_e2 = ref.compute_valid_mask(torch.tensor([H, W]), inv_homography=inv, erosion_radius=2)
    np.savez_compressed(os.path.join(HERE, "warp.npz"), img=img.numpy(), h=hs.numpy(), inv=inv.numpy(),
                        warped=warped.numpy(), nearest=nearest.numpy(), mask=mask.numpy(), mask_e2=mask_e2.numpy())

    # full-size mask statistics on the bench shape (pins nearest-rounding agreement at 240x320, N=100)
    hs100, inv100 = ha_mats(ref, 100, 0)
    m100 = ref.compute_valid_mask(torch.tensor([240, 320]), inv_homography=inv100, erosion_radius=0).numpy()
    np.savez_compressed(os.path.join(HERE, "mask_240x320_n100.npz"), h=hs100.numpy(), inv=inv100.numpy(),
                        mask_bits=np.packbits(m100.astype(np.uint8)), shape=np.array(m100.shape))

    # ---- the reference's own known-answer artefact: notebooks/h2.npz (test/sample_homography.py:85-113)
    h2 = np.load(os.path.join(REF, "notebooks", "h2.npz"))
    np.savez_compressed(os.path.join(HERE, "ref_h2.npz"), homographies=h2["homographies"], masks=h2["masks"])

    # ---- flattenDetection + combine_heatmap (utils/utils.py:480-525, export.py:49-63) -----------
    semi = torch.randn(6, 65, H // 8, W // 8) * 2.0
    heat = ref.flattenDetection(semi, tensor=True)  # [6,1,H,W]
    comb = ref.combine_heatmap(heat, hs[None], mask[:, None], device="cpu")
    np.savez_compressed(os.path.join(HERE, "heat.npz"), semi=semi.numpy(), heat=heat[:, 0].numpy(),
                        h=hs.numpy(), mask=mask.numpy(), combined=comb[0].numpy())

    # ---- getPtsFromHeatmap / nms_fast (models/model_wrap.py:117-180, 252-279), stable-sort contract
    fe = ref.frontend(conf_thresh=0.015, nms_dist=4)
    cases = {}
    rng = np.random.RandomState(5)
    import cv2
    smooth = cv2.GaussianBlur(rng.rand(96, 128).astype(np.float32), (7, 7), 0) * 0.05
    cases["smooth_96x128"] = smooth
    cases["flat_random_240x320"] = (0.0154 + 0.002 * rng.randn(240, 320)).astype(np.float32)
    sparse = np.zeros((64, 80), np.float32)
    idx = rng.choice(64 * 80, 300, replace=False)
    sparse.flat[idx] = rng.uniform(0.01, 0.9, 300).astype(np.float32)
    cases["sparse_64x80"] = sparse
    ties = (np.round(cv2.GaussianBlur(rng.rand(72, 88).astype(np.float32), (5, 5), 0) * 40) / 400).astype(np.float32)
    cases["ties_72x88"] = ties
    cases["empty_40x48"] = np.zeros((40, 48), np.float32)
    one = np.zeros((40, 48), np.float32); one[20, 17] = 0.5
    cases["single_40x48"] = one
    border = np.zeros((40, 48), np.float32); border[2, 2] = 0.9; border[5, 5] = 0.8; border[20, 20] = 0.7; border[39, 47] = 0.3
    cases["border_40x48"] = border
    pts_out = {}
    with refshim.stable_argsort():
        for k, h in cases.items():
            pts_out["heat_" + k] = h
            pts_out["pts_" + k] = fe.getPtsFromHeatmap(h.copy())
        # direct nms_fast with float (sub-pixel) corners and a different distance
        c = np.stack([rng.uniform(0, 79, 400), rng.uniform(0, 63, 400), rng.uniform(0, 1, 400)])
        o, oi = fe.nms_fast(c, 64, 80, dist_thresh=3)
        pts_out["nms_in"], pts_out["nms_out"], pts_out["nms_inds"] = c, o, oi
    np.savez_compressed(os.path.join(HERE, "pts.npz"), **pts_out)

    # ---- networks: three state-dict layouts, eval-mode and batch-mode BN -------------------------
    x = torch.from_numpy(np.stack([structured_image(32, 48, 10 + i) for i in range(3)]))[:, None]
    net_out = {"x": x.numpy()}
    for layout in ("v1", "bnflat", "gauss2"):
        net = ref.nets[layout]()
        net.load_state_dict(synthetic_state_dict(layout, seed=1))
        for mode in (("eval", "batch") if layout != "v1" else ("eval",)):
            net.train(mode == "batch")
            with torch.no_grad():
                o = net(x)
            semi_o, desc_o = (o if isinstance(o, tuple) else (o["semi"], o["desc"]))
            net_out[f"{layout}_{mode}_semi"] = semi_o.numpy()
            net_out[f"{layout}_{mode}_desc"] = desc_o.numpy()
    np.savez_compressed(os.path.join(HERE, "net.npz"), **net_out)

    # ---- sample_desc_from_points (models/model_wrap.py:281-299) ---------------------------------
    coarse = torch.randn(1, 256, 6, 8)
    coarse = coarse / coarse.norm(dim=1, keepdim=True)
    p = np.stack([rng.randint(4, 60, 50).astype(float), rng.randint(4, 44, 50).astype(float), rng.rand(50)])
    d = fe.sample_desc_from_points(coarse, p)
    np.savez_compressed(os.path.join(HERE, "desc.npz"), coarse=coarse.numpy(), pts=p, desc=d)

    # ---- end-to-end HA export of one image (Coco.py:262-284 + export.py:283-328), eval-mode BN ----
    H, W, N = 64, 96, 8
    img = torch.from_numpy(structured_image(H, W, 21))
    hs, inv = ha_mats(ref, N, 9)
    net = ref.nets["gauss2"]()
    net.load_state_dict(synthetic_state_dict("gauss2", seed=2))
    net.eval()
    fe = ref.frontend(conf_thresh=0.015, nms_dist=4)
    fe.net = net
    warped = ref.inv_warp_image_batch(img.repeat(N, 1, 1, 1), inv, mode="bilinear")
    mask = ref.compute_valid_mask(torch.tensor([H, W]), inv_homography=inv, erosion_radius=0)[:, None]
    heat = fe.run(warped, onlyHeatmap=True, train=False)
    agg = ref.combine_heatmap(heat, hs[None], mask, device="cpu")
    with refshim.stable_argsort():
        pts = fe.getPtsFromHeatmap(agg.detach().cpu().squeeze().numpy()).transpose()
    np.savez_compressed(os.path.join(HERE, "ha_e2e.npz"), img=img.numpy(), h=hs.numpy(), inv=inv.numpy(),
                        agg=agg[0].numpy(), pts=pts[:600])

    # ---- descriptor matching artefacts shipped by the reference (models/model_wrap.py:437-480) ----
    k = np.load(os.path.join(REF, "datasets", "kitti", "kitti_test", "0000000000.npz"))
    sel1, sel2 = slice(0, 300), slice(0, 300)
    tr = ref.PointTracker(max_length=2, nn_thresh=0.7)
    d1, d2 = k["desc1"][:, sel1].astype(np.float32), k["desc2"][:, sel2].astype(np.float32)
    m = tr.nn_match_two_way(d1, d2, 0.7)
    np.savez_compressed(os.path.join(HERE, "match.npz"), desc1=d1.astype(np.float16), desc2=d2.astype(np.float16),
                        note=np.array("fp16-rounded copies of kitti_test desc1/desc2[:, :300]; matches recomputed "
                                      "by the reference on the rounded inputs"),
                        matches=ref.PointTracker(2, 0.7).nn_match_two_way(
                            d1.astype(np.float16).astype(np.float32), d2.astype(np.float16).astype(np.float32), 0.7))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
