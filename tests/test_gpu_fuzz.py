"""Randomised parity sweeps (GPU): every operator of the path against the CPU oracle / OpenCV on RANDOM shapes, counts
and parameters -- odd sizes, empty and single-element inputs, ties, extreme homographies -- beyond the fixed cases of
test_gpu_kernels.py / test_gpu_frontend.py.  Seeds are fixed, so a failure reproduces; ``SPB_FUZZ_ITERS`` (default 10
rounds per operator, a few seconds each) widens the sweep for a manual run:

    SPB_FUZZ_ITERS=300 python -m pytest tests/test_gpu_fuzz.py -m gpu -q
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402  (checker only)

ITERS = int(os.environ.get("SPB_FUZZ_ITERS", "10"))


@pytest.fixture(scope="module")
def spb():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import pytorch_superpoint_b200 as pkg
    return pkg


def _random_mats(rng, n):
    """A mix of sampler draws (the export's distribution) and raw perturbations of the identity (strong perspective,
    reflections, near-singular) -> (homographies, inv_homographies) float32 [n,3,3]."""
    from pytorch_superpoint_b200.homographies import make_ha_homographies
    hs, inv = make_ha_homographies(n, seed=int(rng.randint(1 << 30)))
    hs, inv = hs.copy(), inv.copy()
    for i in range(n):
        if rng.rand() < 0.35:
            m = np.eye(3) + rng.randn(3, 3) * np.array([[0.4, 0.4, 0.6], [0.4, 0.4, 0.6], [0.5, 0.5, 0.2]]) * rng.rand()
            inv[i] = m.astype(np.float32)
            with np.errstate(all="ignore"):
                hs[i] = np.linalg.pinv(m).astype(np.float32)
    return hs, inv


def test_fuzz_warp_and_valid_mask(spb):
    """inv_warp_image_batch / compute_valid_mask bit-exact vs the oracle; the fused pass's warp: mask bits identical,
    values close."""
    from pytorch_superpoint_b200 import ops
    rng = np.random.RandomState(101)
    for it in range(ITERS):
        H, W, N = int(rng.randint(3, 130)), int(rng.randint(3, 170)), int(rng.randint(1, 14))
        img = rng.rand(H, W).astype(np.float32)
        _, inv = _random_mats(rng, N)
        got = spb.inv_warp_image_batch(torch.from_numpy(img).cuda(), torch.from_numpy(inv).cuda())[:, 0].cpu().numpy()
        ref = O.inv_warp_image_batch(img[None], inv, "bilinear")
        assert np.array_equal(got, ref, equal_nan=True), (it, H, W, N)
        near = spb.inv_warp_image_batch(torch.from_numpy(img).cuda(), torch.from_numpy(inv).cuda(), mode="nearest")
        assert np.array_equal(near[:, 0].cpu().numpy(), O.inv_warp_image_batch(img[None], inv, "nearest"), equal_nan=True)
        mask = spb.compute_valid_mask((H, W), torch.from_numpy(inv).cuda()).cpu().numpy()
        ref_mask = O.compute_valid_mask((H, W), inv)
        assert np.array_equal(mask, ref_mask), (it, H, W, N)
        if H < 8:  # the fused pass needs H >= 8 and W % 8 == 0 (the network needs multiples of 8 anyway)
            with pytest.raises(ValueError):
                ops.ha_warp_views(torch.from_numpy(img).cuda(), torch.from_numpy(inv).cuda())
            continue
        W8 = max(8, W // 8 * 8)
        img8 = rng.rand(H, W8).astype(np.float32)
        views, bits = ops.ha_warp_views(torch.from_numpy(img8).cuda(), torch.from_numpy(inv).cuda())
        want = np.packbits(O.compute_valid_mask((H, W8), inv).astype(np.uint8), axis=-1, bitorder="little")
        assert np.array_equal(bits.cpu().numpy(), want), (it, H, W8, N)
        assert torch.isfinite(views).all()


def test_fuzz_nms_topk(spb):
    """getPtsFromHeatmap: random sizes, thresholds, NMS distances 0..8, border widths, plateaus of exactly equal
    confidences -- identical points in identical order."""
    rng = np.random.RandomState(202)
    import cv2
    for it in range(ITERS * 2):
        H, W = int(rng.randint(9, 150)), int(rng.randint(9, 200))
        dist, border = int(rng.randint(0, 9)), int(rng.randint(0, 5))
        kind = it % 4
        heat = rng.rand(H, W).astype(np.float32)
        if kind == 1:
            heat = cv2.GaussianBlur(heat, (5, 5), 0)
        elif kind == 2:
            heat = np.round(heat * 6) / 6                      # many exact ties
        elif kind == 3:
            heat = (heat > 0.97).astype(np.float32) * rng.choice([0.5, 1.0], size=(H, W)).astype(np.float32)  # sparse
        thr = float(rng.choice([0.0, 0.015, 0.3, 0.5, 0.9, 2.0]))
        pts = spb.getPtsFromHeatmap(torch.from_numpy(heat.astype(np.float32)).cuda(), thr, dist, border)
        ref = O.get_pts_from_heatmap(heat.astype(np.float32), thr, dist, border)
        assert pts.shape == ref.shape and np.array_equal(pts, ref), (it, H, W, dist, border, thr, kind)
        k = int(rng.randint(1, 40))
        cut = spb.getPtsFromHeatmap(torch.from_numpy(heat.astype(np.float32)).cuda(), thr, dist, border, top_k=k)
        assert np.array_equal(cut, ref[:, :k]), (it, "top_k", k)


def test_fuzz_matcher(spb):
    """nn_match_two_way: random counts (0, 1, odd), duplicated columns (exact ties), thresholds."""
    rng = np.random.RandomState(303)
    for it in range(ITERS * 2):
        D = 256
        K1, K2 = int(rng.choice([0, 1, 2, 7, 63, 64, 65, 130, 301])), int(rng.choice([0, 1, 3, 64, 100, 257]))
        d1 = rng.randn(D, K1).astype(np.float32)
        d2 = rng.randn(D, K2).astype(np.float32)
        if K1 and K2 and it % 2:
            d2[:, : min(K1, K2) // 2] = d1[:, : min(K1, K2) // 2] + 0.05 * rng.randn(D, min(K1, K2) // 2)  # real matches
        if K2 > 4:
            d2[:, 3] = d2[:, 1]                                # exact duplicate: argmin must take the first
        with np.errstate(all="ignore"):
            d1 /= np.linalg.norm(d1, axis=0, keepdims=True)
            d2 /= np.linalg.norm(d2, axis=0, keepdims=True)
        thr = float(rng.choice([0.0, 0.3, 0.7, 1.0, 1.5, 3.0]))
        got = spb.nn_match_two_way(d1, d2, thr)
        ref = O.nn_match_two_way(d1, d2, thr)
        assert got.shape == ref.shape, (it, K1, K2, thr)
        assert np.array_equal(got[:2], ref[:2]), (it, K1, K2, thr)
        assert np.abs(got[2] - ref[2]).max(initial=0.0) < 2e-3   # sqrt(2 - 2 s) amplifies fp32 round-off of s near 1


def test_fuzz_sample_desc(spb):
    """sample_desc_from_points: random grid sizes and point sets (borders, corners, fractional positions)."""
    rng = np.random.RandomState(404)
    for it in range(ITERS):
        Hc, Wc, D = int(rng.randint(2, 40)), int(rng.randint(2, 50)), 256
        K = int(rng.choice([0, 1, 5, 200, 1000]))
        coarse = rng.randn(1, D, Hc, Wc).astype(np.float32)
        pts = np.zeros((3, K))
        pts[0] = rng.randint(0, Wc * 8, K) + (rng.rand(K) < 0.3) * rng.rand(K)
        pts[1] = rng.randint(0, Hc * 8, K) + (rng.rand(K) < 0.3) * rng.rand(K)
        pts[0] = np.minimum(pts[0], Wc * 8 - 1)
        pts[1] = np.minimum(pts[1], Hc * 8 - 1)
        got = spb.sample_desc_from_points(torch.from_numpy(coarse).cuda(), pts)
        ref = O.sample_desc_from_points(coarse, pts)
        assert got.shape == ref.shape, (it, Hc, Wc, K)
        if K:
            assert np.abs(np.asarray(got) - ref).max() < 1e-5, (it, Hc, Wc, K)  # fp32 round-off of the L2 normalisation


def test_fuzz_ingest_and_erosion(spb):
    """Image ingest vs OpenCV (random source / target sizes, up- and down-scaling) and mask erosion vs cv2.erode."""
    import cv2
    from pytorch_superpoint_b200 import ops
    rng = np.random.RandomState(505)
    for it in range(ITERS):
        Hs, Ws = int(rng.randint(17, 400)), int(rng.randint(17, 500))
        H, W = int(rng.randint(2, 40)) * 8, int(rng.randint(2, 50)) * 8
        img = rng.randint(0, 256, (Hs, Ws, 3)).astype(np.uint8)
        want = cv2.cvtColor(cv2.resize(img, (W, H), interpolation=cv2.INTER_AREA), cv2.COLOR_RGB2GRAY).astype("float32") / 255.0
        got = ops.ingest_resize_gray(img, (H, W)).cpu().numpy()
        assert np.array_equal(got, want), (it, Hs, Ws, H, W)
    for it in range(ITERS):
        H, W, N, r = int(rng.randint(5, 90)), int(rng.randint(5, 120)), int(rng.randint(1, 5)), int(rng.randint(1, 8))
        _, inv = _random_mats(rng, N)
        plain = spb.compute_valid_mask((H, W), torch.from_numpy(inv).cuda()).cpu().numpy()
        got = spb.compute_valid_mask((H, W), torch.from_numpy(inv).cuda(), erosion_radius=r).cpu().numpy()
        k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2 * r, 2 * r))
        assert np.array_equal(got, np.stack([cv2.erode(p, k, iterations=1) for p in plain])), (it, H, W, N, r)


def test_fuzz_ha_export_small(spb):
    """Whole HA export on random small shapes / view counts / batch sizes vs the fp32 oracle: heat-map < 1e-2, and the
    points are exactly the reference's NMS / top-k of OUR heat-map."""
    from pytorch_superpoint_b200 import SuperPointNet_b200
    from pytorch_superpoint_b200.export import HAExporter
    from pytorch_superpoint_b200.homographies import make_ha_homographies
    rng = np.random.RandomState(606)
    sd = O.synthetic_state_dict("gauss2", seed=8)
    net = SuperPointNet_b200()
    net.load_state_dict(sd)
    params = O.canonical_params(sd)
    for it in range(max(3, ITERS // 2)):
        H, W = int(rng.randint(2, 12)) * 8, int(rng.randint(2, 14)) * 8
        N, B = int(rng.randint(1, 11)), int(rng.randint(1, 4))
        top_k = int(rng.choice([0, 5, 600]))
        imgs = []
        for _ in range(B):
            img = np.clip(0.4 + 0.05 * rng.randn(H, W), 0, 1).astype(np.float32)
            for _ in range(6):
                x0, y0 = rng.randint(0, W - 4), rng.randint(0, H - 4)
                img[y0:y0 + rng.randint(3, H), x0:x0 + rng.randint(3, W)] = rng.uniform(0, 1)
            imgs.append(img)
        mu, mf = make_ha_homographies(N, seed=int(rng.randint(1 << 30)))
        ex = HAExporter(net, 0.015, 4, top_k)
        out = ex.export_batch([torch.from_numpy(i) for i in imgs], torch.from_numpy(mu), torch.from_numpy(mf))
        heat = ex.last_heat(B, H, W).cpu().numpy()
        for b in range(B):
            _, ref_heat = O.ha_export_one(imgs[b], mu, mf, params, top_k=top_k or 10 ** 9, return_heat=True)
            assert np.abs(heat[b] - ref_heat).max() < 1e-2, (it, H, W, N, B, b)
            want = O.get_pts_from_heatmap(heat[b], 0.015, 4).T
            want = want[:top_k] if top_k else want
            assert np.array_equal(out[b], want), (it, H, W, N, B, b, top_k)
