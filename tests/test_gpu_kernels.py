"""GPU parity tests: every CUDA kernel (through the C ABI) against the CPU oracle on the same seeded
inputs, against the golden fixtures produced by the unmodified reference, and -- at full size -- through
size-independent properties.  Run on the B200 box with ``pytest -m gpu``."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402  (checker only)


@pytest.fixture(scope="module")
def spb():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import pytorch_superpoint_b200 as pkg
    from pytorch_superpoint_b200 import _lib
    assert _lib.lib().spb_version() >= 100  # the in-tree .so really loaded
    return pkg


def _mats(n, seed):
    from pytorch_superpoint_b200.homographies import make_ha_homographies
    return make_ha_homographies(n, seed=seed)


# ---------------------------------------------------------------------------------------------- warp
def test_warp_matches_oracle_bitexact(spb, golden):
    g = golden("warp")
    img, inv = torch.from_numpy(g["img"]).cuda(), torch.from_numpy(g["inv"]).cuda()
    w = spb.inv_warp_image_batch(img, inv, mode="bilinear")[:, 0].cpu().numpy()
    assert np.array_equal(w, O.inv_warp_image_batch(g["img"], g["inv"], "bilinear"))  # same fp32 op order
    assert np.abs(w - g["warped"]).max() < 5e-5  # vs the reference itself (fp32 op-order tolerance)
    n = spb.inv_warp_image_batch(img, inv, mode="nearest")[:, 0].cpu().numpy()
    assert np.array_equal(n, O.inv_warp_image_batch(g["img"], g["inv"], "nearest"))
    m = spb.compute_valid_mask((48, 64), inv).cpu().numpy()
    assert np.array_equal(m, g["mask"])  # reference's own mask, exact
    m2 = spb.compute_valid_mask((48, 64), inv, erosion_radius=2).cpu().numpy()
    assert np.array_equal(m2, g["mask_e2"])


def test_warp_batched_images_and_odd_width(spb):
    rng = np.random.RandomState(0)
    h, inv = _mats(5, 4)
    imgs = rng.rand(5, 37, 53).astype(np.float32)
    w = spb.inv_warp_image_batch(torch.from_numpy(imgs)[:, None].cuda(), torch.from_numpy(inv).cuda())
    assert np.array_equal(w[:, 0].cpu().numpy(), O.inv_warp_image_batch(imgs, inv, "bilinear"))


def test_valid_mask_full_size_vs_reference(spb, golden):
    g = golden("mask_240x320_n100")
    ref = np.unpackbits(g["mask_bits"])[: np.prod(g["shape"])].reshape(g["shape"]).astype(np.float32)
    m = spb.compute_valid_mask((240, 320), torch.from_numpy(g["inv"]).cuda()).cpu().numpy()
    assert np.array_equal(m, O.compute_valid_mask((240, 320), g["inv"]))
    assert int((m != ref).sum()) <= 8


def test_valid_mask_erosion_on_device_vs_reference(spb, golden):
    """utils/utils.py:699-702 on the GPU (spb_erode, no host round trip): the reference's own known-answer artefact
    notebooks/h2.npz (erosion 3 -> space-to-depth -> product over the cell, test/sample_homography.py:85-113) and
    cv2.erode itself for radii 1..6 on random masks (border handling, even-sized elliptic element, anchor)."""
    import cv2
    g = golden("ref_h2")
    inv = torch.inverse(torch.tensor(g["homographies"], dtype=torch.float32))[None]  # as the oracle's KAT test
    m = spb.compute_valid_mask((240, 320), inv.cuda(), erosion_radius=3)
    assert m.is_cuda
    cells = m[0].cpu().numpy().reshape(30, 8, 40, 8).transpose(0, 2, 1, 3).reshape(30, 40, 64).prod(-1)
    assert np.array_equal(cells[None], g["masks"])
    _, mats = _mats(6, 12)
    for r in range(1, 7):
        plain = spb.compute_valid_mask((72, 104), torch.from_numpy(mats).cuda()).cpu().numpy()
        got = spb.compute_valid_mask((72, 104), torch.from_numpy(mats).cuda(), erosion_radius=r).cpu().numpy()
        k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2 * r, 2 * r))
        assert np.array_equal(got, np.stack([cv2.erode(p, k, iterations=1) for p in plain]))


def test_warp_degenerate_homography(spb):
    """w -> 0 / NaN coordinates must produce zeros (zero padding), not crash."""
    M = np.stack([np.eye(3), np.array([[1, 0, 0], [0, 1, 0], [1.0, 0, 0]]), np.zeros((3, 3))]).astype(np.float32)
    img = np.random.RandomState(1).rand(24, 32).astype(np.float32)
    w = spb.inv_warp_image_batch(torch.from_numpy(img).cuda(), torch.from_numpy(M).cuda())[:, 0].cpu().numpy()
    assert np.array_equal(w, O.inv_warp_image_batch(img, M, "bilinear"))
    assert np.all(w[2] == 0)


# ---------------------------------------------------------------------------------------------- heat
def test_flatten_detection(spb, golden):
    g = golden("heat")
    semi = torch.from_numpy(g["semi"]).cuda()
    h = spb.flattenDetection(semi, tensor=True)[:, 0].cpu().numpy()
    assert np.abs(h - g["heat"]).max() < 1e-6
    assert np.abs(h - O.flatten_detection(g["semi"])).max() < 1e-6
    h1 = spb.flattenDetection(semi[0], tensor=True).cpu().numpy()
    assert h1.shape == (1, 48, 64) and np.array_equal(h1[0], h[0])


def test_combine_heatmap(spb, golden):
    g = golden("heat")
    out = spb.combine_heatmap(torch.from_numpy(g["heat"])[:, None].cuda(), torch.from_numpy(g["h"])[None].cuda(),
                              torch.from_numpy(g["mask"])[:, None].cuda())[0].cpu().numpy()
    assert np.array_equal(out, O.combine_heatmap(g["heat"], g["h"], g["mask"]))  # same op order: bit-exact
    assert np.abs(out - g["combined"]).max() < 2e-5


@pytest.mark.parametrize("H,W,N,seed", [(48, 64, 6, 3), (64, 96, 13, 5), (240, 320, 100, 0)])
def test_fused_aggregate(spb, H, W, N, seed):
    """spb_ha_aggregate == flattenDetection * valid_mask -> combine_heatmap of the oracle."""
    from pytorch_superpoint_b200 import ops
    h, inv = _mats(N, seed)
    semi = (torch.randn(N, 65, H // 8, W // 8, generator=torch.Generator().manual_seed(seed)) * 2).numpy()
    nhwc = torch.from_numpy(np.ascontiguousarray(semi.transpose(0, 2, 3, 1))).cuda()
    sums = ops.ha_aggregate(nhwc, torch.from_numpy(h), torch.from_numpy(inv), H, W)
    heat = ops.ha_finalize(sums).cpu().numpy()
    if N <= 16:
        ref = O.combine_heatmap(O.flatten_detection(semi), h, O.compute_valid_mask((H, W), inv))
    else:  # full size: use the (already verified) unfused CUDA kernels as the comparison path
        hm = spb.flattenDetection(torch.from_numpy(semi).cuda(), tensor=True)
        mk = spb.compute_valid_mask((H, W), torch.from_numpy(inv).cuda())[:, None]
        ref = spb.combine_heatmap(hm, torch.from_numpy(h)[None].cuda(), mk)[0].cpu().numpy()
    ok = np.isfinite(ref)
    assert np.array_equal(np.isfinite(heat), ok)
    assert np.abs(heat[ok] - ref[ok]).max() < 2e-6
    # chunked accumulation gives the same sums
    k = N // 2
    s2 = ops.ha_aggregate(nhwc[:k].contiguous(), torch.from_numpy(h[:k]), torch.from_numpy(inv[:k]), H, W)
    s2 = ops.ha_aggregate(nhwc[k:].contiguous(), torch.from_numpy(h[k:]), torch.from_numpy(inv[k:]), H, W, sums=s2,
                          accumulate=True)
    assert torch.allclose(s2, sums, rtol=1e-5, atol=1e-6)


def test_fused_aggregate_extreme_scale_fallback(spb):
    """A strongly magnifying un-warp makes a 16x16 tile cover > 64 cells: exercises the global fallback path."""
    from pytorch_superpoint_b200 import ops
    H, W, N = 96, 128, 3
    s = np.array([[6.0, 0, 0], [0, 6.0, 0], [0, 0, 1]], np.float32)
    h = np.stack([np.eye(3, dtype=np.float32), s, np.linalg.inv(s).astype(np.float32)])
    inv = np.stack([np.linalg.inv(x) for x in h]).astype(np.float32)
    semi = (torch.randn(N, 65, H // 8, W // 8, generator=torch.Generator().manual_seed(2)) * 2).numpy()
    nhwc = torch.from_numpy(np.ascontiguousarray(semi.transpose(0, 2, 3, 1))).cuda()
    heat = ops.ha_finalize(ops.ha_aggregate(nhwc, torch.from_numpy(h), torch.from_numpy(inv), H, W)).cpu().numpy()
    ref = O.combine_heatmap(O.flatten_detection(semi), h, O.compute_valid_mask((H, W), inv))
    assert np.abs(heat - ref).max() < 2e-6


# ------------------------------------------------------------------- fused-pass warp (folded coordinates)
def _packed_mask(spb, H, W, inv):
    m = spb.compute_valid_mask((H, W), torch.from_numpy(inv).cuda()).cpu().numpy().astype(np.uint8)
    return np.packbits(m, axis=-1, bitorder="little")


@pytest.mark.parametrize("H,W,N,seed", [(240, 320, 800, 1), (480, 640, 96, 2), (384, 1248, 64, 3), (48, 64, 40, 4),
                                        (40, 24, 9, 5), (64, 96, 33, 6)])
def test_fast_warp_views_mask_bits_exact_and_values_close(spb, H, W, N, seed):
    """The warp of the fused HA pass (ha_fast.cu: folded homography, reciprocal, magic-number floor) against the
    bit-exact operators: the packed valid masks must be IDENTICAL to compute_valid_mask (ambiguous pixels are
    re-evaluated with the exact expression), the views agree with inv_warp_image_batch to ~1e-3."""
    from pytorch_superpoint_b200 import ops
    rng = np.random.RandomState(seed)
    img = rng.rand(H, W).astype(np.float32)
    _, inv = _mats(N, seed)
    views, bits = ops.ha_warp_views(torch.from_numpy(img).cuda(), torch.from_numpy(inv).cuda())
    assert np.array_equal(bits.cpu().numpy(), _packed_mask(spb, H, W, inv))
    exact = spb.inv_warp_image_batch(torch.from_numpy(img).cuda(), torch.from_numpy(inv).cuda())[:, 0]
    assert (views - exact).abs().max().item() < 2e-3  # white-noise image: gradients of ~1 per pixel


def test_fast_warp_views_reference_masks_and_batched_images(spb, golden):
    """Mask bits against the masks of the UNMODIFIED reference at the bench shape (100 views of 240x320), and a
    batch of images with a per-view image index."""
    from pytorch_superpoint_b200 import ops
    g = golden("mask_240x320_n100")
    ref = np.unpackbits(g["mask_bits"])[: np.prod(g["shape"])].reshape(g["shape"]).astype(np.uint8)
    _, bits = ops.ha_warp_views(torch.zeros(240, 320).cuda(), torch.from_numpy(g["inv"]).cuda())
    got = np.unpackbits(bits.cpu().numpy(), axis=-1, bitorder="little")
    assert np.array_equal(got, O.compute_valid_mask((240, 320), g["inv"]).astype(np.uint8))  # exact vs the oracle
    assert int((got != ref).sum()) <= 8                                                      # and the reference
    rng = np.random.RandomState(7)
    imgs = rng.rand(3, 64, 96).astype(np.float32)
    _, inv = _mats(7, 9)
    vi = np.array([0, 0, 1, 2, 2, 2, 1], np.int32)
    views, bits = ops.ha_warp_views(torch.from_numpy(imgs).cuda(), torch.from_numpy(inv).cuda(), view_image=vi)
    for v in range(7):
        exact = O.inv_warp_image_batch(imgs[vi[v]][None], inv[v:v + 1], "bilinear")[0]
        assert np.abs(views[v].cpu().numpy() - exact).max() < 2e-3
    assert np.array_equal(bits.cpu().numpy(), _packed_mask(spb, 64, 96, inv))


def test_fast_warp_views_extreme_homographies(spb):
    """Strong perspective (denominator close to / through zero inside the frame), huge scales, a singular and a NaN
    matrix: the mask bits stay exactly those of the reference expression (these views fall back to it) and nothing
    reads out of bounds."""
    from pytorch_superpoint_b200 import ops
    H, W = 64, 96
    rng = np.random.RandomState(3)
    mats = [np.eye(3), np.diag([8.0, 8.0, 1.0]), np.diag([0.05, 0.05, 1.0]),
            np.array([[1, 0, 0], [0, 1, 0], [0.9, 0.0, 1.0]]),      # Z = 1 + 0.9 u: down to 0.1
            np.array([[1, 0, 0], [0, 1, 0], [1.2, 0.3, 1.0]]),      # Z changes sign inside the frame
            np.array([[1, 0.2, 0.5], [0.1, 1, -0.7], [0.0, 0.0, 1e-3]]),
            np.zeros((3, 3)), np.full((3, 3), np.nan),
            np.array([[1, 0, 1e-7], [0, 1, 0], [0, 0, 1]]),         # borders within rounding of the decision boundary
            np.array([[1.0000001, 0, 0], [0, 0.9999999, 0], [0, 0, 1]])]
    for _ in range(30):
        m = np.eye(3) + rng.randn(3, 3) * np.array([[0.3, 0.3, 0.5], [0.3, 0.3, 0.5], [0.6, 0.6, 0.2]])
        mats.append(m)
    inv = np.stack(mats).astype(np.float32)
    img = rng.rand(H, W).astype(np.float32)
    views, bits = ops.ha_warp_views(torch.from_numpy(img).cuda(), torch.from_numpy(inv).cuda())
    assert torch.isfinite(views).all()
    assert np.array_equal(bits.cpu().numpy(), _packed_mask(spb, H, W, inv))


# ---------------------------------------------------------------------------------------------- NMS
CASES = ["smooth_96x128", "flat_random_240x320", "sparse_64x80", "ties_72x88", "empty_40x48", "single_40x48",
         "border_40x48"]


@pytest.mark.parametrize("case", CASES)
def test_get_pts_golden_bit_exact(spb, golden, case):
    g = golden("pts")
    pts = spb.getPtsFromHeatmap(torch.from_numpy(g["heat_" + case]).cuda(), 0.015, 4)
    ref = g["pts_" + case]
    assert pts.shape == ref.shape and pts.dtype == np.float64
    assert np.array_equal(pts, ref)


@pytest.mark.parametrize("H,W,dist,thr", [(240, 320, 4, 0.015), (480, 640, 4, 0.015), (384, 1248, 4, 0.015),
                                          (120, 160, 1, 0.3), (64, 64, 0, 0.5), (96, 100, 8, 0.1)])
def test_get_pts_vs_oracle_random(spb, H, W, dist, thr):
    rng = np.random.RandomState(H + W + dist)
    import cv2
    heat = cv2.GaussianBlur(rng.rand(H, W).astype(np.float32), (5, 5), 0)
    heat = (heat * (2 * thr)).astype(np.float32)
    pts = spb.getPtsFromHeatmap(torch.from_numpy(heat).cuda(), thr, dist)
    assert np.array_equal(pts, O.get_pts_from_heatmap(heat, thr, dist))


def test_get_pts_batch_topk_and_properties(spb):
    from pytorch_superpoint_b200 import ops
    rng = np.random.RandomState(3)
    heat = (0.0154 + 0.002 * rng.randn(8, 240, 320)).astype(np.float32)  # random-init-like near-uniform maps
    pts, counts, total = ops.get_pts_from_heatmap_device(torch.from_numpy(heat).cuda(), 0.015, 4, 4, top_k=600)
    pts, counts, total = pts.cpu().numpy(), counts.cpu().numpy(), total.cpu().numpy()
    for b in range(8):
        ref = O.get_pts_from_heatmap(heat[b], 0.015, 4).T
        assert total[b] == ref.shape[0] and counts[b] == min(600, ref.shape[0])
        assert np.array_equal(pts[b, :counts[b]].astype(np.float64), ref[:600])
        p = pts[b, :counts[b]]
        assert np.all(np.diff(p[:, 2]) <= 0)  # sorted
        assert p[:, 0].min() >= 4 and p[:, 0].max() < 316 and p[:, 1].min() >= 4 and p[:, 1].max() < 236
        d = np.abs(p[:, None, :2] - p[None, :, :2]).max(-1) + 1e9 * np.eye(len(p))
        assert d.min() > 4  # pairwise Chebyshev distance > nms_dist
    # idempotence: NMS of the kept points' sparse map keeps exactly those points
    sp = np.zeros((240, 320), np.float32)
    full = O.get_pts_from_heatmap(heat[0], 0.015, 4)
    sp[full[1].astype(int), full[0].astype(int)] = full[2]
    again = spb.getPtsFromHeatmap(torch.from_numpy(sp).cuda(), 0.015, 4)
    assert np.array_equal(again, full)


# ---------------------------------------------------------------------------------------------- descriptors
def test_sample_desc(spb, golden):
    g = golden("desc")
    d = spb.sample_desc_from_points(torch.from_numpy(g["coarse"]).cuda(), g["pts"])
    assert d.shape == g["desc"].shape and d.dtype == np.float32
    assert np.abs(d - g["desc"]).max() < 1e-6
    cos = (d * g["desc"]).sum(0)
    assert cos.min() > 0.999999
    assert spb.sample_desc_from_points(torch.from_numpy(g["coarse"]).cuda(), np.zeros((3, 0))).shape == (256, 0)
    # NHWC device path == NCHW path
    from pytorch_superpoint_b200 import ops
    p = torch.from_numpy(np.ascontiguousarray(g["pts"].T)).float().cuda()
    c = torch.from_numpy(g["coarse"]).cuda()[0]
    a = ops.sample_desc_device(c, p, nhwc=False)
    b = ops.sample_desc_device(c.permute(1, 2, 0).contiguous(), p, nhwc=True)
    assert torch.equal(a, b)


# ---------------------------------------------------------------------------------------------- ingest
@pytest.mark.parametrize("Hs,Ws,H,W", [(480, 640, 240, 320), (375, 500, 240, 320), (427, 640, 240, 320),
                                       (720, 960, 240, 320), (333, 500, 120, 160), (240, 320, 240, 320),
                                       (640, 480, 480, 480), (500, 375, 96, 64), (1080, 1920, 480, 640),
                                       (375, 1242, 384, 1248), (120, 160, 240, 320), (240, 500, 384, 400)])
def test_ingest_resize_gray_equals_opencv(spb, Hs, Ws, H, W):
    """f-4 image ingest on the GPU (ingest.cu) == cv2.resize(INTER_AREA) -> cv2.cvtColor(RGB2GRAY) -> float32 / 255,
    bit for bit (what datasets/Coco.py:165-179 computes on the host), for integer and fractional scale factors."""
    import cv2
    from pytorch_superpoint_b200 import ops
    rng = np.random.RandomState(Hs + W)
    img = rng.randint(0, 256, (Hs, Ws, 3)).astype(np.uint8)
    want = cv2.cvtColor(cv2.resize(img, (W, H), interpolation=cv2.INTER_AREA), cv2.COLOR_RGB2GRAY).astype("float32") / 255.0
    got = ops.ingest_resize_gray(img, (H, W))
    assert got.is_cuda and got.dtype == torch.float32 and tuple(got.shape) == (H, W)
    assert np.array_equal(got.cpu().numpy(), want)
    if Hs * Ws <= 500 * 640:
        assert np.array_equal(got.cpu().numpy(), O.ingest_resize_gray(img, H, W))
    with pytest.raises(ValueError):
        ops.ingest_resize_gray(img[..., :2], (H, W))


@pytest.mark.parametrize("cta", [0, 256, 512, 1024])
@pytest.mark.parametrize("B,H,W,dist", [(3, 120, 160, 4), (40, 48, 72, 4), (2, 96, 100, 2)])
def test_get_pts_every_cta_size_is_bit_identical(spb, cta, B, H, W, dist):
    """spb_get_pts_from_heatmap_ex: the CTA size of the NMS kernel (automatic, or chosen by the caller for a background
    launch) never changes the result -- every image identical to the oracle, for the reference's distance (separable
    window maxima by warp shuffles) and another one (generic path)."""
    from pytorch_superpoint_b200 import _lib
    L = _lib.lib()
    rng = np.random.RandomState(B + H + dist)
    heat = (rng.rand(B, H, W) * 0.04).astype(np.float32)
    heat[:, ::7, ::5] = np.round(heat[:, ::7, ::5] * 200) / 200          # some exact ties
    h = torch.from_numpy(heat).cuda()
    cap = -(-H // (dist + 1)) * -(-W // (dist + 1))
    pts = torch.zeros((B, cap, 3), device="cuda")
    cnt = torch.zeros((B,), dtype=torch.int32, device="cuda")
    n = L.spb_nms_workspace_bytes(B, H, W, dist)
    ws = torch.empty(n, dtype=torch.uint8, device="cuda")
    _lib.check(L.spb_get_pts_from_heatmap_ex(h.data_ptr(), B, H, W, 0.015, dist, 4, 0, cap, pts.data_ptr(), cnt.data_ptr(),
                                             None, ws.data_ptr(), n, cta, torch.cuda.current_stream().cuda_stream), "nms")
    for b in range(B):
        ref = O.get_pts_from_heatmap(heat[b], 0.015, dist, 4)
        k = int(cnt[b])
        assert k == ref.shape[1] and np.array_equal(pts[b, :k].cpu().numpy().astype(np.float64), ref.T)
    if cta == 0:
        assert L.spb_get_pts_from_heatmap_ex(h.data_ptr(), B, H, W, 0.015, dist, 4, 0, cap, pts.data_ptr(), cnt.data_ptr(),
                                             None, ws.data_ptr(), n, 300, None) != 0   # not a CTA size
