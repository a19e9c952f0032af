"""GPU tests of the reference-facing front-end (SuperPointFrontend_b200), the sub-pixel step, the exporter
and the .npz drop-in entry point."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402  (checker only)


def _fe(**kw):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200.frontend import SuperPointFrontend_b200
    fe = SuperPointFrontend_b200({"model": {"subpixel": {"enable": False}}}, "", 4, 0.015, 0.7, load=False, **kw)
    return fe


def test_nms_fast_generic_matches_reference_golden(golden):
    g = golden("pts")
    fe = _fe()
    out, inds = fe.nms_fast(g["nms_in"], 64, 80, 3)
    assert np.array_equal(out, g["nms_out"]) and np.array_equal(inds, g["nms_inds"])
    e, ei = fe.nms_fast(np.zeros((3, 0)), 64, 80, 3)
    assert e.shape == (3, 0) and ei.shape == (0,)
    one, oi = fe.nms_fast(np.array([[3.4], [5.6], [0.9]]), 64, 80, 3)
    assert np.array_equal(one, O.nms_fast(np.array([[3.4], [5.6], [0.9]]), 64, 80, 3)[0])


def test_frontend_run_and_checkpoint_loading(tmp_path):
    from pytorch_superpoint_b200.frontend import SuperPointFrontend_b200
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    sd = O.synthetic_state_dict("gauss2", seed=4)
    tar = str(tmp_path / "superPointNet_100_checkpoint.pth.tar")
    torch.save({"n_iter": 100, "model_state_dict": sd, "optimizer_state_dict": {}, "loss": 0.0}, tar)
    fe = SuperPointFrontend_b200({"model": {"subpixel": {"enable": False}}}, tar, 4, 0.015, 0.7, device="cuda")
    bare = str(tmp_path / "superpoint_v1.pth")
    torch.save(O.synthetic_state_dict("v1", seed=4), bare)
    SuperPointFrontend_b200({"model": {"subpixel": {"enable": False}}}, bare, 4, 0.015, 0.7, device="cuda")
    x = torch.rand(2, 1, 64, 96, generator=torch.Generator().manual_seed(1))
    heat = fe.run(x, onlyHeatmap=True, train=False)
    assert tuple(heat.shape) == (2, 1, 64, 96) and heat.is_cuda
    params = O.canonical_params(sd)
    semi, desc = O.net_forward(params, x)
    ref_heat = O.flatten_detection(semi.numpy())
    assert np.abs(heat[:, 0].cpu().numpy() - ref_heat).max() < 1e-2
    pts, pts_desc, dense, heat2 = fe.run(x, onlyHeatmap=False, train=False)
    # NB the reference's own indexing (models/model_wrap.py:404) yields [256,K] per image, kept as is
    assert len(pts) == 2 and pts[0].shape[0] == 3 and pts_desc[0].shape == (256, pts[0].shape[1])
    assert tuple(dense.shape) == (2, 256, 64, 96)
    # key-points bit-exact given OUR heat-map (the CPU oracle on the same fp32 array)
    for b in range(2):
        assert np.array_equal(pts[b], O.get_pts_from_heatmap(heat2[b, 0].cpu().numpy(), 0.015, 4))
    # sparse descriptors at those points vs the oracle on the fp32 reference descriptors
    d = fe.sample_desc_from_points(torch.from_numpy(desc.numpy()).cuda(), pts[0])
    ref_d = O.sample_desc_from_points(desc.numpy(), pts[0])
    assert np.abs(d - ref_d).max() < 1e-6


def test_subpixel_vs_oracle():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200.subpixel import soft_argmax_points
    rng = np.random.RandomState(0)
    import cv2
    heat = cv2.GaussianBlur(rng.rand(64, 80).astype(np.float32), (5, 5), 0) * 0.05
    pts = O.get_pts_from_heatmap(heat, 0.015, 4, border_remove=0)  # includes points on the border (zero padding)
    out = soft_argmax_points(torch.from_numpy(heat).cuda(), pts, 5)
    ref = O.soft_argmax_points(heat, pts, 5)
    assert out.shape == ref.shape and np.abs(out - ref).max() < 1e-4
    assert np.abs(out[:2] - pts[:2]).max() < 2.0


def test_export_entry_writes_reference_npz(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import SuperPointNet_b200, make_ha_homographies
    from pytorch_superpoint_b200.export import export_detector_homoAdapt_b200
    sd = O.synthetic_state_dict("gauss2", seed=2)
    net = SuperPointNet_b200()
    net.load_state_dict(sd)
    H, W, N = 64, 96, 8
    rng = np.random.RandomState(3)
    samples = []
    for i in range(2):
        mu, mf = make_ha_homographies(N, seed=20 + i)
        samples.append({"image_2D": torch.from_numpy(rng.rand(1, 1, H, W).astype(np.float32)), "name": [f"img{i}"],
                        "homographies": torch.from_numpy(mu)[None], "inv_homographies": torch.from_numpy(mf)[None]})
    cfg = {"model": {"nms": 4, "top_k": 50, "detection_threshold": 0.015, "subpixel": {"enable": False}},
           "data": {"export_folder": "train", "dataset": "Coco",
                    "homography_adaptation": {"num": N, "homographies": {"params": {}}}}}
    n = export_detector_homoAdapt_b200(cfg, str(tmp_path), loader=samples, net=net)
    assert n == 2
    from pytorch_superpoint_b200.export import HAExporter
    single = HAExporter(net, 0.015, 4, 50)
    for i, s in enumerate(samples):
        f = np.load(os.path.join(str(tmp_path), "predictions", "train", f"img{i}.npz"))
        pts = f["pts"]
        assert list(f.keys()) == ["pts"] and pts.dtype == np.float64 and pts.shape[1] == 3 and pts.shape[0] <= 50
        ref, agg = O.ha_export_one(s["image_2D"][0, 0].numpy(), s["homographies"][0].numpy(),
                                   s["inv_homographies"][0].numpy(), O.canonical_params(sd), top_k=50, return_heat=True)
        # (a) the aggregated heat-map behind the file is the oracle's within the bf16 tolerance ...
        again = single.export_image(s["image_2D"][0, 0], s["homographies"][0], s["inv_homographies"][0])
        heat = single.last_heat(1, H, W)[0].cpu().numpy()
        assert np.abs(heat - agg).max() < 1e-2
        # (b) ... the file holds EXACTLY the reference's NMS / ordering / top-k of that heat-map (export.py:311-328) ...
        assert np.array_equal(pts, O.get_pts_from_heatmap(heat, 0.015, 4).T[:50]) and np.array_equal(pts, again)
        # (c) ... and fed the oracle's heat-map the operator returns the oracle's points bit for bit
        from pytorch_superpoint_b200 import getPtsFromHeatmap
        assert np.array_equal(getPtsFromHeatmap(torch.from_numpy(agg).cuda(), 0.015, 4).T[:50], ref)
        # (d) random-init weights give a near-flat heat-map just above the threshold, so the SETS differ where the bf16
        # heat crosses the oracle's ranking; the confidences of the kept points still agree to the heat tolerance
        assert abs(len(pts) - len(ref)) <= 5 and np.abs(np.sort(pts[:, 2])[-10:] - np.sort(ref[:, 2])[-10:]).max() < 1e-2
    # resume: files exist -> nothing re-exported (export.py:302-306)
    assert export_detector_homoAdapt_b200(cfg, str(tmp_path), loader=samples, net=net) == 0


def test_export_entry_train_mode_bn_from_config(tmp_path):
    """``model.b200_bn_mode: 'batch'`` in the YAML of the export entry: the checkpoint is loaded into the train-mode
    BatchNorm path (what the reference's export actually runs, model_wrap.py:111) and the written points are the
    oracle's run with bn_mode='batch' -- heat-map within 1e-3, the same key-point set."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import make_ha_homographies
    from pytorch_superpoint_b200.export import export_detector_homoAdapt_b200
    sd = O.synthetic_state_dict("gauss2", seed=4)
    ckpt = os.path.join(str(tmp_path), "ckpt.pth.tar")
    torch.save({"model_state_dict": sd}, ckpt)
    H, W, N = 64, 96, 16
    rng = np.random.RandomState(5)
    img = np.clip(0.4 + 0.05 * rng.randn(H, W), 0, 1).astype(np.float32)
    for _ in range(8):
        x0, y0 = rng.randint(0, W - 10), rng.randint(0, H - 10)
        img[y0:y0 + rng.randint(5, 20), x0:x0 + rng.randint(5, 30)] = rng.uniform(0, 1)
    mu, mf = make_ha_homographies(N, seed=9)
    samples = [{"image_2D": torch.from_numpy(img)[None, None], "name": ["a"], "homographies": torch.from_numpy(mu)[None],
                "inv_homographies": torch.from_numpy(mf)[None]}]
    cfg = {"model": {"nms": 4, "top_k": 100, "detection_threshold": 0.015, "subpixel": {"enable": False},
                     "b200_bn_mode": "batch"},
           "pretrained": ckpt,
           "data": {"export_folder": "train", "dataset": "Coco",
                    "homography_adaptation": {"num": N, "homographies": {"params": {}}}}}
    assert export_detector_homoAdapt_b200(cfg, str(tmp_path), loader=samples) == 1
    pts = np.load(os.path.join(str(tmp_path), "predictions", "train", "a.npz"))["pts"]
    ref, ref_heat = O.ha_export_one(img, mu, mf, O.canonical_params(sd), top_k=100, bn_mode="batch", return_heat=True)
    ref_eval = O.ha_export_one(img, mu, mf, O.canonical_params(sd), top_k=100, bn_mode="eval")
    a, b = set(map(tuple, pts[:, :2])), set(map(tuple, ref[:, :2]))
    assert len(a & b) >= 0.97 * len(a | b)
    common = sorted(a & b)
    conf = {tuple(p[:2]): p[2] for p in pts}
    conf_ref = {tuple(p[:2]): p[2] for p in ref}
    assert max(abs(conf[k] - conf_ref[k]) for k in common) < 1e-3
    assert len(a & set(map(tuple, ref_eval[:, :2]))) < len(a & b)  # not the eval-mode answer


def test_export_headline_config_vs_oracle():
    """The BASELINE configuration itself: one 240x320 image, 100 homographies, through HAExporter.export_batch
    against the CPU oracle's ha_export_one -- aggregated heat-map within 1e-2, the mask-sum plane within fp32
    round-off of the oracle's (exact mask bits, folded un-warp coordinates), key-points exactly the reference's
    NMS / top-k of our heat-map, and bit-exact when the operator is fed the oracle's heat-map."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import SuperPointNet_b200, getPtsFromHeatmap, make_ha_homographies
    from pytorch_superpoint_b200.export import HAExporter
    H, W, N = 240, 320, 100
    rng = np.random.RandomState(11)
    img = np.full((H, W), 0.35, np.float32) + 0.03 * rng.randn(H, W).astype(np.float32)
    for _ in range(14):
        x0, y0 = rng.randint(0, W - 10), rng.randint(0, H - 10)
        img[y0:y0 + rng.randint(8, H // 3), x0:x0 + rng.randint(8, W // 3)] = rng.uniform(0, 1)
    img = np.clip(img, 0, 1)
    mu, mf = make_ha_homographies(N, seed=0)
    sd = O.synthetic_state_dict("gauss2", seed=1)
    net = SuperPointNet_b200()
    net.load_state_dict(sd)
    exp = HAExporter(net, 0.015, 4, 600)
    pts = exp.export_batch([torch.from_numpy(img)], torch.from_numpy(mu), torch.from_numpy(mf))[0]
    heat = exp.last_heat(1, H, W)[0].cpu().numpy()
    sums = exp._buffers(1, H, W)["slots"][(exp._buffers(1, H, W)["k"] - 1) & 1]["sums"][0].cpu().numpy()
    ref_pts, ref_heat = O.ha_export_one(img, mu, mf, O.canonical_params(sd), return_heat=True)
    assert np.isfinite(heat).all() and np.abs(heat - ref_heat).max() < 1e-2
    mask = O.compute_valid_mask((H, W), mf)
    ref_den = O.inv_warp_image_batch(mask, mu, "bilinear").sum(0)
    assert np.abs(sums[1] - ref_den).max() < 2e-2 and np.abs(sums[1] - ref_den).mean() < 1e-3  # of values up to 100
    assert pts.dtype == np.float64 and pts.shape[1] == 3 and pts.shape[0] <= 600
    assert np.array_equal(pts, O.get_pts_from_heatmap(heat, 0.015, 4).T[:600])
    assert np.array_equal(getPtsFromHeatmap(torch.from_numpy(ref_heat).cuda(), 0.015, 4).T[:600], ref_pts)


def test_export_entry_with_subpixel_refinement(tmp_path):
    """The reference's default export config has model.subpixel.enable: true (export.py:314-319): the refinement runs
    on the device inside the ticket (one launch per owned image on the side stream) and must equal the oracle's
    soft_argmax_points of OUR heat-map at OUR integer points (torchgeometry restated: parity unpinned)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import SuperPointNet_b200, make_ha_homographies
    from pytorch_superpoint_b200.export import HAExporter, export_detector_homoAdapt_b200
    net = SuperPointNet_b200()
    net.load_state_dict(O.synthetic_state_dict("gauss2", seed=2))
    H, W, N = 64, 96, 6
    rng = np.random.RandomState(5)
    samples = []
    for i in range(3):
        mu, mf = make_ha_homographies(N, seed=40 + i)
        samples.append({"image_2D": torch.from_numpy(rng.rand(1, 1, H, W).astype(np.float32)), "name": [f"s{i}"],
                        "homographies": torch.from_numpy(mu)[None], "inv_homographies": torch.from_numpy(mf)[None]})
    cfg = {"model": {"nms": 4, "top_k": 70, "detection_threshold": 0.015, "subpixel": {"enable": True, "patch_size": 5},
                     "b200_batch_images": 2},
           "data": {"export_folder": "train", "dataset": "Coco",
                    "homography_adaptation": {"num": N, "homographies": {"params": {}}}}}
    assert export_detector_homoAdapt_b200(cfg, str(tmp_path), loader=samples, net=net) == 3
    plain = HAExporter(net, 0.015, 4, 70)
    for i, s in enumerate(samples):
        got = np.load(os.path.join(str(tmp_path), "predictions", "train", f"s{i}.npz"))["pts"]
        ints = plain.export_image(s["image_2D"][0, 0], s["homographies"][0], s["inv_homographies"][0])
        heat = plain.last_heat(1, H, W)[0].cpu().numpy()
        ref = O.soft_argmax_points(heat, ints.T, 5).T
        assert got.shape == ref.shape and got.dtype == np.float64
        assert np.abs(got - ref).max() < 1e-4 and np.array_equal(got[:, 2], ints[:, 2])
        assert np.abs(got[:, :2] - ints[:, :2]).max() < 2.0 and np.abs(got[:, :2] - ints[:, :2]).max() > 0
    cfg["data"]["augmentation"] = {"homographic": {"valid_border_margin": 3}}
    with pytest.raises(NotImplementedError):
        export_detector_homoAdapt_b200(cfg, str(tmp_path / "m"), loader=samples, net=net)


def test_nn_match_two_way_vs_reference_golden(golden):
    """PointTracker.nn_match_two_way on the GPU vs the matches of the reference on its own kitti_test artefact."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200.frontend import PointTracker_b200
    g = golden("match")
    tr = PointTracker_b200(nn_thresh=0.7)
    m = tr.nn_match_two_way(g["desc1"].astype(np.float32), g["desc2"].astype(np.float32), 0.7)
    assert m.dtype == np.float64 and np.array_equal(m[:2], g["matches"][:2])
    np.testing.assert_allclose(m[2], g["matches"][2], atol=1e-6)
    assert tr.get_mscores() is m


@pytest.mark.parametrize("K1,K2,D,thr", [(1, 1, 256, 0.7), (37, 130, 256, 0.9), (513, 300, 256, 0.7), (1000, 1000, 256, 1.2),
                                         (65, 64, 8, 2.1)])
def test_nn_match_two_way_vs_oracle(K1, K2, D, thr):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import nn_match_two_way
    rng = np.random.RandomState(K1 + K2)
    d1 = rng.randn(D, K1).astype(np.float32)
    d2 = np.concatenate([d1[:, :min(K1, K2) // 2] + 0.3 * rng.randn(D, min(K1, K2) // 2).astype(np.float32),
                         rng.randn(D, K2 - min(K1, K2) // 2).astype(np.float32)], axis=1)
    d2 = d2[:, rng.permutation(K2)]
    d1 /= np.linalg.norm(d1, axis=0, keepdims=True)
    d2 /= np.linalg.norm(d2, axis=0, keepdims=True)
    got, ref = nn_match_two_way(d1, d2, thr), O.nn_match_two_way(d1, d2, thr)
    assert got.shape == ref.shape and np.array_equal(got[:2], ref[:2])
    np.testing.assert_allclose(got[2], ref[2], atol=2e-6)


def test_nn_match_two_way_edge_cases():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import nn_match_two_way
    assert nn_match_two_way(np.zeros((256, 0), np.float32), np.ones((256, 5), np.float32), 0.7).shape == (3, 0)
    assert nn_match_two_way(np.ones((256, 5), np.float32), np.zeros((256, 0), np.float32), 0.7).shape == (3, 0)
    with pytest.raises(ValueError):
        nn_match_two_way(np.ones((4, 2), np.float32), np.ones((4, 2), np.float32), -0.1)
    # exact ties: identical columns -> np.argmin keeps the first index in both directions
    d = np.zeros((4, 3), np.float32)
    d[0] = 1.0
    m = nn_match_two_way(d, d, 0.5)
    assert np.array_equal(m, O.nn_match_two_way(d, d, 0.5))


def _patch_corners(raw):
    """sample_homography_np's matrix maps the corners of the [-1,1]^2 frame to the sampled patch corners."""
    c = np.array([[-1, -1, 1], [-1, 1, 1], [1, 1, 1], [1, -1, 1]], np.float64)
    q = np.einsum("kij,cj->kci", raw.astype(np.float64), c)
    return (q[..., :2] / q[..., 2:]).reshape(len(raw), 8)


@pytest.mark.parametrize("name", ["export", "noartifacts"])
def test_device_sampler_distribution_vs_reference_draws(golden, name):
    """f-2: the GPU sampler draws from the same distribution as the reference's sample_homography_np (3000 draws of
    the unmodified reference in tests/golden/sampler_draws.npz): per patch-corner coordinate KS two-sample test,
    means and spreads."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from scipy.stats import ks_2samp
    from pytorch_superpoint_b200 import EXPORT_PARAMS, sample_ha_homographies_device
    params = dict(EXPORT_PARAMS)
    if name == "noartifacts":
        params.update(allow_artifacts=False, patch_ratio=0.5, scaling_amplitude=0.4, perspective_amplitude_x=0.3,
                      perspective_amplitude_y=0.3)
    ref = _patch_corners(golden("sampler_draws")[name])
    _, _, raw = sample_ha_homographies_device(20000, 1, seed=11, params=params, return_raw=True)
    got = _patch_corners(raw[0].cpu().numpy())
    assert np.isfinite(got).all()
    for j in range(8):
        ks = ks_2samp(ref[:, j], got[:, j]).statistic
        assert ks < 0.045, f"coordinate {j}: KS statistic {ks:.3f}"  # alpha = 1e-3 critical value is 0.038 + margin
        s = ref[:, j].std()
        assert abs(ref[:, j].mean() - got[:, j].mean()) < 5 * s / np.sqrt(len(ref))
        assert 0.93 < got[:, j].std() / s < 1.07
    if name == "noartifacts":  # without artifacts every patch stays inside the frame
        assert np.abs(got).max() <= 1.0 + 1e-6


def test_device_sampler_structure_and_export_without_loader_matrices(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import SuperPointNet_b200, sample_ha_homographies_device
    from pytorch_superpoint_b200.export import export_detector_homoAdapt_b200
    h, inv, raw = sample_ha_homographies_device(7, 3, seed=5, return_raw=True)
    assert tuple(h.shape) == (3, 7, 3, 3) and h.dtype == torch.float32 and h.is_cuda
    eye = torch.eye(3, device="cuda")
    assert torch.equal(h[:, 0], eye.expand(3, 3, 3)) and torch.equal(inv[:, 0], eye.expand(3, 3, 3))
    assert (h.double() @ inv.double() - eye.double()).abs().max().item() < 1e-5     # Coco.py:276
    assert (raw[:, 1:].double() @ h[:, 1:].double() - eye.double()).abs().max().item() < 1e-5  # Coco.py:268
    h2, _ = sample_ha_homographies_device(7, 3, seed=5)
    assert torch.equal(h, h2)                                                        # (seed, index) reproducible
    h3, _ = sample_ha_homographies_device(21, 1, seed=5)
    assert torch.equal(h3[0, 1:7], h[0, 1:7]) and torch.equal(h3[0, 8:14], h[1, 1:7])  # counter-based: index only
    assert not torch.equal(sample_ha_homographies_device(7, 3, seed=6)[0], h)
    # the .npz entry point draws its own matrices when the loader provides none
    net = SuperPointNet_b200()
    net.load_state_dict(O.synthetic_state_dict("gauss2", seed=2))
    rng = np.random.RandomState(4)
    samples = [{"image_2D": torch.from_numpy(rng.rand(1, 1, 64, 96).astype(np.float32)), "name": [f"a{i}"]}
               for i in range(2)]
    cfg = {"model": {"nms": 4, "top_k": 50, "detection_threshold": 0.015, "subpixel": {"enable": False}},
           "data": {"export_folder": "train", "dataset": "Coco",
                    "homography_adaptation": {"num": 6, "homographies": {"params": {
                        "translation": True, "rotation": True, "scaling": True, "perspective": True,
                        "scaling_amplitude": 0.2, "perspective_amplitude_x": 0.2, "perspective_amplitude_y": 0.2,
                        "allow_artifacts": True, "patch_ratio": 0.85}}}}}
    assert export_detector_homoAdapt_b200(cfg, str(tmp_path), loader=samples, net=net) == 2
    pts = np.load(os.path.join(str(tmp_path), "predictions", "train", "a0.npz"))["pts"]
    assert pts.dtype == np.float64 and pts.shape[1] == 3 and 0 < pts.shape[0] <= 50


def test_superpointnet_process_vs_reference_golden(golden):
    """f-3: SuperPointNet_process (models/model_utils.py) on the GPU vs golden outputs of the unmodified reference:
    NMS mask and roi-pooled patches bit-exact, raw descriptor samples and the cropped/padded feature batch (same
    seeded numpy RNG calls) to fp32 round-off."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200.process import SuperPointNet_process_b200
    g = golden("process")
    proc = SuperPointNet_process_b200(out_num_points=60, patch_size=5, device="cuda:0", nms_dist=4, conf_thresh=0.015)
    heat = torch.from_numpy(g["heat"]).cuda()
    mask = proc.heatmap_to_nms(heat, tensor=True)
    assert mask.is_cuda and np.array_equal(mask.cpu().numpy(), g["mask"])
    assert np.array_equal(proc.heatmap_to_nms(heat, tensor=False), g["mask"])
    out = proc.pred_soft_argmax(mask, heat)
    assert np.array_equal(out["patches"].cpu().numpy(), g["patches"])                      # roi_pool bins, pinned
    pred = out["pred"].cpu().numpy()
    assert pred.shape == (g["idx"].shape[0], 2) and np.isfinite(pred).all() and np.abs(pred).max() <= 2.0
    # the soft-argmax of log(patch) is the probability-weighted centroid of the patch (torchgeometry restated)
    p = g["patches"][:, 0].astype(np.float64)
    gx, gy = np.meshgrid(np.arange(5), np.arange(5))
    cen = np.stack([(p * gx).sum((1, 2)), (p * gy).sum((1, 2))], 1) / p.sum((1, 2))[:, None] - 2
    assert np.abs(pred - cen).max() < 1e-4
    desc = torch.from_numpy(g["desc"]).cuda()
    samp = proc.sample_desc_from_points(desc[:1], torch.from_numpy(g["pts"]).cuda())
    assert tuple(samp.shape) == g["samp"].shape and np.abs(samp.cpu().numpy() - g["samp"]).max() < 2e-6
    np.random.seed(99)
    feat = proc.batch_extract_features(desc, mask, torch.from_numpy(g["residual"]).cuda())
    assert np.array_equal(feat["pts_int"].cpu().numpy(), g["pts_int"])
    assert np.array_equal(feat["pts_offset"].cpu().numpy(), g["pts_offset"])
    assert np.abs(feat["pts_desc"].cpu().numpy() - g["pts_desc"]).max() < 2e-6


def test_process_output_end_to_end():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import SuperPointNet_b200
    from pytorch_superpoint_b200.process import SuperPointNet_process_b200
    net = SuperPointNet_b200()
    net.load_state_dict(O.synthetic_state_dict("gauss2", seed=2))
    x = torch.rand(2, 1, 64, 96, generator=torch.Generator().manual_seed(3)).cuda()
    net(x)
    np.random.seed(0)
    out = net.process_output(SuperPointNet_process_b200(out_num_points=80, device="cuda:0"))
    assert tuple(out["pts_int"].shape) == (2, 80, 2) and tuple(out["pts_offset"].shape) == (2, 80, 2)
    assert tuple(out["pts_desc"].shape) == (2, 80, 256) and torch.isfinite(out["pts_desc"]).all()
    assert out["pts_int"][..., 0].max() < 96 and out["pts_int"][..., 1].max() < 64


@pytest.mark.parametrize("B,Hc,Wc", [(2, 8, 12), (1, 30, 40), (3, 5, 7)])
def test_dense_desc_vs_torch(B, Hc, Wc):
    """f-4 dense-descriptor mode (models/model_wrap.py:393-401): fused x8 bilinear up-sampling + L2 normalisation vs
    torch's F.interpolate + norm on the same fp32 input."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import _lib
    g = torch.Generator().manual_seed(Hc)
    coarse = torch.randn(B, 256, Hc, Wc, generator=g)
    coarse = coarse / coarse.norm(dim=1, keepdim=True)
    ref = torch.nn.functional.interpolate(coarse, scale_factor=(8, 8), mode="bilinear")
    ref = ref / torch.norm(ref, p=2, dim=1, keepdim=True)
    nhwc = coarse.permute(0, 2, 3, 1).contiguous().cuda()
    dense = torch.full((B, 256, Hc * 8, Wc * 8), float("nan"), device="cuda")
    _lib.check(_lib.lib().spb_dense_desc(nhwc.data_ptr(), B, Hc, Wc, 256, 8, dense.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "dense_desc")
    assert torch.isfinite(dense).all() and (dense.cpu() - ref).abs().max().item() < 2e-6


def test_export_entry_batches_and_mixed_shapes(tmp_path):
    """The .npz entry point batches images (b200_batch_images), flushes on a shape change and on the tail, pipelines
    submit/collect and writes one file per image; key-points agree with the one-image-at-a-time export."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import SuperPointNet_b200, make_ha_homographies
    from pytorch_superpoint_b200.export import HAExporter, export_detector_homoAdapt_b200
    net = SuperPointNet_b200()
    net.load_state_dict(O.synthetic_state_dict("gauss2", seed=2))
    rng = np.random.RandomState(8)
    N = 5
    shapes = [(64, 96)] * 6 + [(48, 64)] * 3 + [(64, 96)] * 2
    samples = []
    for i, (H, W) in enumerate(shapes):
        img = np.full((H, W), 0.3, np.float32)
        for _ in range(6):
            x0, y0 = rng.randint(0, W - 12), rng.randint(0, H - 12)
            img[y0:y0 + rng.randint(6, 24), x0:x0 + rng.randint(6, 30)] = rng.uniform(0, 1)
        mu, mf = make_ha_homographies(N, seed=100 + i)
        samples.append({"image_2D": torch.from_numpy(img)[None, None], "name": [f"m{i}"],
                        "homographies": torch.from_numpy(mu)[None], "inv_homographies": torch.from_numpy(mf)[None]})
    cfg = {"model": {"nms": 4, "top_k": 80, "detection_threshold": 0.015, "subpixel": {"enable": False},
                     "b200_batch_images": 4},
           "data": {"export_folder": "val", "dataset": "Coco",
                    "homography_adaptation": {"num": N, "homographies": {"params": {}}}}}
    assert export_detector_homoAdapt_b200(cfg, str(tmp_path), loader=samples, net=net) == len(samples)
    single = HAExporter(net, 0.015, 4, 80)
    for i in (0, 5, 7, 10):
        got = np.load(os.path.join(str(tmp_path), "predictions", "val", f"m{i}.npz"))["pts"]
        ref = single.export_image(samples[i]["image_2D"][0, 0], samples[i]["homographies"][0],
                                  samples[i]["inv_homographies"][0])
        a = {(int(x), int(y)) for x, y, _ in got}
        b = {(int(x), int(y)) for x, y, _ in ref}
        assert got.dtype == np.float64 and len(a & b) / max(len(a | b), 1) > 0.95


def test_extract_batch_matches_per_image_path():
    """Device-resident batched extraction == the reference-facing per-image functions on the same heat-map."""
    fe = _fe(device="cuda")
    from pytorch_superpoint_b200 import SuperPointNet_b200
    fe.net = SuperPointNet_b200()
    fe.net.load_state_dict(O.synthetic_state_dict("gauss2", seed=4))
    x = torch.rand(3, 1, 64, 96, generator=torch.Generator().manual_seed(2)).cuda()
    pts, counts, desc = fe.extract_batch(x, top_k=50)
    assert tuple(pts.shape) == (3, 50, 3) and tuple(desc.shape) == (3, 50, 256)
    semi, dmap = fe.net.forward_nhwc(x)
    for b in range(3):
        k = int(counts[b])
        ref = fe.getPtsFromHeatmap(fe.heatmap[b, 0])[:, :50]
        assert k == ref.shape[1] and np.array_equal(pts[b, :k].cpu().numpy().astype(np.float64).T, ref)
        d = fe.sample_desc_from_points(dmap[b:b + 1].permute(0, 3, 1, 2).contiguous(), ref)
        assert np.abs(desc[b, :k].cpu().numpy().T - d).max() < 1e-6


def test_point_tracker_gpu_vs_reference_golden(golden):
    """PointTracker_b200.update with the GPU matcher on the reference's HPatches artefacts (fixture of the UNMODIFIED
    tracker, tests/golden/make_tracker_golden.py): matches (x1, y1, x2, y2) and the tracks matrix bit-exact."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import PointTracker_b200
    g = golden("tracker")
    for i in range(7):
        tr = PointTracker_b200(max_length=2, nn_thresh=1.0)
        tr.update(g[f"hp{i}_p1"], g[f"hp{i}_d1"].astype(np.float32))
        tr.update(g[f"hp{i}_p2"], g[f"hp{i}_d2"].astype(np.float32))
        assert np.array_equal(tr.get_matches(), g[f"hp{i}_matches"])
        assert np.array_equal(tr.get_mscores()[:2], g[f"hp{i}_mscores"][:2])
        np.testing.assert_allclose(tr.get_mscores()[2], g[f"hp{i}_mscores"][2], atol=2e-6)
        assert np.array_equal(tr.tracks[:, [0, 2, 3]], g[f"hp{i}_tracks"][:, [0, 2, 3]])
    tr = PointTracker_b200(max_length=3, nn_thresh=0.9)
    for k in range(5):
        tr.update(g[f"seq{k}_pts"], g[f"seq{k}_desc"])
        assert np.array_equal(tr.get_matches(), g[f"seq{k}_matches"])
        assert np.array_equal(tr.tracks[:, [0, 2, 3, 4]], g[f"seq{k}_tracks"][:, [0, 2, 3, 4]])
        np.testing.assert_allclose(tr.tracks[:, 1], g[f"seq{k}_tracks"][:, 1], atol=2e-6)


def test_nn_match_pairs_and_batched_descriptor_sampling():
    """The batched device entries (one launch per batch, no host round trip) against the per-image / per-pair ones:
    spb_sample_desc_batch == sample_desc_from_points per image with zero rows behind the count, spb_nn_match_pairs ==
    nn_match_two_way per pair (zero rows never match)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import nn_match_two_way, ops
    g = torch.Generator().manual_seed(3)
    B, Hc, Wc, D, K = 4, 30, 40, 256, 300
    coarse = torch.nn.functional.normalize(torch.randn(B, Hc, Wc, D, generator=g), dim=-1).cuda()
    counts = torch.tensor([300, 120, 0, 257], dtype=torch.int32).cuda()
    pts = torch.stack([torch.randint(4, 316, (B, K), generator=g).float(), torch.randint(4, 236, (B, K), generator=g).float(),
                       torch.rand(B, K, generator=g)], dim=-1).cuda()
    d = ops.sample_desc_batch_device(coarse, pts, counts)
    assert tuple(d.shape) == (B, K, D)
    for b in range(B):
        c = int(counts[b])
        assert (d[b, c:] == 0).all()
        if c:
            ref = ops.sample_desc_device(coarse[b], pts[b, :c].contiguous(), None, nhwc=True)
            assert torch.equal(d[b, :c], ref)
            want = O.sample_desc_from_points(coarse[b].permute(2, 0, 1)[None].cpu().numpy(),
                                             pts[b, :c].cpu().numpy().astype(np.float64).T)
            assert np.abs(d[b, :c].cpu().numpy().T - want).max() < 1e-6
    i1, i2, sc, cnt = ops.nn_match_pairs_device(d[0::2], d[1::2], 0.9)
    for p in range(2):
        a, b_ = d[2 * p], d[2 * p + 1]
        ca, cb = int(counts[2 * p]), int(counts[2 * p + 1])
        ref = nn_match_two_way(a[:ca].T.contiguous(), b_[:cb].T.contiguous(), 0.9) if ca and cb else np.zeros((3, 0))
        n = int(cnt[p])
        assert n == ref.shape[1]
        assert np.array_equal(i1[p, :n].cpu().numpy(), ref[0].astype(np.int32))
        assert np.array_equal(i2[p, :n].cpu().numpy(), ref[1].astype(np.int32))
        np.testing.assert_allclose(sc[p, :n].cpu().numpy(), ref[2], atol=1e-6)
    with pytest.raises(ValueError):
        ops.nn_match_pairs_device(d[0::2], d[1::2], -1.0)


@pytest.mark.parametrize("H,W,B", [(240, 320, 1), (480, 640, 1), (240, 320, 5)])
def test_val_model_heatmap_b200_vs_oracle(tmp_path, H, W, B):
    """Val_model_heatmap_b200 (Val_model_heatmap.py:33-155) at the configs[0] / configs[4] shapes and on a batch:
    heat-map vs the fp32 CPU oracle within 1e-2, points EXACTLY the reference's getPtsFromHeatmap of our heat-map,
    sparse descriptors vs the oracle's sampling of our dense map (1e-6) and of the fp32 map (cosine >= 0.999)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200.val_model import Val_model_heatmap_b200
    sd = O.synthetic_state_dict("gauss2", seed=6)
    tar = str(tmp_path / "superPointNet_170000_checkpoint.pth.tar")
    torch.save({"n_iter": 170000, "model_state_dict": sd, "optimizer_state_dict": {}, "loss": 0.0}, tar)
    cfg = {"name": "SuperPointNet_gauss2", "params": {}, "pretrained": tar, "nms": 4, "detection_threshold": 0.015,
           "nn_thresh": 1.0, "subpixel": {"enable": True, "patch_size": 5}}
    agent = Val_model_heatmap_b200(cfg, device="cuda")
    agent.loadModel()
    rng = np.random.RandomState(H + B)
    x = torch.from_numpy(rng.rand(B, 1, H, W).astype(np.float32))
    heat = agent.run(x)
    assert isinstance(heat, np.ndarray) and heat.shape == (B, 1, H, W) and heat.dtype == np.float32
    semi, desc = O.net_forward(O.canonical_params(sd), x)
    assert np.abs(heat[:, 0] - O.flatten_detection(semi.numpy())).max() < 1e-2
    pts = agent.heatmap_to_pts()
    descs = agent.desc_to_sparseDesc()
    assert len(pts) == B and len(descs) == B
    ours = agent.outs["desc"].contiguous().cpu().numpy()
    assert agent.outs["semi"].shape == (B, 65, H // 8, W // 8) and ours.shape == (B, 256, H // 8, W // 8)
    for b in range(B):
        assert pts[b].dtype == np.float64 and np.array_equal(pts[b], O.get_pts_from_heatmap(heat[b, 0], 0.015, 4))
        assert descs[b].shape == (256, pts[b].shape[1]) and descs[b].dtype == np.float32
        assert np.abs(descs[b] - O.sample_desc_from_points(ours[b:b + 1], pts[b])).max() < 1e-6
        ref_d = O.sample_desc_from_points(desc.numpy()[b:b + 1], pts[b])
        assert (descs[b] * ref_d).sum(0).min() > 0.999
    sub = agent.soft_argmax_points(pts, patch_size=5)
    assert len(sub) == 1 and np.abs(sub[0] - O.soft_argmax_points(heat[0, 0], pts[0], 5)).max() < 1e-4


def test_export_descriptor_b200_writes_reference_npz(tmp_path):
    """export.py:70-192 drop-in on 480x640 pairs (configs[4] shape): the reference's keys, shapes and dtypes; the
    matches are EXACTLY those of the reference tracker logic (oracle matcher) on the written points / descriptors."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import SuperPointNet_b200
    from pytorch_superpoint_b200.val_model import export_descriptor_b200
    sd = O.synthetic_state_dict("gauss2", seed=8)
    net = SuperPointNet_b200()
    net.load_state_dict(sd)
    rng = np.random.RandomState(2)
    H, W = 480, 640
    samples = []
    for i in range(2):
        img = rng.rand(1, 1, H, W).astype(np.float32)
        warped = np.roll(img, (3, -5), axis=(2, 3)) * 0.9 + 0.05  # a second view of the same content
        samples.append({"image": torch.from_numpy(img), "warped_image": torch.from_numpy(warped),
                        "homography": torch.eye(3)[None]})
    cfg = {"model": {"name": "SuperPointNet_gauss2", "params": {}, "pretrained": "", "nms": 4,
                     "detection_threshold": 0.015, "nn_thresh": 1.0, "subpixel": {"enable": False, "patch_size": 5}},
           "data": {"dataset": "hpatches"}}
    assert export_descriptor_b200(cfg, str(tmp_path), loader=samples, net=net) == 2
    for i in range(2):
        f = np.load(os.path.join(str(tmp_path), "predictions", f"{i}.npz"))
        assert sorted(f.files) == sorted(["image", "prob", "desc", "warped_image", "warped_prob", "warped_desc",
                                          "homography", "matches"])
        K1, K2 = f["prob"].shape[0], f["warped_prob"].shape[0]
        assert f["image"].shape == (H, W) and f["prob"].shape == (K1, 3) and f["prob"].dtype == np.float64
        assert f["desc"].shape == (K1, 256) and f["desc"].dtype == np.float32 and f["warped_desc"].shape == (K2, 256)
        assert f["homography"].shape == (3, 3) and f["matches"].shape[1] == 4 and f["matches"].dtype == np.float64
        assert np.all(np.diff(f["prob"][:, 2]) <= 0) and K1 > 0 and K2 > 0
        assert np.abs(np.linalg.norm(f["desc"], axis=1) - 1).max() < 1e-5
        m = O.nn_match_two_way(f["desc"].T, f["warped_desc"].T, 1.0)
        want = np.concatenate((f["prob"][:, :2].T[:, m[0].astype(int)], f["warped_prob"][:, :2].T[:, m[1].astype(int)]))
        assert np.array_equal(f["matches"], want.T)


def test_export_kitti_shape_full_views_vs_oracle():
    """BASELINE configs[3] at full size: ONE 384x1248 image with all 100 homographies through HAExporter -- the pass
    size is bounded by bytes (128 views at this shape), the workspace is sized for the fused first block -- against the
    CPU oracle: heat-map within 1e-2, key-points exactly the reference's NMS / top-k of our heat-map."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import SuperPointNet_b200, _lib, make_ha_homographies
    from pytorch_superpoint_b200.export import HAExporter
    H, W, N = 384, 1248, 100
    rng = np.random.RandomState(3)
    img = np.full((H, W), 0.35, np.float32) + 0.03 * rng.randn(H, W).astype(np.float32)
    for _ in range(30):
        x0, y0 = rng.randint(0, W - 10), rng.randint(0, H - 10)
        img[y0:y0 + rng.randint(8, H // 3), x0:x0 + rng.randint(8, W // 4)] = rng.uniform(0, 1)
    img = np.clip(img, 0, 1)
    mu, mf = make_ha_homographies(N, seed=5)
    sd = O.synthetic_state_dict("gauss2", seed=1)
    net = SuperPointNet_b200()
    net.load_state_dict(sd)
    exp = HAExporter(net, 0.015, 4, 600)
    assert net.ha_pass_views(H, W) == 128
    plan, _ = exp.plan(1, N, H, W, True)
    assert len(plan.passes) == 1 and plan.passes[0].num_views == N
    L = _lib.lib()
    assert L.spb_ha_workspace_bytes_ragged(1, N, N, H, W) < 12 * (1 << 30)  # (round 1 sized 49 GB for the same pass)
    pts = exp.export_batch([torch.from_numpy(img)], torch.from_numpy(mu), torch.from_numpy(mf))[0]
    heat = exp.last_heat(1, H, W)[0].cpu().numpy()
    ref_pts, ref_heat = O.ha_export_one(img, mu, mf, O.canonical_params(sd), return_heat=True)
    assert np.isfinite(heat).all() and np.abs(heat - ref_heat).max() < 1e-2
    assert np.array_equal(pts, O.get_pts_from_heatmap(heat, 0.015, 4).T[:600])
    # a smaller pass size splits the image's views into accumulating slices: same sums up to fp32 order
    small = HAExporter(net, 0.015, 4, 600, chunk=48)
    plan2, _ = small.plan(1, N, H, W, True)
    assert [p.num_views for p in plan2.passes] == [48, 48, 4] and [p.accumulate for p in plan2.passes] == [False, True, True]
    small.export_batch([torch.from_numpy(img)], torch.from_numpy(mu), torch.from_numpy(mf))
    heat2 = small.last_heat(1, H, W)[0].cpu().numpy()
    assert np.abs(heat2 - heat).max() < 1e-5
