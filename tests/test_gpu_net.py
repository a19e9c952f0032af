"""GPU parity tests for the network kernels: tcgen05 implicit-GEMM vs the CUDA-core validation kernels
(same bf16 numerics -> tight), and both vs the fp32 CPU oracle / the reference's golden outputs
(bf16 tolerance: heat-map max-abs 1e-2, descriptor cosine >= 0.999 -- BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402  (checker only)

SHAPES = [(1, 64, 3), (64, 64, 3), (64, 64, 3), (64, 64, 3), (64, 128, 3), (128, 128, 3), (128, 128, 3),
          (128, 128, 3), (128, 256, 3), (256, 65, 1), (128, 256, 3), (256, 256, 1)]
POOL = [0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0]


@pytest.fixture(scope="module")
def net():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pytorch_superpoint_b200 import SuperPointNet_b200
    n = SuperPointNet_b200()
    n.load_state_dict(O.synthetic_state_dict("gauss2", seed=1))
    return n


def _layer(net, li, x, B, H, W, impl, halo=1):
    from pytorch_superpoint_b200 import _lib
    L = _lib.lib()
    h = net.handle(x.device)
    net.set_impl(impl)
    L.spb_conv_tc_set_halo(halo)
    cin, cout, _ = SHAPES[li]
    oh, ow = (H // 2, W // 2) if POOL[li] else (H, W)
    out = torch.full((B, oh, ow, cout), float("nan"), device=x.device,
                     dtype=torch.float32 if li in (9, 11) else torch.bfloat16)
    _lib.check(L.spb_net_layer(h, li, x.data_ptr(), out.data_ptr(), B, H, W, torch.cuda.current_stream().cuda_stream),
               f"net_layer {li}")
    torch.cuda.synchronize()
    L.spb_conv_tc_set_halo(1)
    return out


def _ref_layer(li, x_nhwc, params):
    """fp32 torch reference of one layer on the bf16-rounded inputs/weights (tight comparison target)."""
    name = O.LAYERS[li][0]
    p = O.fold_bn(params)[name]
    w = p["w"].to(torch.bfloat16).float().cuda()
    y = torch.nn.functional.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w, p["b"].cuda(), padding=w.shape[-1] // 2)
    if O.LAYERS[li][4]:
        y = torch.relu(y)
    if O.LAYERS[li][5]:
        y = torch.nn.functional.max_pool2d(y, 2, 2)
    if li == 11:  # convDb rows come back L2-normalised (models/SuperPointNet_pretrained.py:73-74)
        y = y / torch.norm(y, p=2, dim=1, keepdim=True)
    return y.permute(0, 2, 3, 1).contiguous()


# 1x1 layers first: they exercise TMA/UMMA/TMEM plumbing with the canonical (unshifted) operand layout
@pytest.mark.parametrize("halo", [0, 1])
@pytest.mark.parametrize("li,B,H,W", [(9, 2, 32, 40), (11, 2, 32, 40), (9, 3, 30, 40), (1, 2, 32, 48), (2, 2, 32, 40),
                                      (3, 1, 60, 80), (4, 2, 30, 40), (5, 2, 60, 80), (6, 3, 30, 40), (7, 1, 16, 8),
                                      (8, 2, 30, 40), (10, 1, 30, 40), (1, 1, 240, 320)])
def test_tcgen05_layer_vs_reference(net, li, B, H, W, halo):
    if SHAPES[li][2] == 1 and halo == 0:
        pytest.skip("1x1 layers have a single operand layout")
    torch.manual_seed(li * 7 + B)
    cin = SHAPES[li][0]
    x = (torch.randn(B, H, W, cin, device="cuda") * 0.5).to(torch.bfloat16).contiguous()
    params = O.canonical_params(O.synthetic_state_dict("gauss2", seed=1))
    ref = _ref_layer(li, x, params)
    simt = _layer(net, li, x, B, H, W, "simt").float()
    tc = _layer(net, li, x, B, H, W, "tcgen05", halo).float()
    tol = 3e-2 if li not in (9, 11) else 2e-3  # bf16 output rounding (|y| <~ 4) vs fp32 outputs
    assert torch.isfinite(simt).all() and torch.isfinite(tc).all()
    assert (simt - ref).abs().max().item() < tol, "SIMT validation kernel vs torch"
    assert (tc - ref).abs().max().item() < tol, "tcgen05 kernel vs torch"
    # same numerics contract -> the two CUDA paths agree to accumulation-order noise (1 bf16 ulp at most)
    assert (tc - simt).abs().max().item() <= (3.2e-2 if li not in (9, 11) else 1e-4)


@pytest.mark.parametrize("li,B,H,W", [(2, 3, 24, 30), (3, 2, 16, 28), (1, 1, 8, 14), (2, 1, 9, 17), (3, 5, 40, 58),
                                      (2, 2, 120, 160), (3, 2, 120, 160), (2, 1, 3, 2)])
def test_s3_column_stacked_kernel(net, li, B, H, W):
    """64->64 layers on the N=192 column-tap-stacked kernel (conv_s3.cu): ragged tiles (W % 14, H % 8), pooling,
    vs torch fp32, the CUDA-core kernel and the shifted-window tcgen05 kernel."""
    from pytorch_superpoint_b200 import _lib
    L = _lib.lib()
    torch.manual_seed(li * 11 + W)
    x = (torch.randn(B, H, W, 64, device="cuda") * 0.5).to(torch.bfloat16).contiguous()
    params = O.canonical_params(O.synthetic_state_dict("gauss2", seed=1))
    ref = _ref_layer(li, x, params)
    simt = _layer(net, li, x, B, H, W, "simt").float()
    try:
        L.spb_conv_tc_set_s3(0)
        win = _layer(net, li, x, B, H, W, "tcgen05").float()
        L.spb_conv_tc_set_s3(1)
        s3 = _layer(net, li, x, B, H, W, "tcgen05").float()
    finally:
        L.spb_conv_tc_set_s3(1)
    assert torch.isfinite(s3).all()
    assert (s3 - ref).abs().max().item() < 3e-2
    assert (s3 - simt).abs().max().item() <= 3.2e-2 and (s3 - win).abs().max().item() <= 3.2e-2
    assert (s3 != win).float().mean().item() < 0.02  # same products, different fp32 summation order: rare 1-ulp flips


@pytest.mark.parametrize("li,B,H,W", [(5, 3, 60, 80), (5, 1, 16, 8), (6, 5, 30, 40), (7, 7, 30, 40), (6, 1, 16, 8),
                                      (5, 2, 34, 50)])
def test_tile_pairs_equal_single_tiles(net, li, B, H, W):
    """128->128 layers: two tiles per accumulator stage / weight pass (odd tile counts end in a dummy tile that TMA
    zero-fills and clips) == one tile per stage, bit for bit (same per-tile accumulation order)."""
    from pytorch_superpoint_b200 import _lib
    L = _lib.lib()
    torch.manual_seed(li + H)
    x = (torch.randn(B, H, W, 128, device="cuda") * 0.5).to(torch.bfloat16).contiguous()
    try:
        L.spb_conv_tc_set_pair(0)
        one = _layer(net, li, x, B, H, W, "tcgen05")
        L.spb_conv_tc_set_pair(1)
        two = _layer(net, li, x, B, H, W, "tcgen05")
    finally:
        L.spb_conv_tc_set_pair(1)
    assert torch.isfinite(two.float()).all() and torch.equal(one, two)


@pytest.mark.parametrize("li,B,H,W", [(4, 2, 60, 80), (4, 1, 16, 8), (5, 3, 60, 80), (5, 1, 16, 8), (6, 5, 30, 40),
                                      (7, 1, 16, 8), (8, 3, 30, 40), (8, 1, 34, 50), (10, 2, 30, 40), (6, 9, 30, 40)])
def test_cta_pair_kernels_equal_single_cta(net, li, B, H, W):
    """>= 128-channel layers on CTA pairs (tcgen05.mma.cta_group::2, each CTA holds half of every weight block; odd
    tile counts end in dummy tiles) == the single-CTA kernels, bit for bit (same per-row accumulation order)."""
    from pytorch_superpoint_b200 import _lib
    L = _lib.lib()
    torch.manual_seed(li + W)
    x = (torch.randn(B, H, W, SHAPES[li][0], device="cuda") * 0.5).to(torch.bfloat16).contiguous()
    try:
        L.spb_conv_tc_set_cg2(0)
        one = _layer(net, li, x, B, H, W, "tcgen05")
        L.spb_conv_tc_set_cg2(2)  # every layer that has a pair kernel (incl. 64 -> 128, off by default)
        two = _layer(net, li, x, B, H, W, "tcgen05")
    finally:
        L.spb_conv_tc_set_cg2(1)
    assert torch.isfinite(two.float()).all() and torch.equal(one, two)


@pytest.mark.parametrize("B,H,W", [(1, 16, 8), (3, 24, 40), (2, 240, 320), (5, 30, 52)])
def test_conv1a_tensor_core_vs_simt(net, B, H, W):
    """conv1a as a K=16 im2col MMA: same bf16 numerics as the CUDA-core kernel, incl. ragged last tile."""
    x = torch.rand(B, H, W, device="cuda", generator=torch.Generator("cuda").manual_seed(H))
    simt = _layer(net, 0, x, B, H, W, "simt").float()
    tc = _layer(net, 0, x, B, H, W, "tcgen05").float()
    p = O.fold_bn(O.canonical_params(O.synthetic_state_dict("gauss2", seed=1)))["conv1a"]
    ref = torch.relu(torch.nn.functional.conv2d(x[:, None].to(torch.bfloat16).float(),
                                                p["w"].to(torch.bfloat16).float().cuda(), p["b"].cuda(), padding=1))
    ref = ref.permute(0, 2, 3, 1)
    assert torch.isfinite(tc).all()
    assert (tc - ref).abs().max().item() < 2e-2 and (simt - ref).abs().max().item() < 2e-2
    assert (tc - simt).abs().max().item() <= 1.6e-2


@pytest.mark.parametrize("impl", ["simt", "tcgen05"])
@pytest.mark.parametrize("layout", ["gauss2", "bnflat", "v1"])
def test_forward_vs_reference_golden(impl, layout, golden):
    """Full network on the reference's golden input (eval-mode BN): semi/desc within bf16 tolerance."""
    from pytorch_superpoint_b200 import SuperPointNet_b200, flattenDetection
    g = golden("net")
    n = SuperPointNet_b200(impl=impl)
    n.load_state_dict(O.synthetic_state_dict(layout, seed=1))
    out = n(torch.from_numpy(g["x"]).cuda())
    semi, desc = out["semi"], out["desc"]
    assert tuple(semi.shape) == g[f"{layout}_eval_semi"].shape and tuple(desc.shape) == g[f"{layout}_eval_desc"].shape
    ref_semi = torch.from_numpy(g[f"{layout}_eval_semi"]).cuda()
    ref_desc = torch.from_numpy(g[f"{layout}_eval_desc"]).cuda()
    heat, ref_heat = flattenDetection(semi.contiguous(), True), flattenDetection(ref_semi, True)
    assert (heat - ref_heat).abs().max().item() < 1e-2          # north_star: max-abs 1e-2 heat-map in bf16
    cos = (desc * ref_desc).sum(1)
    assert cos.min().item() > 0.999                              # north_star: descriptor cosine >= 0.999
    # what bf16 activations/weights with fp32 accumulation actually deliver on this input (CPU emulation:
    # cos >= 0.99996, |semi| error <= 0.015): anything worse is a kernel bug, not rounding
    assert cos.min().item() > 0.9999
    assert (semi - ref_semi).abs().max().item() < 0.03
    assert (desc.norm(dim=1) - 1).abs().max().item() < 1e-4


def test_forward_full_size_simt_vs_tcgen05(net):
    """240x320 batch: the two CUDA implementations agree; outputs finite; batch independence."""
    x = torch.rand(4, 1, 240, 320, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    net.set_impl("simt")
    s0, d0 = net.forward_nhwc(x)
    net.set_impl("tcgen05")
    s1, d1 = net.forward_nhwc(x)
    assert torch.isfinite(s1).all() and torch.isfinite(d1).all()
    assert (s0 - s1).abs().max().item() < 5e-2
    assert (d0 * d1).sum(-1).min().item() > 0.9995
    s2, _ = net.forward_nhwc(x[2:3].contiguous())
    assert torch.equal(s2[0], s1[2])


def test_ha_end_to_end_vs_reference_golden(golden):
    """Whole HA step (warp -> net -> fused aggregate -> NMS) on the reference's golden image."""
    from pytorch_superpoint_b200 import SuperPointNet_b200, getPtsFromHeatmap
    from pytorch_superpoint_b200.ops import ha_finalize
    g = golden("ha_e2e")
    n = SuperPointNet_b200()
    n.load_state_dict(O.synthetic_state_dict("gauss2", seed=2))
    img = torch.from_numpy(g["img"]).cuda()
    for chunk in (0, 3):
        sums = n.ha_heatmap_sums(img, torch.from_numpy(g["h"]), torch.from_numpy(g["inv"]), chunk=chunk)
        heat = ha_finalize(sums)
        assert (heat - torch.from_numpy(g["agg"]).cuda()).abs().max().item() < 1e-2
    # bit-exact key-points GIVEN the reference's heat-map
    pts = getPtsFromHeatmap(torch.from_numpy(g["agg"]).cuda(), 0.015, 4).T[:600]
    assert np.array_equal(pts, g["pts"])


@pytest.mark.parametrize("chunk", [0, 16, 5])
def test_ha_batched_images_equal_single_image_calls(chunk):
    """spb_ha_heatmap_sums_batch over B images (shared and per-image homography sets, several pass sizes)
    gives the same sums as one call per image."""
    from pytorch_superpoint_b200 import SuperPointNet_b200, make_ha_homographies
    n = SuperPointNet_b200()
    n.load_state_dict(O.synthetic_state_dict("gauss2", seed=2))
    B, H, W, N = 3, 48, 64, 8
    imgs = torch.rand(B, H, W, device="cuda", generator=torch.Generator("cuda").manual_seed(5))
    sets = [make_ha_homographies(N, seed=30 + b) for b in range(B)]
    mu = torch.from_numpy(np.stack([s[0] for s in sets])).cuda()
    mf = torch.from_numpy(np.stack([s[1] for s in sets])).cuda()
    single = torch.stack([n.ha_heatmap_sums(imgs[b], mu[b], mf[b]).clone() for b in range(B)])
    batch = n.ha_heatmap_sums_batch(imgs, mu, mf, chunk=chunk).clone()
    assert torch.allclose(batch, single, rtol=1e-5, atol=1e-6)
    shared_single = torch.stack([n.ha_heatmap_sums(imgs[b], mu[0], mf[0]).clone() for b in range(B)])
    shared_batch = n.ha_heatmap_sums_batch(imgs, mu[0], mf[0], chunk=chunk).clone()
    assert torch.allclose(shared_batch, shared_single, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("B,N,chunk", [(5, 8, 16), (4, 6, 6), (1, 20, 7), (3, 100, 100)])
def test_ha_overlapped_passes_equal_sequential_passes(B, N, chunk):
    """More than one pass: warp(c+1) and aggregate(c-1) on the internal side stream under the net of pass c give
    bit-identical sums to strictly sequential passes (also when one image is split into view chunks), and the
    result is complete in the caller's stream."""
    from pytorch_superpoint_b200 import SuperPointNet_b200, make_ha_homographies
    n = SuperPointNet_b200()
    n.load_state_dict(O.synthetic_state_dict("gauss2", seed=2))
    H, W = 64, 96
    imgs = torch.rand(B, H, W, device="cuda", generator=torch.Generator("cuda").manual_seed(N))
    sets = [make_ha_homographies(N, seed=70 + b) for b in range(B)]
    mu = torch.from_numpy(np.stack([s[0] for s in sets])).cuda()
    mf = torch.from_numpy(np.stack([s[1] for s in sets])).cuda()
    try:
        n.set_ha_overlap(False)
        ref = n.ha_heatmap_sums_batch(imgs, mu, mf, chunk=chunk).clone()
        n.set_ha_overlap(True)
        for _ in range(3):  # back-to-back calls reuse the pass buffers and events
            out = n.ha_heatmap_sums_batch(imgs, mu, mf, chunk=chunk)
            got = out.clone()  # same stream: must already see the side stream's last aggregate
            assert torch.equal(got, ref)
    finally:
        n.set_ha_overlap(False)


@pytest.mark.parametrize("B,H,W", [(1, 16, 8), (2, 24, 40), (3, 64, 96), (2, 240, 320), (1, 40, 24)])
def test_conv1_fused_block_equals_two_kernels(net, B, H, W):
    """conv1a+conv1b+pool fused in one kernel (conv1a's output stays in shared memory) == the two-kernel path."""
    from pytorch_superpoint_b200 import _lib
    L = _lib.lib()
    x = torch.rand(B, 1, H, W, device="cuda", generator=torch.Generator("cuda").manual_seed(W))
    net.set_impl("tcgen05")
    try:
        L.spb_conv_tc_set_s3(0)  # the unfused conv1b must use the same (shifted-window) accumulation order
        L.spb_conv_tc_set_fuse1(0)
        s0, d0 = net.forward_nhwc(x)
        s0, d0 = s0.clone(), d0.clone()
        L.spb_conv_tc_set_fuse1(1)
        s1, d1 = net.forward_nhwc(x)
    finally:
        L.spb_conv_tc_set_fuse1(-1)
        L.spb_conv_tc_set_s3(1)
    assert torch.isfinite(s1).all()
    assert (s0 - s1).abs().max().item() <= 1e-5  # same bf16 intermediate, same accumulation order
    assert (d0 - d1).abs().max().item() <= 1e-6


@pytest.mark.parametrize("H,W,N", [(384, 1248, 3), (480, 640, 4), (120, 160, 7)])
def test_ha_other_config_shapes_vs_oracle(H, W, N):
    """BASELINE configs 3-5 shapes (KITTI 384x1248, HPatches 480x640): whole HA step vs the fp32 CPU oracle --
    aggregated heat-map within the bf16 tolerance, mask sums bit-exact, key-points exact given the oracle's heat."""
    from pytorch_superpoint_b200 import SuperPointNet_b200, getPtsFromHeatmap, make_ha_homographies
    from pytorch_superpoint_b200.ops import ha_finalize
    rng = np.random.RandomState(H + N)
    img = np.full((H, W), 0.3, np.float32) + 0.02 * rng.randn(H, W).astype(np.float32)
    for _ in range(24):
        x0, y0 = rng.randint(0, W - 12), rng.randint(0, H - 12)
        img[y0:y0 + rng.randint(6, H // 3), x0:x0 + rng.randint(6, W // 4)] = rng.uniform(0, 1)
    img = np.clip(img, 0, 1)
    mu, mf = make_ha_homographies(N, seed=N)
    sd = O.synthetic_state_dict("gauss2", seed=3)
    n = SuperPointNet_b200()
    n.load_state_dict(sd)
    sums = n.ha_heatmap_sums(torch.from_numpy(img).cuda(), torch.from_numpy(mu), torch.from_numpy(mf))
    heat = ha_finalize(sums).cpu().numpy()
    ref_pts, ref_heat = O.ha_export_one(img, mu, mf, O.canonical_params(sd), return_heat=True)
    assert np.isfinite(heat).all() and np.abs(heat - ref_heat).max() < 1e-2
    pts = getPtsFromHeatmap(torch.from_numpy(ref_heat).cuda(), 0.015, 4).T[:600]
    assert np.array_equal(pts, ref_pts)


@pytest.mark.parametrize("B,H,W,N,world", [(4, 48, 64, 10, 4), (8, 64, 96, 13, 8), (3, 40, 24, 7, 3)])
def test_ha_ragged_balanced_shards_sum_to_the_full_heatmap(B, H, W, N, world):
    """Balanced homography sharding on one GPU: the `world` ragged partial sums (spb_ha_heatmap_sums_ragged with the
    rotated per-image shares) add up to the unsharded sums, and each equals the per-image uniform calls on its views."""
    from pytorch_superpoint_b200 import SuperPointNet_b200, sample_ha_homographies_device
    from pytorch_superpoint_b200.sharding import balanced_view_plan
    n = SuperPointNet_b200()
    n.load_state_dict(O.synthetic_state_dict("gauss2", seed=2))
    imgs = torch.rand(B, H, W, device="cuda", generator=torch.Generator("cuda").manual_seed(W))
    mu, mf = sample_ha_homographies_device(N, B, seed=W)
    full = n.ha_heatmap_sums_batch(imgs, mu, mf).clone()
    total = torch.zeros_like(full)
    for r in range(world):
        sel, vi, vo, most = balanced_view_plan(B, N, r, world)
        part = n.ha_heatmap_sums_ragged(imgs, mu.reshape(-1, 3, 3)[sel.cuda()], mf.reshape(-1, 3, 3)[sel.cuda()], vi, vo,
                                        most).clone()
        for b in range(B):
            s = (sel[int(vo[b]):int(vo[b + 1])] % N).cuda()
            ref = n.ha_heatmap_sums(imgs[b], mu[b][s], mf[b][s]).clone()
            assert torch.allclose(part[b], ref, rtol=1e-5, atol=1e-6)
        total += part
    assert torch.allclose(total, full, rtol=1e-5, atol=1e-5)            # fp32 summation order differs only


@pytest.mark.parametrize("B,H,W,N", [(2, 64, 96, 9), (1, 120, 160, 16)])
def test_ha_fast_pass_vs_bit_exact_validation_pass(B, H, W, N):
    """The product pass (folded coordinates, fp16 masked planes, tcgen05 convolutions) against the validation pass
    (SPB_IMPL_SIMT: the bit-exact warp / mask / aggregate kernels around CUDA-core convolutions with the same bf16
    numerics): the mask-sum planes agree to the folded un-warp's interpolation error, the heat-maps to a fraction of
    the 1e-2 tolerance."""
    from pytorch_superpoint_b200 import SuperPointNet_b200, sample_ha_homographies_device
    n = SuperPointNet_b200()
    n.load_state_dict(O.synthetic_state_dict("gauss2", seed=2))
    imgs = torch.rand(B, H, W, device="cuda", generator=torch.Generator("cuda").manual_seed(H))
    mu, mf = sample_ha_homographies_device(N, B, seed=H)
    fast = n.ha_heatmap_sums_batch(imgs, mu, mf).clone()
    try:
        n.set_impl("simt")
        exact = n.ha_heatmap_sums_batch(imgs, mu, mf).clone()
    finally:
        n.set_impl("tcgen05")
    assert torch.isfinite(fast).all() and torch.isfinite(exact).all()
    assert (fast[:, 1] - exact[:, 1]).abs().max().item() < 2e-2          # sum of masks (values up to N)
    assert (fast[:, 1] - exact[:, 1]).abs().mean().item() < 1e-3
    heat_f, heat_e = fast[:, 0] / fast[:, 1], exact[:, 0] / exact[:, 1]
    assert (heat_f - heat_e).abs().max().item() < 2e-3


# ---- train-mode BatchNorm (bn_mode='batch'): the as-written reference never calls .eval() --------------------------
@pytest.mark.parametrize("R,C", [(1000, 64), (777, 65), (4096, 128), (50, 256), (3, 64)])
def test_bn_batch_stats_vs_torch(R, C):
    """spb_bn_batch_stats: mean / BIASED variance / scale / shift of fp32 rows vs a float64 torch computation."""
    from pytorch_superpoint_b200 import _lib
    L = _lib.lib()
    g = torch.Generator("cuda").manual_seed(R + C)
    rows = torch.randn(R, C, device="cuda", generator=g) * 3 + torch.arange(C, device="cuda") * 0.1
    gamma, beta = torch.rand(C, device="cuda", generator=g) + 0.5, torch.randn(C, device="cuda", generator=g)
    mean, var, scale, shift = (torch.empty(C, device="cuda") for _ in range(4))
    ws = torch.empty(L.spb_bn_batch_stats_workspace_bytes(C), dtype=torch.uint8, device="cuda")
    _lib.check(L.spb_bn_batch_stats(rows.data_ptr(), R, C, gamma.data_ptr(), beta.data_ptr(), 1e-5, mean.data_ptr(),
                                    var.data_ptr(), scale.data_ptr(), shift.data_ptr(), ws.data_ptr(), ws.numel(),
                                    torch.cuda.current_stream().cuda_stream), "bn_batch_stats")
    r64 = rows.double()
    m, v = r64.mean(0), r64.var(0, unbiased=False)
    sc = gamma.double() / torch.sqrt(v + 1e-5)
    assert torch.allclose(mean.double(), m, rtol=1e-6, atol=1e-6) and torch.allclose(var.double(), v, rtol=1e-5, atol=1e-6)
    assert torch.allclose(scale.double(), sc, rtol=1e-5) and torch.allclose(shift.double(), beta.double() - m * sc, rtol=1e-5, atol=1e-5)
    # the transform reproduces F.batch_norm(training=True)
    y = torch.nn.functional.batch_norm(rows.t()[None], None, None, gamma, beta, training=True, eps=1e-5)[0].t()
    assert (rows * scale + shift - y).abs().max().item() < 1e-4


@pytest.mark.parametrize("B,H,W,cin,cout,ks", [(2, 16, 24, 1, 64, 3), (2, 9, 13, 64, 64, 3), (1, 8, 8, 64, 128, 3),
                                                (3, 6, 10, 128, 128, 3), (1, 5, 7, 128, 256, 3), (2, 4, 6, 256, 65, 1),
                                                (2, 4, 6, 256, 256, 1)])
def test_conv_f32_vs_torch(B, H, W, cin, cout, ks):
    """spb_conv_f32 (the convolution of the train-mode BN path) vs torch's fp32 conv2d on the same operands."""
    from pytorch_superpoint_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(cin + cout + H)
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, ks, ks, generator=g) * (2.0 / (cin * ks * ks)) ** 0.5
    b = torch.randn(cout, generator=g)
    cpad = -(-cout // 16) * 16
    wt = torch.zeros(ks * ks, cin, cpad)
    wt[:, :, :cout] = w.permute(2, 3, 1, 0).reshape(ks * ks, cin, cout)
    bt = torch.zeros(cpad)
    bt[:cout] = b
    xin = x.permute(0, 2, 3, 1).contiguous().cuda()
    out = torch.full((B, H, W, cout), float("nan"), device="cuda")
    wt, bt = wt.cuda(), bt.cuda()
    _lib.check(L.spb_conv_f32(xin.data_ptr(), wt.data_ptr(), bt.data_ptr(), out.data_ptr(), B, H, W, cin, cout, cpad, ks,
                              torch.cuda.current_stream().cuda_stream), "conv_f32")
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=ks // 2).permute(0, 2, 3, 1)
    assert (out.cpu().double() - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,H,W,C,relu,pool,f32", [(2, 8, 12, 64, 1, 1, 0), (3, 6, 10, 128, 1, 0, 0), (1, 4, 6, 65, 0, 0, 1),
                                                    (2, 4, 4, 256, 0, 0, 1), (1, 16, 8, 64, 1, 1, 1)])
def test_bn_apply_vs_torch(B, H, W, C, relu, pool, f32):
    from pytorch_superpoint_b200 import _lib
    L = _lib.lib()
    g = torch.Generator("cuda").manual_seed(B * H + C)
    rows = torch.randn(B, H, W, C, device="cuda", generator=g)
    scale, shift = torch.rand(C, device="cuda", generator=g) + 0.5, torch.randn(C, device="cuda", generator=g)
    oh, ow = (H // 2, W // 2) if pool else (H, W)
    l2 = int(bool(f32) and C == 256)
    out = torch.empty((B, oh, ow, C), device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    _lib.check(L.spb_bn_apply(rows.data_ptr(), B, H, W, C, scale.data_ptr(), shift.data_ptr(), relu, pool, out.data_ptr(),
                              f32, l2, torch.cuda.current_stream().cuda_stream), "bn_apply")
    y = rows * scale + shift
    if pool:
        y = torch.nn.functional.max_pool2d(y.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    if relu:
        y = torch.relu(y)
    if l2:
        y = y / y.norm(dim=-1, keepdim=True)
    if f32:
        assert (out - y).abs().max().item() < 1e-5
    else:
        assert torch.equal(out, y.to(torch.bfloat16))


@pytest.mark.parametrize("layout", ["gauss2", "bnflat", "v1"])
def test_forward_batch_bn_vs_reference_golden(layout, golden):
    """bn_mode='batch' on the reference's golden input: the unmodified reference network left in TRAIN mode (what
    export.py actually runs, model_wrap.py:111) -- semi / desc within the bf16 tolerance; and the two BN modes
    really differ on this input (so the test cannot pass by running eval mode)."""
    from pytorch_superpoint_b200 import SuperPointNet_b200, flattenDetection
    g = golden("net")
    key = "eval" if layout == "v1" else "batch"  # the BN-free network has only one mode
    n = SuperPointNet_b200(bn_mode="batch")
    n.load_state_dict(O.synthetic_state_dict(layout, seed=1))
    out = n(torch.from_numpy(g["x"]).cuda())
    semi, desc = out["semi"], out["desc"]
    ref_semi = torch.from_numpy(g[f"{layout}_{key}_semi"]).cuda()
    ref_desc = torch.from_numpy(g[f"{layout}_{key}_desc"]).cuda()
    assert tuple(semi.shape) == tuple(ref_semi.shape) and tuple(desc.shape) == tuple(ref_desc.shape)
    heat, ref_heat = flattenDetection(semi.contiguous(), True), flattenDetection(ref_semi, True)
    assert (heat - ref_heat).abs().max().item() < 1e-4          # fp32 end to end: far inside the bf16 tolerance
    assert (semi - ref_semi).abs().max().item() < 2e-3
    assert (desc * ref_desc).sum(1).min().item() > 0.99999
    assert (desc.norm(dim=1) - 1).abs().max().item() < 1e-4
    if layout != "v1":
        other = torch.from_numpy(g[f"{layout}_eval_semi"]).cuda()
        assert (other - ref_semi).abs().max().item() > 10 * (semi - ref_semi).abs().max().item()


def test_forward_batch_bn_vs_oracle_larger_batch():
    """100 views of 64x96 (the export's batch: all views of one image): batch-statistics network vs the fp32 oracle."""
    from pytorch_superpoint_b200 import SuperPointNet_b200, flattenDetection
    sd = O.synthetic_state_dict("gauss2", seed=5)
    x = torch.rand(100, 1, 64, 96, generator=torch.Generator().manual_seed(3))
    n = SuperPointNet_b200(bn_mode="batch")
    n.load_state_dict(sd)
    semi, desc = n.forward_nhwc(x.cuda())
    ref_semi, ref_desc = O.net_forward(O.canonical_params(sd), x, bn_mode="batch")
    heat = flattenDetection(semi.permute(0, 3, 1, 2).contiguous(), True).cpu()
    ref_heat = flattenDetection(ref_semi.cuda(), True).cpu()
    assert (heat - ref_heat).abs().max().item() < 1e-4
    assert (desc.permute(0, 3, 1, 2).cpu() * ref_desc).sum(1).min().item() > 0.99999


def test_ha_export_batch_bn_vs_oracle():
    """Whole HA export of one image with train-mode BN (the as-written reference): aggregated heat-map vs the fp32
    oracle run with bn_mode='batch'; a plan that would split an image's views is refused, not silently changed."""
    from pytorch_superpoint_b200 import SuperPointNet_b200, make_ha_homographies
    from pytorch_superpoint_b200.export import HAExporter
    H, W, N = 64, 96, 24
    rng = np.random.RandomState(11)
    img = np.clip(0.4 + 0.05 * rng.randn(H, W), 0, 1).astype(np.float32)
    for _ in range(10):
        x0, y0 = rng.randint(0, W - 10), rng.randint(0, H - 10)
        img[y0:y0 + rng.randint(5, 20), x0:x0 + rng.randint(5, 30)] = rng.uniform(0, 1)
    mu, mf = make_ha_homographies(N, seed=4)
    sd = O.synthetic_state_dict("gauss2", seed=6)
    n = SuperPointNet_b200(bn_mode="batch")
    n.load_state_dict(sd)
    ex = HAExporter(n, conf_thresh=0.015, nms_dist=4, top_k=600)
    ti, tu, tf = torch.from_numpy(img), torch.from_numpy(mu), torch.from_numpy(mf)
    pts = ex.export_image(ti, tu, tf)
    heat = ex.last_heat(1, H, W)[0].cpu().numpy()
    ref_pts, ref_heat = O.ha_export_one(img, mu, mf, O.canonical_params(sd), bn_mode="batch", return_heat=True)
    _, eval_heat = O.ha_export_one(img, mu, mf, O.canonical_params(sd), bn_mode="eval", return_heat=True)
    err = np.abs(heat - ref_heat).max()
    assert err < 1e-3 and np.abs(eval_heat - ref_heat).max() > 5 * err
    assert np.array_equal(pts[:, :2], ref_pts[:, :2]) or len(set(map(tuple, pts[:, :2])) & set(map(tuple, ref_pts[:, :2]))) >= 0.98 * len(ref_pts)
    assert pts.shape[1] == 3 and pts.dtype == np.float64
    with pytest.raises(NotImplementedError):
        HAExporter(n, chunk=8).export_image(ti, tu, tf)
