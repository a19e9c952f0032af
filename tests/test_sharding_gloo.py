"""world_size-2 (and 3, 4) gloo tests of the homography-sharding host logic on CPU.  They drive the functions the
product's ``HAExporter.submit_resident`` executes -- ``sharding.plan_batch`` (rotated view blocks, passes),
``sharding.sharded_heatmap_sums`` and ``sharding.exchange_sums`` (reduce-scatter over the image dimension, or
all-reduce + ownership slice) -- with the CPU oracle standing in for the CUDA kernels of a pass: the images each rank
ends up owning must equal the unsharded oracle up to fp32 summation order, every image exactly once."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(B, N, H, W):
    from pytorch_superpoint_b200.homographies import make_ha_homographies
    sets = [make_ha_homographies(N, seed=4 + b) for b in range(B)]
    mu, mf = np.stack([s[0] for s in sets]), np.stack([s[1] for s in sets])  # [B,N,3,3]: one set per image
    semi = (torch.randn(B, N, 65, H // 8, W // 8, generator=torch.Generator().manual_seed(0)) * 2).numpy()
    return mu, mf, semi


def _worker(rank, world, port, q, B, N, cap, async_op):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from pytorch_superpoint_b200.sharding import sharded_heatmap_sums
        torch.set_num_threads(1)
        H, W = 48, 64
        mu, mf, semi = _inputs(B, N, H, W)
        mu_flat, mf_flat, semi_flat = mu.reshape(-1, 3, 3), mf.reshape(-1, 3, 3), semi.reshape((B * N,) + semi.shape[2:])

        def partial_fn(p, sums):  # what net.ha_heatmap_sums_ragged does on the GPU, for one pass
            vi, vm, vo = p.view_image.numpy(), p.view_mat.numpy(), p.view_offsets.numpy()
            assert len(vi) == len(vm) == vo[-1] and len(vo) == p.nb + 1
            for j in range(p.nb):
                sel = vm[vo[j]:vo[j + 1]]
                assert (vi[vo[j]:vo[j + 1]] == j).all() and (sel // N == p.b0 + j).all() and len(sel) <= p.max_views
                part = torch.zeros(2, H, W)
                if len(sel):
                    heat = O.flatten_detection(semi_flat[sel])
                    mask = O.compute_valid_mask((H, W), mf_flat[sel])
                    part[0] = torch.from_numpy(O.inv_warp_image_batch(heat * mask, mu_flat[sel], "bilinear").sum(0))
                    part[1] = torch.from_numpy(O.inv_warp_image_batch(mask, mu_flat[sel], "bilinear").sum(0))
                if p.accumulate:
                    sums[p.b0 + j] += part
                else:
                    sums[p.b0 + j] = part

        sums = torch.full((B, 2, H, W), float("nan"))  # every plane must be written (or zeroed) by the step
        owned, work, plan = sharded_heatmap_sums(partial_fn, B, N, sums, max_pass_views=cap, async_op=async_op)
        if work is not None:
            work.wait()
        lo, hi = plan.own
        assert owned.shape[0] == hi - lo and plan.scatter == (B % world == 0)
        out = {lo + j: (owned[j, 0] / owned[j, 1]).numpy() for j in range(hi - lo)}
        q.put((rank, out, plan.local_views, len(plan.passes)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B,N,cap,async_op", [(2, 4, 7, 0, False), (3, 3, 7, 4, True), (2, 3, 5, 2, False),
                                                    (4, 4, 3, 0, True)])
def test_sharded_equals_unsharded(world, B, N, cap, async_op):
    """(2,4,7): reduce-scatter, uneven shares rotating; (3,3,7,cap 4): several passes; (2,3,5,cap 2): all-reduce
    (B % world != 0) with images split over accumulating passes; (4,4,3): more ranks than views (idle shares)."""
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, B, N, cap, async_op)) for r in range(world)]
    for p in procs:
        p.start()
    got, views = {}, []
    for _ in range(world):
        rank, out, local_views, npasses = q.get(timeout=180)
        assert not (set(out) & set(got))  # every image finalised by exactly one rank
        got.update(out)
        views.append(local_views)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    H, W = 48, 64
    mu, mf, semi = _inputs(B, N, H, W)
    assert sorted(got) == list(range(B)) and sum(views) == B * N
    if B % world == 0:
        assert max(views) == min(views) or N < world  # rotated shares: every rank has the same number of views
    for b in range(B):
        ref = O.combine_heatmap(O.flatten_detection(semi[b]), mu[b], O.compute_valid_mask((H, W), mf[b]))
        assert np.abs(got[b] - ref).max() < 1e-6
        a = O.get_pts_from_heatmap(got[b], 0.015, 4)
        r = O.get_pts_from_heatmap(ref, 0.015, 4)
        sa, sr = {tuple(c[:2]) for c in a.T}, {tuple(c[:2]) for c in r.T}
        assert len(sa & sr) >= 0.97 * max(len(sa | sr), 1)  # threshold/tie flips from summation order only


def test_plan_batch_covers_every_view_once_and_respects_the_pass_size():
    from pytorch_superpoint_b200.sharding import owned_images, owner_of_image, plan_batch
    for B, N, world, cap in [(32, 100, 8, 800), (32, 100, 1, 800), (8, 100, 8, 128), (4, 100, 1, 30), (6, 10, 4, 0),
                             (5, 7, 3, 4), (3, 2, 4, 0), (64, 100, 8, 800)]:
        seen = torch.zeros(B * N, dtype=torch.int64)
        owned = []
        for r in range(world):
            plan = plan_batch(B, N, r, world, cap)
            owned.append(plan.own)
            assert plan.own == owned_images(B, r, world)
            touched = set()
            for p in plan.passes:
                assert 0 < p.num_views <= (cap or 1 << 30) and p.view_offsets[-1] == p.num_views
                assert len(p.view_offsets) == p.nb + 1 and p.max_views == int((p.view_offsets[1:] - p.view_offsets[:-1]).max())
                seen[p.view_mat.long()] += 1
                assert (p.view_mat // N == p.view_image + p.b0).all()
                imgs = set(range(p.b0, p.b0 + p.nb))
                if p.accumulate:  # a slice of a split image: alone in its pass, and an earlier pass wrote the image
                    assert p.nb == 1 and p.b0 in touched
                else:
                    assert not (imgs & touched)
                touched |= imgs
            assert plan.local_views == sum(p.num_views for p in plan.passes)
        assert (seen == 1).all()
        assert owned[0][0] == 0 and owned[-1][1] == B and all(a[1] == b[0] for a, b in zip(owned, owned[1:]))
        for b in range(B):
            r = owner_of_image(b, world, B)
            assert owned[r][0] <= b < owned[r][1]


def test_balanced_view_plan_partitions_every_pair_once():
    """Rotated shares: every (image, view) pair belongs to exactly one rank, every image is spread over all ranks, and
    with B % world == 0 all ranks get the same number of views per step (100 views over 8 ranks: 400 each, not 416/384)."""
    import torch
    from pytorch_superpoint_b200.sharding import balanced_view_plan, shard_bounds
    for B, N, world in [(32, 100, 8), (8, 100, 8), (6, 10, 4), (5, 7, 3), (16, 100, 3)]:
        seen = torch.zeros(B * N, dtype=torch.int64)
        totals = []
        for r in range(world):
            sel, vi, vo, most = balanced_view_plan(B, N, r, world)
            seen[sel] += 1
            totals.append(len(sel))
            assert vo[0] == 0 and vo[-1] == len(sel) and len(vo) == B + 1 and len(vi) == len(sel)
            for b in range(B):
                lo, hi = int(vo[b]), int(vo[b + 1])
                assert (vi[lo:hi] == b).all() and hi - lo <= most
                assert (sel[lo:hi] // N == b).all() and hi > lo            # every rank touches every image
                want = shard_bounds(N, (r + b) % world, world)
                assert sel[lo] % N == want[0] and sel[hi - 1] % N == want[1] - 1
        assert (seen == 1).all()
        if B % world == 0:
            assert max(totals) == min(totals) == B * N // world
        assert max(totals) - min(totals) <= max(1, B % world)


def test_exchange_hook_replaces_the_collective():
    """``sharded_heatmap_sums(..., exchange=fn)``: the partial sums are computed exactly as for the collective path and
    handed to ``fn`` (the peer-memory exchange of the GPU path, sharding.PeerExchange) instead of exchange_sums."""
    from pytorch_superpoint_b200.sharding import plan_batch, sharded_heatmap_sums
    B, N, H, W = 3, 5, 4, 6
    sums = torch.full((B, 2, H, W), 7.0)
    seen = {}

    def partial_fn(p, out):
        for j in range(p.nb):
            n = int(p.view_offsets[j + 1] - p.view_offsets[j])
            if p.accumulate:
                out[p.b0 + j] += n
            else:
                out[p.b0 + j] = n

    def exchange(s, plan):
        seen["sums"], seen["plan"] = s.clone(), plan
        return "owned", "work"

    owned, work, plan = sharded_heatmap_sums(partial_fn, B, N, sums, plan=plan_batch(B, N, 0, 1, 2), exchange=exchange)
    assert (owned, work) == ("owned", "work") and seen["plan"] is plan
    assert torch.equal(seen["sums"], torch.full((B, 2, H, W), float(N)))  # every view of every image counted once
