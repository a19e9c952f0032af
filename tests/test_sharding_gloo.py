"""world_size-2 (and 3) gloo test of the homography-sharding host logic on CPU: each rank computes the partial
(sum heat*mask, sum mask) of its block of views with the CPU oracle, one all_reduce, then every rank divides
and extracts key-points -- the result must equal the unsharded oracle up to fp32 summation order."""
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


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from pytorch_superpoint_b200.homographies import make_ha_homographies
        from pytorch_superpoint_b200.sharding import owner_of_image, sharded_heatmap_sums
        torch.set_num_threads(1)
        H, W, N, B = 48, 64, 7, 3
        mu, mf = make_ha_homographies(N, seed=4)
        semi = (torch.randn(B, N, 65, H // 8, W // 8, generator=torch.Generator().manual_seed(0)) * 2).numpy()

        def partial(b, lo, hi):
            if hi <= lo:
                return torch.zeros(2, H, W)
            heat = O.flatten_detection(semi[b, lo:hi])
            mask = O.compute_valid_mask((H, W), mf[lo:hi])
            num = O.inv_warp_image_batch(heat * mask, mu[lo:hi], "bilinear").sum(0)
            den = O.inv_warp_image_batch(mask, mu[lo:hi], "bilinear").sum(0)
            return torch.from_numpy(np.stack([num, den]))

        sums = sharded_heatmap_sums(partial, B, N)
        out = {}
        for b in range(B):
            if owner_of_image(b, world) == rank:
                heat = (sums[b, 0] / sums[b, 1]).numpy()
                out[b] = heat
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded(world):
    from oracle import oracle as O
    from pytorch_superpoint_b200.homographies import make_ha_homographies
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, out = q.get(timeout=120)
        got.update(out)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    H, W, N, B = 48, 64, 7, 3
    mu, mf = make_ha_homographies(N, seed=4)
    semi = (torch.randn(B, N, 65, H // 8, W // 8, generator=torch.Generator().manual_seed(0)) * 2).numpy()
    assert sorted(got) == list(range(B))  # every image finalised by exactly one rank
    for b in range(B):
        ref = O.combine_heatmap(O.flatten_detection(semi[b]), mu, O.compute_valid_mask((H, W), mf))
        assert np.abs(got[b] - ref).max() < 1e-6
        a = O.get_pts_from_heatmap(got[b], 0.015, 4)
        r = O.get_pts_from_heatmap(ref, 0.015, 4)
        sa, sr = {tuple(c[:2]) for c in a.T}, {tuple(c[:2]) for c in r.T}
        assert len(sa & sr) >= 0.97 * max(len(sa | sr), 1)  # threshold/tie flips from summation order only
