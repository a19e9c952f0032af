"""The peer-memory exchange of the sharded HA path (sharding.PeerExchange + csrc/peer.cu: CUDA IPC mappings, ready /
consumed counters, fused reduce + divide) with TWO ranks as two processes on the one GPU of the test box: CUDA IPC
works between processes on the same device, the process group is gloo (NCCL refuses two ranks on one GPU).  Checked
against the unsharded exporter of rank 0 and across a change of batch shape (mappings released and rebuilt)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _images(B, H, W, seed):
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(B):
        img = np.clip(0.4 + 0.05 * rng.randn(H, W), 0, 1).astype(np.float32)
        for _ in range(8):
            x0, y0 = rng.randint(0, W - 8), rng.randint(0, H - 8)
            img[y0:y0 + rng.randint(4, 24), x0:x0 + rng.randint(4, 30)] = rng.uniform(0, 1)
        out.append(torch.from_numpy(img))
    return out


def _worker(rank, world, port, q, exchange):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["SPB200_EXCHANGE"] = exchange
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        from pytorch_superpoint_b200 import SuperPointNet_b200, make_ha_homographies
        from pytorch_superpoint_b200.export import HAExporter
        from pytorch_superpoint_b200.synthetic import synthetic_state_dict
        net = SuperPointNet_b200()
        net.load_state_dict(synthetic_state_dict("gauss2", 3))
        ex = HAExporter(net, 0.015, 4, 200)
        local = HAExporter(net, 0.015, 4, 200, sharded=False)
        res = []
        H, W, N = 48, 64, 7
        mu, mf = (torch.from_numpy(m) for m in make_ha_homographies(N, seed=2))
        # three batches of 4 (slot reuse: the consumed counters gate the third), then a tail batch of 3 (new shape)
        for step, B in enumerate([4, 4, 4, 3]):
            imgs = _images(B, H, W, seed=10 + step)
            pts = ex.export_batch(imgs, mu, mf)
            heat = ex.last_heat(B, H, W).clone()
            lo, hi = ex._buffers(B, H, W)["slots"][0]["own"]
            ref_pts = local.export_batch(imgs, mu, mf)
            ref_heat = local.last_heat(B, H, W)[lo:hi]
            err = float((heat - ref_heat).abs().max()) if hi > lo else 0.0
            jac = []
            for b in range(lo, hi):  # fp32 summation order flips a few points at the threshold of a flat heat-map
                a, c = {tuple(p[:2]) for p in pts[b]}, {tuple(p[:2]) for p in ref_pts[b]}
                jac.append(len(a & c) / max(len(a | c), 1))
            same = min(jac, default=1.0) >= 0.97
            owned_only = all((pts[b] is None) == (not lo <= b < hi) for b in range(B))
            res.append((step, B, (lo, hi), err, same, owned_only))
        kind = ex.exchange
        ex.close()
        q.put((rank, kind, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["peer"])
def test_peer_exchange_two_ranks_on_one_gpu(exchange):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, exchange)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = {}
    for rank, kind, res in got:
        assert kind == exchange
        for step, B, (lo, hi), err, same, owned_only in res:
            assert err < 1e-6 and owned_only, (rank, step, err)   # fp32 summation order only
            assert same, (rank, step)                             # key-points of the unsharded exporter (Jaccard >= 0.97)
            owned.setdefault(step, []).extend(range(lo, hi))
    assert {s: sorted(v) for s, v in owned.items()} == {0: [0, 1, 2, 3], 1: [0, 1, 2, 3], 2: [0, 1, 2, 3], 3: [0, 1, 2]}
