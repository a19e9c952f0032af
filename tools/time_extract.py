"""Where single-frame extraction (bench.py --workload c3: batch 64 at 240x320, top-k 1000) spends its time: the whole
call next to its pieces (CUDA events, device resident).   python tools/time_extract.py [B=64]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # structured_images
from pytorch_superpoint_b200 import SuperPointNet_b200, _lib, ops
from pytorch_superpoint_b200.frontend import SuperPointFrontend_b200
from pytorch_superpoint_b200.synthetic import synthetic_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
net = SuperPointNet_b200(); net.load_state_dict(synthetic_state_dict("gauss2", seed=1))
fe = SuperPointFrontend_b200({"model": {"subpixel": {"enable": False}}}, "", 4, 0.015, 0.7, load=False, device="cuda:0")
fe.net = net
x = torch.from_numpy(bench.structured_images(B, 240, 320, seed=5)[:, None]).to(dev)
L = _lib.lib()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


semi, desc = net.forward_nhwc(x, want_desc=True)
heat = torch.empty((B, 240, 320), device=dev)
flat = lambda: _lib.check(L.spb_flatten_detection(semi.data_ptr(), 65, heat.data_ptr(), B, 30, 40, torch.cuda.current_stream().cuda_stream))
flat()
pts, counts, _ = ops.get_pts_from_heatmap_device(heat, 0.015, 4, 4, 1000)
print(f"B={B}: candidates above the threshold per image {float((heat >= 0.015).float().sum() / B):.0f}, key-points kept {float(counts.float().mean()):.0f}")
print(f"extract_batch (whole call)        {timed(lambda: fe.extract_batch(x, top_k=1000)):.4f} ms")
print(f"  network (detector + descriptor) {timed(lambda: net.forward_nhwc(x, want_desc=True)):.4f} ms")
print(f"  flatten_detection               {timed(flat):.4f} ms")
print(f"  threshold + NMS + top-k         {timed(lambda: ops.get_pts_from_heatmap_device(heat, 0.015, 4, 4, 1000)):.4f} ms")
print(f"  descriptor sampling             {timed(lambda: ops.sample_desc_batch_device(desc, pts, counts)):.4f} ms")
