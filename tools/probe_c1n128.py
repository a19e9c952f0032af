"""Timeline probe of conv1_n128_kernel (CTA 0): per-tile clock64 stamps of the MMA issuers, stage 1 and the conv1b
epilogue (see N1_PROBE in csrc/conv1_n128.cu)."""
import os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_superpoint_b200 import SuperPointNet_b200, _lib
from pytorch_superpoint_b200.synthetic import synthetic_state_dict
L = _lib.lib()
net = SuperPointNet_b200(); net.load_state_dict(synthetic_state_dict("gauss2", 1))
L.spb_conv_tc_set_fuse1(3)
x = torch.rand(100, 1, 240, 320, device="cuda")
net.forward_nhwc(x, want_desc=False); torch.cuda.synchronize()
buf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
L.spb_conv1_fused_set_probe(buf.data_ptr())
net.forward_nhwc(x, want_desc=False); torch.cuda.synchronize()
L.spb_conv1_fused_set_probe(None)
d = buf.cpu().view(64, 16)
names = ["mma.top", "mma.acc2_empty", "mma.a2_full", "mma.token", "mma.B_issued", "mma.A_issued", "s1.top",
         "s1.built", "s1.waited", "s1.done", "e2.top", "e2.acc2_full", "e2.converted", "e2.done"]
t0 = int(d[20, 0])
for it in range(20, 25):
    print(f"tile {it}: " + "  ".join(f"{names[k]}={int(d[it, k]) - t0}" for k in range(14)))
print("period per tile (s1.top)", statistics.median(int(d[i + 1, 6] - d[i, 6]) for i in range(10, 60)))
for a, b, nm in [(0, 1, "mma wait acc2_empty"), (1, 2, "mma wait a2_full"), (2, 3, "mma wait token"),
                 (3, 4, "mma issue conv1b"), (4, 5, "mma waits + issue conv1a"), (6, 7, "s1 patch wait + build"),
                 (7, 8, "s1 waits"), (8, 9, "s1 convert"), (10, 11, "e2 wait acc2_full"), (11, 12, "e2 convert"),
                 (12, 13, "e2 barrier + store")]:
    v = [int(d[i, b] - d[i, a]) for i in range(10, 60)]
    print(f"{nm:32s} median {statistics.median(v):8.0f}  max {max(v)}")
