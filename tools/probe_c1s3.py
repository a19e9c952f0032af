"""Timeline probe of conv1_s3_kernel (CTA 0): per-tile clock64 stamps of the MMA issuers, the conv1a epilogue and
the conv1b epilogue / im2col groups (see C1_PROBE in csrc/conv1_s3.cu)."""
import os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_superpoint_b200 import SuperPointNet_b200, _lib
from pytorch_superpoint_b200.synthetic import synthetic_state_dict
L = _lib.lib()
net = SuperPointNet_b200(); net.load_state_dict(synthetic_state_dict("gauss2", 1))
L.spb_conv_tc_set_fuse1(2)
x = torch.rand(100, 1, 240, 320, device="cuda")
net.forward_nhwc(x, want_desc=False); torch.cuda.synchronize()
buf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
L.spb_conv1_fused_set_probe(buf.data_ptr())
net.forward_nhwc(x, want_desc=False); torch.cuda.synchronize()
L.spb_conv1_fused_set_probe(None)
d = buf.cpu().view(64, 16)
names = ["mma.top", "mma.A_issued", "mma.acc2e+a2f", "mma.token", "mma.B_issued", "-", "s1.top",
         "s1.waited", "s1.done", "e2.top", "e2.built", "e2.acc2_full", "e2.done"]
t0 = int(d[20, 0])
for it in range(20, 26):
    print(f"tile {it}: " + "  ".join(f"{names[k]}={int(d[it, k]) - t0}" for k in range(13)))
print("period per tile (s1.top)", statistics.median(int(d[i + 1, 6] - d[i, 6]) for i in range(10, 60)))
for a, b, nm in [(0, 1, "mma waits + issue conv1a(i+1)"), (1, 2, "mma wait acc2_empty, a2_full"), (2, 3, "mma wait token"),
                 (3, 4, "mma issue conv1b"), (6, 7, "s1 waits"),
                 (7, 8, "s1 convert"), (9, 10, "e2 build (waits + im2col)"), (10, 11, "e2 wait acc2_full"),
                 (11, 12, "e2 convert + store")]:
    v = [int(d[i, b] - d[i, a]) for i in range(10, 60)]
    print(f"{nm:32s} median {statistics.median(v):8.0f}  max {max(v)}")
