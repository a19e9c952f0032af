"""Timeline probe of conv1_fused_kernel / conv1_pair_kernel (CTA 0): per-iteration clock64 stamps of the MMA, stage-1
and epilogue roles.   python tools/probe_fused.py [fuse1 mode: 1 single CTA (default), 2 CTA pairs]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_superpoint_b200 import SuperPointNet_b200, _lib
from pytorch_superpoint_b200.synthetic import synthetic_state_dict
L = _lib.lib()
L.spb_conv_tc_set_fuse1(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
net = SuperPointNet_b200(); net.load_state_dict(synthetic_state_dict("gauss2", 1))
x = torch.rand(100, 1, 240, 320, device="cuda")
net.forward_nhwc(x, want_desc=False); torch.cuda.synchronize()
buf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
L.spb_conv1_fused_set_probe(buf.data_ptr())
net.forward_nhwc(x, want_desc=False); torch.cuda.synchronize()
L.spb_conv1_fused_set_probe(None)
d = buf.cpu().view(64, 16)
t0 = int(d[8, 0])
names = ["mma:iter", "mma:acc2_empty", "mma:a2_full", "mma:token", "mma:conv1b_issued", "s1:iter", "s1:built",
         "s1:acc1_full", "s1:a2_empty", "s1:done", "e2:iter", "e2:acc2_full", "e2:done", "s1:bar", "s1:sts", "s1:arrived"]
for it in range(8, 14):
    print(f"it {it}: " + "  ".join(f"{names[k].replace(':', '.')}={int(d[it, k]) - t0}" for k in range(16)))
import statistics
per = [int(d[i + 1, 5] - d[i, 5]) for i in range(10, 60)]
print("S1 period median", statistics.median(per))
for a, b, nm in [(5, 13, "s1 wait patch"), (13, 14, "s1 im2col lds/sts"),
                 (14, 15, "s1 fence+arrive"), (15, 6, "s1 patch issue"), (6, 8, "s1 2 waits in parallel"), (8, 9, "s1 epi1"), (0, 1, "mma wait acc2_empty"),
                 (1, 2, "mma wait a2_full"), (2, 3, "mma wait token"), (3, 4, "mma issue conv1b"),
                 (10, 11, "e2 wait acc2_full"), (11, 12, "e2 work")]:
    v = [int(d[i, b] - d[i, a]) for i in range(10, 60)]
    print(f"{nm:32s} median {statistics.median(v):8.0f}  max {max(v)}")
v = [int(d[i + 2, 0] - d[i, 4]) for i in range(10, 58)]
print(f"{'mma issuer: conv1b done -> next':32s} median {statistics.median(v):8.0f}  max {max(v)}")
