"""Timeline probe of conv1_ws_kernel (CTA 0): per-tile clock64 stamps of the pair issuers, the conv1a issuer, stage 1
and the epilogue (slots: see WS_PROBE in csrc/conv1_ws.cu)."""
import os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_superpoint_b200 import SuperPointNet_b200, _lib
from pytorch_superpoint_b200.synthetic import synthetic_state_dict
L = _lib.lib()
L.spb_conv_tc_set_fuse1(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
net = SuperPointNet_b200(); net.load_state_dict(synthetic_state_dict("gauss2", 1))
x = torch.rand(100, 1, 240, 320, device="cuda")
net.forward_nhwc(x, want_desc=False); torch.cuda.synchronize()
buf = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
L.spb_conv1_fused_set_probe(buf.data_ptr())
net.forward_nhwc(x, want_desc=False); torch.cuda.synchronize()
L.spb_conv1_fused_set_probe(None)
d = buf.cpu().view(64, 16)
t0 = int(d[8, 6])
names = ["iss.top", "iss.ready", "iss.token", "iss.issued", "c1a.ready", "c1a.issued", "s1.top", "s1.built", "s1.waited",
         "s1.done", "e2.top", "e2.acc", "e2.done"]
for j in range(8, 20):
    print(f"tile {j}: " + "  ".join(f"{names[k]}={int(d[j, k]) - t0 if int(d[j, k]) else 0}" for k in range(13)))
med = lambda v: statistics.median(v)
print("S1 period per tile       ", med([int(d[j + 1, 6] - d[j, 6]) for j in range(10, 60)]))
print("s1 build next            ", med([int(d[j, 7] - d[j, 6]) for j in range(10, 60)]))
print("s1 wait acc1_full/a2_empty", med([int(d[j, 8] - d[j, 7]) for j in range(10, 60)]))
print("s1 convert + store       ", med([int(d[j, 9] - d[j, 8]) for j in range(10, 60)]))
ev = list(range(12, 60, 2))
print("issuer pair period       ", med([int(d[j + 2, 0] - d[j, 0]) for j in range(12, 56, 2)]) , "(two issuers alternate: per-issuer period / 2 pairs)")
print("issuer wait operands     ", med([int(d[j, 1] - d[j, 0]) for j in ev]))
print("issuer wait token        ", med([int(d[j, 2] - d[j, 1]) for j in ev]))
print("issuer issue 72 MMAs     ", med([int(d[j, 3] - d[j, 2]) for j in ev]))
print("pair start -> next pair start (token to token)", med([int(d[j + 2, 2] - d[j, 2]) for j in range(12, 56, 2)]))
print("c1a wait                 ", med([int(d[j, 4] - d[j - 1, 5]) for j in range(10, 60)]))
print("e2 wait acc              ", med([int(d[j, 11] - d[j, 10]) for j in range(10, 60)]))
print("e2 work                  ", med([int(d[j, 12] - d[j, 11]) for j in range(10, 60)]))
print("a2 written -> pair token ", med([int(d[j, 2] - d[j + 1, 9]) for j in ev]))
