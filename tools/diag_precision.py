"""Per-layer and whole-network numerical error of the CUDA paths vs fp32 torch (diagnostic)."""
import os, sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from pytorch_superpoint_b200 import SuperPointNet_b200, _lib
SHAPES = [(1, 64, 3), (64, 64, 3), (64, 64, 3), (64, 64, 3), (64, 128, 3), (128, 128, 3), (128, 128, 3),
          (128, 128, 3), (128, 256, 3), (256, 65, 1), (128, 256, 3), (256, 256, 1)]
POOL = [0, 1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0]
L = _lib.lib()
sd = O.synthetic_state_dict("gauss2", 1)
net = SuperPointNet_b200(); net.load_state_dict(sd)
p = O.fold_bn(O.canonical_params(sd))
def layer(li, x, B, H, W, impl):
    h = net.handle(x.device); net.set_impl(impl)
    cin, cout, _ = SHAPES[li]
    oh, ow = (H // 2, W // 2) if POOL[li] else (H, W)
    out = torch.empty((B, oh, ow, cout), device="cuda", dtype=torch.float32 if li in (9, 11) else torch.bfloat16)
    _lib.check(L.spb_net_layer(h, li, x.data_ptr(), out.data_ptr(), B, H, W, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize(); return out.float()
torch.manual_seed(0)
for li in range(1, 12):
    cin = SHAPES[li][0]; B, H, W = 2, 32, 40
    x = (torch.randn(B, H, W, cin, device="cuda") * 0.5).to(torch.bfloat16)
    name = O.LAYERS[li][0]
    w = p[name]["w"].to(torch.bfloat16).double().cuda()
    y = F.conv2d(x.double().permute(0, 3, 1, 2), w, p[name]["b"].double().cuda(), padding=w.shape[-1] // 2)
    if O.LAYERS[li][4]: y = torch.relu(y)
    if O.LAYERS[li][5]: y = F.max_pool2d(y, 2, 2)
    if li == 11: y = y / y.norm(dim=1, keepdim=True)
    y = y.permute(0, 2, 3, 1).float()
    for impl in ("simt", "tcgen05"):
        o = layer(li, x, B, H, W, impl)
        err = (o - y).abs()
        rel = err.max().item() / y.abs().max().item()
        print(f"layer {li:2d} {name} {impl:8s} max|err| {err.max().item():.3e}  mean|err| {err.mean().item():.3e}  max|y| {y.abs().max().item():.2f} rel {rel:.2e}")
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "net.npz"))
x = torch.from_numpy(g["x"]).cuda()
ref = torch.from_numpy(g["gauss2_eval_desc"]).cuda(); refs = torch.from_numpy(g["gauss2_eval_semi"]).cuda()
for impl in ("simt", "tcgen05"):
    for fuse in (0, 1):
        net.set_impl(impl); L.spb_conv_tc_set_fuse1(fuse)
        o = net(x)
        c = (o["desc"] * ref).sum(1)
        print(impl, "fuse", fuse, "desc cos min", c.min().item(), "mean", c.mean().item(), "semi max err", (o["semi"] - refs).abs().max().item())
