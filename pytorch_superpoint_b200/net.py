"""``SuperPointNet_b200`` -- drop-in for the reference networks behind ``utils/loader.py:155-164 modelLoader``.

Same constructor / ``forward`` / ``load_state_dict`` surface as ``models/SuperPointNet_gauss2.py:11-70``
(and ``models/SuperPointNet.py``, ``models/SuperPointNet_pretrained.py``): ``forward(x[B,1,H,W])`` returns
``{'semi': [B,65,H/8,W/8], 'desc': [B,256,H/8,W/8]}``.  Any of the three reference state-dict layouts is
accepted; BatchNorm runs in eval mode (running statistics), folded into the conv weights -- the product path,
see DESIGN.md.  ``bn_mode='batch'`` is the opt-in compatibility path for the as-written reference, which never
calls ``.eval()`` (models/model_wrap.py:111): every BN then uses the statistics of the current batch (CUDA-core
kernels + three passes per layer, ``csrc/bn.cu``).  The math runs in libspb200.so; PyTorch only owns the buffers.
"""
from __future__ import annotations

import collections
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib

LAYER_NAMES = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b",
               "convPa", "convPb", "convDa", "convDb"]
LAYER_SHAPES = [(1, 64, 3), (64, 64, 3), (64, 64, 3), (64, 64, 3), (64, 128, 3), (128, 128, 3), (128, 128, 3),
                (128, 128, 3), (128, 256, 3), (256, 65, 1), (128, 256, 3), (256, 256, 1)]

# SuperPointNet_gauss2 / unet_parts naming (models/unet_parts.py:14-21,41-44): (conv, bn) module paths
_GAUSS2 = {
    "conv1a": ("inc.conv.conv.0", "inc.conv.conv.1"), "conv1b": ("inc.conv.conv.3", "inc.conv.conv.4"),
    "conv2a": ("down1.mpconv.1.conv.0", "down1.mpconv.1.conv.1"),
    "conv2b": ("down1.mpconv.1.conv.3", "down1.mpconv.1.conv.4"),
    "conv3a": ("down2.mpconv.1.conv.0", "down2.mpconv.1.conv.1"),
    "conv3b": ("down2.mpconv.1.conv.3", "down2.mpconv.1.conv.4"),
    "conv4a": ("down3.mpconv.1.conv.0", "down3.mpconv.1.conv.1"),
    "conv4b": ("down3.mpconv.1.conv.3", "down3.mpconv.1.conv.4"),
    "convPa": ("convPa", "bnPa"), "convPb": ("convPb", "bnPb"),
    "convDa": ("convDa", "bnDa"), "convDb": ("convDb", "bnDb"),
}


def _norm_device(device=None) -> torch.device:
    """Explicit CUDA device index (``cuda`` != ``cuda:0`` for torch.device comparisons)."""
    d = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if d.type != "cuda":
        raise RuntimeError("SuperPointNet_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())


def _layer_params(state_dict):
    """Any reference layout -> per layer (weight OIHW, bias, (gamma, beta, running_mean, running_var) or None), float64."""
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
    gauss2 = "inc.conv.conv.0.weight" in sd
    for name, (cin, cout, ks) in zip(LAYER_NAMES, LAYER_SHAPES):
        cname, bname = _GAUSS2[name] if gauss2 else (name, "bn" + name[4:])
        if cname + ".weight" not in sd:
            raise KeyError(f"state_dict has no '{cname}.weight' (not a SuperPointNet layout)")
        w = sd[cname + ".weight"].detach().to("cpu", torch.float64)
        b = sd[cname + ".bias"].detach().to("cpu", torch.float64)
        if tuple(w.shape) != (cout, cin, ks, ks):
            raise ValueError(f"{cname}.weight has shape {tuple(w.shape)}, expected {(cout, cin, ks, ks)}")
        bn = None
        if bname + ".running_mean" in sd:
            bn = tuple(sd[bname + s].detach().to("cpu", torch.float64)
                       for s in (".weight", ".bias", ".running_mean", ".running_var"))
        yield w, b, bn


def _as_np(t):
    return np.ascontiguousarray(t.to(torch.float32).numpy())


def fold_state_dict(state_dict, eps: float = 1e-5):
    """Any reference layout -> 12 (weight OIHW fp32, bias fp32) numpy pairs with eval-mode BN folded in."""
    ws, bs = [], []
    for w, b, bn in _layer_params(state_dict):
        if bn is not None:
            g, beta, m, v = bn
            s = g / torch.sqrt(v + eps)
            w = w * s[:, None, None, None]
            b = (b - m) * s + beta
        ws.append(_as_np(w))
        bs.append(_as_np(b))
    return ws, bs


def raw_state_dict(state_dict):
    """Any reference layout -> 12 RAW (weight, bias) numpy pairs and the per-layer BN affine ``(gamma, beta)`` (or
    None for a layer without BN): what the train-mode (batch statistics) path needs."""
    ws, bs, affine = [], [], []
    for w, b, bn in _layer_params(state_dict):
        ws.append(_as_np(w))
        bs.append(_as_np(b))
        affine.append(None if bn is None else (bn[0].to(torch.float32), bn[1].to(torch.float32)))
    return ws, bs, affine


class SuperPointNet_b200(nn.Module):
    """B200-native SuperPoint network.  ``SuperPointNet_b200(subpixel_channel=1)`` like the reference."""

    def __init__(self, subpixel_channel=1, impl: str = "tcgen05", bn_mode: str = "eval"):
        super().__init__()
        if bn_mode not in ("eval", "batch"):
            raise ValueError(f"bn_mode must be 'eval' or 'batch', got {bn_mode!r}")
        self.bn_mode = bn_mode
        self.bn_eps = 1e-5  # nn.BatchNorm2d default, what every reference network uses
        self._sd = collections.OrderedDict()
        self._handle = None
        self._handle_dev = None
        self._raw = None  # bn_mode='batch': fp32 raw conv weights + BN affine parameters on the device
        self._ws = None
        self._impl = impl
        self._ha_overlap = False  # C-handle settings live here and are re-applied whenever the handle is rebuilt
        self._profile = False
        self.output = None
        # default initialisation = torch's Conv2d default, gauss2 naming (reference default model)
        for name, (cin, cout, ks) in zip(LAYER_NAMES, LAYER_SHAPES):
            conv = nn.Conv2d(cin, cout, ks, padding=ks // 2)
            cname, bname = _GAUSS2[name]
            self._sd[cname + ".weight"] = conv.weight.detach().clone()
            self._sd[cname + ".bias"] = conv.bias.detach().clone()
            self._sd[bname + ".weight"] = torch.ones(cout)
            self._sd[bname + ".bias"] = torch.zeros(cout)
            self._sd[bname + ".running_mean"] = torch.zeros(cout)
            self._sd[bname + ".running_var"] = torch.ones(cout)
            self._sd[bname + ".num_batches_tracked"] = torch.tensor(0)

    # ---- state dict: keep whatever reference layout was loaded ------------------------------------------
    def state_dict(self, *args, **kwargs):
        return collections.OrderedDict(self._sd)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        fold_state_dict(state_dict)  # validates names and shapes
        self._sd = collections.OrderedDict((k, v.detach().clone().cpu()) for k, v in state_dict.items())
        self._release()
        return torch.nn.modules.module._IncompatibleKeys([], [])

    # ---- C handle ---------------------------------------------------------------------------------------
    def _release(self):
        if self._handle is not None:
            _lib.lib().spb_net_destroy(self._handle)
        self._handle = None
        self._raw = None
        self._ws = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def set_impl(self, impl: str):
        self._impl = impl
        if self._handle is not None:
            _lib.check(_lib.lib().spb_net_set_impl(self._handle, {"tcgen05": 0, "simt": 1}[impl]), "net_set_impl")

    def handle(self, dev):
        dev = _norm_device(dev)
        if self._handle is not None and self._handle_dev != dev:
            self._release()
        if self._handle is None:
            L = _lib.lib()
            ws, bs = fold_state_dict(self._sd)
            arr = (C.c_void_p * 12)(*[w.ctypes.data for w in ws])
            brr = (C.c_void_p * 12)(*[b.ctypes.data for b in bs])
            h = C.c_void_p()
            with torch.cuda.device(dev):
                _lib.check(L.spb_net_create(C.byref(h), arr, brr), "net_create")
            self._handle, self._handle_dev = h, dev
            self.set_impl(self._impl)
            _lib.check(L.spb_net_set_ha_overlap(h, int(self._ha_overlap)), "set_ha_overlap")
            _lib.check(L.spb_net_profile(h, int(self._profile)), "net_profile")
        return self._handle

    def set_profile(self, on: bool, device=None) -> None:
        """Per-stage CUDA-event timing inside the library (bench.py's roofline); read with spb_net_profile_read."""
        self._profile = bool(on)
        if self._handle is not None or device is not None:
            _lib.check(_lib.lib().spb_net_profile(self.handle(device if device is not None else self._handle_dev),
                                                  int(self._profile)), "net_profile")

    def _workspace(self, nbytes: int, dev):
        dev = _norm_device(dev)
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        return self._ws

    # ---- forward ----------------------------------------------------------------------------------------
    def forward_nhwc(self, x: torch.Tensor, want_desc: bool = True):
        """x [B,1,H,W] cuda fp32 -> (semi [B,Hc,Wc,65] fp32, desc [B,Hc,Wc,256] fp32 or None)."""
        if not x.is_cuda:
            if not torch.cuda.is_available():
                raise RuntimeError("SuperPointNet_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            x = x.cuda()
        x = x.to(torch.float32).contiguous()
        B, ch, H, W = x.shape
        if ch != 1 or H % 8 or W % 8:
            raise ValueError(f"expected [B,1,H,W] with H, W multiples of 8, got {tuple(x.shape)}")
        dev = x.device
        if self.bn_mode == "batch":
            return self._forward_batch_bn(x.reshape(B, H, W), want_desc)
        h = self.handle(dev)
        L = _lib.lib()
        semi = torch.empty((B, H // 8, W // 8, 65), device=dev, dtype=torch.float32)
        desc = torch.empty((B, H // 8, W // 8, 256), device=dev, dtype=torch.float32) if want_desc else None
        nbytes = L.spb_net_workspace_bytes(B, H, W, int(want_desc))
        ws = self._workspace(nbytes, dev)
        with torch.cuda.device(dev):
            _lib.check(L.spb_net_forward(h, x.data_ptr(), B, H, W, semi.data_ptr(),
                                         desc.data_ptr() if want_desc else None, ws.data_ptr(), ws.numel(),
                                         torch.cuda.current_stream().cuda_stream), "net_forward")
        return semi, desc

    # ---- train-mode BatchNorm (batch statistics): the as-written reference, opt-in ------------------------
    def _raw_weights(self, dev):
        """fp32 device copies of the RAW conv weights as [tap][cin][cout_pad] (+ bias [cout_pad]) and the BN affine
        parameters of every layer: what spb_conv_f32 / spb_bn_batch_stats read."""
        dev = _norm_device(dev)
        if self._raw is None or self._raw["dev"] != dev:
            ws, bs, affine = raw_state_dict(self._sd)
            layers = []
            for (cin, cout, ks), w, b, a in zip(LAYER_SHAPES, ws, bs, affine):
                cpad = -(-cout // 16) * 16
                wt = torch.zeros((ks * ks, cin, cpad), dtype=torch.float32)
                wt[:, :, :cout] = torch.from_numpy(w).permute(2, 3, 1, 0).reshape(ks * ks, cin, cout)
                bt = torch.zeros(cpad, dtype=torch.float32)
                bt[:cout] = torch.from_numpy(b)
                layers.append(dict(w=wt.to(dev), b=bt.to(dev), cpad=cpad,
                                   affine=None if a is None else (a[0].to(dev).contiguous(), a[1].to(dev).contiguous())))
            self._raw = dict(dev=dev, layers=layers)
        return self._raw

    def _forward_batch_bn(self, x: torch.Tensor, want_desc: bool):
        """x [B,H,W] cuda fp32 -> (semi, desc) NHWC fp32 like forward_nhwc, every BatchNorm normalising with the
        statistics of THIS batch (F.batch_norm(training=True): biased variance), as models/unet_parts.py:14-21 /
        SuperPointNet_gauss2.py:56-69 do when the module is left in train mode.  Per layer, all fp32: convolution ->
        per-channel statistics -> scale / shift -> normalise + ReLU + 2x2 max-pool (csrc/bn.cu)."""
        dev = x.device
        raw = self._raw_weights(dev)
        L = _lib.lib()
        B, H, W = x.shape
        st = torch.cuda.current_stream().cuda_stream
        f32 = dict(device=dev, dtype=torch.float32)
        ws = torch.empty(L.spb_bn_batch_stats_workspace_bytes(256), device=dev, dtype=torch.uint8)

        def block(li, inp, hh, ww, relu, pool, l2norm=False):
            cin, cout, ks = LAYER_SHAPES[li]
            lay = raw["layers"][li]
            rows = torch.empty((B * hh * ww, cout), **f32)
            _lib.check(L.spb_conv_f32(inp.data_ptr(), lay["w"].data_ptr(), lay["b"].data_ptr(), rows.data_ptr(), B, hh, ww,
                                      cin, cout, lay["cpad"], ks, st), "conv_f32")
            scale, shift = torch.empty(cout, **f32), torch.empty(cout, **f32)
            if lay["affine"] is None:  # a network without BN (superpoint_v1 layout): identity transform
                scale.fill_(1.0)
                shift.zero_()
            else:
                g, beta = lay["affine"]
                _lib.check(L.spb_bn_batch_stats(rows.data_ptr(), rows.shape[0], cout, g.data_ptr(), beta.data_ptr(),
                                                self.bn_eps, None, None, scale.data_ptr(), shift.data_ptr(),
                                                ws.data_ptr(), ws.numel(), st), "bn_batch_stats")
            oh, ow = (hh // 2, ww // 2) if pool else (hh, ww)
            out = torch.empty((B, oh, ow, cout), **f32)
            _lib.check(L.spb_bn_apply(rows.data_ptr(), B, hh, ww, cout, scale.data_ptr(), shift.data_ptr(), int(relu),
                                      int(pool), out.data_ptr(), 1, int(l2norm), st), "bn_apply")
            return out, oh, ow

        with torch.cuda.device(dev):
            t, hh, ww = x.contiguous(), H, W  # [B,H,W] = NHWC with one channel
            for li in range(8):
                t, hh, ww = block(li, t, hh, ww, True, li in (1, 3, 5))
            pa, _, _ = block(8, t, hh, ww, True, False)
            semi, _, _ = block(9, pa, hh, ww, False, False)
            desc = None
            if want_desc:
                da, _, _ = block(10, t, hh, ww, True, False)
                desc, _, _ = block(11, da, hh, ww, False, False, l2norm=True)
        return semi, desc

    def _ha_sums_batch_bn(self, img, mu, mf, sums, accumulate):
        """One image, ALL of its views in one batch (their statistics are the BN statistics): the reference's own
        warp (spb_inv_warp_image_batch) -> batch-statistics network -> fused aggregate."""
        from . import ops
        H, W = img.shape
        views = ops.inv_warp_image_batch(img, mf)  # [N,1,H,W]
        semi, _ = self.forward_nhwc(views, want_desc=False)
        return ops.ha_aggregate(semi, mu, mf, H, W, sums=sums, accumulate=accumulate)

    def forward(self, x):
        semi, desc = self.forward_nhwc(x, want_desc=True)
        output = {"semi": semi.permute(0, 3, 1, 2), "desc": desc.permute(0, 3, 1, 2)}
        self.output = output
        return output

    def process_output(self, sp_processer):
        """models/SuperPointNet_gauss2.py:72-99: after forward(), flatten -> NMS -> soft-argmax residuals -> fixed
        number of (points, residuals, descriptors) per image; ``sp_processer`` is a SuperPointNet_process_b200."""
        from .ops import flattenDetection
        output = self.output
        semi, desc = output["semi"], output["desc"]
        heatmap = flattenDetection(semi.contiguous(), True)
        heatmap_nms_batch = sp_processer.heatmap_to_nms(heatmap, tensor=True)
        residual = sp_processer.pred_soft_argmax(heatmap_nms_batch, heatmap)["pred"]
        output.update(sp_processer.batch_extract_features(desc, heatmap_nms_batch, residual))
        self.output = output
        return output

    # ---- whole HA step (datasets/Coco.py:262-284 + export.py:309-310) -----------------------------------
    def ha_heatmap_sums(self, img, mats_unwarp, mats_fwd, sums=None, accumulate=False, chunk: int = 0):
        """img [H,W] cuda fp32; mats [N,3,3] (this rank's views) -> sums [2,H,W] (sum heat*mask, sum mask)."""
        H, W = img.shape[-2:]
        out = self.ha_heatmap_sums_batch(img.reshape(1, H, W), mats_unwarp, mats_fwd,
                                         None if sums is None else sums.reshape(1, 2, H, W), accumulate, chunk)
        return out[0]

    def ha_pass_views(self, H: int, W: int, device=None) -> int:
        """Views the library puts into one pass at this resolution (bounded by bytes; spb_ha_pass_views)."""
        dev = _norm_device(device)
        return int(_lib.lib().spb_ha_pass_views(self.handle(dev), int(H), int(W)))

    def ha_heatmap_sums_ragged(self, imgs, mats_unwarp, mats_fwd, view_image, view_offsets, max_views, sums=None,
                               accumulate=False, view_mat=None):
        """Homography sharding, ONE pass: imgs [B,H,W]; V views as a flat list -- ``view_image`` int32 [V] (image of
        every view, grouped by image), ``view_offsets`` int32 [B+1], ``max_views`` = the largest per-image count;
        ``view_mat`` int32 [V] indexes the matrix arrays (any [M,3,3] / [B,N,3,3], flattened), or None with mats
        [V,3,3] listed in view order -> sums [B,2,H,W]."""
        dev = imgs.device
        B, H, W = imgs.shape
        mu = mats_unwarp.to(dev, torch.float32).contiguous()
        mf = mats_fwd.to(dev, torch.float32).contiguous()
        vi = view_image.to(dev, torch.int32).contiguous()
        vo = view_offsets.to(dev, torch.int32).contiguous()
        V = vi.numel()
        vm = None if view_mat is None else view_mat.to(dev, torch.int32).contiguous()
        if vo.numel() != B + 1 or (vm is None and mu.reshape(-1, 3, 3).shape[0] != V) or (vm is not None and vm.numel() != V):
            raise ValueError("view_image / view_mat must be [V] and view_offsets [B+1]")
        if self.bn_mode == "batch":
            # the caller guarantees that the pass holds EVERY view of each of its images (HAExporter checks its plan)
            if sums is None:
                sums = torch.empty((B, 2, H, W), device=dev, dtype=torch.float32)
                accumulate = False
            off = vo.cpu().tolist()
            mu3, mf3 = mu.reshape(-1, 3, 3), mf.reshape(-1, 3, 3)
            for b in range(B):
                idx = torch.arange(off[b], off[b + 1], device=dev) if vm is None else vm[off[b]:off[b + 1]].long()
                if idx.numel():
                    self._ha_sums_batch_bn(imgs[b], mu3[idx], mf3[idx], sums[b], accumulate)
                elif not accumulate:
                    sums[b].zero_()
            return sums
        h = self.handle(dev)
        L = _lib.lib()
        nbytes = L.spb_ha_workspace_bytes_ragged(B, V, int(max_views), H, W)
        ws = self._workspace(nbytes, dev)
        if sums is None:
            sums = torch.empty((B, 2, H, W), device=dev, dtype=torch.float32)
            accumulate = False
        assert sums.is_contiguous()
        with torch.cuda.device(dev):
            _lib.check(L.spb_ha_heatmap_sums_ragged(h, imgs.contiguous().data_ptr(), mu.data_ptr(), mf.data_ptr(),
                                                    vi.data_ptr(), None if vm is None else vm.data_ptr(), vo.data_ptr(),
                                                    B, V, int(max_views), H, W, sums.data_ptr(),
                                                    int(bool(accumulate)), ws.data_ptr(), ws.numel(),
                                                    torch.cuda.current_stream().cuda_stream), "ha_heatmap_sums_ragged")
        return sums

    def set_ha_overlap(self, on: bool, device=None) -> None:
        """Overlap the image warp / aggregate of neighbouring passes with the detector net (default off)."""
        self._ha_overlap = bool(on)
        if self._handle is not None or device is not None:
            dev = _norm_device(device if device is not None else self._handle_dev)
            _lib.check(_lib.lib().spb_net_set_ha_overlap(self.handle(dev), int(self._ha_overlap)), "set_ha_overlap")

    def ha_heatmap_sums_batch(self, imgs, mats_unwarp, mats_fwd, sums=None, accumulate=False, chunk: int = 0):
        """imgs [B,H,W] cuda fp32; mats [N,3,3] shared by the batch or [B,N,3,3] -> sums [B,2,H,W].
        All B*N views go through the kernels in passes of at most ``chunk`` views (0 = library default)."""
        dev = imgs.device
        B, H, W = imgs.shape
        shared = mats_unwarp.dim() == 3
        mu = mats_unwarp.to(dev, torch.float32).contiguous()
        mf = mats_fwd.to(dev, torch.float32).contiguous()
        N = mu.shape[-3]
        if not shared and mu.shape[0] != B:
            raise ValueError("per-image homographies must be [B,N,3,3]")
        if self.bn_mode == "batch":  # per image, all N views at once (chunk is ignored: the statistics need them all)
            if sums is None:
                sums = torch.empty((B, 2, H, W), device=dev, dtype=torch.float32)
                accumulate = False
            for b in range(B):
                self._ha_sums_batch_bn(imgs[b], mu if shared else mu[b], mf if shared else mf[b], sums[b], accumulate)
            return sums
        h = self.handle(dev)
        L = _lib.lib()
        _lib.check(L.spb_net_set_ha_chunk(h, int(chunk)), "set_ha_chunk")
        nbytes = L.spb_ha_workspace_bytes_batch(h, B, N, H, W)
        ws = self._workspace(nbytes, dev)
        if sums is None:
            sums = torch.empty((B, 2, H, W), device=dev, dtype=torch.float32)
            accumulate = False
        assert sums.is_contiguous()
        with torch.cuda.device(dev):
            _lib.check(L.spb_ha_heatmap_sums_batch(h, imgs.contiguous().data_ptr(), mu.data_ptr(), mf.data_ptr(),
                                                   int(shared), B, N, H, W, sums.data_ptr(), int(bool(accumulate)),
                                                   ws.data_ptr(), ws.numel(),
                                                   torch.cuda.current_stream().cuda_stream), "ha_heatmap_sums_batch")
        return sums


class SuperPointNet_pretrained_b200(SuperPointNet_b200):
    """Drop-in for ``models/SuperPointNet_pretrained.py:10-76`` (MagicLeap ``superpoint_v1.pth``): the same network,
    ``forward`` returns the TUPLE ``(semi, desc)`` of that class instead of the dict of the other two layouts."""

    def forward(self, x):
        out = super().forward(x)
        return out["semi"], out["desc"]
