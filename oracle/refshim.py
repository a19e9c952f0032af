"""Import prelude for the UNMODIFIED reference at /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` (fixture generation) and by
the ``-m "not gpu"`` tests that cross-check the oracle live when ``/root/reference`` is mounted.
Nothing here runs on the GPU box (the reference does not travel); no reference file is edited.

Harness-side shims:
  * stub ``imageio`` / ``tensorboardX`` (export.py:18,20 -- I/O and logging only);
  * ``collections.Mapping`` alias (utils/tools.py:18);
  * ``stable_argsort()``: the reference sorts with numpy's default (unstable) argsort
    (models/model_wrap.py:146,177,271), so ties are implementation-defined; the parity
    contract is the reference *with a stable sort*.
"""
from __future__ import annotations

import collections
import collections.abc
import contextlib
import os
import sys
import types

REF = os.environ.get("SPB200_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "export.py"))


def load():
    """Returns a namespace with the reference callables on the hot path."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REF}")
    sys.dont_write_bytecode = True  # /root/reference is read-only
    if REF not in sys.path:
        sys.path.insert(0, REF)
    collections.Mapping = collections.abc.Mapping
    for name in ("imageio", "tensorboardX"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.imread = lambda *a, **k: None
            m.SummaryWriter = object
            sys.modules[name] = m
    ns = types.SimpleNamespace()
    import export as ref_export
    from models.model_wrap import PointTracker, SuperPointFrontend_torch
    from models.SuperPointNet import SuperPointNet as BNFlat
    from models.SuperPointNet_gauss2 import SuperPointNet_gauss2
    from models.SuperPointNet_pretrained import SuperPointNet as V1
    from utils.homographies import sample_homography_np
    from utils.utils import compute_valid_mask, flattenDetection, inv_warp_image_batch

    ns.export = ref_export
    ns.combine_heatmap = ref_export.combine_heatmap
    ns.SuperPointFrontend_torch = SuperPointFrontend_torch
    ns.PointTracker = PointTracker
    ns.nets = {"v1": V1, "bnflat": BNFlat, "gauss2": SuperPointNet_gauss2}
    ns.sample_homography_np = sample_homography_np
    ns.compute_valid_mask = compute_valid_mask
    ns.flattenDetection = flattenDetection
    ns.inv_warp_image_batch = inv_warp_image_batch

    def frontend(conf_thresh=0.015, nms_dist=4, nn_thresh=0.7, subpixel=False):
        return SuperPointFrontend_torch({"model": {"subpixel": {"enable": subpixel}}}, "", nms_dist,
                                        conf_thresh, nn_thresh, load=False, device="cpu")

    ns.frontend = frontend
    return ns


@contextlib.contextmanager
def stable_argsort():
    import numpy as np
    orig = np.argsort

    def stable(a, *args, **kw):
        kw["kind"] = "stable"
        return orig(a, *args, **kw)

    np.argsort = stable
    try:
        yield
    finally:
        np.argsort = orig
