"""ctypes binding of libspb200.so (the C ABI declared in include/spb200.h).

There is no CPU fallback: ``lib()`` raises if the library is missing or cannot be loaded, and
``check()`` turns every non-zero return code into a ``RuntimeError`` carrying
``spb_last_error_string()``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPB200_LIB") or os.path.join(_HERE, "libspb200.so")  # SPB200_LIB: A/B builds of the kernels

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_z = C.c_size_t
_ll = C.c_longlong

# name -> (restype, argtypes); mirrors include/spb200.h one to one
PROTOTYPES = {
    "spb_version": (_i, []),
    "spb_last_error_string": (C.c_char_p, []),
    "spb_launch_count": (C.c_ulonglong, []),
    "spb_inv_warp_image_batch": (_i, [_p, _ll, _p, _p, _i, _i, _i, _i, _p]),
    "spb_compute_valid_mask": (_i, [_p, _p, _i, _i, _i, _p]),
    "spb_erode": (_i, [_p, _p, _i, _i, _i, _p, _i, _i, _i, _p]),
    "spb_flatten_detection": (_i, [_p, _i, _p, _i, _i, _i, _p]),
    "spb_combine_heatmap": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "spb_ha_aggregate_workspace_bytes": (_z, [_i, _i, _i]),
    "spb_ha_aggregate": (_i, [_p, _i, _p, _p, _i, _i, _i, _p, _i, _p, _z, _p]),
    "spb_ha_finalize": (_i, [_p, _p, _i, _i, _p]),
    "spb_ha_finalize_batch": (_i, [_p, _p, _i, _i, _i, _p]),
    "spb_peer_alloc": (_i, [_p, _z]),
    "spb_peer_free": (_i, [_p]),
    "spb_peer_export": (_i, [_p, _p]),
    "spb_peer_open": (_i, [_p, _p]),
    "spb_peer_close": (_i, [_p]),
    "spb_peer_signal": (_i, [_p, _i, C.c_uint, _p]),
    "spb_peer_wait": (_i, [_p, _i, C.c_uint, C.c_uint, _p]),
    "spb_ha_finalize_peers": (_i, [_p, _i, _i, _i, _p, _i, _i, _p]),
    "spb_nms_workspace_bytes": (_z, [_i, _i, _i, _i]),
    "spb_get_pts_from_heatmap": (_i, [_p, _i, _i, _i, _f, _i, _i, _i, _i, _p, _p, _p, _p, _z, _p]),
    "spb_get_pts_from_heatmap_ex": (_i, [_p, _i, _i, _i, _f, _i, _i, _i, _i, _p, _p, _p, _p, _z, _i, _p]),
    "spb_sample_desc_from_points": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "spb_sample_desc_from_points_ex": (_i, [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "spb_roi_soft_argmax": (_i, [_p, _i, _i, _i, _p, _i, _p, _p, _p]),
    "spb_gather_dense_desc": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "spb_dense_desc": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "spb_soft_argmax_points": (_i, [_p, _i, _i, _p, _p, _i, _i, _p, _p]),
    "spb_soft_argmax_points_batch": (_i, [_p, _i, _i, _i, _p, _p, _i, _i, _p, _p]),
    "spb_sample_desc_batch": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "spb_nn_match_pairs_workspace_bytes": (_z, [_i, _i]),
    "spb_nn_match_pairs": (_i, [_p, _p, _i, _i, _i, _f, _p, _p, _p, _p, _p, _z, _p]),
    "spb_nn_match_workspace_bytes": (_z, [_i, _i]),
    "spb_nn_match_two_way": (_i, [_p, _i, _p, _i, _i, _f, _p, _p, _p, _p, _p, _z, _p]),
    "spb_sample_ha_homographies": (_i, [C.c_ulonglong, _i, _i, _i, _i, _i, _i, _i, _i, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.c_double, _i, C.c_double, _p, _p, _p, _p]),
    "spb_ingest_resize_gray": (_i, [_p, _i, _i, _p, _i, _i, _p]),
    "spb_net_create": (_i, [C.POINTER(_p), C.POINTER(_p), C.POINTER(_p)]),
    "spb_net_destroy": (_i, [_p]),
    "spb_net_set_impl": (_i, [_p, _i]),
    "spb_net_set_ha_chunk": (_i, [_p, _i]),
    "spb_net_set_ha_overlap": (_i, [_p, _i]),
    "spb_net_profile": (_i, [_p, _i]),
    "spb_net_profile_read": (_i, [_p, _p, _p]),
    "spb_net_workspace_bytes": (_z, [_i, _i, _i, _i]),
    "spb_net_forward": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _z, _p]),
    "spb_net_layer": (_i, [_p, _i, _p, _p, _i, _i, _i, _p]),
    "spb_conv_f32": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "spb_bn_batch_stats_workspace_bytes": (_z, [_i]),
    "spb_bn_batch_stats": (_i, [_p, _ll, _i, _p, _p, _f, _p, _p, _p, _p, _p, _z, _p]),
    "spb_bn_apply": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _i, _p, _i, _i, _p]),
    "spb_ha_workspace_bytes": (_z, [_i, _i, _i]),
    "spb_ha_heatmap_sums": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _i, _p, _z, _p]),
    "spb_ha_workspace_bytes_batch": (_z, [_p, _i, _i, _i, _i]),
    "spb_ha_workspace_bytes_ragged": (_z, [_i, _i, _i, _i, _i]),
    "spb_ha_pass_views": (_i, [_p, _i, _i]),
    "spb_ha_warp_views_workspace_bytes": (_z, [_i, _i, _i, _i]),
    "spb_ha_warp_views": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _z, _p]),
    "spb_ha_heatmap_sums_ragged": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _i, _p, _z, _p]),
    "spb_ha_heatmap_sums_batch": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _i, _p, _z, _p]),
}
# debug knobs exported by the library but not part of the public header
_EXTRA = {"spb_conv_tc_set_halo": (_i, [_i]), "spb_conv_tc_set_fuse1": (_i, [_i]), "spb_conv_tc_set_s3": (_i, [_i]), "spb_conv_tc_set_pair": (_i, [_i]), "spb_conv_tc_set_cg2": (_i, [_i]),
          "spb_conv1_fused_set_probe": (_i, [_p])}

IMPL_TCGEN05, IMPL_SIMT = 0, 1
MODE_BILINEAR, MODE_NEAREST = 0, 1

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m pytorch_superpoint_b200.build` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in {**PROTOTYPES, **_EXTRA}.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().spb_last_error_string().decode(errors="replace")
        raise RuntimeError(f"libspb200 {what} failed (rc={rc}): {msg}")
