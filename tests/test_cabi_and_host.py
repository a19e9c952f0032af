"""CPU-side tests: the C-ABI library loads and exports every symbol include/spb200.h declares (no compute
calls), the ctypes prototypes cover the header, and the host logic (state-dict folding, sharding)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "spb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from pytorch_superpoint_b200.build import build
    lib = ctypes.CDLL(build())
    syms = _header_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), f"libspb200.so does not export {s}"
    lib.spb_version.restype = ctypes.c_int
    assert lib.spb_version() >= 100
    lib.spb_nms_workspace_bytes.restype = ctypes.c_size_t
    assert lib.spb_nms_workspace_bytes(1, 240, 320, 4) >= 48 * 64 * 8  # pure host arithmetic


def test_ctypes_prototypes_cover_header():
    from pytorch_superpoint_b200 import _lib
    assert set(_header_symbols()) == set(_lib.PROTOTYPES)


def test_ops_fail_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import pytorch_superpoint_b200 as pkg
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.inv_warp_image_batch(torch.zeros(1, 1, 8, 8), torch.eye(3)[None])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.SuperPointNet_b200()(torch.zeros(1, 1, 8, 8))


@pytest.mark.parametrize("layout", ["v1", "bnflat", "gauss2"])
def test_fold_state_dict_matches_oracle(layout):
    from oracle import oracle as O
    from pytorch_superpoint_b200 import fold_state_dict
    sd = O.synthetic_state_dict(layout, seed=3)
    ws, bs = fold_state_dict(sd)
    ref = O.fold_bn(O.canonical_params(sd))
    for (name, *_), w, b in zip(O.LAYERS, ws, bs):
        np.testing.assert_allclose(w, ref[name]["w"].numpy(), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(b, ref[name]["b"].numpy(), rtol=1e-6, atol=1e-7)
    sd2 = {"module." + k: v for k, v in sd.items()}  # nn.DataParallel prefix
    ws2, _ = fold_state_dict(sd2)
    assert all(np.array_equal(a, b) for a, b in zip(ws, ws2))
    with pytest.raises(KeyError):
        fold_state_dict({"foo.weight": torch.zeros(1)})


def test_net_state_dict_roundtrip():
    from oracle import oracle as O
    from pytorch_superpoint_b200 import SuperPointNet_b200
    n = SuperPointNet_b200()
    assert len(n.state_dict()) == 84  # gauss2 layout by default, like the reference's default model
    sd = O.synthetic_state_dict("v1", seed=0)
    n.load_state_dict(sd)
    assert set(n.state_dict()) == set(sd)


def test_shard_bounds():
    from pytorch_superpoint_b200.sharding import shard_bounds
    for n, w in [(100, 8), (100, 1), (20, 8), (5, 8), (100, 3)]:
        spans = [shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert [b - a for a, b in (shard_bounds(100, r, 8) for r in range(8))] == [13] * 4 + [12] * 4
    with pytest.raises(ValueError):
        shard_bounds(10, 8, 8)


def test_synthetic_weights_match_oracle_generator():
    from oracle import oracle as O
    from pytorch_superpoint_b200.synthetic import synthetic_state_dict
    for layout in ("v1", "bnflat", "gauss2"):
        a, b = synthetic_state_dict(layout, 5), O.synthetic_state_dict(layout, 5)
        assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)


def test_cabi_argument_errors_without_a_gpu():
    """Argument validation happens before any CUDA call: the error convention of the C ABI (negative code +
    spb_last_error_string, never a throw) can be checked on the CPU box for the entry points added for the f rows."""
    from pytorch_superpoint_b200 import _lib
    L = _lib.lib()
    buf = (ctypes.c_float * 64)()
    ibuf = (ctypes.c_int * 8)()
    p = lambda a: ctypes.cast(a, ctypes.c_void_p)  # noqa: E731
    # models/model_wrap.py:455-456: a negative threshold is an error (ValueError in the reference)
    rc = L.spb_nn_match_two_way(p(buf), 2, p(buf), 2, 4, -0.5, p(ibuf), p(ibuf), p(buf), p(ibuf), p(buf), 1024, None)
    assert rc < 0 and b"nn_thresh" in L.spb_last_error_string()
    assert L.spb_nn_match_workspace_bytes(10, 20) >= 30 * 8
    # sampler: n_scales outside 1..16, patch_ratio outside (0, 1]
    args = [0, 4, 2, 1, 1, 1, 1]
    rc = L.spb_sample_ha_homographies(*args, 0, 25, 0.2, 0.2, 0.2, 0.85, 1.57, 1, 0.0, p(buf), p(buf), None, None)
    assert rc < 0 and b"n_scales" in L.spb_last_error_string()
    rc = L.spb_sample_ha_homographies(*args, 5, 25, 0.2, 0.2, 0.2, 1.5, 1.57, 1, 0.0, p(buf), p(buf), None, None)
    assert rc < 0 and b"patch_ratio" in L.spb_last_error_string()
    # dense descriptors: D must be a multiple of 8 up to 256
    assert L.spb_dense_desc(p(buf), 1, 2, 2, 7, 8, p(buf), None) < 0
    assert L.spb_roi_soft_argmax(None, 1, 8, 8, None, 3, p(buf), None, None) < 0
    assert L.spb_sample_desc_from_points_ex(p(buf), 1, p(buf), 1, None, 4, 4, 4, 8, 8, 0, p(buf), None) < 0  # stride < 2
    # ragged HA entry: a NULL network handle
    assert L.spb_ha_heatmap_sums_ragged(None, p(buf), p(buf), p(buf), p(ibuf), p(ibuf), p(ibuf), 1, 1, 1, 8, 8, p(buf), 0,
                                        p(buf), 64, None) < 0
    assert L.spb_ha_workspace_bytes_ragged(0, 1, 1, 8, 8) == 0
    # pass size: bounded by bytes (800 views of 240x320), at least one view
    assert L.spb_ha_pass_views(None, 240, 320) == 800 and L.spb_ha_pass_views(None, 480, 640) == 200
    assert L.spb_ha_pass_views(None, 384, 1248) == 128 and L.spb_ha_pass_views(None, 4096, 4096) == 3
    # fused-pass warp on its own: W must be a multiple of 8, workspace checked
    assert L.spb_ha_warp_views(p(buf), p(buf), p(ibuf), None, 1, 1, 8, 12, p(buf), p(buf), p(buf), 1 << 20, None) < 0
    assert L.spb_ha_warp_views(p(buf), p(buf), p(ibuf), None, 1, 1, 8, 8, p(buf), p(buf), p(buf), 16, None) == -3
    assert L.spb_ha_warp_views_workspace_bytes(1, 0, 8, 8) == 0
    # pass-overlap switch and pass size: a NULL network handle
    assert L.spb_net_set_ha_overlap(None, 1) < 0
    assert L.spb_net_set_ha_chunk(None, 100) < 0
    with pytest.raises(RuntimeError, match="libspb200"):
        _lib.check(-1, "demo")
