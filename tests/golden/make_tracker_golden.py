"""Golden vectors for PointTracker (models/model_wrap.py:411-605) -- run in the build container, where the reference
is mounted at /root/reference:

    python tests/golden/make_tracker_golden.py

1. HPatches replay: the 7 prediction files the reference ships under
   logs/superpoint_coco_heat2_0_170k_nms4_det0.015/predictions/{0..6}.npz hold key-points, descriptors and the matches
   its own export_descriptor wrote (nn_thresh 1.0).  First the full files are replayed through the UNMODIFIED
   PointTracker.update and must reproduce the stored matches bit for bit (asserted here).  The fixture then keeps the
   220 most confident points of every image with fp16-rounded descriptors (1.6 MB instead of 6.3 MB) together with
   the matches the unmodified PointTracker produces on exactly those inputs.
2. A 5-frame sequence (max_length 3, nn_thresh 0.9) of synthetic unit descriptors with partial overlap between
   consecutive frames: tracks / matches / offsets of the unmodified tracker after every update.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

ref = refshim.load()
PRED = os.path.join(refshim.REF, "logs", "superpoint_coco_heat2_0_170k_nms4_det0.015", "predictions")
out = {}
KEEP = 220
for i in range(7):
    f = np.load(os.path.join(PRED, f"{i}.npz"))
    tr = ref.PointTracker(max_length=2, nn_thresh=1.0)
    tr.update(f["prob"].T, f["desc"].T)
    tr.update(f["warped_prob"].T, f["warped_desc"].T)
    assert np.array_equal(tr.get_matches().T, f["matches"]), f"file {i}: replay differs from the stored matches"
    p1, p2 = f["prob"][:KEEP].T.copy(), f["warped_prob"][:KEEP].T.copy()
    d1 = f["desc"][:KEEP].T.astype(np.float16)
    d2 = f["warped_desc"][:KEEP].T.astype(np.float16)
    tr = ref.PointTracker(max_length=2, nn_thresh=1.0)
    tr.update(p1, d1.astype(np.float32))
    tr.update(p2, d2.astype(np.float32))
    out.update({f"hp{i}_p1": p1, f"hp{i}_p2": p2, f"hp{i}_d1": d1, f"hp{i}_d2": d2,
                f"hp{i}_matches": tr.get_matches().copy(), f"hp{i}_mscores": tr.get_mscores().copy(),
                f"hp{i}_tracks": tr.tracks.copy()})
    print(f"file {i}: full replay exact ({f['matches'].shape[0]} matches); fixture keeps {tr.get_matches().shape[1]}")

rng = np.random.RandomState(7)
D = 64
base = rng.randn(D, 90).astype(np.float32)
tr = ref.PointTracker(max_length=3, nn_thresh=0.9)
for k in range(5):
    # frame k sees a sliding window of the base descriptors, perturbed, in random order, plus a few new ones
    idx = rng.permutation(np.arange(10 * k, 10 * k + 50))
    d = base[:, idx] + 0.15 * rng.randn(D, len(idx)).astype(np.float32)
    d = np.concatenate([d, rng.randn(D, 6).astype(np.float32)], axis=1)
    d /= np.linalg.norm(d, axis=0, keepdims=True)
    p = np.vstack([rng.uniform(4, 300, d.shape[1]), rng.uniform(4, 220, d.shape[1]), rng.uniform(0.02, 1, d.shape[1])])
    tr.update(p, d)
    out.update({f"seq{k}_pts": p, f"seq{k}_desc": d, f"seq{k}_tracks": tr.tracks.copy(),
                f"seq{k}_matches": tr.get_matches().copy(), f"seq{k}_offsets": tr.get_offsets().copy(),
                f"seq{k}_long": tr.get_tracks(2).copy()})
out["seq_count"] = np.array(tr.track_count)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tracker.npz"), **out)
print("wrote tests/golden/tracker.npz", os.path.getsize(os.path.join(ROOT, "tests", "golden", "tracker.npz")))
