"""``PointTracker_b200`` -- mirror of ``models/model_wrap.py:411-583 PointTracker``.

``export_descriptor`` (export.py:118,147-151) feeds two frames through ``update`` and stores ``get_matches()``:
rows ``(x1, y1, x2, y2)`` of the mutual nearest neighbours.  The descriptor contraction and the two arg-mins run on
the GPU (``ops.nn_match_two_way`` -> match.cu); the track bookkeeping is a few hundred integers per frame and stays on
the host, restated here with whole-array numpy operations instead of the reference's per-match Python loop.  Same
attributes (``tracks`` rows ``[track id, average score, point id_0 .. id_L-1]``, ``all_pts``, ``last_desc``,
``matches``, ``mscores``), same return conventions and error behaviour.
"""
from __future__ import annotations

import numpy as np

from . import ops

__all__ = ["PointTracker_b200"]


class PointTracker_b200(object):
    def __init__(self, max_length=2, nn_thresh=0.7):
        if max_length < 2:
            raise ValueError("max_length must be greater than or equal to 2.")
        self.maxl = max_length
        self.nn_thresh = nn_thresh
        self.all_pts = [np.zeros((2, 0)) for _ in range(self.maxl)]
        self.last_desc = None
        self.tracks = np.zeros((0, self.maxl + 2))
        self.track_count = 0
        self.max_score = 9999
        self.matches = None
        self.last_pts = None
        self.mscores = None

    # ---- models/model_wrap.py:437-480 ---------------------------------------------------------------------
    def nn_match_two_way(self, desc1, desc2, nn_thresh):
        """desc1 [D,K1], desc2 [D,K2] -> float64 [3,L] rows (index in desc1, index in desc2, L2 distance)."""
        matches = ops.nn_match_two_way(desc1, desc2, nn_thresh)
        self.mscores = matches
        return matches

    # ---- models/model_wrap.py:482-505 ---------------------------------------------------------------------
    def get_offsets(self):
        """Global id of the first point of every stored frame (the last frame's size is not needed)."""
        sizes = [0] + [p.shape[1] for p in self.all_pts[:-1]]
        return np.cumsum(np.array(sizes))

    def get_matches(self):
        return self.matches

    def get_mscores(self):
        return self.mscores

    def clear_desc(self):
        self.last_desc = None

    # ---- models/model_wrap.py:507-583 ---------------------------------------------------------------------
    def update(self, pts, desc):
        """pts [3,N] (x, y, conf), desc [D,N]: match against the previous frame and extend / open tracks."""
        if pts is None or desc is None:
            print("PointTracker: Warning, no points were added to tracker.")
            return
        pts, desc = np.asarray(pts), np.asarray(desc)
        assert pts.shape[1] == desc.shape[1]
        if self.last_desc is None:
            self.last_desc = np.zeros((desc.shape[0], 0))
        # the oldest frame leaves: its points, its track column, and its share of every global point id
        gone = self.all_pts[0].shape[1]
        self.all_pts = self.all_pts[1:] + [pts]
        t = np.delete(self.tracks, 2, axis=1)
        t[:, 2:] -= gone
        ids = t[:, 2:]
        ids[ids < -1] = -1
        offsets = self.get_offsets()
        t = np.hstack((t, -np.ones((t.shape[0], 1))))
        matches = self.nn_match_two_way(self.last_desc, desc, self.nn_thresh)
        self.matches = matches
        if self.last_pts is not None:  # export.py:147-151 reads these (x1, y1, x2, y2) columns
            i1, i2 = matches[0].astype(int), matches[1].astype(int)
            self.matches = np.concatenate((self.last_pts[:, i1], pts[:2, i2]), axis=0)
        # extend the tracks whose head is a matched point of the previous frame
        matched = np.zeros(pts.shape[1], dtype=bool)
        if matches.shape[1] and t.shape[0]:
            id1 = matches[0].astype(int) + offsets[-2]
            id2 = matches[1].astype(int) + offsets[-1]
            heads = t[:, -2]
            order = np.argsort(heads, kind="stable")
            pos = np.searchsorted(heads[order], id1)
            pos[pos >= len(order)] = len(order) - 1
            hit = heads[order][pos] == id1
            rows = order[pos[hit]]
            matched[matches[1, hit].astype(int)] = True
            t[rows, -1] = id2[hit]
            score = matches[2, hit]
            fresh = t[rows, 1] == self.max_score
            # running average over the track's links (may include links that already left the window, as upstream)
            links = (t[rows, 2:] != -1).sum(axis=1) - 1.0
            frac = 1.0 / links
            t[rows, 1] = np.where(fresh, score, (1.0 - frac) * t[rows, 1] + frac * score)
        # open a track for every unmatched point
        new_ids = (np.arange(pts.shape[1]) + offsets[-1])[~matched]
        fresh_tracks = -np.ones((new_ids.shape[0], self.maxl + 2))
        fresh_tracks[:, -1] = new_ids
        fresh_tracks[:, 0] = self.track_count + np.arange(new_ids.shape[0])
        fresh_tracks[:, 1] = self.max_score
        t = np.vstack((t, fresh_tracks))
        self.track_count += new_ids.shape[0]
        self.tracks = t[np.any(t[:, 2:] >= 0, axis=1), :]
        self.last_desc = desc.copy()
        self.last_pts = pts[:2, :].copy()
        return

    # ---- models/model_wrap.py:585-605 ---------------------------------------------------------------------
    def get_tracks(self, min_length):
        """Tracks with at least ``min_length`` observations that reach the most recent frame."""
        if min_length < 1:
            raise ValueError("'min_length' too small.")
        long_enough = np.sum(self.tracks[:, 2:] != -1, axis=1) >= min_length
        has_head = self.tracks[:, -1] != -1
        return self.tracks[long_enough & has_head, :].copy()
