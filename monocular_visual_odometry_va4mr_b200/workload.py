"""Synthetic per-frame KLT + PnP workloads shaped like the reference's datasets (numpy only).

A workload is what ``continuous_operation`` (reference ``VisualOdometryPipeLine.py:326-373``)
hands to its hot path on every frame: the previous and the new frame, the tracked landmark
keypoints with their 3-D landmarks, and the candidate keypoints.  Frames come from
``synth.render_sequence``; landmarks are back-projected from the rendered depth, with pixel noise
and a fraction of gross outliers, so that the PnP-RANSAC sees what it sees in the reference.
Sequences visit their frames forwards then backwards (0,1,..,F-1,F-2,..,1,0,1,..), so every step
is a small-baseline pair and the "new frame becomes the previous frame" recursion holds.
"""
from __future__ import annotations

import numpy as np

from . import synth

# the reference's per-dataset options that reach the hot path (main.py:20-44, 50-74, 80-104)
REFERENCE_OPTIONS = {
    "kitti": dict(win=(15, 15), max_level=5, criteria=(3, 50, 0.01), pnp_err=8.0, pnp_iters=500, pnp_conf=0.99,
                  min_dist=1.0, max_dist=150.0),
    "malaga": dict(win=(15, 15), max_level=10, criteria=(3, 50, 0.01), pnp_err=5.0, pnp_iters=500, pnp_conf=0.99,
                   min_dist=0.0, max_dist=100.0),
    "parking": dict(win=(15, 15), max_level=10, criteria=(3, 50, 0.02), pnp_err=5.0, pnp_iters=500, pnp_conf=0.99,
                    min_dist=1.0, max_dist=50.0),
}


def frame_at(n_frames: int, t: int) -> int:
    """Frame index visited at step t >= 0 (forwards, then backwards, ...): defined for EVERY t, so loops
    that look one or two steps ahead never run off the end of a precomputed list."""
    if n_frames <= 1:
        return 0
    period = 2 * n_frames - 2
    k = t % period
    return k if k < n_frames else period - k


def frame_order(n_frames: int, n_steps: int) -> list[int]:
    """Frame index visited at step t = 0..n_steps (forwards, then backwards, ...)."""
    return [frame_at(n_frames, t) for t in range(n_steps + 1)]


class TrackWorkload:
    """``batch`` sequences over ``n_distinct`` rendered scenes, ``n_frames`` frames each."""

    def __init__(self, shape="kitti", batch=64, n_frames=4, n_landmarks=1000, n_candidates=1000, n_distinct=2,
                 seed=0, outlier_frac=0.1, noise_px=0.3, width=None, height=None, cap_landmarks=None,
                 cap_candidates=None, first_index=None):
        """first_index: global index of this workload's first sequence.  When given, sequence g = first_index + s
        uses scene g % n_distinct and its own random stream, so the sequence is the same however the batch is
        sharded over ranks (contiguous blocks of sequences per GPU)."""
        K, w, h, step, _ = synth.SHAPES[shape]
        self.shape, self.K = shape, K.copy()
        self.w, self.h = width or w, height or h
        self.batch, self.F = batch, n_frames
        self.L = cap_landmarks or n_landmarks
        self.Cn = cap_candidates if cap_candidates is not None else n_candidates
        rng = np.random.default_rng(seed)
        g0 = first_index or 0
        used = sorted({(g0 + s) % n_distinct for s in range(batch)})
        from concurrent.futures import ThreadPoolExecutor
        import os
        with ThreadPoolExecutor(max_workers=min(len(used), os.cpu_count() or 1)) as pool:
            rendered = list(pool.map(lambda d: synth.render_sequence(shape, n_frames, seed=seed * 131 + d, width=self.w,
                                                                     height=self.h), used))
        seqs = dict(zip(used, rendered))
        self.seqs = seqs
        self.n_distinct, self.first_index = n_distinct, g0
        self.frames = np.empty((n_frames, batch, self.h, self.w), np.uint8)
        self.lm_pts = np.zeros((n_frames, batch, self.L, 2), np.float32)
        self.lm_obj = np.zeros((n_frames, batch, self.L, 3), np.float32)
        self.n_lm = np.zeros((n_frames, batch), np.int32)
        self.cand_pts = np.zeros((n_frames, batch, max(self.Cn, 1), 2), np.float32)
        self.n_cand = np.zeros((n_frames, batch), np.int32)
        pool = {}
        for d in used:
            for f in range(n_frames):
                pool[d, f] = synth.grid_corners(seqs[d]["frames"][f], 2 * (n_landmarks + n_candidates), seed=seed + 7 * d + f)
        for s in range(batch):
            d = (g0 + s) % n_distinct
            sq = seqs[d]
            if first_index is not None:
                rng = np.random.default_rng([seed, g0 + s])
            for f in range(n_frames):
                self.frames[f, s] = sq["frames"][f]
                pts = pool[d, f][rng.permutation(len(pool[d, f]))]
                nl = min(self.L, max(4, n_landmarks - ((g0 + s) % 5) * 7))          # ragged live counts
                nc = min(self.Cn, max(0, n_candidates - ((g0 + s) % 3) * 11)) if self.Cn else 0
                # the reference only keeps landmarks within [min_dist, max_dist] of the camera
                # (main.py:22-23 ...; VisualOdometryPipeLine.py:168): landmarks come from that depth band
                zi = sq["depth"][f][np.clip(np.rint(pts[:, 1]).astype(int), 0, self.h - 1),
                                    np.clip(np.rint(pts[:, 0]).astype(int), 0, self.w - 1)]
                near = (zi > REFERENCE_OPTIONS[shape]["min_dist"]) & (zi < 0.6 * REFERENCE_OPTIONS[shape]["max_dist"])
                pts = np.concatenate([pts[near], pts[~near]])
                nl = min(nl, int(near.sum()))
                lp = pts[:nl]
                # landmark = back-projection of the keypoint displaced by the triangulation noise
                X = synth.backproject(self.K, sq["R_cw"][f], sq["c"][f],
                                      lp.astype(np.float64) + rng.normal(0, noise_px, lp.shape), sq["depth"][f])
                n_out = int(outlier_frac * nl)
                if n_out:
                    oi = rng.choice(nl, n_out, replace=False)
                    X[oi] += rng.uniform(-4.0, 4.0, (n_out, 3))
                self.lm_pts[f, s, :nl] = lp
                self.lm_obj[f, s, :nl] = X.astype(np.float32)
                self.n_lm[f, s] = nl
                if nc:
                    self.cand_pts[f, s, :nc] = pts[nl:nl + nc]
                    self.n_cand[f, s] = nc

    def bytes_per_step(self):
        h2d = self.frames[0].nbytes + self.lm_pts[0].nbytes + self.lm_obj[0].nbytes + self.n_lm[0].nbytes
        if self.Cn:
            h2d += self.cand_pts[0].nbytes + self.n_cand[0].nbytes
        b, L, Cn = self.batch, self.L, self.Cn
        d2h = b * L * 8 + b * L + b * Cn * 8 + b * Cn + b * 48 + b + b * L + b * 4
        return h2d, d2h

    def true_pose(self, s: int, f: int):
        """World->camera (R, t) of frame f of sequence s (what solvePnPRansac should recover)."""
        sq = self.seqs[(self.first_index + s) % self.n_distinct]
        R_cw, c = sq["R_cw"][f], sq["c"][f]
        return R_cw.T, -R_cw.T @ c
