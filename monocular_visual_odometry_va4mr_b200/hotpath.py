"""Host side of the two components next to the hot path -- numpy in the reference,
one or two kernel launches here.  Same argument meaning as the reference's own state arrays, so the
binding is a two-line replacement:

* ``min_distance_mask``      -> ``valid_dist`` of ``feature_adding``       (VisualOdometryPipeLine.py:258)
* ``triangulate_landmarks``  -> the candidate loop of ``triangulate_landmarks`` (:170-204)
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import c_f32p, c_f64p, c_i32p, c_u8p


def _p(a, t):
    return a.ctypes.data_as(t)


def _chk(ctx, rc, what):
    if rc == -2:   # B200VO_E_UNSUPPORTED
        raise NotImplementedError(ctx.last_error())
    if rc != 0:
        raise _lib.B200VOError(f"{what} failed ({rc}): {ctx.last_error()}")


def min_distance_mask(pts, potential_keys, min_dist, ctx: _lib.Context | None = None) -> np.ndarray:
    """bool (n,): corner i is farther than ``min_dist`` from every existing candidate (ref :258)."""
    ctx = ctx or _lib.default_context(0)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    ex = np.ascontiguousarray(potential_keys, np.float32).reshape(-1, 2)
    valid = np.zeros(len(pts), np.uint8)
    if len(pts):
        _chk(ctx, ctx.lib.b200vo_min_distance_mask(ctx.h, _p(pts, c_f32p), len(pts), _p(ex, c_f32p) if len(ex) else None, len(ex),
                                                   C.c_float(min_dist), _p(valid, c_u8p)), "b200vo_min_distance_mask")
    return valid.astype(bool)


def pack_poses(transforms) -> np.ndarray:
    """``self.transforms`` ([(R_CW (3,3), t_CW (3,1)), ...]) -> float64 (n,12) rows (R row-major | t)."""
    return np.ascontiguousarray([np.hstack([np.reshape(R, 9), np.reshape(t, 3)]) for R, t in transforms], np.float64).reshape(-1, 12)


def triangulate_landmarks(K, options, potential_first_keys, potential_keys, potential_transforms, transforms,
                          R_current_CW, t_current_CW, ctx: _lib.Context | None = None):
    """The candidate loop of the reference's ``triangulate_landmarks`` (:170-204).

    ``options`` needs ``min_dist_landmarks, max_dist_landmarks, min_baseline_angle, min_baseline_frames``
    (main.py:21-25); ``transforms`` is ``self.transforms`` (a list of (R_CW, t_CW)) or a packed (n,12) array.
    Returns ``(too_short_baseline bool (n,), new_landmarks float32 (k,3), new_keypoints float32 (k,2))``:
    the mask the reference passes to ``filter_potential`` (:206) and the rows it appends to
    ``matched_landmarks`` / ``matched_keypoints`` (:196-202)."""
    fk = np.ascontiguousarray(potential_first_keys, np.float32).reshape(-1, 2)
    k = np.ascontiguousarray(potential_keys, np.float32).reshape(-1, 2)
    fp = np.ascontiguousarray(np.asarray(potential_transforms).reshape(-1), np.int32)
    poses = transforms if isinstance(transforms, np.ndarray) else pack_poses(transforms)
    poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 12)
    cur = np.ascontiguousarray(np.hstack([np.reshape(R_current_CW, 9), np.reshape(t_current_CW, 3)]), np.float64)
    Kf = np.ascontiguousarray(K, np.float64).reshape(9)
    n = len(k)
    if len(fk) != n or len(fp) != n:
        raise ValueError("potential_first_keys, potential_keys and potential_transforms must have one row per candidate")
    ctx = ctx or _lib.default_context(0)
    keep = np.zeros(n, np.uint8)
    lm = np.zeros((max(n, 1), 3), np.float32)
    kp = np.zeros((max(n, 1), 2), np.float32)
    cnt = C.c_int(0)
    if n:
        _chk(ctx, ctx.lib.b200vo_triangulate_landmarks(
            ctx.h, _p(Kf, c_f64p), C.c_double(options['min_dist_landmarks']), C.c_double(options['max_dist_landmarks']),
            C.c_double(options['min_baseline_angle']), int(options['min_baseline_frames']), _p(fk, c_f32p), _p(k, c_f32p),
            _p(fp, c_i32p), n, _p(poses, c_f64p), len(poses), _p(cur, c_f64p), _p(keep, c_u8p), _p(lm, c_f32p), _p(kp, c_f32p),
            C.byref(cnt)), "b200vo_triangulate_landmarks")
    return keep.astype(bool), lm[:cnt.value].copy(), kp[:cnt.value].copy()


def find_essential_mat_samples(p1, p2, K, samples, prob=0.999, threshold=1.0, want_models=False,
                               ctx: _lib.Context | None = None):
    """``cv2.findEssentialMat(p1, p2, K, RANSAC, prob, threshold)`` (ref :308) on the CALLER's 5-subsets
    (``samples`` int32 (iters,5)): the same-hypothesis-set form north_star's mask criterion is stated for.
    Returns a dict: E (3,3)|None, mask (n,1) uint8|None, nmodels int32 (iters,), counts int32 (iters,10),
    models float64 (iters,10,3,3)|None, winner (sample*10+model, -1: none), iters_run."""
    ctx = ctx or _lib.default_context(0)
    p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
    Kf = np.ascontiguousarray(K, np.float64).reshape(9)
    smp = np.ascontiguousarray(samples, np.int32).reshape(-1, 5)
    n, iters = len(p1), len(smp)
    E = np.zeros(9)
    mask = np.zeros((n, 1), np.uint8)
    nmodels = np.zeros(iters, np.int32)
    counts = np.zeros((iters, 10), np.int32)
    models = np.zeros((iters, 10, 3, 3)) if want_models else None
    found, winner, run = C.c_int(0), C.c_int(-1), C.c_int(0)
    _chk(ctx, ctx.lib.b200vo_find_essential_mat_ransac_samples(
        ctx.h, _p(p1, c_f32p), _p(p2, c_f32p), n, _p(Kf, c_f64p), _p(smp, c_i32p), iters, float(prob), float(threshold),
        _p(E, c_f64p), _p(mask, c_u8p), C.byref(found), _p(nmodels, c_i32p), _p(counts, c_i32p),
        _p(models, c_f64p) if want_models else None, C.byref(winner), C.byref(run)), "b200vo_find_essential_mat_ransac_samples")
    return dict(E=E.reshape(3, 3) if found.value else None, mask=mask if found.value else None, nmodels=nmodels, counts=counts,
                models=models, winner=winner.value, iters_run=run.value)


def batch_min_distance_mask(pts, n, potential_keys, m, min_dist, ctx: _lib.Context | None = None) -> np.ndarray:
    """``min_distance_mask`` for every sequence of a batch in one launch: ``pts`` float32 (batch, n_cap, 2) with ``n``
    (batch,) live rows, ``potential_keys`` float32 (batch, m_cap, 2) with ``m`` (batch,) live rows.
    -> bool (batch, n_cap); rows beyond n[s] are False."""
    ctx = ctx or _lib.default_context(0)
    pts = np.ascontiguousarray(pts, np.float32)
    ex = np.ascontiguousarray(potential_keys, np.float32)
    batch, n_cap = pts.shape[0], pts.shape[1]
    m_cap = ex.shape[1] if ex.ndim == 3 else 0
    n = np.ascontiguousarray(n, np.int32).reshape(batch)
    m = np.ascontiguousarray(m, np.int32).reshape(batch)
    valid = np.zeros((batch, n_cap), np.uint8)
    _chk(ctx, ctx.lib.b200vo_batch_min_distance_mask(ctx.h, batch, _p(pts, c_f32p), _p(n, c_i32p), n_cap,
                                                     _p(ex, c_f32p) if m_cap else None, _p(m, c_i32p), m_cap, C.c_float(min_dist),
                                                     _p(valid, c_u8p)), "b200vo_batch_min_distance_mask")
    return valid.astype(bool)


def batch_triangulate_landmarks(K, options, potential_first_keys, potential_keys, potential_transforms, n, poses, n_poses, cur_poses,
                                ctx: _lib.Context | None = None):
    """``triangulate_landmarks`` for every sequence of a batch (two launches).  Arrays carry a leading batch axis and a fixed
    capacity: first_keys / keys float32 (batch, cap, 2), potential_transforms int (batch, cap), ``n`` (batch,) live candidates;
    ``poses`` float64 (batch, pose_cap, 12) (see ``pack_poses``) with ``n_poses`` (batch,), ``cur_poses`` float64 (batch, 12).
    -> (too_short_baseline bool (batch, cap), new_landmarks float32 (batch, cap, 3), new_keypoints float32 (batch, cap, 2),
    n_new int32 (batch,)); the new rows of sequence s are [:n_new[s]]."""
    ctx = ctx or _lib.default_context(0)
    fk = np.ascontiguousarray(potential_first_keys, np.float32)
    k = np.ascontiguousarray(potential_keys, np.float32)
    batch, cap = k.shape[0], k.shape[1]
    fp = np.ascontiguousarray(potential_transforms, np.int32).reshape(batch, cap)
    n = np.ascontiguousarray(n, np.int32).reshape(batch)
    poses = np.ascontiguousarray(poses, np.float64)
    pose_cap = poses.shape[1]
    n_poses = np.ascontiguousarray(n_poses, np.int32).reshape(batch)
    cur = np.ascontiguousarray(cur_poses, np.float64).reshape(batch, 12)
    Kf = np.ascontiguousarray(K, np.float64).reshape(9)
    keep = np.zeros((batch, cap), np.uint8)
    lm = np.zeros((batch, cap, 3), np.float32)
    kp = np.zeros((batch, cap, 2), np.float32)
    n_new = np.zeros(batch, np.int32)
    _chk(ctx, ctx.lib.b200vo_batch_triangulate_landmarks(
        ctx.h, _p(Kf, c_f64p), C.c_double(options['min_dist_landmarks']), C.c_double(options['max_dist_landmarks']),
        C.c_double(options['min_baseline_angle']), int(options['min_baseline_frames']), batch, cap, _p(fk, c_f32p), _p(k, c_f32p),
        _p(fp, c_i32p), _p(n, c_i32p), _p(poses, c_f64p), _p(n_poses, c_i32p), pose_cap, _p(cur, c_f64p), _p(keep, c_u8p),
        _p(lm, c_f32p), _p(kp, c_f32p), _p(n_new, c_i32p)), "b200vo_batch_triangulate_landmarks")
    return keep.astype(bool), lm, kp, n_new
