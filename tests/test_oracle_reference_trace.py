"""The oracle against the recorded hot-path calls of the unmodified reference class (cv2 4.13 outputs)."""
from types import SimpleNamespace

import numpy as np

import oracle
import reference_trace


def _klt(prev, nxt, pts, win, ml, crit):
    p, st, _ = oracle.calc_optical_flow_pyr_lk(prev, nxt, pts, win, ml, crit)
    return p, st


def _knn(q, t):
    idx, dist, _ = oracle.knn2_ratio(q, t, 0.8)
    return idx, dist


def _emat(p1, p2, K, prob, thr):
    E, mask, _ = oracle.find_essential_mat(p1, p2, K, prob, thr, 1000)
    return E, mask


def _pnp(obj, img, K, iters, err, conf):
    ok, rv, tv, inl, _ = oracle.solve_pnp_ransac_p3p(obj, img, K, iters, err, conf)
    return ok, rv, tv, inl


def test_oracle_replays_reference_trace():
    impl = SimpleNamespace(klt=_klt, gftt=lambda img, mc, q, md, bs: oracle.good_features_to_track(img, mc, q, md, bs),
                           knn=_knn, emat=_emat, pnp=_pnp, tri=oracle.triangulate_landmarks, fadd=oracle.min_distance_mask,
                           rpose=oracle.recover_pose)
    seen = reference_trace.replay(impl)
    assert sum(seen.values()) >= 10
