"""The CUDA path (through the cv2-shaped shim over the C ABI) against the recorded hot-path calls of the
unmodified reference class: same inputs the reference passed, compared with what cv2 returned to it."""
from types import SimpleNamespace

import numpy as np
import pytest

import reference_trace
from monocular_visual_odometry_va4mr_b200 import cv2_compat, hotpath

pytestmark = pytest.mark.gpu


def _klt(prev, nxt, pts, win, ml, crit):
    p, st, _ = cv2_compat.calcOpticalFlowPyrLK(prev, nxt, pts, None, winSize=win, maxLevel=ml, criteria=crit)   # :281 / :287
    return p, st


def _gftt(img, mc, q, md, bs):
    return cv2_compat.goodFeaturesToTrack(img, maxCorners=mc, qualityLevel=q, minDistance=md, blockSize=bs,
                                          useHarrisDetector=False, mask=None)                                  # :256


def _knn(q, t):
    rows = cv2_compat.BFMatcher().knnMatch(q, t, k=2)                                                           # :36, :229
    return (np.array([[m.trainIdx for m in r] for r in rows], np.int32),
            np.array([[m.distance for m in r] for r in rows], np.float32))


def _emat(p1, p2, K, prob, thr):
    return cv2_compat.findEssentialMat(p1, p2, K, method=cv2_compat.RANSAC, prob=prob, threshold=thr)           # :308


def _pnp(obj, img, K, iters, err, conf):
    return cv2_compat.solvePnPRansac(obj, img, K, np.zeros(4), flags=cv2_compat.SOLVEPNP_P3P, confidence=conf,
                                     reprojectionError=err, iterationsCount=iters)                              # :343


def _tri(K, cfg, first_keys, keys, first_pose, poses, cur):                                                      # :170-204
    opt = dict(min_dist_landmarks=cfg[0], max_dist_landmarks=cfg[1], min_baseline_angle=cfg[2], min_baseline_frames=int(cfg[3]))
    return hotpath.triangulate_landmarks(K, opt, first_keys, keys, first_pose, poses, cur[:9].reshape(3, 3), cur[9:])


def test_cuda_replays_reference_trace():
    seen = reference_trace.replay(SimpleNamespace(klt=_klt, gftt=_gftt, knn=_knn, emat=_emat, pnp=_pnp, tri=_tri,
                                                  fadd=hotpath.min_distance_mask,                              # :258
                                                  rpose=cv2_compat.recoverPose))                               # :315
    assert sum(seen.values()) >= 10
