"""Free-running integration (SURVEY.md 8c tier ii): the restated per-frame data flow (tests/mini_vo.py, validated
against the recorded reference run by tests/test_oracle_free_running.py) driven by the CUDA path only --
findEssentialMat, recoverPose, triangulation, KLT x2, solvePnPRansac, goodFeaturesToTrack and the candidate
distance filter feed each other for five frames; the result is compared with the run of the unmodified
reference class on cv2."""
from types import SimpleNamespace

import numpy as np
import pytest

from mini_vo import run_on_trace
from monocular_visual_odometry_va4mr_b200 import cv2_compat, hotpath

pytestmark = pytest.mark.gpu


def test_free_running_pipeline_on_cuda_path():
    def klt(prev, nxt, pts, win, ml, crit):
        p, st, _ = cv2_compat.calcOpticalFlowPyrLK(prev, nxt, pts, None, winSize=win, maxLevel=ml, criteria=crit)
        return p, st

    ops = SimpleNamespace(
        klt=klt,
        gftt=lambda img, mc, q, md, bs: cv2_compat.goodFeaturesToTrack(img, maxCorners=mc, qualityLevel=q, minDistance=md, blockSize=bs,
                                                                        useHarrisDetector=False, mask=None),
        findEssentialMat=lambda p1, p2, K, prob, thr: cv2_compat.findEssentialMat(p1, p2, K, method=cv2_compat.RANSAC, prob=prob, threshold=thr),
        recoverPose=lambda E, p1, p2, K: cv2_compat.recoverPose(E, p1, p2, K),
        solvePnPRansac=lambda obj, img, K, it, err, conf: cv2_compat.solvePnPRansac(obj, img, K, np.zeros(4), flags=cv2_compat.SOLVEPNP_P3P,
                                                                                    confidence=conf, reprojectionError=err, iterationsCount=it),
        triangulate=lambda K, o, first, keys, tr, transforms, R, t: hotpath.triangulate_landmarks(K, o, first, keys, tr, transforms, R, t),
        min_distance=hotpath.min_distance_mask)
    g, vo = run_on_trace(ops)
    ref = [int(v) for v in g["num_pts"]]
    n = sum(1 for k in g.files if k.startswith("tri") and k.endswith("_cur"))
    assert len(vo.num_pts) == len(ref) and len(vo.poses) == n
    # bootstrap: identical inlier count (masks are bit-exact); afterwards the tracker's <= 0.02 px differences on a few
    # percent of the points may move single points in or out of an inlier set
    assert vo.num_pts[0] == ref[0]
    for a, b in zip(vo.num_pts[1:], ref[1:]):
        assert abs(a - b) <= max(3, 0.03 * b), (vo.num_pts, ref)
    for i in range(n):
        R, t = vo.poses[i][:9].reshape(3, 3), vo.poses[i][9:]
        Rr, tr_ = g[f"tri{i}_cur"][:9].reshape(3, 3), g[f"tri{i}_cur"][9:]
        assert np.abs(R - Rr).max() < 2e-3 and np.abs(t - tr_).max() < 2e-2 * max(1.0, np.abs(tr_).max()), (i, np.abs(R - Rr).max(), np.abs(t - tr_).max())
    print("free-running num_pts", vo.num_pts, "reference", ref,
          "max pose dev", max(np.abs(vo.poses[i] - g[f"tri{i}_cur"]).max() for i in range(n)))
