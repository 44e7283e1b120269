"""The restated per-frame data flow (tests/mini_vo.py) driven by REAL cv2 (+ the oracle for the two numpy
loops) must reproduce the recorded run of the unmodified reference class exactly -- this validates the
driver that tests/test_free_running_gpu.py then runs on the CUDA path."""
from types import SimpleNamespace

import numpy as np
import pytest

import oracle
from mini_vo import run_long, run_on_trace


def _cv2_ops():
    cv2 = pytest.importorskip("cv2")

    def klt(prev, nxt, pts, win, ml, crit):
        p, st, _ = cv2.calcOpticalFlowPyrLK(prev, nxt, pts, None, winSize=win, maxLevel=ml, criteria=crit)
        return p, st

    def tri(K, o, first, keys, tr, transforms, R, t):
        return oracle.triangulate_landmarks(K, (o['min_dist_landmarks'], o['max_dist_landmarks'], o['min_baseline_angle'], o['min_baseline_frames']),
                                            first, keys, tr, oracle.pack_poses(transforms), np.hstack([np.reshape(R, 9), np.reshape(t, 3)]))

    ops = SimpleNamespace(
        klt=klt,
        gftt=lambda img, mc, q, md, bs: cv2.goodFeaturesToTrack(img, maxCorners=mc, qualityLevel=q, minDistance=md, blockSize=bs,
                                                                 useHarrisDetector=False, mask=None),
        findEssentialMat=lambda p1, p2, K, prob, thr: cv2.findEssentialMat(p1, p2, K, method=cv2.RANSAC, prob=prob, threshold=thr),
        recoverPose=lambda E, p1, p2, K: cv2.recoverPose(E, p1, p2, K),
        solvePnPRansac=lambda obj, img, K, it, err, conf: cv2.solvePnPRansac(obj, img, K, np.zeros(4), flags=cv2.SOLVEPNP_P3P, confidence=conf,
                                                                             reprojectionError=err, iterationsCount=it),
        triangulate=tri, min_distance=oracle.min_distance_mask)
    return ops


def test_driver_reproduces_long_reference_run_with_cv2():
    g, vo, poses = run_long(_cv2_ops())
    assert vo.num_pts == [int(v) for v in g["num_pts"]]
    assert poses.shape == g["poses"].shape and np.abs(poses - g["poses"]).max() < 1e-8


def test_driver_reproduces_reference_run_with_cv2():
    ops = _cv2_ops()
    g, vo = run_on_trace(ops)
    assert vo.num_pts == [int(v) for v in g["num_pts"]]
    n = sum(1 for k in g.files if k.startswith("tri") and k.endswith("_cur"))
    assert len(vo.poses) == n
    for i in range(n):
        assert np.abs(vo.poses[i] - g[f"tri{i}_cur"]).max() < 1e-9, i
