"""Host-side contract of the cv2-shaped shim (no GPU needed): argument names / order / defaults follow cv2's
documented signatures, argument patterns the reference never uses raise NotImplementedError
BEFORE any device work (never a CPU fallback), cv2's own assertion conditions raise cv2.error-like errors,
install() / uninstall() patch and restore exactly the six names, and unknown attributes forward to cv2."""
import inspect

import numpy as np
import pytest

from monocular_visual_odometry_va4mr_b200 import cv2_compat, hotpath


def _params(f):
    return [(p.name, p.default) for p in inspect.signature(f).parameters.values()]


def test_signatures_follow_cv2():
    P = inspect.Parameter.empty
    assert _params(cv2_compat.calcOpticalFlowPyrLK) == [
        ("prevImg", P), ("nextImg", P), ("prevPts", P), ("nextPts", P), ("status", None), ("err", None), ("winSize", (21, 21)),
        ("maxLevel", 3), ("criteria", (3, 30, 0.01)), ("flags", 0), ("minEigThreshold", 1e-4)]
    assert _params(cv2_compat.goodFeaturesToTrack) == [
        ("image", P), ("maxCorners", P), ("qualityLevel", P), ("minDistance", P), ("corners", None), ("mask", None), ("blockSize", 3),
        ("useHarrisDetector", False), ("k", 0.04)]
    assert _params(cv2_compat.findEssentialMat)[:7] == [
        ("points1", P), ("points2", P), ("cameraMatrix", None), ("method", cv2_compat.RANSAC), ("prob", 0.999), ("threshold", 1.0), ("maxIters", 1000)]
    assert [n for n, _ in _params(cv2_compat.recoverPose)][:4] == ["E", "points1", "points2", "cameraMatrix"]
    assert _params(cv2_compat.solvePnPRansac) == [
        ("objectPoints", P), ("imagePoints", P), ("cameraMatrix", P), ("distCoeffs", P), ("rvec", None), ("tvec", None),
        ("useExtrinsicGuess", False), ("iterationsCount", 100), ("reprojectionError", 8.0), ("confidence", 0.99), ("inliers", None),
        ("flags", cv2_compat.SOLVEPNP_ITERATIVE)]
    assert [n for n, _ in _params(cv2_compat.BFMatcher.knnMatch)][1:4] == ["queryDescriptors", "trainDescriptors", "k"]
    try:
        import cv2
    except ImportError:
        return
    assert (cv2_compat.RANSAC, cv2_compat.SOLVEPNP_P3P, cv2_compat.SOLVEPNP_ITERATIVE, cv2_compat.NORM_L2) == (
        cv2.RANSAC, cv2.SOLVEPNP_P3P, cv2.SOLVEPNP_ITERATIVE, cv2.NORM_L2)
    assert cv2_compat.TERM_CRITERIA_COUNT + cv2_compat.TERM_CRITERIA_EPS == cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS == 3


def test_unsupported_patterns_raise_before_any_device_work():
    img = np.zeros((40, 60), np.uint8)
    pts = np.zeros((3, 2), np.float32)
    K = np.eye(3)
    with pytest.raises(NotImplementedError):
        cv2_compat.calcOpticalFlowPyrLK(img, img, pts, np.zeros((3, 2), np.float32))          # initial flow / preallocated output
    with pytest.raises(NotImplementedError):
        cv2_compat.goodFeaturesToTrack(img, 10, 0.1, 5, mask=np.ones_like(img))
    with pytest.raises(NotImplementedError):
        cv2_compat.goodFeaturesToTrack(img, 10, 0.1, 5, useHarrisDetector=True)
    with pytest.raises(NotImplementedError):
        cv2_compat.findEssentialMat(pts, pts, K, method=cv2_compat.LMEDS)
    with pytest.raises(NotImplementedError):
        cv2_compat.findEssentialMat(pts, pts)                                                  # focal / pp overload
    with pytest.raises(NotImplementedError):
        cv2_compat.recoverPose(np.eye(3), pts, pts)
    with pytest.raises(NotImplementedError):
        cv2_compat.solvePnPRansac(np.zeros((5, 3), np.float32), np.zeros((5, 2), np.float32), K, np.zeros(4))   # flags=ITERATIVE
    with pytest.raises(NotImplementedError):
        cv2_compat.solvePnPRansac(np.zeros((5, 3), np.float32), np.zeros((5, 2), np.float32), K, np.float64([0.1, 0, 0, 0]),
                                  flags=cv2_compat.SOLVEPNP_P3P)
    with pytest.raises(NotImplementedError):
        cv2_compat.BFMatcher(crossCheck=True)
    with pytest.raises(NotImplementedError):
        cv2_compat.BFMatcher().knnMatch(np.zeros((2, 128), np.float32), np.zeros((3, 128), np.float32), k=3)


def test_cv2_assertion_conditions_raise_cv2_style_errors():
    img = np.zeros((40, 60), np.uint8)
    with pytest.raises(cv2_compat.error):
        cv2_compat.calcOpticalFlowPyrLK(img, np.zeros((41, 60), np.uint8), np.zeros((3, 2), np.float32), None)   # size mismatch
    with pytest.raises(cv2_compat.error):
        cv2_compat.calcOpticalFlowPyrLK(img, img, np.zeros((3, 2), np.float64), None)                          # CV_32F points only
    with pytest.raises(cv2_compat.error):
        cv2_compat.goodFeaturesToTrack(img, 10, 0.0, 5)                                                       # qualityLevel > 0
    with pytest.raises(cv2_compat.error):
        cv2_compat.findEssentialMat(np.zeros((6, 2), np.float32), np.zeros((5, 2), np.float32), np.eye(3))
    assert cv2_compat.calcOpticalFlowPyrLK(img, img, np.zeros((0, 2), np.float32), None) == (None, None, None)   # cv2: N == 0
    assert cv2_compat.findEssentialMat(np.zeros((4, 2), np.float32), np.zeros((4, 2), np.float32), np.eye(3)) == (None, None)


def test_install_patches_and_restores_seven_names():
    cv2 = pytest.importorskip("cv2")
    names = ("calcOpticalFlowPyrLK", "goodFeaturesToTrack", "BFMatcher", "findEssentialMat", "recoverPose", "solvePnPRansac", "SIFT_create")
    orig = {n: getattr(cv2, n) for n in names}
    untouched = {n: getattr(cv2, n) for n in ("triangulatePoints", "Rodrigues", "pyrDown")}
    try:
        cv2_compat.install()
        for n in names:
            assert getattr(cv2, n) is getattr(cv2_compat, n), n
        for n, f in untouched.items():
            assert getattr(cv2, n) is f, n
    finally:
        cv2_compat.uninstall()
    for n in names:
        assert getattr(cv2, n) is orig[n], n
    assert cv2_compat.Rodrigues is cv2.Rodrigues                     # unknown attributes forward to the real cv2


def test_hotpath_argument_checks():
    K = np.eye(3)
    opt = dict(min_dist_landmarks=1, max_dist_landmarks=150, min_baseline_angle=2, min_baseline_frames=2)
    with pytest.raises(ValueError):
        hotpath.triangulate_landmarks(K, opt, np.zeros((3, 2), np.float32), np.zeros((2, 2), np.float32), np.zeros(3),
                                      [(np.eye(3), np.zeros((3, 1)))], np.eye(3), np.zeros((3, 1)))
    assert hotpath.pack_poses([(np.eye(3), np.zeros((3, 1))), (2 * np.eye(3), np.ones((3, 1)))]).shape == (2, 12)
