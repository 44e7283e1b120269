"""Oracle for the two components next to the hot path (candidate min-distance filter :258,
triangulate_landmarks :107-206) against the recorded behaviour of the unmodified reference class
(tests/golden/reference_trace.npz) and against live cv2 / the reference's own numpy expressions."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "reference_trace.npz"))


def test_min_distance_mask_vs_reference_trace(g):
    n = sum(1 for k in g.files if k.startswith("fadd") and k.endswith("_valid"))
    assert n >= 3
    for i in range(n):
        pts = g[f"gftt{int(g[f'fadd{i}_gftt'])}_out"].reshape(-1, 2)
        valid = oracle.min_distance_mask(pts, g[f"fadd{i}_existing"], float(g[f"fadd{i}_min_dist"]))
        assert np.array_equal(valid, g[f"fadd{i}_valid"].astype(bool)), i


def test_min_distance_mask_vs_numpy_expression():
    rng = np.random.default_rng(5)
    for n, m in ((300, 700), (1, 1), (50, 0), (0, 40)):
        pts = np.rint(rng.uniform(0, 400, (n, 2))).astype(np.float32)           # gFTT corners are integer-valued
        ex = rng.uniform(0, 400, (m, 2)).astype(np.float32)
        if n > 1 and m > 1:   # plant exact ties: distance exactly 10 (not > 10) and the next float above
            ex[0] = pts[0] + np.float32([6, 8])
            ex[1] = pts[1] + np.float32([6, np.nextafter(np.float32(8), np.float32(9))])
        ref = np.array([np.all(np.linalg.norm(pts[i, :] - ex, axis=1) > 10) for i in range(n)], bool)   # ref :258
        assert np.array_equal(oracle.min_distance_mask(pts, ex, 10.0), ref)


def _replay_tri(g, i, fn):
    keep, lm, kp = fn(g["K"], g["tri_cfg"], g[f"tri{i}_first_keys"], g[f"tri{i}_keys"], g[f"tri{i}_first_pose"],
                      g[f"tri{i}_poses"], g[f"tri{i}_cur"])
    assert np.array_equal(keep, g[f"tri{i}_keep"].astype(bool)), f"tri{i}: too_short_baseline mask"
    ref_lm = g[f"tri{i}_landmarks"].reshape(-1, 3)
    assert lm.shape == ref_lm.shape and lm.dtype == np.float32
    assert np.array_equal(kp, g[f"tri{i}_keypoints"].reshape(-1, 2))
    if len(lm):
        # float32 results of a float64 solve: equal up to one float32 ulp (observed: bit-equal)
        assert np.all(np.abs(lm - ref_lm) <= np.spacing(np.abs(ref_lm))), f"tri{i}: landmarks"
    return int((lm == ref_lm).sum()), lm.size


def test_triangulate_vs_reference_trace(g):
    n = sum(1 for k in g.files if k.startswith("tri") and k.endswith("_keep"))
    assert n >= 4
    same = tot = 0
    for i in range(n):
        s, t = _replay_tri(g, i, oracle.triangulate_landmarks)
        same += s; tot += t
    assert tot > 300 and same >= 0.999 * tot


def test_triangulate_points_vs_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(9)
    K = np.array([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1]])
    n = 400
    Xw = np.column_stack([rng.uniform(-8, 8, n), rng.uniform(-2, 1.6, n), rng.uniform(5, 120, n)])
    ang = 0.03
    R1 = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])   # world -> camera 1
    t1 = np.array([[0.1], [0.02], [-1.6]])
    # the reference stores camera-in-world style (R_CW, t_CW) and inverts them (:60-76)
    poses = oracle.pack_poses([(np.eye(3), np.zeros((3, 1))), (R1.T, -R1.T @ t1)])
    p0 = (K @ Xw.T).T; p0 = (p0[:, :2] / p0[:, 2:]).astype(np.float32)
    x1 = (K @ (R1 @ Xw.T + t1)).T; p1 = (x1[:, :2] / x1[:, 2:] + rng.normal(0, 0.3, (n, 2))).astype(np.float32)
    keep, lm, kp = oracle.triangulate_landmarks(K, (1, 150, 0.0, 0), p0, p1, np.zeros(n, np.int32), poses, poses[1])
    P0 = K @ np.hstack([np.eye(3), np.zeros((3, 1))]); P1 = K @ np.hstack([R1, t1])
    k = 0
    for i in range(n):
        X = cv2.triangulatePoints(P0, P1, p0[i].reshape(-1, 1), p1[i].reshape(-1, 1))
        L = (X[:3] / X[3]).ravel()
        z0, z1 = float(L[2]), float((R1 @ L.astype(np.float64) + t1.ravel())[2])
        ok = 1 < z0 < 150 and 1 < z1 < 150
        assert keep[i] == (not ok), i
        if ok:
            assert np.all(np.abs(lm[k] - L) <= np.spacing(np.abs(L))), i
            k += 1
    assert k == len(lm) and k > 200


def test_recover_pose_vs_live_cv2():
    """ref :315 -- count, mask (0/255) and the chosen (R, t) as cv2.recoverPose returns them."""
    cv2 = pytest.importorskip("cv2")
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden import make_emat_pair
    for n, of, seed in ((600, 0.2, 7), (2500, 0.3, 40), (300, 0.6, 3), (50, 0.0, 9), (8, 0.0, 1)):
        p1, p2, K = make_emat_pair(n, of, seed)
        E, m = cv2.findEssentialMat(p1, p2, K, method=cv2.RANSAC, prob=0.99, threshold=1)
        E = E[:3]
        good, R, t, mask = cv2.recoverPose(E, p1, p2, K)
        go, Ro, to, mo = oracle.recover_pose(E, p1, p2, K)
        assert go == good and np.array_equal(mo, mask) and mo.dtype == np.uint8
        assert np.abs(R - Ro).max() < 1e-12 and np.abs(t - to).max() < 1e-12
