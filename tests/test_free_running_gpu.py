"""Free-running integration (SURVEY.md 8c tier ii): the restated per-frame data flow (tests/mini_vo.py, validated
against the recorded reference run by tests/test_oracle_free_running.py) driven by the CUDA path only --
findEssentialMat, recoverPose, triangulation, KLT x2, solvePnPRansac, goodFeaturesToTrack and the candidate
distance filter feed each other for five frames; the result is compared with the run of the unmodified
reference class on cv2."""
from types import SimpleNamespace

import numpy as np
import pytest

from mini_vo import run_long, run_on_trace
from monocular_visual_odometry_va4mr_b200 import cv2_compat, hotpath

pytestmark = pytest.mark.gpu


def _cuda_ops():
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
    return ops


def test_free_running_26_frames_on_cuda_path():
    """completes without ValueError; over the first 10 frames the camera positions stay within
    0.5 % of the path length of the reference's; over the whole run within the reference's own sensitivity
    envelope (+-0.01 px tracker noise moves the reference itself by 2.5-10 % of the path length)."""
    g, vo, poses = run_long(_cuda_ops())
    ref = g["poses"]
    assert poses.shape == ref.shape and len(vo.num_pts) == len(g["num_pts"])
    path = np.concatenate([[0.0], np.cumsum(np.linalg.norm(np.diff(ref[:, 9:], axis=0), axis=1))])
    dev = np.linalg.norm(poses[:, 9:] - ref[:, 9:], axis=1)
    assert np.all(dev[:11] <= 0.005 * np.maximum(path[:11], 1.0)), dev[:11]
    assert np.sqrt((dev ** 2).mean()) <= 0.10 * path[-1], (dev.max(), path[-1])
    same = sum(int(a == b) for a, b in zip(vo.num_pts, g["num_pts"]))
    print("26-frame free run: identical inlier counts in", same, "of", len(vo.num_pts), "frames; max position deviation",
          float(dev.max()), "of path", float(path[-1]), "max rotation deviation", float(np.abs(poses[:, :9] - ref[:, :9]).max()))


def _oracle_ops():
    import oracle

    def klt(prev, nxt, pts, win, ml, crit):
        p, st, _ = oracle.calc_optical_flow_pyr_lk(prev, nxt, pts, win, ml, crit)
        return p, st

    def emat(p1, p2, K, prob, thr):
        E, m, _ = oracle.find_essential_mat(p1, p2, K, prob, thr, 1000)
        return E, m

    def pnp(obj, img, K, it, err, conf):
        ok, rv, tv, inl, _ = oracle.solve_pnp_ransac_p3p(obj, img, K, it, err, conf)
        return ok, rv, tv, inl

    def tri(K, o, first, keys, tr, transforms, R, t):
        return oracle.triangulate_landmarks(K, (o['min_dist_landmarks'], o['max_dist_landmarks'], o['min_baseline_angle'], o['min_baseline_frames']),
                                            first, keys, tr, oracle.pack_poses(transforms), np.hstack([np.reshape(R, 9), np.reshape(t, 3)]))

    return SimpleNamespace(klt=klt, gftt=lambda img, mc, q, md, bs: oracle.good_features_to_track(img, mc, q, md, bs), findEssentialMat=emat,
                           recoverPose=lambda E, p1, p2, K: oracle.recover_pose(E, p1, p2, K), solvePnPRansac=pnp, triangulate=tri,
                           min_distance=oracle.min_distance_mask)


def test_free_running_cuda_equals_free_running_oracle():
    """Where the 26-frame CUDA run leaves the reference's trajectory it does so because of the tracker's documented
    <= 0.02 px differences from cv2 (exact integer sums vs float32 SIMD lanes), not because of the CUDA code: the same
    loop driven by the ORACLE alone gives the CUDA run's inlier counts and poses."""
    g, vo_c, poses_c = run_long(_cuda_ops())
    _, vo_o, poses_o = run_long(_oracle_ops())
    assert vo_c.num_pts == vo_o.num_pts
    assert np.abs(poses_c - poses_o).max() < 1e-6


def test_free_running_pipeline_on_cuda_path():
    ops = _cuda_ops()
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


def test_bootstrap_from_raw_frames_on_cuda_path():
    """The whole bootstrap of `initialization` (:293-323) on the CUDA path, starting from the two raw frames:
    SIFT detectAndCompute x2 (:226-227) -> knnMatch(k=2) + ratio test (:229, :218-224) -> findEssentialMat (:308).
    The matched point pairs must be the ones the unmodified reference class produced with cv2 (recorded in
    reference_trace.npz): CUDA SIFT keypoint positions are bit-equal to cv2's and ~99.9 % of the descriptors are identical,
    so at most a pair or two may differ where one descriptor entry is off by one at a ratio-test tie."""
    import reference_trace
    g, frames = reference_trace.load()
    b0, b1 = (int(v) for v in g["bootstrap"])
    sift = cv2_compat.SIFT_create()
    kp0, d0 = sift.detectAndCompute(frames[b0], None)
    kp1, d1 = sift.detectAndCompute(frames[b1], None)
    matches = cv2_compat.BFMatcher().knnMatch(d0, d1, k=2)
    ratio = 0.8                                              # the reference's KITTI option 'feature_ratio' (main.py:28)
    good = [m for m, n in matches if m.distance < ratio * n.distance]
    pts0 = np.float32([kp0[m.queryIdx].pt for m in good]).reshape(-1, 2)
    pts1 = np.float32([kp1[m.trainIdx].pt for m in good]).reshape(-1, 2)
    want = {tuple(np.r_[a, b].tolist()) for a, b in zip(g["emat0_p1"], g["emat0_p2"])}
    got = {tuple(np.r_[a, b].tolist()) for a, b in zip(pts0, pts1)}
    assert abs(len(got) - len(want)) <= max(2, len(want) // 200), (len(got), len(want))
    assert len(got & want) >= len(want) - max(2, len(want) // 200), (len(got & want), len(want))
    E, mask = cv2_compat.findEssentialMat(pts0, pts1, g["K"], method=cv2_compat.RANSAC, prob=0.99, threshold=1.0)
    assert abs(int(mask.sum()) - int(g["num_pts"][0])) <= max(3, 0.01 * int(g["num_pts"][0]))
    print("bootstrap pairs", len(got), "recorded", len(want), "common", len(got & want), "E inliers", int(mask.sum()), "recorded", int(g["num_pts"][0]))
