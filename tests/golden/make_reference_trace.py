"""Records what the UNMODIFIED reference class does at its six hot-path call sites.

Imports VisualOdometryPipeLine from /root/reference (read-only, never copied), runs
`initialization` + `continuous_operation` on synthetic frames (the renderer in
monocular_visual_odometry_va4mr_b200/synth.py; the datasets of utils.py are not available offline)
with the real cv2, and wraps the five cv2 names the class calls on the hot path
(VisualOdometryPipeLine.py:229, :256, :281, :287, :308, :343) so that every call's inputs and cv2's
outputs are written to tests/golden/reference_trace.npz.  The two numpy components next to the hot
path are recorded at method level: `triangulate_landmarks` (:107-206; inputs = the
candidate state + poses, outputs = the mask handed to filter_potential and the appended rows) and
`feature_adding` (:248-268; the min-distance mask of :258).  The parity tests replay each call
("teacher-forced") through the oracle (CPU) and through the C-ABI (GPU).

Frames are NOT stored: the fixture keeps the render parameters and a CRC32 per frame; the tests
re-render and check the CRC.  Descriptors are integer-valued (SIFT) and stored as uint8.

Run in the build container:  python tests/golden/make_reference_trace.py
"""
import os
import sys
import zlib

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from monocular_visual_odometry_va4mr_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_trace.npz")

# the reference's KITTI options (main.py:20-44)
OPTIONS = {
    'min_dist_landmarks': 1, 'max_dist_landmarks': 150, 'min_baseline_angle': 2, 'min_baseline_frames': 2,
    'feature_ratio': 0.8, 'feature_max_corners': 1400, 'feature_quality_level': 0.1, 'feature_min_dist': 10,
    'feature_block_size': 3, 'feature_use_harris': False,
    'winSize': (15, 15), 'maxLevel': 5, 'criteria': (cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 50, 0.01),
    'PnP_conf': 0.99, 'PnP_error': 8, 'PnP_iterations': 500,
}
RENDER = dict(shape="kitti", seed=2, n_frames=7, bootstrap=(0, 2))   # main.py:18 bootstrap_frames = [0, 2]


def main():
    if not hasattr(np, "bool"):          # the reference uses np.bool (:346); numpy 1.24-1.26 dropped it
        np.bool = bool
    from VisualOdometryPipeLine import VisualOdometryPipeLine

    s = synth.render_sequence(RENDER["shape"], RENDER["n_frames"], seed=RENDER["seed"])
    frames = s["frames"]
    frame_id = {id(f): i for i, f in enumerate(frames)}
    frame_list = [frames[i] for i in range(len(frames))]     # stable objects: the class keeps references
    frame_id = {id(f): i for i, f in enumerate(frame_list)}

    rec = dict(render_shape=np.array(RENDER["shape"]), render_seed=np.array(RENDER["seed"]),
               render_n=np.array(RENDER["n_frames"]), bootstrap=np.array(RENDER["bootstrap"]),
               frame_crc=np.array([zlib.crc32(f.tobytes()) for f in frame_list], np.uint32),
               K=s["K"], cv2_version=np.array(cv2.__version__))
    calls = []   # (kind, index within kind)
    n = dict(klt=0, gftt=0, knn=0, emat=0, pnp=0, tri=0, fadd=0, rpose=0)
    real = dict(klt=cv2.calcOpticalFlowPyrLK, gftt=cv2.goodFeaturesToTrack, bf=cv2.BFMatcher,
                emat=cv2.findEssentialMat, pnp=cv2.solvePnPRansac, rpose=cv2.recoverPose)

    def klt(prev, nxt, pts, nextPts, **kw):
        out = real["klt"](prev, nxt, pts, nextPts, **kw)
        i = n["klt"]; n["klt"] += 1
        rec[f"klt{i}_prev"] = np.array(frame_id[id(prev)]); rec[f"klt{i}_next"] = np.array(frame_id[id(nxt)])
        rec[f"klt{i}_pts"] = np.array(pts, copy=True)
        rec[f"klt{i}_out"] = out[0]; rec[f"klt{i}_status"] = out[1]
        rec[f"klt{i}_err"] = np.where(out[1] == 1, out[2], 0).astype(np.float32)
        rec[f"klt{i}_cfg"] = np.array([kw["winSize"][0], kw["winSize"][1], kw["maxLevel"], kw["criteria"][0], kw["criteria"][1]], np.int32)
        rec[f"klt{i}_eps"] = np.array(kw["criteria"][2])
        calls.append(("klt", i))
        return out

    def gftt(img, **kw):
        out = real["gftt"](img, **kw)
        i = n["gftt"]; n["gftt"] += 1
        rec[f"gftt{i}_frame"] = np.array(frame_id[id(img)])
        rec[f"gftt{i}_cfg"] = np.array([kw["maxCorners"], kw["qualityLevel"], kw["minDistance"], kw["blockSize"]], np.float64)
        rec[f"gftt{i}_out"] = out
        calls.append(("gftt", i))
        return out

    class Matcher:
        def __init__(self, *a, **kw):
            self._m = real["bf"](*a, **kw)

        def knnMatch(self, q, t, k):
            out = self._m.knnMatch(q, t, k=k)
            i = n["knn"]; n["knn"] += 1
            assert np.array_equal(q, np.rint(q)) and q.max() <= 255 and np.array_equal(t, np.rint(t)) and t.max() <= 255
            rec[f"knn{i}_q"] = q.astype(np.uint8); rec[f"knn{i}_t"] = t.astype(np.uint8)
            rec[f"knn{i}_idx"] = np.array([[m.trainIdx for m in row] for row in out], np.int32)
            rec[f"knn{i}_dist"] = np.array([[m.distance for m in row] for row in out], np.float32)
            calls.append(("knn", i))
            return out

    def emat(p1, p2, K, **kw):
        out = real["emat"](p1, p2, K, **kw)
        i = n["emat"]; n["emat"] += 1
        rec[f"emat{i}_p1"] = np.array(p1, copy=True); rec[f"emat{i}_p2"] = np.array(p2, copy=True)
        rec[f"emat{i}_cfg"] = np.array([kw["prob"], kw["threshold"]])
        rec[f"emat{i}_E"] = out[0]; rec[f"emat{i}_mask"] = out[1]
        calls.append(("emat", i))
        return out

    def pnp(obj, img, K, dist, **kw):
        out = real["pnp"](obj, img, K, dist, **kw)
        i = n["pnp"]; n["pnp"] += 1
        rec[f"pnp{i}_obj"] = np.array(obj, copy=True); rec[f"pnp{i}_img"] = np.array(img, copy=True)
        rec[f"pnp{i}_cfg"] = np.array([kw["iterationsCount"], kw["reprojectionError"], kw["confidence"]], np.float64)
        rec[f"pnp{i}_ok"] = np.array(out[0]); rec[f"pnp{i}_rvec"] = out[1]; rec[f"pnp{i}_tvec"] = out[2]
        rec[f"pnp{i}_inliers"] = out[3] if out[3] is not None else np.zeros((0, 1), np.int32)
        calls.append(("pnp", i))
        return out

    def rpose(E, p1, p2, K):                    # :315, "next" row f3
        out = real["rpose"](E, p1, p2, K)
        i = n["rpose"]; n["rpose"] += 1
        rec[f"rpose{i}_E"] = np.array(E, copy=True); rec[f"rpose{i}_p1"] = np.array(p1, copy=True); rec[f"rpose{i}_p2"] = np.array(p2, copy=True)
        # copies: the reference flips t in place right after the call (:317)
        rec[f"rpose{i}_good"] = np.array(out[0]); rec[f"rpose{i}_R"] = np.array(out[1], copy=True)
        rec[f"rpose{i}_t"] = np.array(out[2], copy=True); rec[f"rpose{i}_mask"] = np.array(out[3], copy=True)
        calls.append(("rpose", i))
        return out

    # ---- the two "next" rows of SURVEY.md 8f, recorded at method level (they are numpy, not cv2 calls) ----
    orig_tri = VisualOdometryPipeLine.triangulate_landmarks
    orig_add = VisualOdometryPipeLine.feature_adding

    def tri(self, R_cur, t_cur):
        i = n["tri"]; n["tri"] += 1
        nk = len(self.potential_keys)
        rec[f"tri{i}_first_keys"] = np.array(self.potential_first_keys, copy=True)
        rec[f"tri{i}_keys"] = np.array(self.potential_keys, copy=True)
        rec[f"tri{i}_first_pose"] = np.array(self.potential_transforms, copy=True).reshape(-1).astype(np.int32)
        rec[f"tri{i}_poses"] = np.array([np.hstack([R.reshape(9), np.reshape(t, 3)]) for R, t in self.transforms], np.float64)
        rec[f"tri{i}_cur"] = np.hstack([np.reshape(R_cur, 9), np.reshape(t_cur, 3)]).astype(np.float64)
        n_lm0 = len(self.matched_landmarks)
        grabbed = {}
        orig_filter = self.filter_potential

        def grab(mask):
            grabbed["mask"] = np.array(mask, copy=True)
            return orig_filter(mask)

        self.filter_potential = grab
        try:
            out = orig_tri(self, R_cur, t_cur)
        finally:
            del self.filter_potential
        assert len(grabbed["mask"]) == nk
        rec[f"tri{i}_keep"] = grabbed["mask"].astype(np.uint8)
        rec[f"tri{i}_landmarks"] = np.array(self.matched_landmarks[n_lm0:], copy=True)
        rec[f"tri{i}_keypoints"] = np.array(self.matched_keypoints[n_lm0:], copy=True)
        rec[f"tri{i}_lm_dtype"] = np.array(str(np.asarray(self.matched_landmarks).dtype))
        calls.append(("tri", i))
        return out

    def fadd(self, img):
        i = n["fadd"]; n["fadd"] += 1
        existing = np.array(self.potential_keys, copy=True)
        n_gftt = n["gftt"]
        out = orig_add(self, img)
        pts = rec[f"gftt{n_gftt}_out"].squeeze()
        # the reference's own expression (:258), evaluated on the recorded inputs
        valid = np.array([np.all(np.linalg.norm(pts[j, :] - existing, axis=1) > OPTIONS['feature_min_dist']) for j in range(pts.shape[0])])
        assert np.array_equal(np.asarray(self.potential_keys)[len(existing):], pts[valid])
        rec[f"fadd{i}_existing"] = existing
        rec[f"fadd{i}_gftt"] = np.array(n_gftt)
        rec[f"fadd{i}_valid"] = valid.astype(np.uint8)
        rec[f"fadd{i}_min_dist"] = np.array(float(OPTIONS['feature_min_dist']))
        calls.append(("fadd", i))
        return out

    VisualOdometryPipeLine.triangulate_landmarks = tri
    VisualOdometryPipeLine.feature_adding = fadd
    rec["tri_cfg"] = np.array([OPTIONS['min_dist_landmarks'], OPTIONS['max_dist_landmarks'], OPTIONS['min_baseline_angle'],
                               OPTIONS['min_baseline_frames']], np.float64)

    cv2.calcOpticalFlowPyrLK, cv2.goodFeaturesToTrack, cv2.BFMatcher = klt, gftt, Matcher
    cv2.findEssentialMat, cv2.solvePnPRansac, cv2.recoverPose = emat, pnp, rpose
    try:
        vo = VisualOdometryPipeLine(s["K"], OPTIONS)
        b0, b1 = RENDER["bootstrap"]
        vo.initialization(frame_list[b0], frame_list[b1])
        for i in range(b1 + 1, len(frame_list)):
            vo.continuous_operation(frame_list[i])
    finally:
        cv2.calcOpticalFlowPyrLK, cv2.goodFeaturesToTrack, cv2.BFMatcher = real["klt"], real["gftt"], real["bf"]
        cv2.findEssentialMat, cv2.solvePnPRansac, cv2.recoverPose = real["emat"], real["pnp"], real["rpose"]
        VisualOdometryPipeLine.triangulate_landmarks, VisualOdometryPipeLine.feature_adding = orig_tri, orig_add
    rec["calls"] = np.array([f"{k}{i}" for k, i in calls])
    rec["num_pts"] = np.array(vo.num_pts)
    np.savez_compressed(OUT, **rec)
    print("calls:", {k: v for k, v in n.items()}, "num_pts:", vo.num_pts, "->", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
