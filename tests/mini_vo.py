"""Test infrastructure: a minimal restatement of the DATA FLOW of the reference's per-frame loop
(VisualOdometryPipeLine.py `initialization` :295-330 after the SIFT matching, `continuous_operation`
:333-373, `feature_tracking` :271-290, `feature_adding` :248-268, `triangulate_landmarks` :107-206) with
every numerical operation injected.  It exists so that the hot-path implementations can be run
FREE-RUNNING (each call fed by the previous calls' outputs) and compared with the recorded run of the
unmodified reference class (tests/golden/reference_trace.npz): with ops = real cv2 + the oracle the driver
must reproduce the recording exactly (that validates the driver), with ops = the CUDA shim it must stay
within the tolerance SURVEY.md 8c (ii) states.  Not a product component."""
import numpy as np


def rodrigues(rvec):
    r = np.asarray(rvec, np.float64).ravel()
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * Kx


class MiniVO:
    def __init__(self, K, options, ops):
        self.K, self.o, self.ops = np.asarray(K, np.float64), options, ops
        self.transforms = [(np.eye(3), np.zeros((3, 1)))]
        self.lm = np.zeros((0, 3), np.float32)
        self.lm_kp = np.zeros((0, 2), np.float32)
        self.num_pts = []
        self.poses = []          # (R_CW | t_CW) handed to the triangulation of each frame

    def _filter_potential(self, mask):
        self.pot_keys, self.pot_first, self.pot_tr = self.pot_keys[mask], self.pot_first[mask], self.pot_tr[mask]

    def _triangulate(self, R, t):
        self.poses.append(np.hstack([np.reshape(R, 9), np.reshape(t, 3)]))
        too_short, new_lm, new_kp = self.ops.triangulate(self.K, self.o, self.pot_first, self.pot_keys, self.pot_tr, self.transforms, R, t)
        self.lm = np.concatenate([self.lm, new_lm.astype(np.float32)])
        self.lm_kp = np.concatenate([self.lm_kp, new_kp.astype(np.float32)])
        self._filter_potential(too_short)

    def initialize(self, first_keys, keys, frame1):
        E, mask = self.ops.findEssentialMat(first_keys, keys, self.K, 0.99, 1.0)
        inl = mask.ravel() == 1
        self.pot_first, self.pot_keys = first_keys[inl], keys[inl]
        self.pot_tr = np.zeros(int(inl.sum()), np.int32)
        _, R, t, _ = self.ops.recoverPose(E, self.pot_first, self.pot_keys, self.K)
        t = t * np.sign(t[2])
        self._triangulate(R, t)
        self.transforms.append((R, t))
        self.num_pts = [int(inl.sum())]
        self.frame = frame1

    def step(self, img):
        o = self.o
        p, st = self.ops.klt(self.frame, img, self.lm_kp, o['winSize'], o['maxLevel'], o['criteria'])
        tr = st.ravel() == 1
        self.lm_kp, self.lm = p[tr], self.lm[tr]
        if len(self.pot_keys) > 1:
            p, st = self.ops.klt(self.frame, img, self.pot_keys, o['winSize'], o['maxLevel'], o['criteria'])
            self.pot_keys = p
            self._filter_potential(st.ravel() == 1)
        if len(self.lm_kp) < 8:
            raise ValueError("Not enough keypoints for PnP")
        ok, rvec, tvec, inliers = self.ops.solvePnPRansac(self.lm, self.lm_kp, self.K, o['PnP_iterations'], o['PnP_error'], o['PnP_conf'])
        if not ok:
            raise ValueError("PnP failed")
        mask = np.isin(np.arange(len(self.lm)), np.asarray(inliers).ravel())
        self.lm, self.lm_kp = self.lm[mask], self.lm_kp[mask]
        R_WC = rodrigues(rvec)
        R_CW, t_CW = R_WC.T, -R_WC.T @ np.reshape(tvec, (3, 1))
        if len(self.pot_keys) > 1:
            self._triangulate(R_CW, t_CW)
        pts = np.asarray(self.ops.gftt(img, o['feature_max_corners'], o['feature_quality_level'], o['feature_min_dist'],
                                       o['feature_block_size'])).reshape(-1, 2)
        pts = pts[self.ops.min_distance(pts, self.pot_keys, o['feature_min_dist'])]
        self.pot_keys = np.concatenate([self.pot_keys, pts])
        self.pot_first = np.concatenate([self.pot_first, pts])
        self.pot_tr = np.concatenate([self.pot_tr, np.full(len(pts), len(self.transforms), np.int32)])
        self.transforms.append((R_CW, t_CW))
        self.num_pts.append(len(inliers))
        self.frame = img


REFERENCE_KITTI_OPTIONS = {      # main.py:20-44
    'min_dist_landmarks': 1, 'max_dist_landmarks': 150, 'min_baseline_angle': 2, 'min_baseline_frames': 2,
    'feature_ratio': 0.8, 'feature_max_corners': 1400, 'feature_quality_level': 0.1, 'feature_min_dist': 10,
    'feature_block_size': 3, 'feature_use_harris': False, 'winSize': (15, 15), 'maxLevel': 5, 'criteria': (3, 50, 0.01),
    'PnP_conf': 0.99, 'PnP_error': 8, 'PnP_iterations': 500,
}


def run_on_trace(ops):
    """Bootstrap from the recorded ratio-test survivors, then free-run the recorded frames."""
    import reference_trace
    g, frames = reference_trace.load()
    vo = MiniVO(g["K"], REFERENCE_KITTI_OPTIONS, ops)
    b0, b1 = (int(v) for v in g["bootstrap"])
    vo.initialize(g["emat0_p1"], g["emat0_p2"], frames[b1])
    for i in range(b1 + 1, len(frames)):
        vo.step(frames[i])
    return g, vo


def run_long(ops):
    """The 26-frame run of tests/golden/reference_long.npz (made by make_reference_long.py): only the bootstrap
    matches, the inlier counts and the poses of the reference are stored; frames are re-rendered."""
    import os
    import zlib

    from conftest import GOLDEN
    from monocular_visual_odometry_va4mr_b200 import synth
    g = np.load(os.path.join(GOLDEN, "reference_long.npz"))
    frames = synth.render_sequence(str(g["render_shape"]), int(g["render_n"]), seed=int(g["render_seed"]))["frames"]
    assert np.array_equal(np.array([zlib.crc32(f.tobytes()) for f in frames], np.uint32), g["frame_crc"]), "renderer drifted"
    vo = MiniVO(g["K"], REFERENCE_KITTI_OPTIONS, ops)
    b0, b1 = (int(v) for v in g["bootstrap"])
    vo.initialize(g["p1"], g["p2"], frames[b1])
    for i in range(b1 + 1, len(frames)):
        vo.step(frames[i])
    poses = np.array([np.hstack([np.reshape(R, 9), np.reshape(t, 3)]) for R, t in vo.transforms], np.float64)
    return g, vo, poses
