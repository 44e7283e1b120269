"""Runs the UNMODIFIED reference class (imported from /root/reference, never copied) for a longer synthetic
sequence with the real cv2 and stores only its trajectory: per-frame inlier counts and camera poses
(tests/golden/reference_long.npz, a few KB).  tests/test_free_running_gpu.py free-runs the CUDA path over the
same frames (re-rendered from the stored parameters) and compares.  The SIFT bootstrap matches (the inputs of
findEssentialMat) are stored too, since SIFT itself stays on the host.

Run in the build container:  python tests/golden/make_reference_long.py
"""
import os
import sys
import zlib

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_reference_trace import OPTIONS  # noqa: E402
from monocular_visual_odometry_va4mr_b200 import synth  # noqa: E402

RENDER = dict(shape="kitti", seed=5, n_frames=26, bootstrap=(0, 2))


def main():
    if not hasattr(np, "bool"):
        np.bool = bool
    from VisualOdometryPipeLine import VisualOdometryPipeLine
    s = synth.render_sequence(RENDER["shape"], RENDER["n_frames"], seed=RENDER["seed"])
    frames = [s["frames"][i] for i in range(RENDER["n_frames"])]
    real = cv2.findEssentialMat
    grabbed = {}

    def emat(p1, p2, K, **kw):
        grabbed["p1"], grabbed["p2"] = np.array(p1, copy=True), np.array(p2, copy=True)
        return real(p1, p2, K, **kw)

    cv2.findEssentialMat = emat
    try:
        vo = VisualOdometryPipeLine(s["K"], OPTIONS)
        b0, b1 = RENDER["bootstrap"]
        vo.initialization(frames[b0], frames[b1])
        for i in range(b1 + 1, len(frames)):
            vo.continuous_operation(frames[i])
    finally:
        cv2.findEssentialMat = real
    poses = np.array([np.hstack([np.reshape(R, 9), np.reshape(t, 3)]) for R, t in vo.transforms], np.float64)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_long.npz")
    np.savez_compressed(out, render_shape=np.array(RENDER["shape"]), render_seed=np.array(RENDER["seed"]), render_n=np.array(RENDER["n_frames"]),
                        bootstrap=np.array(RENDER["bootstrap"]), frame_crc=np.array([zlib.crc32(f.tobytes()) for f in frames], np.uint32),
                        K=s["K"], p1=grabbed["p1"], p2=grabbed["p2"], num_pts=np.array(vo.num_pts), poses=poses, cv2_version=np.array(cv2.__version__))
    print("frames", len(frames), "num_pts", [int(v) for v in vo.num_pts], "->", out, os.path.getsize(out), "bytes")
    print("path length", float(np.linalg.norm(np.diff(poses[:, 9:], axis=0), axis=1).sum()))


if __name__ == "__main__":
    main()
