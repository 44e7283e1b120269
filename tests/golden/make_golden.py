"""Generates the committed golden vectors under tests/golden/ from the INSTALLED cv2
(4.13.0 in the build container) -- the third-party library that holds the reference's
hot-path arithmetic (reference requirements.txt:6 pins opencv-python==4.6.0.66; see
SURVEY.md 8c).  Run from the repo root:  python tests/golden/make_golden.py [name ...]

The fixtures travel to the GPU box; /root/reference and (possibly) cv2 do not.
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from monocular_visual_odometry_va4mr_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def klt_small():
    s = synth.render_sequence("kitti", 2, seed=3, width=320, height=240)
    f0, f1 = s["frames"]
    rng = np.random.default_rng(0)
    pts = cv2.goodFeaturesToTrack(f0, 300, 0.01, 4).reshape(-1, 2)
    pts = pts + rng.uniform(-0.5, 0.5, pts.shape).astype(np.float32)
    edge = np.float32([[0, 0], [-5, -5], [319.4, 239.2], [400, 100], [3.2, 236.9], [160, -30], [0.5, 120.25],
                       [319, 0], [-20.9, -20.9], [339.5, 100]])
    pts = np.ascontiguousarray(np.concatenate([pts, edge]), np.float32)
    out = dict(f0=f0, f1=f1, pts=pts)
    for tag, win, ml, crit in (("w21", (21, 21), 3, (3, 30, 0.01)), ("w15", (15, 15), 5, (3, 50, 0.01)),
                               ("w9x13", (9, 13), 2, (1, 10, 0.01))):
        p, st, err = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None, winSize=win, maxLevel=ml, criteria=crit)
        out[f"{tag}_next"] = p
        out[f"{tag}_status"] = st
        out[f"{tag}_err"] = np.where(st == 1, err, 0).astype(np.float32)
        out[f"{tag}_cfg"] = np.array([win[0], win[1], ml, crit[0], crit[1]], np.int32)
        out[f"{tag}_eps"] = np.array([crit[2]])
    out["pyr1"] = cv2.pyrDown(f0)
    out["pyr2"] = cv2.pyrDown(out["pyr1"])
    out["scharr_x"] = cv2.Scharr(f0, cv2.CV_16S, 1, 0)
    out["scharr_y"] = cv2.Scharr(f0, cv2.CV_16S, 0, 1)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(OUT, "klt_small.npz"), **out)


ALL = dict(klt_small=klt_small)

if __name__ == "__main__":
    names = sys.argv[1:] or list(ALL)
    for nm in names:
        ALL[nm]()
        print("wrote", nm)
