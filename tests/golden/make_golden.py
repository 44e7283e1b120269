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


def make_pnp_case(n, out_frac, seed, noise=0.3, shape="kitti"):
    """Landmarks in a corridor-like volume, projected with the frame-7 pose, N(0,noise) px noise,
    `out_frac` gross outliers (+-60 px uniform).  float32 arrays as the reference passes them."""
    rng = np.random.default_rng(seed)
    K, w, h, step, _ = synth.SHAPES[shape]
    Xc = np.column_stack([rng.uniform(-8, 8, n), rng.uniform(-2, 1.6, n), rng.uniform(4, 80, n)])
    R_cw, c = synth.Corridor.pose(7, step)
    Xw = c[None] + Xc @ R_cw.T
    uv = synth.project(K, R_cw, c, Xw) + rng.normal(0, noise, (n, 2))
    no = int(out_frac * n)
    idx = rng.choice(n, no, replace=False)
    uv[idx] += rng.uniform(-60, 60, (no, 2))
    return Xw.astype(np.float32), uv.astype(np.float32), K.copy()


PNP_CASES = [  # n, outlier fraction, seed, iterationsCount, reprojectionError
    (2000, 0.1, 0, 500, 8.0), (2000, 0.5, 1, 500, 8.0), (1000, 0.3, 2, 500, 5.0), (983, 0.2, 3, 2000, 8.0),
    (2000, 0.3, 5, 500, 8.0), (500, 0.7, 6, 500, 5.0), (73, 0.2, 7, 500, 8.0), (8, 0.0, 8, 500, 8.0),
    (300, 0.85, 9, 500, 8.0), (4, 0.0, 10, 500, 8.0), (5000, 0.6, 4, 2000, 8.0), (40, 1.0, 11, 50, 2.0),
]


def pnp():
    out = dict(cases=np.array(PNP_CASES, np.float64))
    for ci, (n, of, seed, iters, thr) in enumerate(PNP_CASES):
        obj, img, K = make_pnp_case(n, of, seed)
        ok, rv, tv, inl = cv2.solvePnPRansac(obj, img, K, np.zeros(4), flags=cv2.SOLVEPNP_P3P, confidence=0.99,
                                             reprojectionError=thr, iterationsCount=iters)
        out[f"c{ci}_obj"], out[f"c{ci}_img"], out[f"c{ci}_K"] = obj, img, K
        out[f"c{ci}_ok"] = np.array(bool(ok))
        out[f"c{ci}_rvec"], out[f"c{ci}_tvec"] = rv, tv
        out[f"c{ci}_inliers"] = inl if inl is not None else np.zeros((0, 1), np.int32)
    # minimal-solver vectors: cv2.solvePnP(flags=P3P) on 4-point samples (some with an outlier)
    obj, img, K = make_pnp_case(400, 0.3, 21)
    rng = np.random.default_rng(5)
    S = np.stack([rng.choice(400, 4, replace=False) for _ in range(300)]).astype(np.int32)
    oks, rvs, tvs = [], [], []
    for smp in S:
        try:
            ok, rv, tv = cv2.solvePnP(obj[smp], img[smp], K, np.zeros(4), flags=cv2.SOLVEPNP_P3P)
        except cv2.error:
            ok, rv, tv = False, np.zeros((3, 1)), np.zeros((3, 1))
        fin = bool(ok) and np.all(np.isfinite(tv)) and np.all(np.isfinite(rv))
        oks.append(fin); rvs.append(rv.ravel() if fin else np.zeros(3)); tvs.append(tv.ravel() if fin else np.zeros(3))
    out.update(min_obj=obj, min_img=img, min_K=K, min_samples=S, min_ok=np.array(oks), min_rvec=np.array(rvs),
               min_tvec=np.array(tvs))
    # reprojection errors (cv2.projectPoints, float32 output) for one pose
    ok, rv, tv, _ = cv2.solvePnPRansac(obj, img, K, np.zeros(4), flags=cv2.SOLVEPNP_P3P, reprojectionError=8.0, iterationsCount=100)
    proj, _ = cv2.projectPoints(obj, rv, tv, K, np.zeros(4))
    d = img - proj.reshape(-1, 2)
    out.update(err_rvec=rv, err_tvec=tv, err_vals=(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32))
    # EPnP refit vectors (float64 inputs, as solvePnPRansac feeds it)
    for ei, (n, seed, noise) in enumerate([(200, 0, 0.3), (1000, 1, 0.3), (6, 4, 0.3), (1000, 5, 2.0), (50, 3, 0.0)]):
        o, i, K = make_pnp_case(n, 0.0, seed, noise)
        ok, rv, tv = cv2.solvePnP(o.astype(np.float64), i.astype(np.float64), K, np.zeros(4), flags=cv2.SOLVEPNP_EPNP)
        out[f"e{ei}_obj"], out[f"e{ei}_img"], out[f"e{ei}_K"], out[f"e{ei}_rvec"], out[f"e{ei}_tvec"] = o, i, K, rv, tv
    # cv::RNG known answers and SVD sign convention
    svd_in = np.random.default_rng(3).normal(0, 3, (20, 3, 3))
    svd_in = np.einsum("nij,nkj->nik", svd_in, svd_in)
    out["svd_in"] = svd_in
    out["svd_u"] = np.stack([cv2.SVDecomp(m)[1] for m in svd_in])
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(OUT, "pnp.npz"), **out)


def gftt():
    out = {}
    s = synth.render_sequence("kitti", 1, seed=11, width=400, height=200)
    img = s["frames"][0]
    blobs = np.full((200, 300), 50, np.uint8)
    for (x, y) in [(20, 20), (100, 30), (200, 50), (50, 100), (150, 120), (250, 150), (30, 170)]:
        blobs[y:y + 6, x:x + 6] = 200     # seven identical blobs: exact ties, order = descending address
    out["img"], out["blobs"] = img, blobs
    out["eig"] = cv2.cornerMinEigenVal(img, 3, ksize=3)
    cases = [(1400, 0.1, 10.0), (1400, 0.03, 10.0), (2000, 0.001, 3.0), (0, 0.2, 7.5), (500, 0.05, 0.0), (300, 0.01, 1.4), (5, 0.01, 10.0)]
    out["cases"] = np.array(cases)
    for name, im in (("img", img), ("blobs", blobs)):
        for ci, (mc, q, md) in enumerate(cases):
            c = cv2.goodFeaturesToTrack(im, int(mc), q, md, blockSize=3, useHarrisDetector=False)
            out[f"{name}_c{ci}"] = c if c is not None else np.zeros((0, 1, 2), np.float32)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(OUT, "gftt.npz"), **out)


def sift_like(n, seed):
    """Integer-valued 0..255 descriptors with |d| ~ 512, like cv2 SIFT output."""
    r = np.random.default_rng(seed)
    x = r.gamma(0.6, 30, (n, 128))
    x = x / np.linalg.norm(x, axis=1, keepdims=True) * 512
    return np.clip(np.rint(x), 0, 255).astype(np.float32)


def knn():
    out = {}
    rng = np.random.default_rng(0)
    q, t = sift_like(700, 1), sift_like(900, 2)
    t[50] = t[100]; q[7] = t[50]          # planted duplicates: tie -> lower train index first
    t[10] = q[8]; t[800] = q[8]
    t[300:340] = np.clip(q[100:140] + rng.integers(-6, 7, (40, 128)), 0, 255)   # true matches (ratio accepts)
    q[20] = 255; t[21] = 255; t[22] = 0    # extreme norms: d^2 = 0 and 128*255^2
    m = cv2.BFMatcher().knnMatch(q, t, k=2)
    out["q"], out["t"] = q, t
    out["idx"] = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
    out["dist"] = np.array([[a.distance, b.distance] for a, b in m], np.float32)
    out["accept"] = np.array([a.distance < 0.8 * b.distance for a, b in m], np.uint8)
    # real SIFT descriptors of a synthetic pair (cropped to keep the fixture small)
    s = synth.render_sequence("kitti", 3, seed=2, width=500, height=250)
    sift = cv2.SIFT_create()
    _, d0 = sift.detectAndCompute(s["frames"][0], None)
    _, d1 = sift.detectAndCompute(s["frames"][2], None)
    m = cv2.BFMatcher().knnMatch(d0, d1, k=2)
    out["sq"], out["st"] = d0.astype(np.uint8), d1.astype(np.uint8)
    out["sidx"] = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
    out["sdist"] = np.array([[a.distance, b.distance] for a, b in m], np.float32)
    out["saccept"] = np.array([a.distance < 0.8 * b.distance for a, b in m], np.uint8)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(OUT, "knn.npz"), **out)


def make_emat_pair(n, out_frac, seed, noise=0.3):
    """Two views of random landmarks under a KITTI-like forward motion, N(0,noise) px noise and
    `out_frac` gross outliers in the second view; float32 pixel arrays as the reference passes them."""
    rng = np.random.default_rng(seed)
    K = synth.K_KITTI
    Rg, _ = cv2.Rodrigues(rng.normal(0, 0.03, 3))
    tg = np.array([0.05, 0.02, -1.6]) + rng.normal(0, 0.05, 3)
    X = np.column_stack([rng.uniform(-8, 8, n), rng.uniform(-2, 1.6, n), rng.uniform(4, 80, n)])
    X2 = X @ Rg.T + tg
    p1 = (X[:, :2] / X[:, 2:3]) * [K[0, 0], K[1, 1]] + [K[0, 2], K[1, 2]]
    p2 = (X2[:, :2] / X2[:, 2:3]) * [K[0, 0], K[1, 1]] + [K[0, 2], K[1, 2]]
    p1 = p1 + rng.normal(0, noise, p1.shape)
    p2 = p2 + rng.normal(0, noise, p2.shape)
    no = int(out_frac * n)
    idx = rng.choice(n, no, replace=False)
    p2[idx] += rng.uniform(-60, 60, (no, 2))
    return p1.astype(np.float32), p2.astype(np.float32), K.copy()


EMAT_CASES = [(1912, 0.1, 0, 0.99, 1.0), (2000, 0.4, 1, 0.99, 1.0), (500, 0.6, 2, 0.99, 1.0), (3000, 0.25, 3, 0.99, 1.0),
              (200, 0.0, 4, 0.999, 1.0), (50, 0.5, 5, 0.99, 2.0), (8, 0.0, 6, 0.99, 1.0)]


def emat():
    out = dict(cases=np.array(EMAT_CASES))
    for ci, (n, of, seed, pr, thr) in enumerate(EMAT_CASES):
        p1, p2, K = make_emat_pair(n, of, seed)
        E, m = cv2.findEssentialMat(p1, p2, K, method=cv2.RANSAC, prob=pr, threshold=thr)
        out[f"c{ci}_p1"], out[f"c{ci}_p2"], out[f"c{ci}_K"], out[f"c{ci}_E"], out[f"c{ci}_mask"] = p1, p2, K, E, m
    # minimal-solver candidate sets: cv2.findEssentialMat on exactly 5 points returns all models stacked
    P1, P2, Es, Ns = [], [], [], []
    for t in range(60):
        p1, p2, K = make_emat_pair(5, 0.0, 100 + t, 0.0)
        E, _ = cv2.findEssentialMat(p1, p2, K, method=cv2.RANSAC, prob=0.99, threshold=1)
        k = 0 if E is None else E.shape[0] // 3
        Epad = np.zeros((30, 3))
        if k:
            Epad[:3 * k] = E
        P1.append(p1); P2.append(p2); Es.append(Epad); Ns.append(k)
    out.update(min_p1=np.array(P1), min_p2=np.array(P2), min_E=np.array(Es), min_n=np.array(Ns), min_K=K)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(OUT, "emat.npz"), **out)


def sift():
    """cv2.SIFT_create().detectAndCompute(img, None) (reference :35, :226-227) on three small synthetic frames: keypoints
    (x, y, size, angle, response), packed octave codes, descriptors (integer-valued, stored as uint8)."""
    out = {}
    for ci, (shape, seed, w, h) in enumerate((("kitti", 2, 416, 160), ("parking", 1, 320, 240), ("malaga", 4, 257, 193))):
        img = synth.render_sequence(shape, 1, seed=seed, width=w, height=h)["frames"][0]
        kps, des = cv2.SIFT_create().detectAndCompute(img, None)
        out[f"c{ci}_img"] = img
        out[f"c{ci}_kp"] = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in kps], np.float32)
        out[f"c{ci}_octave"] = np.array([k.octave for k in kps], np.int32)
        assert np.array_equal(des, np.round(des)) and des.min() >= 0 and des.max() <= 255
        out[f"c{ci}_desc"] = des.astype(np.uint8)
    out["n_cases"] = np.array(3)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(OUT, "sift.npz"), **out)


ALL = dict(klt_small=klt_small, pnp=pnp, gftt=gftt, knn=knn, emat=emat, sift=sift)

if __name__ == "__main__":
    names = sys.argv[1:] or list(ALL)
    for nm in names:
        ALL[nm]()
        print("wrote", nm)
