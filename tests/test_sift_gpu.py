"""CUDA SIFT (b200vo_sift_detect_and_compute; reference :35, :226-227) through the C ABI vs the C oracle
(same arithmetic: exact keypoint list expected), the golden cv2 vectors and live cv2 (tolerances: tests/sift_compare.py)."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from monocular_visual_odometry_va4mr_b200 import cv2_compat, synth
from sift_compare import compare

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "sift.npz"))


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_vs_golden_cv2_and_oracle(g, ci):
    img = g[f"c{ci}_img"]
    kp, octv, des = cv2_compat.sift_detect_and_compute(img)
    compare(kp, octv, des, g[f"c{ci}_kp"], g[f"c{ci}_octave"], g[f"c{ci}_desc"].astype(np.float32))
    assert len(kp) == len(g[f"c{ci}_kp"]) and np.array_equal(octv, g[f"c{ci}_octave"])
    assert np.array_equal(kp[:, :2], g[f"c{ci}_kp"][:, :2])          # count, order, octave codes, x, y exact against cv2
    okp, ooct, odes = oracle.sift_detect_and_compute(img)
    assert np.array_equal(kp, okp) and np.array_equal(octv, ooct) and np.array_equal(des, odes)   # bit for bit against the oracle


@pytest.mark.parametrize("shape,seed", [("kitti", 0), ("parking", 2), ("malaga", 1)])
def test_full_size_frames_vs_oracle(shape, seed):
    img = synth.render_sequence(shape, 1, seed=seed)["frames"][0]
    kp, octv, des = cv2_compat.sift_detect_and_compute(img)
    okp, ooct, odes = oracle.sift_detect_and_compute(img)
    assert len(kp) > 500
    assert np.array_equal(kp, okp) and np.array_equal(octv, ooct)
    assert np.array_equal(des, odes)


def test_live_cv2_and_reference_call_pattern():
    cv2 = pytest.importorskip("cv2")
    img = synth.render_sequence("parking", 1, seed=5)["frames"][0]
    ckps, cdes = cv2.SIFT_create().detectAndCompute(img, None)
    ck = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in ckps], np.float32)
    co = np.array([k.octave for k in ckps], np.int32)
    kps, des = cv2_compat.SIFT_create().detectAndCompute(img, None)          # as the reference calls it (:226)
    kp = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in kps], np.float32)
    octv = np.array([k.octave for k in kps], np.int32)
    r = compare(kp, octv, des, ck, co, cdes, min_match=0.99, min_desc_rows=0.97)   # live cv2 on whatever host CPU the box has (its SIMD dispatch decides the last ulp)
    assert des.dtype == np.float32 and des.shape == (len(kps), 128)
    # downstream use in the reference (:229-233): the matcher takes these descriptors, .pt is read per match
    idx, dist, acc = cv2_compat.knn2_ratio(des, des, 0.8)
    assert (dist[:, 0] == 0).all() and (idx[:, 0] == np.arange(len(des))).mean() > 0.9     # every descriptor finds itself (or an identical twin)
    assert np.float32([kps[3].pt]).shape == (1, 2)


def test_edge_cases():
    flat = np.full((64, 80), 128, np.uint8)
    kp, octv, des = cv2_compat.sift_detect_and_compute(flat)
    assert len(kp) == 0 and des.shape == (0, 128)
    kps, d = cv2_compat.SIFT_create().detectAndCompute(flat, None)
    assert kps == () and d is None
    tiny = np.arange(7 * 9, dtype=np.uint8).reshape(7, 9) * 3
    kp, octv, des = cv2_compat.sift_detect_and_compute(tiny)
    okp, ooct, odes = oracle.sift_detect_and_compute(tiny)
    assert len(kp) == len(okp)
    view = synth.render_sequence("kitti", 1, seed=4, width=300, height=150)["frames"][0][:, 10:250]   # strided view
    a = cv2_compat.sift_detect_and_compute(view)
    b = oracle.sift_detect_and_compute(np.ascontiguousarray(view))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    with pytest.raises(NotImplementedError):
        cv2_compat.SIFT_create(nfeatures=500)
    with pytest.raises(NotImplementedError):
        cv2_compat.SIFT_create().detectAndCompute(flat, np.ones_like(flat))
