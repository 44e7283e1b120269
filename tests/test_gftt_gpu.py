"""B200 Shi-Tomasi detector through the C-ABI: ordered corner list identical to cv2 / the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from monocular_visual_odometry_va4mr_b200 import cv2_compat, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "gftt.npz"))


@pytest.mark.parametrize("name", ["img", "blobs"])
def test_vs_golden_cv2(g, name):
    for ci, (mc, q, md) in enumerate(g["cases"]):
        c = cv2_compat.goodFeaturesToTrack(g[name], int(mc), q, md, blockSize=3, useHarrisDetector=False, mask=None)
        ref = g[f"{name}_c{ci}"]
        if len(ref) == 0:
            assert c is None
        else:
            assert c.shape == ref.shape and c.dtype == np.float32
            assert np.array_equal(c, ref), (name, ci, int((c != ref).any(axis=(1, 2)).sum()))


def test_vs_oracle_full_size():
    import oracle
    for shape, seed in (("kitti", 0), ("parking", 1), ("malaga", 2)):
        f = synth.render_sequence(shape, 1, seed=seed)["frames"][0]
        for mc, q, md in ((1400, 0.1, 10), (1400, 0.03, 10), (2000, 0.001, 3), (0, 0.3, 12.5), (100000, 0.0001, 0)):
            a = oracle.good_features_to_track(f, mc, q, md, 3)
            b = cv2_compat.goodFeaturesToTrack(f, mc, q, md, blockSize=3)
            assert (a is None) == (b is None)
            if a is not None:
                assert a.shape == b.shape and np.array_equal(a, b), (shape, mc, q, md)


def test_live_cv2():
    cv2 = pytest.importorskip("cv2")
    for shape, seed in (("kitti", 5), ("parking", 6)):
        f = synth.render_sequence(shape, 1, seed=seed)["frames"][0]
        a = cv2.goodFeaturesToTrack(f, maxCorners=1400, qualityLevel=0.1, minDistance=10, blockSize=3, useHarrisDetector=False, mask=None)
        b = cv2_compat.goodFeaturesToTrack(f, maxCorners=1400, qualityLevel=0.1, minDistance=10, blockSize=3, useHarrisDetector=False, mask=None)
        assert np.array_equal(a, b)
        assert np.array_equal(a.squeeze(), b.squeeze())      # the reference squeezes immediately (:256)


def test_edge_cases():
    import oracle
    assert cv2_compat.goodFeaturesToTrack(np.full((60, 80), 9, np.uint8), 10, 0.1, 5) is None
    with pytest.raises(cv2_compat.error):
        cv2_compat.goodFeaturesToTrack(np.zeros((60, 80), np.uint8), 10, 0.0, 5)
    with pytest.raises(NotImplementedError):
        cv2_compat.goodFeaturesToTrack(np.zeros((60, 80), np.uint8), 10, 0.1, 5, useHarrisDetector=True)
    with pytest.raises(NotImplementedError):
        cv2_compat.goodFeaturesToTrack(np.zeros((60, 80), np.uint8), 10, 0.1, 5, mask=np.ones((60, 80), np.uint8))
    # dense checkerboard of tied maxima: > 32768 candidates -> bitonic path; no min distance
    img = np.zeros((300, 400), np.uint8)
    img[::2, ::2] = 255
    c = cv2_compat.goodFeaturesToTrack(img, 0, 0.5, 0)
    assert np.array_equal(c, oracle.good_features_to_track(img, 0, 0.5, 0, 3))
