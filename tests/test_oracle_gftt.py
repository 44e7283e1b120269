"""CPU oracle for cv2.goodFeaturesToTrack (oracle/gftt_oracle.c) vs golden cv2 vectors and live cv2."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "gftt.npz"))


def test_min_eig_map_bit_equal(g):
    assert np.array_equal(oracle.min_eig_map(g["img"], 3), g["eig"])


@pytest.mark.parametrize("name", ["img", "blobs"])
def test_corner_lists_golden(g, name):
    for ci, (mc, q, md) in enumerate(g["cases"]):
        c = oracle.good_features_to_track(g[name], int(mc), q, md, 3)
        ref = g[f"{name}_c{ci}"]
        if len(ref) == 0:
            assert c is None
        else:
            assert c.shape == ref.shape and c.dtype == np.float32 and np.array_equal(c, ref), (name, ci)


def test_flat_image_returns_none():
    assert oracle.good_features_to_track(np.full((60, 80), 9, np.uint8), 10, 0.1, 5) is None


def test_live_cv2_full_size():
    cv2 = pytest.importorskip("cv2")
    from monocular_visual_odometry_va4mr_b200 import synth
    for shape, seed in (("kitti", 0), ("parking", 1), ("malaga", 2)):
        f = synth.render_sequence(shape, 1, seed=seed)["frames"][0]
        assert np.array_equal(oracle.min_eig_map(f, 3), cv2.cornerMinEigenVal(f, 3, ksize=3))
        for mc, q, md in ((1400, 0.1, 10), (1400, 0.03, 10), (2000, 0.001, 3)):
            a = cv2.goodFeaturesToTrack(f, mc, q, md, blockSize=3, useHarrisDetector=False)
            b = oracle.good_features_to_track(f, mc, q, md, 3)
            assert np.array_equal(a, b)
