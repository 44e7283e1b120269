"""CPU oracle for BFMatcher.knnMatch(k=2) + ratio test vs golden cv2 vectors and live cv2."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "knn.npz"))


def test_planted_cases_golden(g):
    idx, dist, acc = oracle.knn2_ratio(g["q"], g["t"], 0.8)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"]) and np.array_equal(acc, g["accept"])
    assert idx[7].tolist() == [50, 100] and idx[8].tolist() == [10, 800]      # ties: lower train index first
    assert dist[20, 0] == 0.0 and acc.sum() >= 40


def test_real_sift_descriptors_golden(g):
    idx, dist, acc = oracle.knn2_ratio(g["sq"].astype(np.float32), g["st"].astype(np.float32), 0.8)
    assert np.array_equal(idx, g["sidx"]) and np.array_equal(dist, g["sdist"]) and np.array_equal(acc, g["saccept"])


def test_short_train_sets():
    q = np.float32(np.arange(256).reshape(2, 128) % 200)
    idx, dist, acc = oracle.knn2_ratio(q, q[:1], 0.8)
    assert idx.tolist() == [[0, -1], [0, -1]] and not acc.any()


def test_live_cv2():
    cv2 = pytest.importorskip("cv2")
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden import sift_like
    q, t = sift_like(1500, 5), sift_like(1300, 6)
    m = cv2.BFMatcher().knnMatch(q, t, k=2)
    idx, dist, acc = oracle.knn2_ratio(q, t, 0.8)
    assert np.array_equal(idx, np.array([[a.trainIdx, b.trainIdx] for a, b in m]))
    assert np.array_equal(dist, np.array([[a.distance, b.distance] for a, b in m], np.float32))
