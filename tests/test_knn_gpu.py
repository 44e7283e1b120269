"""tcgen05/TMEM/TMA brute-force kNN(k=2) + ratio test through the C-ABI: indices, distances and
accept flags bit-exact against cv2 (golden + live) and the oracle."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN
from monocular_visual_odometry_va4mr_b200 import cv2_compat

pytestmark = pytest.mark.gpu
sys.path.insert(0, GOLDEN)


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "knn.npz"))


def test_planted_cases_vs_golden(g):
    idx, dist, acc = cv2_compat.knn2_ratio(g["q"], g["t"], 0.8)
    assert np.array_equal(idx, g["idx"]), int((idx != g["idx"]).any(1).sum())
    assert np.array_equal(dist, g["dist"])
    assert np.array_equal(acc, g["accept"])


def test_real_sift_vs_golden(g):
    idx, dist, acc = cv2_compat.knn2_ratio(g["sq"].astype(np.float32), g["st"].astype(np.float32), 0.8)
    assert np.array_equal(idx, g["sidx"]) and np.array_equal(dist, g["sdist"]) and np.array_equal(acc, g["saccept"])


@pytest.mark.parametrize("nq,nt", [(1, 2), (5, 1), (127, 255), (128, 256), (129, 257), (1000, 3000), (4739, 4705)])
def test_ragged_sizes_vs_oracle(nq, nt):
    import oracle
    from make_golden import sift_like
    q, t = sift_like(nq, nq), sift_like(nt, nt + 1)
    if nt > 40 and nq > 3:
        t[nt - 1] = q[0]; t[3] = q[0]; t[nt // 2] = q[1]
    a = oracle.knn2_ratio(q, t, 0.8)
    b = cv2_compat.knn2_ratio(q, t, 0.8)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_dmatch_api_matches_reference_usage(g):
    matcher = cv2_compat.BFMatcher()
    matches = matcher.knnMatch(g["q"], g["t"], k=2)
    assert isinstance(matches, tuple) and len(matches) == len(g["q"]) and len(matches[0]) == 2
    good = [m for m, n in matches if m.distance < 0.8 * n.distance]          # the reference's ratio_test (:218-224)
    assert [m.queryIdx for m in good] == np.flatnonzero(g["accept"]).tolist()
    assert [m.trainIdx for m in good] == g["idx"][g["accept"] == 1, 0].tolist()


def test_full_size_properties():
    """BASELINE config 3 size (8192 x 8192): checked through size-independent properties."""
    from make_golden import sift_like
    q = sift_like(8192, 11)
    perm = np.random.default_rng(3).permutation(8192)
    t = q[perm].copy()                       # every query has an exact duplicate somewhere in the train set
    idx, dist, acc = cv2_compat.knn2_ratio(q, t, 0.8)
    inv = np.empty(8192, np.int64); inv[perm] = np.arange(8192)
    assert np.array_equal(idx[:, 0], inv) and np.all(dist[:, 0] == 0) and acc.all()
    assert np.all(dist[:, 1] > 0) and np.all(idx[:, 1] != idx[:, 0])
    # second neighbour agrees with the exact float64 numpy computation on a sample of rows
    rows = np.arange(0, 8192, 257)
    d2 = ((q[rows, None, :].astype(np.float64) - t[None, :, :]) ** 2).sum(-1)
    d2[np.arange(len(rows)), inv[rows]] = np.inf
    assert np.array_equal(idx[rows, 1], d2.argmin(1))
    assert np.array_equal(dist[rows, 1], np.sqrt(d2.min(1).astype(np.float32)))


def test_full_size_8192_vs_live_cv2_every_row():
    """BASELINE config 3 as stated: 8192 x 8192 descriptors, ALL rows against cv2's own BFMatcher (79 ms on the host)
    and against the C oracle: indices, distances and accept flags bit-exact."""
    import oracle
    from make_golden import sift_like
    q, t = sift_like(8192, 21), sift_like(8192, 22)
    t[100] = q[5]; t[4000] = q[5]            # planted ties: the lower train index must win
    t[8191] = q[8191]
    idx, dist, acc = cv2_compat.knn2_ratio(q, t, 0.8)
    oi, od, oa = oracle.knn2_ratio(q, t, 0.8)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od) and np.array_equal(acc, oa)
    assert tuple(idx[5]) == (100, 4000) and idx[8191, 0] == 8191
    cv2 = pytest.importorskip("cv2")
    m = cv2.BFMatcher().knnMatch(q, t, k=2)
    ci = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
    cd = np.array([[a.distance, b.distance] for a, b in m], np.float32)
    ca = np.array([a.distance < 0.8 * b.distance for a, b in m], np.uint8)      # the reference's ratio loop (:218-224)
    assert np.array_equal(idx, ci) and np.array_equal(dist, cd) and np.array_equal(acc, ca)


def test_errors():
    from make_golden import sift_like
    q = sift_like(10, 1)
    with pytest.raises(cv2_compat.error):
        cv2_compat.knn2_ratio(q, q[:, :64], 0.8)
    with pytest.raises(NotImplementedError):
        cv2_compat.knn2_ratio(q + 0.5, q, 0.8)          # not integer-valued: fp16 contraction would not be exact
    with pytest.raises(NotImplementedError):
        cv2_compat.BFMatcher(crossCheck=True)
