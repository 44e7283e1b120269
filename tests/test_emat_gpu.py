"""B200 five-point essential-matrix RANSAC through the C-ABI: inlier mask identical to cv2 (golden,
live) and to the oracle; E identical up to sign."""
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
    return np.load(os.path.join(GOLDEN, "emat.npz"))


def _same_E(E, Ec, tol=1e-8):
    return min(np.abs(E - Ec).max(), np.abs(E + Ec).max()) < tol


def _explains_mask(E, p1, p2, K, mask, thr=1.0):
    """Two candidates of one sample can be nearly identical (close roots) and tie on the inlier count;
    which of them cv2 keeps depends on its root order.  Such an E must still reproduce the mask."""
    import oracle
    n1 = np.column_stack([(p1[:, 0].astype(np.float64) - K[0, 2]) / K[0, 0], (p1[:, 1].astype(np.float64) - K[1, 2]) / K[1, 1]])
    n2 = np.column_stack([(p2[:, 0].astype(np.float64) - K[0, 2]) / K[0, 0], (p2[:, 1].astype(np.float64) - K[1, 2]) / K[1, 1]])
    t = thr / ((K[0, 0] + K[1, 1]) / 2)
    return np.array_equal((oracle.sampson_errors(n1, n2, E) <= np.float32(t * t)).astype(np.uint8), mask.ravel())


def test_vs_golden_cv2(g):
    for ci, (n, of, seed, pr, thr) in enumerate(g["cases"]):
        E, m = cv2_compat.findEssentialMat(g[f"c{ci}_p1"], g[f"c{ci}_p2"], g[f"c{ci}_K"], method=cv2_compat.RANSAC, prob=pr, threshold=thr)
        assert E.shape == (3, 3) and E.dtype == np.float64 and m.shape == (int(n), 1) and m.dtype == np.uint8
        assert np.array_equal(m, g[f"c{ci}_mask"]), (ci, int((m != g[f"c{ci}_mask"]).sum()))
        if n >= 20:   # tiny sets: candidates of one sample tie on the count, cv2's root order decides (not pinned)
            assert _same_E(E, g[f"c{ci}_E"]), ci


def test_vs_oracle_and_live_cv2():
    import oracle
    from make_golden import make_emat_pair
    for n, of, seed in ((2500, 0.2, 40), (800, 0.5, 41), (4000, 0.3, 42), (120, 0.7, 43)):
        p1, p2, K = make_emat_pair(n, of, seed)
        E, m = cv2_compat.findEssentialMat(p1, p2, K, method=cv2_compat.RANSAC, prob=0.99, threshold=1)
        Eo, mo, _ = oracle.find_essential_mat(p1, p2, K, 0.99, 1.0, 1000)
        assert np.array_equal(m, mo) and (_same_E(E, Eo, 1e-9) or _explains_mask(E, p1, p2, K, m))
        try:
            import cv2
        except ImportError:
            continue
        Ec, mc = cv2.findEssentialMat(p1, p2, K, method=cv2.RANSAC, prob=0.99, threshold=1)
        assert np.array_equal(m, mc) and (_same_E(E, Ec) or _explains_mask(E, p1, p2, K, m))


def test_reference_call_pattern_and_edges():
    from make_golden import make_emat_pair
    p1, p2, K = make_emat_pair(600, 0.2, 7)
    E, ransac_mask = cv2_compat.findEssentialMat(p1, p2, K, method=cv2_compat.RANSAC, prob=0.99, threshold=1)   # :308
    inliers = ransac_mask.ravel() == 1                                                                         # :310
    assert 350 < inliers.sum() <= 600
    assert cv2_compat.findEssentialMat(p1[:4], p2[:4], K, method=cv2_compat.RANSAC) == (None, None)
    with pytest.raises(NotImplementedError):
        cv2_compat.findEssentialMat(p1, p2, K, method=cv2_compat.LMEDS)
    with pytest.raises(cv2_compat.error):
        cv2_compat.findEssentialMat(p1, p2[:-1], K, method=cv2_compat.RANSAC)
