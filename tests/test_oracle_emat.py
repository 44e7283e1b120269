"""CPU oracle for cv2.findEssentialMat(RANSAC) (oracle/emat_oracle.c) vs golden cv2 vectors / live cv2."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "emat.npz"))


def _norm(p, K):
    return np.column_stack([(p[:, 0].astype(np.float64) - K[0, 2]) / K[0, 0], (p[:, 1].astype(np.float64) - K[1, 2]) / K[1, 1]])


def test_minimal_solver_candidate_sets(g):
    K = g["min_K"]
    diffs, agree = [], 0
    for p1, p2, Epad, k in zip(g["min_p1"], g["min_p2"], g["min_E"], g["min_n"]):
        mine = oracle.five_point(_norm(p1, K), _norm(p2, K))
        agree += len(mine) == k
        for m in range(int(k)):
            Ek = Epad[3 * m:3 * m + 3]
            diffs.append(min([min(np.abs(Ek - M).max(), np.abs(Ek + M).max()) for M in mine], default=9.0))
    diffs = np.array(diffs)
    assert agree >= len(g["min_n"]) - 1              # same number of real solutions
    assert np.median(diffs) < 1e-10 and np.quantile(diffs, 0.95) < 1e-6   # same candidates up to sign


def test_sampson_matches_formula(g):
    p1, p2, K, E = g["c0_p1"], g["c0_p2"], g["c0_K"], g["c0_E"]
    x1, x2 = _norm(p1, K), _norm(p2, K)
    err = oracle.sampson_errors(x1, x2, E)
    thr = 1.0 / ((K[0, 0] + K[1, 1]) / 2)
    assert np.array_equal((err <= np.float32(thr * thr)).astype(np.uint8), g["c0_mask"].ravel())


def test_full_call_golden(g):
    for ci, (n, of, seed, pr, thr) in enumerate(g["cases"]):
        E, m, run = oracle.find_essential_mat(g[f"c{ci}_p1"], g[f"c{ci}_p2"], g[f"c{ci}_K"], pr, thr, 1000)
        assert np.array_equal(m, g[f"c{ci}_mask"]), ci                     # inlier mask identical
        Ec = g[f"c{ci}_E"]
        if n >= 20:   # with a handful of points several candidates of one sample tie on the count; cv2 breaks
            assert min(np.abs(E - Ec).max(), np.abs(E + Ec).max()) < 1e-8, ci  # the tie by its root order


def test_too_few_points():
    p = np.zeros((4, 2), np.float32)
    assert oracle.find_essential_mat(p, p, np.eye(3))[0] is None


def test_live_cv2():
    cv2 = pytest.importorskip("cv2")
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden import make_emat_pair
    for n, of, seed in ((2500, 0.2, 40), (800, 0.5, 41), (4000, 0.3, 42)):
        p1, p2, K = make_emat_pair(n, of, seed)
        Ec, mc = cv2.findEssentialMat(p1, p2, K, method=cv2.RANSAC, prob=0.99, threshold=1)
        E, m, _ = oracle.find_essential_mat(p1, p2, K, 0.99, 1.0, 1000)
        assert np.array_equal(m, mc) and min(np.abs(E - Ec).max(), np.abs(E + Ec).max()) < 1e-8
