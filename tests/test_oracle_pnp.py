"""CPU oracle for the PnP-RANSAC path (oracle/ransac_oracle.c, oracle/pnp_oracle.c) against the
golden cv2 vectors (tests/golden/pnp.npz) and the live cv2 when importable."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "pnp.npz"))


def test_rng_known_answers():
    # cv::RNG((uint64)-1) + getSubset
    assert oracle.ransac_subsets(2000, 4, 1)[0].tolist() == [1605, 1004, 940, 1173]
    assert oracle.ransac_subsets(1000, 4, 1)[0].tolist() == [605, 4, 940, 173]
    assert oracle.ransac_subsets(20000, 4, 1)[0].tolist() == [3605, 19004, 8940, 7173]
    s = oracle.ransac_subsets(9, 5, 400)          # tiny count: duplicate rejection is exercised
    assert all(len(set(r)) == 5 for r in s.tolist())


def test_update_num_iters():
    import math
    for p, ep, m, mx in [(0.99, 0.1, 4, 500), (0.99, 0.5, 4, 500), (0.99, 0.9, 4, 500), (0.99, 0.3, 5, 1000), (0.999, 0.6, 5, 1000)]:
        want = min(mx, round(math.log(1 - p) / math.log(1 - (1 - ep) ** m)))
        assert oracle.ransac_update_num_iters(p, ep, m, mx) == want
    assert oracle.ransac_update_num_iters(0.99, 0.0, 4, 500) == 0          # denom < DBL_MIN
    assert oracle.ransac_update_num_iters(0.99, 1.0, 4, 500) == 500        # log(1) = 0 -> maxIters


def test_select_replays_sequential_loop():
    counts = np.array([3, 10, 10, 50, 49, 900, 900, 901] + [0] * 92, np.int32)
    best, run = oracle.ransac_select(counts, None, 1, 100, 1000, 4, 0.99)
    assert best == 5 and run == 6   # 900/1000 inliers: niters shrinks to 4, the loop ends after iteration 5


def test_minimal_solver_golden(g):
    obj, img, K, S = g["min_obj"], g["min_img"], g["min_K"], g["min_samples"]
    n_ok = 0
    for k, smp in enumerate(S):
        ok, rv, tv = oracle.pnp_minimal(obj[smp], img[smp], K)
        assert ok == bool(g["min_ok"][k])
        if ok:
            n_ok += 1
            assert np.abs(rv - g["min_rvec"][k]).max() < 1e-7
            assert np.abs(tv - g["min_tvec"][k]).max() < 1e-6 * max(1.0, np.abs(tv).max())
    assert n_ok > 250


def test_reprojection_errors_bit_exact(g):
    err = oracle.pnp_errors(g["min_obj"], g["min_img"], g["min_K"], g["err_rvec"], g["err_tvec"])
    assert np.array_equal(err, g["err_vals"])


@pytest.mark.parametrize("ei", range(5))
def test_epnp_golden(g, ei):
    ok, R, t = oracle.epnp(g[f"e{ei}_obj"].astype(np.float64), g[f"e{ei}_img"].astype(np.float64), g[f"e{ei}_K"])
    assert ok
    rv = oracle.R_to_rodrigues(R)
    assert np.abs(rv - g[f"e{ei}_rvec"].ravel()).max() < 1e-9
    assert np.abs(t - g[f"e{ei}_tvec"].ravel()).max() < 1e-9 * max(1.0, np.abs(t).max())


def test_full_call_golden(g):
    for ci, (n, of, seed, iters, thr) in enumerate(g["cases"]):
        ok, rv, tv, inl, run = oracle.solve_pnp_ransac_p3p(g[f"c{ci}_obj"], g[f"c{ci}_img"], g[f"c{ci}_K"], int(iters), thr, 0.99)
        assert ok == bool(g[f"c{ci}_ok"]), ci
        if ok:
            assert np.array_equal(inl, g[f"c{ci}_inliers"]), ci            # inlier set identical
            assert np.abs(rv - g[f"c{ci}_rvec"]).max() < 1e-6, ci
            assert np.abs(tv - g[f"c{ci}_tvec"]).max() < 1e-6 * max(1.0, np.abs(tv).max()), ci


def test_rodrigues_roundtrip():
    rng = np.random.default_rng(0)
    for _ in range(50):
        rv = rng.normal(0, 1, 3)
        assert np.abs(oracle.R_to_rodrigues(oracle.rodrigues_to_R(rv)) - rv).max() < 1e-12


def test_full_call_live_cv2():
    cv2 = pytest.importorskip("cv2")
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden import make_pnp_case
    for n, of, seed, iters, thr in [(1500, 0.2, 31, 500, 8.0), (600, 0.55, 32, 500, 5.0), (20000, 0.6, 4, 2000, 8.0)]:
        obj, img, K = make_pnp_case(n, of, seed)
        ok, rv, tv, inl = cv2.solvePnPRansac(obj, img, K, np.zeros(4), flags=cv2.SOLVEPNP_P3P, confidence=0.99,
                                             reprojectionError=thr, iterationsCount=iters)
        ok2, rv2, tv2, inl2, _ = oracle.solve_pnp_ransac_p3p(obj, img, K, iters, thr, 0.99)
        assert ok == ok2 and np.array_equal(inl, inl2)
        assert np.abs(rv - rv2).max() < 1e-6 and np.abs(tv - tv2).max() < 1e-6 * max(1, np.abs(tv).max())
