"""B200 P3P-RANSAC (+EPnP refit) through the C-ABI against the golden cv2 vectors, the oracle and
live cv2.  north_star bar: inlier masks bit-exact given the same hypothesis sample set."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN
from monocular_visual_odometry_va4mr_b200 import _lib, cv2_compat

pytestmark = pytest.mark.gpu
sys.path.insert(0, GOLDEN)


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "pnp.npz"))


def _pose_close(rv, tv, rv0, tv0, tol=1e-6):
    assert np.abs(np.ravel(rv) - np.ravel(rv0)).max() < tol
    assert np.abs(np.ravel(tv) - np.ravel(tv0)).max() < tol * max(1.0, np.abs(tv0).max())


def test_full_call_vs_golden_cv2(g):
    for ci, (n, of, seed, iters, thr) in enumerate(g["cases"]):
        ok, rv, tv, inl = cv2_compat.solvePnPRansac(g[f"c{ci}_obj"], g[f"c{ci}_img"], g[f"c{ci}_K"], np.zeros(4),
                                                    flags=cv2_compat.SOLVEPNP_P3P, confidence=0.99,
                                                    reprojectionError=thr, iterationsCount=int(iters))
        assert ok == bool(g[f"c{ci}_ok"]), ci
        if not ok:
            assert inl is None
            continue
        assert inl.dtype == np.int32 and inl.shape == g[f"c{ci}_inliers"].shape, ci
        assert np.array_equal(inl, g[f"c{ci}_inliers"]), ci
        assert rv.shape == (3, 1) and tv.shape == (3, 1) and rv.dtype == np.float64
        _pose_close(rv, tv, g[f"c{ci}_rvec"], g[f"c{ci}_tvec"])


def test_same_samples_counts_and_mask_vs_oracle():
    import oracle
    from make_golden import make_pnp_case
    ctx = _lib.default_context(0)
    for n, of, seed, iters, thr in [(700, 0.4, 41, 200, 8.0), (1500, 0.15, 42, 120, 5.0)]:
        obj, img, K = make_pnp_case(n, of, seed)
        samples = oracle.ransac_subsets(n, 4, iters)
        counts = np.zeros(iters, np.int32)
        rv, tv = np.zeros(3), np.zeros(3)
        inl = np.empty(n, np.int32)
        n_in, ok, win, run = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        Kc = np.ascontiguousarray(K.reshape(9))
        rc = ctx.lib.b200vo_solve_pnp_ransac_p3p_samples(
            ctx.h, obj.ctypes.data_as(_lib.c_f32p), img.ctypes.data_as(_lib.c_f32p), n, Kc.ctypes.data_as(_lib.c_f64p),
            samples.ctypes.data_as(_lib.c_i32p), iters, thr, 0.99, rv.ctypes.data_as(_lib.c_f64p),
            tv.ctypes.data_as(_lib.c_f64p), inl.ctypes.data_as(_lib.c_i32p), C.byref(n_in), C.byref(ok), counts.ctypes.data_as(_lib.c_i32p),
            C.byref(win), C.byref(run))
        assert rc == 0, ctx.last_error()
        # oracle: per-hypothesis inlier counts for the same sample set
        ref = np.zeros(iters, np.int32)
        thr2 = np.float32(thr * thr)
        for it in range(iters):
            o, r, t = oracle.pnp_minimal(obj[samples[it]], img[samples[it]], K)
            if o:
                ref[it] = int((oracle.pnp_errors(obj, img, K, r, t) <= thr2).sum())
        assert np.array_equal(counts, ref), f"{(counts != ref).sum()} hypothesis counts differ"
        best, run_ref = oracle.ransac_select(ref, None, 1, iters, n, 4, 0.99)
        assert win.value == best and run.value == run_ref
        o, r, t = oracle.pnp_minimal(obj[samples[best]], img[samples[best]], K)
        mask_ref = oracle.pnp_errors(obj, img, K, r, t) <= thr2
        assert np.array_equal(inl[:n_in.value], np.flatnonzero(mask_ref).astype(np.int32))


def test_device_sampler_matches_cv_rng():
    """With the library drawing the samples itself the result must equal the supplied-samples run."""
    import oracle
    from make_golden import make_pnp_case
    for n in (5, 9, 64, 300, 2000):
        obj, img, K = make_pnp_case(n, 0.3, 50 + n)
        a = cv2_compat.solvePnPRansac(obj, img, K, np.zeros(4), flags=cv2_compat.SOLVEPNP_P3P, iterationsCount=300,
                                      reprojectionError=8.0, confidence=0.99)
        b = oracle.solve_pnp_ransac_p3p(obj, img, K, 300, 8.0, 0.99)
        assert a[0] == b[0]
        if a[0]:
            assert np.array_equal(a[3], b[3]), n
            if len(a[3]) >= 6:   # EPnP on < 6 points is under-determined (2n < 12): cv2's own answer then
                _pose_close(a[1], a[2], b[1], b[2])   # depends on its random null-space fill; not pinned


def test_live_cv2_large():
    cv2 = pytest.importorskip("cv2")
    from make_golden import make_pnp_case
    for n, of, seed, iters, thr in [(20000, 0.3, 4, 2000, 8.0), (2000, 0.1, 77, 500, 8.0), (1000, 0.5, 78, 500, 5.0)]:
        obj, img, K = make_pnp_case(n, of, seed)
        ok, rv, tv, inl = cv2.solvePnPRansac(obj, img, K, np.zeros(4), flags=cv2.SOLVEPNP_P3P, confidence=0.99,
                                             reprojectionError=thr, iterationsCount=iters)
        ok2, rv2, tv2, inl2 = cv2_compat.solvePnPRansac(obj, img, K, np.zeros(4), flags=cv2.SOLVEPNP_P3P, confidence=0.99,
                                                        reprojectionError=thr, iterationsCount=iters)
        assert ok == ok2 and np.array_equal(inl, inl2)
        _pose_close(rv2, tv2, rv, tv)


def test_edge_cases():
    from make_golden import make_pnp_case
    obj, img, K = make_pnp_case(50, 0.0, 3)
    with pytest.raises(cv2_compat.error):
        cv2_compat.solvePnPRansac(obj[:3], img[:3], K, np.zeros(4), flags=cv2_compat.SOLVEPNP_P3P)
    ok, rv, tv, inl = cv2_compat.solvePnPRansac(obj[:4], img[:4], K, np.zeros(4), flags=cv2_compat.SOLVEPNP_P3P)
    assert ok and inl.ravel().tolist() == [0, 1, 2, 3]
    with pytest.raises(NotImplementedError):
        cv2_compat.solvePnPRansac(obj, img, K, np.zeros(4))            # default flags = ITERATIVE
    with pytest.raises(NotImplementedError):
        cv2_compat.solvePnPRansac(obj, img, K, np.array([0.1, 0, 0, 0]), flags=cv2_compat.SOLVEPNP_P3P)
    # pure noise: no model reaches 4 inliers at a 0.01 px threshold -> (False, ..., None)
    rng = np.random.default_rng(0)
    bad = rng.uniform(0, 1000, img.shape).astype(np.float32)
    ok, _, _, inl = cv2_compat.solvePnPRansac(obj, bad, K, np.zeros(4), flags=cv2_compat.SOLVEPNP_P3P, reprojectionError=0.01,
                                              iterationsCount=50)
    assert not ok and inl is None
    # (N,1,3)/(N,1,2) inputs as cv2 accepts
    ok, rv, tv, inl = cv2_compat.solvePnPRansac(obj.reshape(-1, 1, 3), img.reshape(-1, 1, 2), K, np.zeros(4),
                                                flags=cv2_compat.SOLVEPNP_P3P, iterationsCount=100)
    assert ok and len(inl) == 50

