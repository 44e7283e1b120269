"""CPU oracle (oracle/klt_oracle.c) against the golden vectors and the live cv2."""
import numpy as np
import pytest

import oracle

CASES = ("w21", "w15", "w9x13")


def test_pyrdown_scharr_golden(small_pair):
    g = small_pair
    p1 = oracle.pyr_down(g["f0"])
    assert np.array_equal(p1, g["pyr1"])
    assert np.array_equal(oracle.pyr_down(p1), g["pyr2"])
    d = oracle.scharr(g["f0"])
    assert np.array_equal(d[..., 0], g["scharr_x"])
    assert np.array_equal(d[..., 1], g["scharr_y"])


@pytest.mark.parametrize("tag", CASES)
def test_klt_golden(small_pair, tag):
    g = small_pair
    ww, wh, ml, ct, cm = (int(v) for v in g[f"{tag}_cfg"])
    eps = float(g[f"{tag}_eps"][0])
    p, st, err = oracle.calc_optical_flow_pyr_lk(g["f0"], g["f1"], g["pts"], (ww, wh), ml, (ct, cm, eps))
    assert np.array_equal(st, g[f"{tag}_status"])
    ok = st.ravel() == 1
    d = np.abs(p - g[f"{tag}_next"])[ok].max(axis=1)
    assert (d <= 0.05).mean() >= 0.99 and d.max() < 0.05       # north_star tolerance
    assert np.quantile(d, 0.9) == 0.0                            # ~all points bit-exact
    assert np.abs(err - g[f"{tag}_err"])[ok].max() < 0.05


def test_pyr_levels_rule():
    # SURVEY A.1 probed level counts
    assert oracle.pyr_levels(1241, 376, (15, 15), 5) == 5
    assert oracle.pyr_levels(1241, 376, (15, 15), 10) == 5
    assert oracle.pyr_levels(1241, 376, (21, 21), 3) == 4
    assert oracle.pyr_levels(640, 480, (15, 15), 10) == 5
    assert oracle.pyr_levels(1024, 768, (15, 15), 10) == 6


def test_klt_live_cv2(kitti_pair):
    cv2 = pytest.importorskip("cv2")
    f0, f1 = kitti_pair["frames"]
    pts = cv2.goodFeaturesToTrack(f0, 400, 0.01, 8).reshape(-1, 2)
    pts = pts + np.random.default_rng(1).uniform(-0.5, 0.5, pts.shape).astype(np.float32)
    pts = np.concatenate([pts, np.float32([[0, 0], [-5, -5], [1240.4, 375.2], [1300, 100]])])
    for win, ml, crit in (((21, 21), 3, (3, 30, 0.01)), ((15, 15), 5, (3, 50, 0.01))):
        p1, s1, e1 = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None, winSize=win, maxLevel=ml, criteria=crit)
        p2, s2, e2 = oracle.calc_optical_flow_pyr_lk(f0, f1, pts, win, ml, crit)
        assert np.array_equal(s1, s2)
        ok = s1.ravel() == 1
        assert np.abs(p1 - p2)[ok].max() < 0.05
        assert np.abs(e1 - e2)[ok].max() < 0.05
