"""B200 tracker through the C-ABI against the oracle, the golden cv2 vectors and live cv2.

north_star bar: status flags identical; positions within 0.05 px for >= 99 % of points."""
import ctypes as C

import numpy as np
import pytest

from monocular_visual_odometry_va4mr_b200 import _lib, cv2_compat, synth

pytestmark = pytest.mark.gpu
CASES = ("w21", "w15", "w9x13")


def _check(p, st, err, rp, rst, rerr, exact_frac=0.5):
    assert st.shape == rst.shape and st.dtype == np.uint8
    assert np.array_equal(st, rst), f"{(st != rst).sum()} status flags differ"
    ok = rst.ravel() == 1
    d = np.abs(p - rp)[ok].max(axis=1)
    assert (d <= 0.05).mean() >= 0.99, f"only {(d <= 0.05).mean():.4f} within 0.05 px (max {d.max()})"
    # err (mean |I - J| over the window, tens of grey levels per pixel of shift on texture) is compared where the
    # two trackers stopped at the very same position
    close = np.zeros(len(ok), bool)
    close[np.flatnonzero(ok)[d == 0]] = True
    assert close.sum() >= 0.5 * ok.sum()
    assert np.abs(err - rerr)[close].max() < 0.05
    return d


def test_pyramid_levels_bit_exact(ctx, small_pair):
    import oracle
    g = small_pair
    f0 = np.ascontiguousarray(g["f0"])
    rc = ctx.lib.b200vo_frame_upload(ctx.h, 0, f0.ctypes.data_as(_lib.c_u8p), f0.shape[0], f0.shape[1], f0.shape[1], 15, 15, 4)
    assert rc == 0, ctx.last_error()
    w, h, nl = C.c_int(), C.c_int(), C.c_int()
    ref = f0
    ctx.lib.b200vo_frame_download_level(ctx.h, 0, 0, None, C.byref(w), C.byref(h), C.byref(nl))
    assert nl.value == oracle.pyr_levels(f0.shape[1], f0.shape[0], (15, 15), 4)
    for lvl in range(nl.value):
        ctx.lib.b200vo_frame_download_level(ctx.h, 0, lvl, None, C.byref(w), C.byref(h), C.byref(nl))
        out = np.empty((h.value, w.value), np.uint8)
        rc = ctx.lib.b200vo_frame_download_level(ctx.h, 0, lvl, out.ctypes.data_as(_lib.c_u8p), C.byref(w), C.byref(h), C.byref(nl))
        assert rc == 0
        assert np.array_equal(out, ref), f"level {lvl}"
        ref = oracle.pyr_down(ref)
    assert np.array_equal(oracle.pyr_down(f0), g["pyr1"])


@pytest.mark.parametrize("tag", CASES)
def test_klt_vs_golden_cv2(small_pair, tag):
    g = small_pair
    ww, wh, ml, ct, cm = (int(v) for v in g[f"{tag}_cfg"])
    eps = float(g[f"{tag}_eps"][0])
    p, st, err = cv2_compat.calcOpticalFlowPyrLK(g["f0"], g["f1"], g["pts"], None, winSize=(ww, wh), maxLevel=ml,
                                                 criteria=(ct, cm, eps))
    assert p.shape == g["pts"].shape and p.dtype == np.float32 and err.shape == (len(p), 1)
    _check(p, st, err, g[f"{tag}_next"], g[f"{tag}_status"], g[f"{tag}_err"])


@pytest.mark.parametrize("win,ml,crit", [((21, 21), 3, (3, 30, 0.01)), ((15, 15), 5, (3, 50, 0.01)),
                                          ((15, 15), 10, (3, 50, 0.02)), ((21, 21), 0, (3, 30, 0.01)),
                                          ((9, 13), 2, (1, 10, 0.01)), ((31, 31), 3, (2, 30, 0.03))])
def test_klt_vs_oracle_kitti(kitti_pair, win, ml, crit):
    import oracle
    f0, f1 = kitti_pair["frames"]
    pts = synth.grid_corners(f0, 2000, seed=0)
    edge = np.float32([[0, 0], [-5, -5], [1240.4, 375.2], [1300, 100], [3.2, 370.9], [620, -30], [1240, 0]])
    pts = np.ascontiguousarray(np.concatenate([pts, edge]))
    p, st, err = cv2_compat.calcOpticalFlowPyrLK(f0, f1, pts, None, winSize=win, maxLevel=ml, criteria=crit)
    rp, rst, rerr = oracle.calc_optical_flow_pyr_lk(f0, f1, pts, win, ml, crit)
    d = _check(p, st, err, rp, rst, rerr)
    # integer sums are exact on both sides: the CUDA path must equal the oracle bit for bit
    assert d.max() == 0.0, f"max |delta| vs oracle {d.max()}"
    assert np.array_equal(err[rst == 1], rerr[rst == 1])
    assert st.mean() > 0.9


def test_klt_live_cv2(kitti_pair):
    cv2 = pytest.importorskip("cv2")
    f0, f1 = kitti_pair["frames"]
    pts = synth.grid_corners(f0, 2000, seed=1)
    for win, ml, crit in (((21, 21), 3, (3, 30, 0.01)), ((15, 15), 5, (3, 50, 0.01))):
        rp, rst, rerr = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None, winSize=win, maxLevel=ml, criteria=crit)
        p, st, err = cv2_compat.calcOpticalFlowPyrLK(f0, f1, pts, None, winSize=win, maxLevel=ml, criteria=crit)
        _check(p, st, err, rp, rst, rerr)


def test_klt_shapes_and_errors(small_pair):
    g = small_pair
    f0, f1, pts = g["f0"], g["f1"], g["pts"]
    p3, st3, _ = cv2_compat.calcOpticalFlowPyrLK(f0, f1, pts.reshape(-1, 1, 2), None)
    p2, st2, _ = cv2_compat.calcOpticalFlowPyrLK(f0, f1, pts, None)
    assert p3.shape == (len(pts), 1, 2) and st3.shape == (len(pts), 1)
    assert np.array_equal(p3.reshape(-1, 2), p2) and np.array_equal(st3, st2)
    assert cv2_compat.calcOpticalFlowPyrLK(f0, f1, np.zeros((0, 2), np.float32), None) == (None, None, None)
    with pytest.raises(cv2_compat.error):
        cv2_compat.calcOpticalFlowPyrLK(f0, f1, pts.astype(np.float64), None)
    with pytest.raises(cv2_compat.error):
        cv2_compat.calcOpticalFlowPyrLK(f0, f1[:-1], pts, None)
    with pytest.raises(cv2_compat.error):
        cv2_compat.calcOpticalFlowPyrLK(f0, f1, pts, None, winSize=(2, 2))
    with pytest.raises(NotImplementedError):
        cv2_compat.calcOpticalFlowPyrLK(f0, f1, pts, None, flags=4)
    # non-contiguous view (cv2 accepts it via step)
    big = np.zeros((f0.shape[0], f0.shape[1] + 40), np.uint8)
    big[:, 7:7 + f0.shape[1]] = f0
    pv, stv, _ = cv2_compat.calcOpticalFlowPyrLK(big[:, 7:7 + f0.shape[1]], f1, pts, None)
    assert np.array_equal(pv, p2) and np.array_equal(stv, st2)


def test_klt_flat_image_rejected():
    flat = np.full((120, 160), 77, np.uint8)
    pts = np.float32([[40, 40], [80, 60.5]])
    p, st, err = cv2_compat.calcOpticalFlowPyrLK(flat, flat, pts, None)
    assert not st.any()
    assert np.array_equal(p, pts)  # minEig rejection at every level: nextPts == prevPts


def test_identity_cache_same_results(kitti_pair):
    f0, f1 = kitti_pair["frames"]
    pts = synth.grid_corners(f0, 500, seed=2)
    ref = cv2_compat.calcOpticalFlowPyrLK(f0, f1, pts, None, winSize=(15, 15), maxLevel=5, criteria=(3, 50, 0.01))
    cv2_compat.set_frame_cache("identity")
    try:
        for _ in range(2):
            out = cv2_compat.calcOpticalFlowPyrLK(f0, f1, pts, None, winSize=(15, 15), maxLevel=5, criteria=(3, 50, 0.01))
            for a, b in zip(out, ref):
                assert np.array_equal(a, b)
    finally:
        cv2_compat.set_frame_cache("strict")


@pytest.mark.parametrize("shape", ["parking", "malaga"])
def test_klt_other_dataset_shapes_vs_oracle_and_cv2(shape):
    """BASELINE configs 2/3 geometry (640x480: 5 pyramid levels, 1024x768: 6) with the reference's own options."""
    import oracle
    from monocular_visual_odometry_va4mr_b200 import workload
    o = workload.REFERENCE_OPTIONS[shape]
    s = synth.render_sequence(shape, 2, seed=4)
    f0, f1 = s["frames"]
    pts = synth.grid_corners(f0, 1000, seed=2)
    p, st, err = cv2_compat.calcOpticalFlowPyrLK(f0, f1, pts, None, winSize=o["win"], maxLevel=o["max_level"], criteria=o["criteria"])
    rp, rst, rerr = oracle.calc_optical_flow_pyr_lk(f0, f1, pts, o["win"], o["max_level"], o["criteria"])
    d = _check(p, st, err, rp, rst, rerr)
    assert d.max() == 0.0 and np.array_equal(err[rst == 1], rerr[rst == 1])
    try:
        import cv2
    except ImportError:
        return
    cp, cst, cerr = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None, winSize=o["win"], maxLevel=o["max_level"], criteria=o["criteria"])
    _check(p, st, err, cp, cst, cerr)


def test_klt_stress_config4_vs_live_cv2(kitti_pair):
    """BASELINE config 4 tracker part: 20 000 points, 21x21, maxLevel 3 at 1241x376."""
    cv2 = pytest.importorskip("cv2")
    import oracle
    f0, f1 = kitti_pair["frames"]
    pts = synth.grid_corners(f0, 20000, seed=0)
    kw = dict(winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01))
    p, st, err = cv2_compat.calcOpticalFlowPyrLK(f0, f1, pts, None, **kw)
    cp, cst, cerr = cv2.calcOpticalFlowPyrLK(f0, f1, pts, None, **kw)
    _check(p, st, err, cp, cst, cerr)
    sub = np.ascontiguousarray(pts[::10])                     # the oracle on every tenth point: bit-equal
    rp, rst, rerr = oracle.calc_optical_flow_pyr_lk(f0, f1, sub, (21, 21), 3, (3, 30, 0.01))
    assert np.array_equal(st[::10], rst) and np.array_equal(p[::10][rst.ravel() == 1], rp[rst.ravel() == 1])


def test_klt_negative_bilinear_weight(kitti_pair):
    """iw11 = 2^14 - iw00 - iw01 - iw10 is -1 when the three rounded weights add up to 2^14 + 1 (fractional
    parts of ~3e-5); cv2 carries the -1 through its signed integer arithmetic.  Regression for the packed
    16-bit-weight path (the weights must be treated as signed)."""
    import oracle
    f0, f1 = kitti_pair["frames"]
    s = np.float32(16384)
    one = np.float32(1)
    pts = []
    for base in range(30, 340):                       # x == y: both fractional parts are the same tiny value
        x = np.float32(base)
        for _ in range(12):
            x = np.nextafter(x, np.float32(1e9))
            a = np.float32(x - np.float32(7)) - np.floor(np.float32(x - np.float32(7)))
            w = 16384 - int(np.rint((one - a) * (one - a) * s)) - 2 * int(np.rint(a * (one - a) * s))
            if w < 0:
                pts.append((x, x))
                break
    pts = np.ascontiguousarray(np.float32(pts))
    assert len(pts) >= 50, len(pts)
    for win, ml in (((15, 15), 0), ((21, 21), 0), ((15, 15), 3)):
        p, st, err = cv2_compat.calcOpticalFlowPyrLK(f0, f1, pts, None, winSize=win, maxLevel=ml, criteria=(3, 30, 0.01))
        rp, rst, rerr = oracle.calc_optical_flow_pyr_lk(f0, f1, pts, win, ml, (3, 30, 0.01))
        assert np.array_equal(st, rst)
        ok = rst.ravel() == 1
        assert np.array_equal(p[ok], rp[ok]) and np.array_equal(err[rst == 1], rerr[rst == 1])
