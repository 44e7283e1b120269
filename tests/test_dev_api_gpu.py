"""the `_dev` forms of the five call sites (device pointers in, device pointers out, asynchronous on the ctx
stream) give exactly what the host forms give."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN
from monocular_visual_odometry_va4mr_b200 import _lib, cv2_compat, synth

pytestmark = pytest.mark.gpu
sys.path.insert(0, GOLDEN)


@pytest.fixture(scope="module")
def env():
    import torch
    ctx = _lib.default_context(0)
    dev = torch.device("cuda", 0)

    def up(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def done():
        ctx.sync()
        torch.cuda.synchronize()
    return ctx, torch, dev, up, done


def _K(K):
    return np.ascontiguousarray(K, np.float64).reshape(9).ctypes.data_as(_lib.c_f64p)


def test_klt_dev_equals_host_form(env):
    ctx, torch, dev, up, done = env
    s = synth.render_sequence("kitti", 2, seed=3, width=640, height=240)
    f0, f1 = s["frames"]
    pts = synth.grid_corners(f0, 700, seed=1)
    want = cv2_compat.calcOpticalFlowPyrLK(f0, f1, pts, None, winSize=(15, 15), maxLevel=5, criteria=(3, 50, 0.01))
    d0, d1, dp = up(f0), up(f1), up(pts)
    n = len(pts)
    o_p = torch.zeros((n, 2), dtype=torch.float32, device=dev)
    o_s = torch.zeros((n,), dtype=torch.uint8, device=dev)
    o_e = torch.zeros((n,), dtype=torch.float32, device=dev)
    rc = ctx.lib.b200vo_calc_optical_flow_pyr_lk_dev(ctx.h, d0.data_ptr(), d1.data_ptr(), 240, 640, 640, 640, dp.data_ptr(), n, 15, 15, 5,
                                                     3, 50, 0.01, 0, 1e-4, o_p.data_ptr(), o_s.data_ptr(), o_e.data_ptr())
    assert rc == 0, ctx.last_error()
    done()
    assert np.array_equal(o_s.cpu().numpy(), want[1].ravel())
    ok = want[1].ravel() == 1
    assert np.array_equal(o_p.cpu().numpy()[ok], want[0][ok]) and np.array_equal(o_e.cpu().numpy()[ok], want[2].ravel()[ok])
    # pitched device image (a view into a wider buffer): same result
    wide = torch.zeros((240, 768), dtype=torch.uint8, device=dev)
    wide[:, :640] = d1
    o_p.zero_(); o_s.zero_()
    rc = ctx.lib.b200vo_calc_optical_flow_pyr_lk_dev(ctx.h, d0.data_ptr(), wide.data_ptr(), 240, 640, 640, 768, dp.data_ptr(), n, 15, 15, 5,
                                                     3, 50, 0.01, 0, 1e-4, o_p.data_ptr(), o_s.data_ptr(), o_e.data_ptr())
    assert rc == 0, ctx.last_error()
    done()
    assert np.array_equal(o_s.cpu().numpy(), want[1].ravel()) and np.array_equal(o_p.cpu().numpy()[ok], want[0][ok])


def test_gftt_dev_equals_host_form(env):
    ctx, torch, dev, up, done = env
    img = synth.render_sequence("kitti", 1, seed=6)["frames"][0]
    want = cv2_compat.goodFeaturesToTrack(img, 1400, 0.1, 10, blockSize=3)
    h, w = img.shape
    d = up(img)
    o_c = torch.full((1400, 2), -1.0, dtype=torch.float32, device=dev)
    o_n = torch.zeros((1,), dtype=torch.int32, device=dev)
    rc = ctx.lib.b200vo_good_features_to_track_dev(ctx.h, d.data_ptr(), h, w, w, 1400, 0.1, 10.0, 3, o_c.data_ptr(), o_n.data_ptr())
    assert rc == 0, ctx.last_error()
    done()
    n = int(o_n.item())
    assert n == len(want) and np.array_equal(o_c.cpu().numpy()[:n], want.reshape(-1, 2))


def test_knn_dev_equals_host_form(env):
    ctx, torch, dev, up, done = env
    from make_golden import sift_like
    q, t = sift_like(1500, 5), sift_like(1111, 6)
    wi, wd, wa = cv2_compat.knn2_ratio(q, t, 0.8)
    dq, dt = up(q), up(t)
    o_i = torch.zeros((1500, 2), dtype=torch.int32, device=dev)
    o_d = torch.zeros((1500, 2), dtype=torch.float32, device=dev)
    o_a = torch.zeros((1500,), dtype=torch.uint8, device=dev)
    bad = torch.ones((1,), dtype=torch.int32, device=dev)
    rc = ctx.lib.b200vo_knn2_ratio_dev(ctx.h, dq.data_ptr(), 1500, dt.data_ptr(), 1111, 128, 0.8, o_i.data_ptr(), o_d.data_ptr(),
                                       o_a.data_ptr(), bad.data_ptr())
    assert rc == 0, ctx.last_error()
    done()
    assert int(bad.item()) == 0
    assert np.array_equal(o_i.cpu().numpy(), wi) and np.array_equal(o_d.cpu().numpy(), wd) and np.array_equal(o_a.cpu().numpy(), wa)
    dq2 = up(q + 0.25)        # not integer-valued: flagged, not silently wrong
    rc = ctx.lib.b200vo_knn2_ratio_dev(ctx.h, dq2.data_ptr(), 1500, dt.data_ptr(), 1111, 128, 0.8, o_i.data_ptr(), o_d.data_ptr(),
                                       o_a.data_ptr(), bad.data_ptr())
    done()
    assert rc == 0 and int(bad.item()) == 1


def test_emat_dev_equals_host_form(env):
    ctx, torch, dev, up, done = env
    from make_golden import make_emat_pair
    p1, p2, K = make_emat_pair(1800, 0.3, 61)
    E, m = cv2_compat.findEssentialMat(p1, p2, K, method=cv2_compat.RANSAC, prob=0.99, threshold=1)
    d1, d2 = up(p1), up(p2)
    o_E = torch.zeros((9,), dtype=torch.float64, device=dev)
    o_m = torch.zeros((1800,), dtype=torch.uint8, device=dev)
    o_f = torch.zeros((1,), dtype=torch.int32, device=dev)
    rc = ctx.lib.b200vo_find_essential_mat_ransac_dev(ctx.h, d1.data_ptr(), d2.data_ptr(), 1800, _K(K), 0.99, 1.0, 1000, o_E.data_ptr(),
                                                      o_m.data_ptr(), o_f.data_ptr())
    assert rc == 0, ctx.last_error()
    done()
    assert int(o_f.item()) == 1 and np.array_equal(o_m.cpu().numpy(), m.ravel()) and np.array_equal(o_E.cpu().numpy().reshape(3, 3), E)


def test_pnp_dev_equals_host_form(env):
    ctx, torch, dev, up, done = env
    g = np.load(os.path.join(GOLDEN, "pnp.npz"))
    ci = 0
    obj, img, K = g[f"c{ci}_obj"], g[f"c{ci}_img"], g[f"c{ci}_K"]
    n = len(obj)
    ok, rv, tv, inl = cv2_compat.solvePnPRansac(obj, img, K, np.zeros(4), flags=cv2_compat.SOLVEPNP_P3P, confidence=0.99,
                                                reprojectionError=8.0, iterationsCount=500)
    do, di = up(obj.astype(np.float32)), up(img.astype(np.float32))
    o_pose = torch.zeros((6,), dtype=torch.float64, device=dev)
    o_inl = torch.zeros((n,), dtype=torch.int32, device=dev)
    o_n = torch.zeros((1,), dtype=torch.int32, device=dev)
    o_ok = torch.zeros((1,), dtype=torch.uint8, device=dev)
    rc = ctx.lib.b200vo_solve_pnp_ransac_p3p_dev(ctx.h, do.data_ptr(), di.data_ptr(), n, _K(K), 500, C.c_float(8.0), 0.99, o_pose.data_ptr(),
                                                 o_inl.data_ptr(), o_n.data_ptr(), o_ok.data_ptr())
    assert rc == 0, ctx.last_error()
    done()
    assert bool(o_ok.item()) == ok
    m = int(o_n.item())
    assert m == len(inl) and np.array_equal(o_inl.cpu().numpy()[:m], inl.ravel())
    assert np.array_equal(o_pose.cpu().numpy(), np.concatenate([rv.ravel(), tv.ravel()]))
