#!/usr/bin/env python
"""Where the fused pose-chain kernel (pnp_fused_kernel) spends its time: phase stamps (clock64 of thread 0) on a
bench-shaped problem (~900 tracked landmarks of one sequence, 10 % outliers).  python benchmarks/pose_phases.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

NAMES = ["compaction", "ransac chunks (subsets + 32 P3P + scoring + replay)", "winner mask + inlier list",
         "epnp: centroid, covariance, 3x3 SVD, control points", "epnp: barycentric pass + M^T M", "epnp: 12x12 Jacobi",
         "epnp: betas + Gauss-Newton + 3 poses", "epnp: reprojection errors + choice"]


def main():
    import torch
    from monocular_visual_odometry_va4mr_b200 import _lib
    args = bench.parse(["--batch", "1", "--distinct", "1"])
    wl = bench.make_workload(args, 1, 0)
    ctx = _lib.Context(0)
    dev = torch.device("cuda", 0)
    n = int(wl.n_lm[0, 0])
    obj = torch.from_numpy(wl.lm_obj[0, 0, :n].copy()).to(dev)
    img = torch.from_numpy(wl.lm_pts[0, 0, :n].copy()).to(dev)     # the points themselves: every non-corrupted landmark is an inlier
    K = np.ascontiguousarray(wl.K, np.float64).reshape(9)
    clk = np.zeros(16, np.int64)
    ms = C.c_float(0)
    rows = []
    for rep in range(12):
        rc = ctx.lib.b200vo_debug_pose_phases(ctx.h, obj.data_ptr(), img.data_ptr(), n, K.ctypes.data_as(_lib.c_f64p), 500, C.c_float(8.0), 0.99,
                                              clk.ctypes.data_as(C.POINTER(C.c_longlong)), C.byref(ms))
        assert rc == 0, ctx.last_error()
        if rep >= 2:
            rows.append((clk.copy(), ms.value))
    mhz = 1965.0
    c = np.median(np.array([r[0] for r in rows]), axis=0)
    print(f"pose_phases: N = {n}, kernel {np.median([r[1] for r in rows]) * 1e3:.1f} us by CUDA events (launch + memset included); SM clock assumed {mhz:.0f} MHz")
    for k, name in enumerate(NAMES):
        print(f"  {name:58s} {(c[k + 1] - c[k]) / mhz:7.1f} us")
    print(f"  {'last chunk: subsets / 32 minimal solves / scoring':58s} {(c[9] - c[1]) / mhz:7.1f} / {(c[10] - c[9]) / mhz:.1f} / {(c[11] - c[10]) / mhz:.1f} us")
    print(f"  {'total (first to last stamp)':58s} {(c[8] - c[0]) / mhz:7.1f} us")


if __name__ == "__main__":
    main()
