"""Finds and dissects a findEssentialMat case where the CUDA mask and the oracle's differ (fuzz record: n = 1002,
outlier fraction 0.12188663426661836, 4 points): which sample / model the two loops choose and how the per-model
inlier counts compare on the SAME sample set.  python tests/tools/emat_probe.py [n of [seed]]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from make_golden import make_emat_pair  # noqa: E402
from monocular_visual_odometry_va4mr_b200 import cv2_compat, hotpath  # noqa: E402


def dissect(n, of, seed):
    p1, p2, K = make_emat_pair(n, of, seed)
    E, m = cv2_compat.findEssentialMat(p1, p2, K, method=cv2_compat.RANSAC, prob=0.99, threshold=1)
    Eo, mo, iters_o = oracle.find_essential_mat(p1, p2, K, 0.99, 1.0, 1000)
    print(f"case n={n} of={of} seed={seed}: mask differs in {int((m != mo).sum())} points; inliers cuda {int(m.sum())} oracle {int(mo.sum())}; oracle loop ran {iters_o}")
    smp = oracle.ransac_subsets(n, 5, 1000)
    out = hotpath.find_essential_mat_samples(p1, p2, K, smp, prob=0.99, threshold=1.0, want_models=True)
    n1 = np.column_stack([(p1[:, 0].astype(np.float64) - K[0, 2]) / K[0, 0], (p1[:, 1].astype(np.float64) - K[1, 2]) / K[1, 1]])
    n2 = np.column_stack([(p2[:, 0].astype(np.float64) - K[0, 2]) / K[0, 0], (p2[:, 1].astype(np.float64) - K[1, 2]) / K[1, 1]])
    t = 1.0 / ((K[0, 0] + K[1, 1]) / 2)
    thr = np.float32(t * t)
    run = out["iters_run"]
    print(f"  cuda: winner {out['winner']} (sample {out['winner'] // 10}, model {out['winner'] % 10}), loop ran {run}")
    o_cnt = np.zeros((run + 2, 10), np.int32)
    o_nm = np.zeros(run + 2, np.int32)
    for i in range(min(run + 2, len(smp))):
        Es = oracle.five_point(n1[smp[i]], n2[smp[i]])
        o_nm[i] = len(Es)
        for k, Em in enumerate(Es):
            err = oracle.sampson_errors(n1, n2, Em)
            o_cnt[i, k] = int((err <= thr).sum())
    bo, ro = oracle.ransac_select(o_cnt, o_nm, 10, run + 2, n, 5, 0.99)
    print(f"  oracle on the same subsets: winner {bo} (sample {bo // 10}, model {bo % 10}), loop ran {ro}")
    d = np.argwhere(out["counts"][:run + 2] != o_cnt)
    for i, k in d[:10]:
        g_err = oracle.sampson_errors(n1, n2, out["models"][i, k])
        err = oracle.sampson_errors(n1, n2, oracle.five_point(n1[smp[i]], n2[smp[i]])[k])
        fl = np.where((err <= thr) != (g_err <= thr))[0]
        print(f"  sample {i} model {k}: counts cuda {out['counts'][i, k]} oracle {o_cnt[i, k]}; flipped points {fl.tolist()}, "
              f"|err - thr| / thr = {np.abs(err[fl].astype(np.float64) - float(thr)) / float(thr)}, model diff {np.abs(out['models'][i, k] - oracle.five_point(n1[smp[i]], n2[smp[i]])[k]).max():.2e}")
    print("  model-count rows differ:", int((out["nmodels"][:run + 2] != o_nm).sum()))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1002
    of = float(sys.argv[2]) if len(sys.argv) > 2 else 0.12188663426661836
    if len(sys.argv) > 3:
        return dissect(n, of, int(sys.argv[3]))
    for seed in range(10000):
        p1, p2, K = make_emat_pair(n, of, seed)
        E, m = cv2_compat.findEssentialMat(p1, p2, K, method=cv2_compat.RANSAC, prob=0.99, threshold=1)
        Eo, mo, _ = oracle.find_essential_mat(p1, p2, K, 0.99, 1.0, 1000)
        if (E is None) != (Eo is None) or (E is not None and (m != mo).any()):
            dissect(n, of, seed)
            if int((m != mo).sum()) > 1:
                break


if __name__ == "__main__":
    main()
