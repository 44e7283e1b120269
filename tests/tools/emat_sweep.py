"""Three-way sweep of findEssentialMat: CUDA path vs the C oracle vs cv2 on random synthetic pairs.
python tests/tools/emat_sweep.py [cases]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from make_golden import make_emat_pair  # noqa: E402
from monocular_visual_odometry_va4mr_b200 import cv2_compat  # noqa: E402

import cv2  # noqa: E402


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    rng = np.random.default_rng(0)
    todo = [(1002, 0.12188663426661836, 1797), (1002, 0.12188663426661836, 105)]
    for _ in range(cases):
        todo.append((int(rng.integers(20, 3000)), float(rng.uniform(0, 0.6)), int(rng.integers(0, 10000))))
    go, gc, oc = [], [], []
    for n, of, seed in todo:
        p1, p2, K = make_emat_pair(n, of, seed)
        Eg, mg = cv2_compat.findEssentialMat(p1, p2, K, method=cv2_compat.RANSAC, prob=0.99, threshold=1)
        Eo, mo, _ = oracle.find_essential_mat(p1, p2, K, 0.99, 1.0, 1000)
        Ec, mc = cv2.findEssentialMat(p1, p2, K, method=cv2.RANSAC, prob=0.99, threshold=1)

        def diff(a, b):
            return -1 if (a is None) != (b is None) else 0 if a is None else int((a != b).sum())
        d1, d2, d3 = diff(mg, mo), diff(mg, mc), diff(mo, mc)
        if d1: go.append((n, round(of, 3), seed, d1))
        if d2: gc.append((n, round(of, 3), seed, d2))
        if d3: oc.append((n, round(of, 3), seed, d3))
    print(f"{len(todo)} cases")
    print("cuda vs oracle:", len(go), go)
    print("cuda vs cv2   :", len(gc), gc)
    print("oracle vs cv2 :", len(oc), oc)


if __name__ == "__main__":
    main()
