"""Shared comparison of two SIFT results (test infrastructure).

Tolerances (floating point; SURVEY 8f row f4 -- stated here once, used by the oracle-vs-cv2 and the CUDA-vs-oracle tests):
  * keypoints are matched one to one in cv2's output order when the counts agree, else by nearest (x, y, angle);
  * at least ``min_match`` of the reference keypoints must have a counterpart with |dx|, |dy| <= 1e-3 px, the same packed
    octave code (octave, layer, quantised sub-layer offset), size within 1e-5 relative, angle within 1e-2 degrees
    (mod 360), response within 1e-5 relative;
  * of the matched descriptors at least ``min_desc_rows`` must be identical and no entry may differ by more than 2
    (descriptors are integers 0..255; one ulp in a Gaussian weight moves an entry that sits on .5 by one).
"""
import numpy as np


def compare(kp, octave, desc, ref_kp, ref_octave, ref_desc, min_match=0.995, min_desc_rows=0.97):
    kp, ref_kp = np.asarray(kp, np.float64), np.asarray(ref_kp, np.float64)
    n_ref = len(ref_kp)
    if n_ref == 0:
        assert len(kp) == 0
        return dict(matched=0, n_ref=0)
    if len(kp) == n_ref and np.abs(kp[:, :2] - ref_kp[:, :2]).max() <= 1e-3:
        idx = np.arange(n_ref)
    else:
        from scipy.spatial import cKDTree
        assert len(kp) > 0
        tree = cKDTree(np.c_[kp[:, 0], kp[:, 1], kp[:, 3] * 0.01])
        _, idx = tree.query(np.c_[ref_kp[:, 0], ref_kp[:, 1], ref_kp[:, 3] * 0.01])
    a, b = kp[idx], ref_kp
    dang = np.abs(a[:, 3] - b[:, 3])
    dang = np.minimum(dang, 360.0 - dang)
    ok = ((np.abs(a[:, 0] - b[:, 0]) <= 1e-3) & (np.abs(a[:, 1] - b[:, 1]) <= 1e-3)
          & (np.asarray(octave)[idx] == np.asarray(ref_octave))
          & (np.abs(a[:, 2] - b[:, 2]) <= 1e-5 * np.abs(b[:, 2])) & (dang <= 1e-2)
          & (np.abs(a[:, 4] - b[:, 4]) <= 1e-5 * np.abs(b[:, 4]) + 1e-9))
    frac = ok.mean()
    assert frac >= min_match, f"only {ok.sum()} of {n_ref} keypoints reproduced ({len(kp)} found)"
    assert abs(len(kp) - n_ref) <= max(2, int(0.005 * n_ref)), f"{len(kp)} keypoints against {n_ref}"
    out = dict(matched=int(ok.sum()), n_ref=n_ref, n=len(kp), bit_equal_rows=int((a == b).all(axis=1)[ok].sum()))
    if desc is not None and ref_desc is not None:
        d = np.abs(np.asarray(desc, np.float32)[idx][ok] - np.asarray(ref_desc, np.float32)[ok])
        rows_equal = (d.max(axis=1) == 0).mean()
        assert d.max() <= 2, f"descriptor entries differ by up to {d.max()}"
        assert rows_equal >= min_desc_rows, f"only {rows_equal:.3f} of the matched descriptors identical"
        out.update(desc_rows_equal=float(rows_equal), desc_max_diff=float(d.max()))
    return out
