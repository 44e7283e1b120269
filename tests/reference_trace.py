"""Shared replay of tests/golden/reference_trace.npz (made by tests/golden/make_reference_trace.py from the
unmodified reference class running on cv2): every hot-path call of the reference, with the reference's
actual inputs ("teacher-forced"), checked against what cv2 returned to the reference.

`impl` is a namespace of callables with the oracle's / the shim's shapes; see the two test modules."""
import os
import zlib

import numpy as np

from conftest import GOLDEN
from monocular_visual_odometry_va4mr_b200 import synth


def load():
    g = np.load(os.path.join(GOLDEN, "reference_trace.npz"))
    s = synth.render_sequence(str(g["render_shape"]), int(g["render_n"]), seed=int(g["render_seed"]))
    frames = s["frames"]
    crc = np.array([zlib.crc32(f.tobytes()) for f in frames], np.uint32)
    assert np.array_equal(crc, g["frame_crc"]), "synthetic renderer drifted from the frames the trace was recorded on"
    return g, frames


def same_E(E, Ec, tol=1e-8):
    return min(np.abs(E - Ec).max(), np.abs(E + Ec).max()) < tol


def check_klt(g, frames, i, run):
    w, h, ml, ct, cc = (int(v) for v in g[f"klt{i}_cfg"])
    p, st = run(frames[int(g[f"klt{i}_prev"])], frames[int(g[f"klt{i}_next"])], g[f"klt{i}_pts"], (w, h), ml, (ct, cc, float(g[f"klt{i}_eps"])))
    assert np.array_equal(st.ravel(), g[f"klt{i}_status"].ravel()), f"klt{i}: status"
    ok = g[f"klt{i}_status"].ravel() == 1
    d = np.abs(p.reshape(-1, 2) - g[f"klt{i}_out"].reshape(-1, 2))[ok].max(axis=1)
    assert (d <= 0.05).mean() >= 0.99 and d.max() < 0.5, (f"klt{i}", float(d.max()))     # north_star: <= 0.05 px for >= 99 %
    return float(d.max()), float((d == 0).mean())


def check_gftt(g, frames, i, run):
    mc, q, md, bs = g[f"gftt{i}_cfg"]
    out = run(frames[int(g[f"gftt{i}_frame"])], int(mc), float(q), float(md), int(bs))
    assert np.array_equal(np.asarray(out).reshape(-1, 2), g[f"gftt{i}_out"].reshape(-1, 2)), f"gftt{i}"


def check_knn(g, i, run):
    q, t = g[f"knn{i}_q"].astype(np.float32), g[f"knn{i}_t"].astype(np.float32)
    idx, dist = run(q, t)
    assert np.array_equal(idx, g[f"knn{i}_idx"]), f"knn{i}: indices"
    assert np.array_equal(dist, g[f"knn{i}_dist"]), f"knn{i}: distances"


def check_emat(g, i, run):
    pr, thr = g[f"emat{i}_cfg"]
    E, mask = run(g[f"emat{i}_p1"], g[f"emat{i}_p2"], g["K"], float(pr), float(thr))
    assert np.array_equal(mask.ravel(), g[f"emat{i}_mask"].ravel()), f"emat{i}: mask"
    assert same_E(E, g[f"emat{i}_E"]), f"emat{i}: E"


def check_pnp(g, i, run):
    it, err, conf = g[f"pnp{i}_cfg"]
    ok, rv, tv, inl = run(g[f"pnp{i}_obj"], g[f"pnp{i}_img"], g["K"], int(it), float(err), float(conf))
    assert bool(ok) == bool(g[f"pnp{i}_ok"]), f"pnp{i}: success"
    if ok:
        assert np.array_equal(np.asarray(inl).ravel(), g[f"pnp{i}_inliers"].ravel()), f"pnp{i}: inliers"
        assert np.abs(np.ravel(rv) - g[f"pnp{i}_rvec"].ravel()).max() < 1e-6
        assert np.abs(np.ravel(tv) - g[f"pnp{i}_tvec"].ravel()).max() < 1e-6 * max(1.0, np.abs(g[f"pnp{i}_tvec"]).max())


def check_tri(g, i, run):
    keep, lm, kp = run(g["K"], g["tri_cfg"], g[f"tri{i}_first_keys"], g[f"tri{i}_keys"], g[f"tri{i}_first_pose"],
                       g[f"tri{i}_poses"], g[f"tri{i}_cur"])
    assert np.array_equal(keep, g[f"tri{i}_keep"].astype(bool)), f"tri{i}: too_short_baseline"
    ref = g[f"tri{i}_landmarks"].reshape(-1, 3)
    assert lm.shape == ref.shape and np.array_equal(kp, g[f"tri{i}_keypoints"].reshape(-1, 2))
    if len(ref):
        assert np.all(np.abs(lm - ref) <= np.spacing(np.abs(ref))), f"tri{i}: landmarks"     # <= 1 float32 ulp


def check_fadd(g, i, run):
    pts = g[f"gftt{int(g[f'fadd{i}_gftt'])}_out"].reshape(-1, 2)
    valid = run(pts, g[f"fadd{i}_existing"], float(g[f"fadd{i}_min_dist"]))
    assert np.array_equal(valid, g[f"fadd{i}_valid"].astype(bool)), f"fadd{i}"


def check_rpose(g, i, run):
    good, R, t, mask = run(g[f"rpose{i}_E"], g[f"rpose{i}_p1"], g[f"rpose{i}_p2"], g["K"])
    assert int(good) == int(g[f"rpose{i}_good"]), f"rpose{i}: count"
    assert np.array_equal(np.asarray(mask).ravel(), g[f"rpose{i}_mask"].ravel()), f"rpose{i}: mask"
    assert np.abs(R - g[f"rpose{i}_R"]).max() < 1e-12 and np.abs(np.ravel(t) - g[f"rpose{i}_t"].ravel()).max() < 1e-12


def replay(impl):
    g, frames = load()
    seen = dict(klt=0, gftt=0, knn=0, emat=0, pnp=0, tri=0, fadd=0, rpose=0)
    for name in g["calls"]:
        name = str(name)
        kind, i = name.rstrip("0123456789"), int(name[len(name.rstrip("0123456789")):])
        seen[kind] += 1
        if kind == "klt":
            check_klt(g, frames, i, impl.klt)
        elif kind == "gftt":
            check_gftt(g, frames, i, impl.gftt)
        elif kind == "knn":
            check_knn(g, i, impl.knn)
        elif kind == "emat":
            check_emat(g, i, impl.emat)
        elif kind == "pnp":
            check_pnp(g, i, impl.pnp)
        elif kind == "tri":
            check_tri(g, i, impl.tri)
        elif kind == "fadd":
            check_fadd(g, i, impl.fadd)
        elif kind == "rpose":
            check_rpose(g, i, impl.rpose)
    assert seen["klt"] >= 2 and seen["gftt"] >= 1 and seen["knn"] == 1 and seen["emat"] == 1 and seen["pnp"] >= 1
    assert seen["tri"] >= 2 and seen["fadd"] >= 1 and seen["rpose"] == 1
    return seen
