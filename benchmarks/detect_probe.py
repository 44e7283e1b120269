"""Batched Shi-Tomasi detection on 64 resident KITTI-shaped frames: wall time per call (for ncu launch lists)."""
import sys, time
sys.path.insert(0, '.')
from monocular_visual_odometry_va4mr_b200 import workload
from monocular_visual_odometry_va4mr_b200.batch import SequenceBatch
wl = workload.TrackWorkload("kitti", batch=64, n_frames=2, n_landmarks=100, n_candidates=100, n_distinct=2, seed=0, cap_landmarks=128, cap_candidates=128)
sb = SequenceBatch(wl.batch, wl.h, wl.w, wl.K, max_landmarks=wl.L, max_candidates=wl.Cn)
sb.prime(wl.frames[0])
for _ in range(3):
    t0 = time.perf_counter(); c, n = sb.good_features(1400, 0.1, 10.0); print("ms", 1e3 * (time.perf_counter() - t0), int(n.mean()))
