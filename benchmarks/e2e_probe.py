"""Per-step wall / device time of the host-buffer batch API in both forms (call by call, prefetch) with the
in-library stage events -- the probe used to find the copy-engine ordering issue described in DESIGN.md section 6."""
import time, numpy as np, torch, ctypes as C, sys
sys.path.insert(0, '.')
from monocular_visual_odometry_va4mr_b200 import _lib, workload
from monocular_visual_odometry_va4mr_b200.batch import SequenceBatch
opts = workload.REFERENCE_OPTIONS["kitti"]
ctx = _lib.Context(0)
wl = workload.TrackWorkload("kitti", batch=64, n_frames=6, n_landmarks=1000, n_candidates=1000, n_distinct=2, seed=0, cap_landmarks=1024, cap_candidates=1024)
sb = SequenceBatch(wl.batch, wl.h, wl.w, wl.K, win=opts["win"], max_level=opts["max_level"], criteria=opts["criteria"],
                   pnp_iters=opts["pnp_iters"], pnp_reproj_err=opts["pnp_err"], pnp_conf=opts["pnp_conf"], max_landmarks=wl.L, max_candidates=wl.Cn, ctx=ctx)
hf = sb.pinned_frames(wl.F); hf[:] = wl.frames
P = [sb.pinned_like(x) for x in (wl.lm_pts, wl.lm_obj, wl.n_lm, wl.cand_pts, wl.n_cand)]
order = workload.frame_order(wl.F, 200)
# raw H2D bandwidth
d = torch.empty(hf[0].size, dtype=torch.uint8, device='cuda')
src = torch.from_numpy(hf[0].reshape(-1))
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); d.copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize()
    print("H2D %.1f MB in %.3f ms = %.1f GB/s (is_pinned=%s)" % (src.numel()/1e6, e0.elapsed_time(e1), src.numel()/e0.elapsed_time(e1)/1e6, src.is_pinned()))
last = C.c_float
ctx.lib.b200vo_last_gpu_ms.restype = C.c_float
for mode in ("sync", "prefetch"):
    sb.prime(hf[order[0]])
    if mode == "prefetch": sb.submit_frames(hf[order[1]])
    ctx.lib.b200vo_batch_profile(sb.h, 1)
    walls, gms, subs = [], [], []
    for t in range(60):
        f, g = order[t], order[t+1]
        t0 = time.perf_counter()
        if mode == "prefetch":
            sb.submit_frames(hf[order[t+2]]); t1 = time.perf_counter()
            sb.step(None, P[0][f], P[1][f], P[2][f], P[3][f], P[4][f])
        else:
            t1 = t0
            sb.step(hf[g], P[0][f], P[1][f], P[2][f], P[3][f], P[4][f])
        walls.append(time.perf_counter() - t0); subs.append(t1 - t0)
        gms.append(ctx.lib.b200vo_last_gpu_ms(ctx.h))
    if mode == "prefetch": sb.step(None, P[0][0], P[1][0], P[2][0], P[3][0], P[4][0])
    st = np.zeros(3, np.float32); n = np.zeros(1, np.int32)
    ctx.lib.b200vo_batch_profile_read(sb.h, st.ctypes.data_as(_lib.c_f32p), n.ctypes.data_as(_lib.c_intp))
    ctx.lib.b200vo_batch_profile(sb.h, 0)
    print(mode, "wall ms/step %.3f  submit %.3f  gpu ev0->ev1 %.3f  stages" % (1e3*np.median(walls[10:]), 1e3*np.median(subs[10:]), np.median(gms[10:])), st / max(n[0],1), n[0])
