#!/usr/bin/env python
"""Benchmark of the per-frame correspondence-and-pose hot path (BASELINE.json metric:
frames/s of KLT + PnP-RANSAC on 1241x376 frames with 2k tracked points).

One *step* = one new frame for each of `--batch` independent synthetic KITTI-shaped sequences
on this rank: pyramid build, KLT on ~1000 landmark keypoints and ~1000 candidate keypoints
(reference VisualOdometryPipeLine.py:281,:287), P3P-RANSAC + EPnP on the tracked landmarks
(:343).  Sequences are sharded across ranks with no data-path collective (weak scaling: the
per-GPU batch is fixed); the trajectories are gathered over NCCL at the end of the timed region.

  python bench.py --gpus N --steps K --warmup W            # B200 arm (libb200vo.so)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own cv2 CPU path

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/s KLT+PnP-RANSAC 1241x376 2k pts"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="independent sequences per GPU")
    ap.add_argument("--shape", default="kitti", choices=["kitti", "parking", "malaga"])
    ap.add_argument("--frames", type=int, default=6, help="distinct frames per sequence (visited back and forth)")
    ap.add_argument("--landmarks", type=int, default=1000)
    ap.add_argument("--candidates", type=int, default=1000)
    ap.add_argument("--cpu-seqs", type=int, default=0, help="sequences per step in the bounded CPU sample (0 = one per host core, at least 4)")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single", action="store_true", help="skip the extra one-sequence-per-call measurement")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
def clocks_sampler(stop, out, dev):
    """Samples SM clock / throttle reasons DURING the timed regions (NVML; nvidia-smi as a fallback)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        R = pynvml
        while not stop.is_set():
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") \
                else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            flags = ["Active" if rs & getattr(R, nm, 0) else "Not Active" for nm in (
                "nvmlClocksThrottleReasonHwSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown",
                "nvmlClocksThrottleReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwPowerCap")]
            out.append([str(sm), str(mx), "0", hex(rs)] + flags)
            stop.wait(0.02)
        return
    except Exception:
        pass
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--id={dev}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            if r.returncode == 0 and r.stdout.strip():
                out.append([c.strip() for c in r.stdout.strip().splitlines()[0].split(",")])
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = [float(s[0]) for s in samples if s[0].replace(".", "").isdigit()]
    mx = [float(s[1]) for s in samples if s[1].replace(".", "").isdigit()]
    reasons = set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for s in samples:
        for k, nm in enumerate(names):
            if len(s) > 4 + k and s[4 + k].lower().startswith("active"):
                reasons.add(nm)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons), "samples": len(samples)}


def cpu_reference_step(cv2, wl, opts, f, g, seqs):
    """The reference's own cv2 calls for one frame of each sequence in `seqs` (VisualOdometryPipeLine.py
    :281, :287, :343 with its KITTI options)."""
    n = 0
    for s in seqs:
        nl, nc = int(wl.n_lm[f, s]), int(wl.n_cand[f, s])
        prev, nxt = wl.frames[f, s], wl.frames[g, s]
        p, st, _ = cv2.calcOpticalFlowPyrLK(prev, nxt, wl.lm_pts[f, s, :nl], None, winSize=opts["win"],
                                            maxLevel=opts["max_level"], criteria=opts["criteria"])
        keep = (st == 1).squeeze()
        if nc > 1:
            cv2.calcOpticalFlowPyrLK(prev, nxt, wl.cand_pts[f, s, :nc], None, winSize=opts["win"],
                                     maxLevel=opts["max_level"], criteria=opts["criteria"])
        kp, lm = p[keep], wl.lm_obj[f, s, :nl][keep]
        if len(kp) >= 8:
            cv2.solvePnPRansac(lm, kp, wl.K, np.zeros(4), flags=cv2.SOLVEPNP_P3P, confidence=opts["pnp_conf"],
                               reprojectionError=opts["pnp_err"], iterationsCount=opts["pnp_iters"])
        n += 1
    return n


def time_cpu_reference(wl, opts, n_seqs, steps, warmup):
    """-> (frames/s, cores, kind, sample description).  cv2 is the reference's own CPU implementation of the
    path; two ways of using all host threads are timed and the FASTER one is reported: (a) the reference as
    written -- sequences one after the other, cv2's internal parallel_for over all cores; (b) one thread per
    sequence (cv2 releases the GIL), cv2 single-threaded inside.  If cv2 is not importable the C oracle port is
    timed instead (kind 'port', 1 core)."""
    from concurrent.futures import ThreadPoolExecutor
    from monocular_visual_odometry_va4mr_b200 import workload
    order = workload.frame_order(wl.F, steps + warmup)
    seqs = list(range(min(n_seqs, wl.batch)))
    try:
        import cv2
        cores = os.cpu_count() or 1

        def run(mode):
            if mode == "internal":
                cv2.setNumThreads(cores)
                step = lambda f, g: cpu_reference_step(cv2, wl, opts, f, g, seqs)
            else:
                cv2.setNumThreads(1)
                pool = ThreadPoolExecutor(max_workers=cores)
                step = lambda f, g: sum(pool.map(lambda s_: cpu_reference_step(cv2, wl, opts, f, g, [s_]), seqs))
            for t in range(warmup):
                step(order[t], order[t + 1])
            t0 = time.perf_counter()
            n = 0
            for t in range(warmup, warmup + steps):
                n += step(order[t], order[t + 1])
            return n / (time.perf_counter() - t0)

        fps_a, fps_b = run("internal"), run("pool")
        cv2.setNumThreads(cores)
        mode = "cv2 internal threading" if fps_a >= fps_b else "one thread per sequence"
        return max(fps_a, fps_b), cores, "reference", (
            f"cv2 {cv2.__version__} calcOpticalFlowPyrLK x2 + solvePnPRansac(P3P), {len(seqs)} sequences x {steps} frames, "
            f"{cores} threads; best of cv2-internal threading ({fps_a:.0f} f/s) and one thread per sequence ({fps_b:.0f} f/s): {mode}")
    except ImportError:
        import oracle
        t0 = time.perf_counter()
        n = 0
        for t in range(steps):
            f, g = order[t], order[t + 1]
            for s in seqs:
                nl, nc = int(wl.n_lm[f, s]), int(wl.n_cand[f, s])
                p, st, _ = oracle.calc_optical_flow_pyr_lk(wl.frames[f, s], wl.frames[g, s], wl.lm_pts[f, s, :nl],
                                                           opts["win"], opts["max_level"], opts["criteria"])
                oracle.calc_optical_flow_pyr_lk(wl.frames[f, s], wl.frames[g, s], wl.cand_pts[f, s, :nc],
                                                opts["win"], opts["max_level"], opts["criteria"])
                keep = st.ravel() == 1
                oracle.solve_pnp_ransac_p3p(wl.lm_obj[f, s, :nl][keep], p[keep], wl.K, opts["pnp_iters"], opts["pnp_err"], opts["pnp_conf"])
                n += 1
        dt = time.perf_counter() - t0
        return n / dt, 1, "port", f"C oracle port, {len(seqs)} sequences x {steps} frames, 1 thread"


# ------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.cpu_seqs <= 0:
        args.cpu_seqs = max(4, os.cpu_count() or 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from monocular_visual_odometry_va4mr_b200 import workload
    opts = workload.REFERENCE_OPTIONS[args.shape]
    cfg_common = {
        "workload": f"{args.batch} independent synthetic {args.shape}-shaped sequences per GPU (BASELINE.json configs[4] = 64 x the configs[0] sequence shape; SURVEY's C5 = 64 x C1), "
                    f"one new frame each per step: KLT {opts['win'][0]}x{opts['win'][1]} maxLevel {opts['max_level']} criteria {opts['criteria']} on "
                    f"~{args.landmarks} landmark + ~{args.candidates} candidate keypoints, P3P-RANSAC {opts['pnp_iters']} it / {opts['pnp_err']} px + EPnP",
        "shape": args.shape, "sequences_per_gpu": args.batch, "landmarks": args.landmarks, "candidates": args.candidates,
        "l2": f"inputs larger than L2: {args.frames} frame sets x batch rotate through HBM",
    }

    # ---------------- reference arm: the reference's cv2 CPU path, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        wl = workload.TrackWorkload(args.shape, batch=args.cpu_seqs, n_frames=args.frames, n_landmarks=args.landmarks,
                                    n_candidates=args.candidates, n_distinct=min(2, args.cpu_seqs), seed=0,
                                    cap_landmarks=1024 if args.landmarks <= 1024 else args.landmarks,
                                    cap_candidates=1024 if args.candidates <= 1024 else args.candidates)
        ref_steps = min(args.steps, 20)   # each step is a bounded sample; the whole arm stays within a few minutes
        fps, cores, kind, sample = time_cpu_reference(wl, opts, args.cpu_seqs, ref_steps, min(args.warmup, 3))
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.cpu_seqs / fps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32/f32+f64", "data": "synthetic",
            "config": dict(cfg_common, sample=f"each step = {args.cpu_seqs} of the sequences (bounded sample)"),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

    # ---------------- B200 arm ----------------
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference for the cv2 path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from monocular_visual_odometry_va4mr_b200 import _lib, sharding
    from monocular_visual_odometry_va4mr_b200.batch import SequenceBatch
    ctx = _lib.Context(local)
    capL = 1024 if args.landmarks <= 1024 else args.landmarks
    capC = 1024 if args.candidates <= 1024 else args.candidates
    wl = workload.TrackWorkload(args.shape, batch=args.batch, n_frames=args.frames, n_landmarks=args.landmarks,
                                n_candidates=args.candidates, n_distinct=2, seed=sharding.sequence_seed(rank * args.batch),
                                cap_landmarks=capL, cap_candidates=capC)
    sb = SequenceBatch(wl.batch, wl.h, wl.w, wl.K, win=opts["win"], max_level=opts["max_level"], criteria=opts["criteria"],
                       pnp_iters=opts["pnp_iters"], pnp_reproj_err=opts["pnp_err"], pnp_conf=opts["pnp_conf"],
                       max_landmarks=wl.L, max_candidates=wl.Cn, ctx=ctx)
    K, W = args.steps, max(args.warmup, 3)
    order = workload.frame_order(wl.F, K + W + 1)
    dev = torch.device("cuda", local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    # ---- (1) device-resident: every input already in HBM ----
    d_frames = torch.from_numpy(wl.frames).to(dev)
    d_lm_pts = torch.from_numpy(wl.lm_pts).to(dev)
    d_lm_obj = torch.from_numpy(wl.lm_obj).to(dev)
    d_n_lm = torch.from_numpy(wl.n_lm).to(dev)
    d_cand = torch.from_numpy(wl.cand_pts).to(dev)
    d_n_cand = torch.from_numpy(wl.n_cand).to(dev)
    b, L, Cn = wl.batch, wl.L, wl.Cn
    d_out = dict(lm_next=torch.empty((b, L, 2), dtype=torch.float32, device=dev), lm_status=torch.empty((b, L), dtype=torch.uint8, device=dev),
                 cand_next=torch.empty((b, Cn, 2), dtype=torch.float32, device=dev), cand_status=torch.empty((b, Cn), dtype=torch.uint8, device=dev),
                 pose=torch.zeros((K + W, b, 6), dtype=torch.float64, device=dev), pnp_ok=torch.empty((b,), dtype=torch.uint8, device=dev),
                 inlier_mask=torch.empty((b, L), dtype=torch.uint8, device=dev), n_inliers=torch.empty((b,), dtype=torch.int32, device=dev))
    torch.cuda.synchronize()

    def dev_step(t):
        f, g = order[t], order[t + 1]
        o = {k: v.data_ptr() for k, v in d_out.items()}
        o["pose"] = d_out["pose"][t].data_ptr()
        sb.step_dev(d_frames[g].data_ptr(), d_lm_pts[f].data_ptr(), d_lm_obj[f].data_ptr(), d_n_lm[f].data_ptr(),
                    d_cand[f].data_ptr(), d_n_cand[f].data_ptr(), o)

    sb.prime(wl.frames[order[0]])
    for t in range(W):
        dev_step(t)
    if world > 1:   # the trajectory gather of the timed region runs once untimed first (NCCL sets its channels up lazily)
        stream.synchronize()
        sharding.gather_trajectories(d_out["pose"], world)
        torch.cuda.current_stream().synchronize()
    barrier()
    clk_samples, stop = [], threading.Event()
    th = threading.Thread(target=clocks_sampler, args=(stop, clk_samples, local), daemon=True)
    th.start()
    ctx.lib.b200vo_batch_profile(sb.h, 1)
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(W, W + K):
        dev_step(t)
    if world > 1:   # gather the trajectories (poses) of every rank's sequences over NCCL
        stream.synchronize()
        gathered = sharding.gather_trajectories(d_out["pose"], world)
        torch.cuda.current_stream().synchronize()
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    stage_ms = np.zeros(3, np.float32)
    nprof = np.zeros(1, np.int32)
    ctx.lib.b200vo_batch_profile_read(sb.h, stage_ms.ctypes.data_as(_lib.c_f32p), nprof.ctypes.data_as(_lib.c_intp))
    ctx.lib.b200vo_batch_profile(sb.h, 0)
    n_ok = int(d_out["pnp_ok"].sum().item())
    if world > 1:
        tms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dev_ms = float(tms.item())
    value = world * wl.batch * K / (dev_ms * 1e-3)

    # ---- (2) end to end through the C-ABI with HOST buffers (pinned frames; H2D + D2H in the timed region) ----
    h_frames = sb.pinned_frames(wl.F)
    h_frames[:] = wl.frames
    # the caller's point / landmark arrays live in page-locked memory too (b200vo_host_alloc)
    h_lm_pts, h_lm_obj, h_n_lm = sb.pinned_like(wl.lm_pts), sb.pinned_like(wl.lm_obj), sb.pinned_like(wl.n_lm)
    h_cand, h_n_cand = sb.pinned_like(wl.cand_pts), sb.pinned_like(wl.n_cand)
    def timed_host_loop(prefetch):
        """K steps through b200vo_batch_step with host buffers.  prefetch: the frames of step t+1 are handed
        to b200vo_batch_submit_frames before step t is called (what a video reader does), so their upload and
        pyramid build overlap step t's kernels; every step still uploads one frame set and its point arrays
        and reads every result back inside the timed region."""
        sb.prime(h_frames[order[0]])
        if prefetch:
            sb.submit_frames(h_frames[order[1]])
        for t in range(W + K):
            if t == W:
                barrier()
                t0 = time.perf_counter()
            f, g = order[t], order[t + 1]
            if prefetch:
                sb.submit_frames(h_frames[order[t + 2]])
                sb.step(None, h_lm_pts[f], h_lm_obj[f], h_n_lm[f], h_cand[f], h_n_cand[f])
            else:
                sb.step(h_frames[g], h_lm_pts[f], h_lm_obj[f], h_n_lm[f], h_cand[f], h_n_cand[f])
        barrier()
        dt = time.perf_counter() - t0
        if prefetch:     # consume the frame set submitted ahead of the last timed step
            sb.step(None, h_lm_pts[order[W + K]], h_lm_obj[order[W + K]], h_n_lm[order[W + K]], h_cand[order[W + K]], h_n_cand[order[W + K]])
        return dt

    e2e_sync_s = timed_host_loop(False)
    e2e_s = timed_host_loop(True)
    stop.set()
    th.join(timeout=2)
    if world > 1:
        tms = torch.tensor([e2e_s, e2e_sync_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_s, e2e_sync_s = float(tms[0].item()), float(tms[1].item())
    e2e_value = world * wl.batch * K / e2e_s
    h2d, d2h = wl.bytes_per_step()

    # ---- roofline of the dominant kernel (klt_kernel), timed live with CUDA events on the ctx stream ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    levels = 1
    w_, h_ = wl.w, wl.h
    P = w_ * h_
    for _ in range(opts["max_level"]):
        w_, h_ = (w_ + 1) // 2, (h_ + 1) // 2
        if w_ <= opts["win"][0] or h_ <= opts["win"][1]:
            break
        P += w_ * h_
        levels += 1
    # the step launches the tracker twice (landmark set, then candidate set beside the pose chain); the roofline is
    # quoted for the LANDMARK launch, which runs alone between two CUDA events of the ctx stream
    n_pts_total = int(wl.n_lm[0].sum())
    klt_bytes = wl.batch * 2 * P + n_pts_total * 21           # both pyramids once + this launch's points in/out
    klt_ms = float(stage_ms[1]) / max(int(nprof[0]), 1)
    achieved = klt_bytes / (klt_ms * 1e-3) / 1e9 if klt_ms > 0 else 0.0
    # measured DRAM traffic of that kernel (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch):
    # taken from the committed capture when it was made on this very workload, else null
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "klt_dram_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("workload") == {"shape": args.shape, "batch": args.batch, "landmarks": args.landmarks, "candidates": args.candidates}:
            traffic, traffic_src = tj["dram_bytes_read"] + tj["dram_bytes_write"], tj["source"]
    roofline = {"bound": "hbm", "kernel": "klt_kernel_v2 (landmark launch)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": klt_bytes, "avg_launch_ms": klt_ms,
                "stage_ms_per_step": {"pyramid": float(stage_ms[0]) / max(int(nprof[0]), 1), "klt_landmarks": klt_ms,
                                      "pnp_beside_klt_candidates": float(stage_ms[2]) / max(int(nprof[0]), 1)},
                "note": "klt_kernel_v2 is instruction-issue bound (smsp issue active 82 %, profiles/r1p_klt_kernel_full.txt): integer bilinear taps from shared-memory-staged windows; DRAM traffic ~ the algorithmic bytes; see DESIGN.md"}

    # ---- extra (not the headline): Shi-Tomasi detection (reference :256, feature_adding runs it every frame) for the
    # whole batch on the resident frames ----
    detect = None
    if rank == 0 and world == 1 and not args.no_single:
        sb.good_features(1400, 0.1, 10.0)
        t0 = time.perf_counter()
        for _ in range(20):
            sb.good_features(1400, 0.1, 10.0)
        dt = (time.perf_counter() - t0) / 20
        detect = {"ms_per_step": 1e3 * dt, "frames_per_s": wl.batch / dt,
                  "what": f"b200vo_batch_good_features(1400, 0.1, 10) on the {wl.batch} resident frames, corner lists read back to the host"}
        if not args.no_cpu_baseline:
            try:
                import cv2
                from concurrent.futures import ThreadPoolExecutor
                cores = os.cpu_count() or 1
                cv2.setNumThreads(1)
                imgs = [wl.frames[0, s_ % wl.batch] for s_ in range(cores)]
                with ThreadPoolExecutor(max_workers=cores) as pool:
                    list(pool.map(lambda im: cv2.goodFeaturesToTrack(im, 1400, 0.1, 10, blockSize=3), imgs))
                    t0 = time.perf_counter()
                    for _ in range(5):
                        list(pool.map(lambda im: cv2.goodFeaturesToTrack(im, 1400, 0.1, 10, blockSize=3), imgs))
                    detect["cv2_frames_per_s"] = 5 * len(imgs) / (time.perf_counter() - t0)
                cv2.setNumThreads(cores)
                detect["cv2_what"] = f"cv2.goodFeaturesToTrack, one thread per image on {cores} cores"
            except ImportError:
                pass

    # ---- extra (not the headline): ONE sequence per call -- BASELINE config 0's shape, latency-bound ----
    single = None
    if rank == 0 and world == 1 and not args.no_single:
        wl1 = workload.TrackWorkload(args.shape, batch=1, n_frames=args.frames, n_landmarks=args.landmarks, n_candidates=args.candidates,
                                     n_distinct=1, seed=sharding.sequence_seed(0), cap_landmarks=capL, cap_candidates=capC)
        sb1 = SequenceBatch(1, wl1.h, wl1.w, wl1.K, win=opts["win"], max_level=opts["max_level"], criteria=opts["criteria"],
                            pnp_iters=opts["pnp_iters"], pnp_reproj_err=opts["pnp_err"], pnp_conf=opts["pnp_conf"],
                            max_landmarks=wl1.L, max_candidates=wl1.Cn, ctx=ctx)
        f1 = sb1.pinned_frames(wl1.F)
        f1[:] = wl1.frames
        a1 = [sb1.pinned_like(x) for x in (wl1.lm_pts, wl1.lm_obj, wl1.n_lm, wl1.cand_pts, wl1.n_cand)]
        n1 = min(K, 200)
        o1 = workload.frame_order(wl1.F, n1 + W + 1)
        sb1.prime(f1[o1[0]])
        sb1.submit_frames(f1[o1[1]])
        for t in range(W + n1):
            if t == W:
                ctx.sync()
                t0 = time.perf_counter()
            sb1.submit_frames(f1[o1[t + 2]])
            sb1.step(None, a1[0][o1[t]], a1[1][o1[t]], a1[2][o1[t]], a1[3][o1[t]], a1[4][o1[t]])
        dt1 = time.perf_counter() - t0
        sb1.step(None, a1[0][0], a1[1][0], a1[2][0], a1[3][0], a1[4][0])
        single = {"value": n1 / dt1, "unit": "frames/s", "ms_per_frame": 1e3 * dt1 / n1,
                  "what": "one sequence per b200vo_batch_step call (batch = 1), host buffers in and out, next frame prefetched: "
                          "the latency-bound shape of BASELINE config 0; a single sequence cannot be sharded (frame i needs frame i-1)"}
        if not args.no_cpu_baseline:
            try:
                import cv2
                cv2.setNumThreads(os.cpu_count() or 1)
                nc1 = 30
                for t in range(3):
                    cpu_reference_step(cv2, wl1, opts, o1[t], o1[t + 1], [0])
                t0 = time.perf_counter()
                for t in range(3, 3 + nc1):
                    cpu_reference_step(cv2, wl1, opts, o1[t], o1[t + 1], [0])
                single["cv2_value"] = nc1 / (time.perf_counter() - t0)
                single["cv2_what"] = f"cv2 {cv2.__version__}, same calls, one sequence, cv2-internal threading on {os.cpu_count()} cores"
            except ImportError:
                pass
        sb1.close()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            fps, cores, kind, sample = time_cpu_reference(wl, opts, args.cpu_seqs, args.cpu_steps, 2)
            cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "i32/f32+f64", "data": "synthetic",
            "config": dict(cfg_common, pyramid_levels=levels, pnp_ok_last_step=n_ok, parallelism=f"sequences sharded x{world}, NCCL all_gather of poses"),
            "clocks": summarize_clocks(clk_samples),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / K,
                    "api": "b200vo_batch_submit_frames(t+1) + b200vo_batch_step(t): page-locked host buffers in and out, "
                           "next frames uploaded while the current step runs",
                    "call_by_call": {"value": world * wl.batch * K / e2e_sync_s, "ms_per_step": 1e3 * e2e_sync_s / K,
                                     "api": "b200vo_batch_step(frames) only: upload, kernels and read-back serialised per call"}},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "single_sequence": single,
            "detect": detect,
        }
        print(json.dumps(line))
    sb.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
