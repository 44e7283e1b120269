#!/usr/bin/env python
"""Short device-resident run of the batched per-frame step (the bench.py workload, nothing else) for ncu:
    python benchmarks/prof_step.py [--batch 64] [--steps 3] [--warmup 3]
Prints the CUDA-event time per step; a number printed under ncu is never a bench value."""
import argparse
import os
import sys


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--distinct", type=int, default=4)
    ap.add_argument("--shape", default="kitti")
    ap.add_argument("--no-lookahead", action="store_true")
    ap.add_argument("--first", type=int, default=0, help="global index of the shard's first sequence (which scenes it gets)")
    a = ap.parse_args()
    args = bench.parse(["--batch", str(a.batch), "--steps", str(a.steps), "--warmup", str(a.warmup), "--distinct", str(a.distinct),
                        "--shape", a.shape] + (["--no-lookahead"] if a.no_lookahead else []))
    import torch
    from monocular_visual_odometry_va4mr_b200 import _lib, workload
    torch.cuda.set_device(0)
    opts = workload.REFERENCE_OPTIONS[args.shape]
    ctx = _lib.Context(0)
    wl = bench.make_workload(args, a.batch, a.first)
    arm = bench.Arm(args, opts, ctx, wl, 1, 0, a.batch)
    res = arm.device_resident([a.batch])   # timed loop without the profiler, stage times from the pass behind it
    n = max(res["nprof"], 1)
    print(f"prof_step: batch {a.batch} (first sequence {a.first}): {res['dev_ms'] / arm.K:.4f} ms/step, {a.batch * arm.K / (res['dev_ms'] * 1e-3):.0f} frames/s, "
          f"stages (ms) pyramid {res['stage_ms'][0] / n:.4f} klt_landmarks {res['stage_ms'][1] / n:.4f} pose|cand {res['stage_ms'][2] / n:.4f}, "
          f"launches/step {res['launches'] / arm.K:.1f}, pnp_ok {res['n_ok']}/{a.batch}")
    arm.close()


if __name__ == "__main__":
    main()
