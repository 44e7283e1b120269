#!/usr/bin/env python
"""GPU wall-clock picture of the end-to-end loop (host buffers, next frames prefetched): run with
B200VO_TRACE_FILE=<path>; the file lists when each tracker launch claimed its first feature and retired its last warp.
    B200VO_TRACE_FILE=gpurun_out/e2e_trace.txt python benchmarks/e2e_trace.py [--batch 64] [--steps 30]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--call-by-call", action="store_true")
    a = ap.parse_args()
    args = bench.parse(["--batch", str(a.batch), "--steps", str(a.steps), "--warmup", "5"])
    import torch
    from monocular_visual_odometry_va4mr_b200 import _lib, workload
    torch.cuda.set_device(0)
    opts = workload.REFERENCE_OPTIONS[args.shape]
    ctx = _lib.Context(0)
    wl = bench.make_workload(args, a.batch, 0)
    arm = bench.Arm(args, opts, ctx, wl, 1, 0, a.batch)
    s, wall, devms = arm.timed_host_loop(not a.call_by_call)
    print(f"e2e_trace: batch {a.batch}: {1e3 * s / arm.K:.4f} ms/step wall, device ev0->ev1 median {sorted(devms)[len(devms) // 2]:.4f} ms")
    arm.close()


if __name__ == "__main__":
    main()
