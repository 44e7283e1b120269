"""Turns ncu outputs brought back from gpurun into the small text summaries committed here.

  python profiles/summarize.py launches gpurun_out/launches_r1.csv > profiles/r1_launches_summary.txt
  python profiles/summarize.py kernel gpurun_out/klt_r1.ncu-rep   > profiles/r1_klt_kernel_full.txt
  python profiles/summarize.py source gpurun_out/r2a_klt.ncu-rep  > profiles/r2a_klt_source_hot.txt   (needs -lineinfo + --import-source on)
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    kn, mv = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[hdr + 1:]:
        if len(r) > mv:
            agg[r[kn].split("(")[0]].append(float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: gpu__time_duration.sum per launch (ns), cold-cache serialised; compare SHARES")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:70]:70s} n={len(v):4d} total_us={sum(v) / 1e3:11.1f} avg_us={sum(v) / len(v) / 1e3:9.1f} share={sum(v) / tot:.3f}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H = rows[0]
    for V in rows[2:]:
        d = dict(zip(H, V))
        print(f"# {path}: {d.get('Kernel Name', '?')}  (ncu --set full --clock-control none)")
        for k in KEYS:
            if k in d:
                print(f"{k} = {d[k]}")
        stalls = sorted(((float(v), k) for k, v in d.items() if k.startswith("smsp__average_warp") and k.endswith("_per_issue_active.ratio") and v), reverse=True)
        for v, k in stalls[:8]:
            print(f"{k} = {v}")


def source(path, top=45):
    """Per SOURCE LINE: share of executed warp instructions, shared-memory wavefronts (actual / ideal) and stall samples."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    recs, cur, hdr, kname = [], None, None, "?"
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Function Name":
            kname = r[1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0].isdigit():
            recs.append((cur, int(r[0]), r[1].strip(), dict(zip(hdr, r))))

    def num(x):
        try:
            return float(x)
        except (TypeError, ValueError):
            return 0.0
    ti = sum(num(d["Instructions Executed"]) for *_, d in recs) or 1.0
    ts = sum(num(d["# Samples"]) for *_, d in recs) or 1.0
    tw = sum(num(d["L1 Wavefronts Shared"]) for *_, d in recs)
    te = sum(num(d["L1 Wavefronts Shared Excessive"]) for *_, d in recs)
    print(f"# {path}: {kname}: per source line (ncu --set full --import-source on; compiled with -lineinfo)")
    print(f"# warp instructions executed {ti:.0f}; shared-memory wavefronts {tw:.0f} of which excessive (bank conflicts) {te:.0f} = {100 * te / max(tw, 1):.1f} %")
    print("# -- by instructions executed")
    for f, l, src, d in sorted(recs, key=lambda o: -num(o[3]["Instructions Executed"]))[:top]:
        print(f"{f}:{l:<4d} inst {100 * num(d['Instructions Executed']) / ti:5.2f} %  samples {100 * num(d['# Samples']) / ts:5.2f} %  | {src[:120]}")
    print("# -- by shared-memory wavefronts")
    for f, l, src, d in sorted(recs, key=lambda o: -num(o[3]["L1 Wavefronts Shared"]))[:16]:
        if num(d["L1 Wavefronts Shared"]) <= 0:
            break
        print(f"{f}:{l:<4d} wavefronts {num(d['L1 Wavefronts Shared']) / 1e6:7.2f} M  ideal {num(d['L1 Wavefronts Shared Ideal']) / 1e6:7.2f} M  "
              f"excessive {num(d['L1 Wavefronts Shared Excessive']) / 1e6:7.2f} M  | {src[:100]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "source": source}[sys.argv[1]](sys.argv[2])
