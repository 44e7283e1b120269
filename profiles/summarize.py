"""Turns ncu outputs brought back from gpurun into the small text summaries committed here.

  python profiles/summarize.py launches gpurun_out/launches_r1.csv > profiles/r1_launches_summary.txt
  python profiles/summarize.py kernel gpurun_out/klt_r1.ncu-rep   > profiles/r1_klt_kernel_full.txt
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


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
