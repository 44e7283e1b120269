set -x
mkdir -p gpurun_out
T=r2k
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_read.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio --clock-control none -k regex:knn_gemm --csv --log-file gpurun_out/${T}_knn_v1.csv python benchmarks/bench_components.py --only knn --no-cv2 --reps 1 > /dev/null 2>&1
tail -12 gpurun_out/${T}_knn_v1.csv | awk -F'","' '{print $(NF-2), $NF}'
