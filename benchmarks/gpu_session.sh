set -x
mkdir -p gpurun_out
T=r2s
timeout 600 python -m pytest tests/test_knn_gpu.py tests/test_dev_api_gpu.py -m gpu -q > gpurun_out/${T}_pytest_knn.log 2>&1; echo "pytest knn rc=$?"; tail -3 gpurun_out/${T}_pytest_knn.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:knn_gemm --csv --log-file gpurun_out/${T}_knn.csv python benchmarks/bench_components.py --only knn --no-cv2 --reps 1 > /dev/null 2>&1
grep "m256" gpurun_out/${T}_knn.csv | awk -F'","' '{print $(NF-2), $NF}' | tail -4
