set -x
mkdir -p gpurun_out
T=r3c
timeout 600 python -m pytest tests/test_knn_gpu.py tests/test_dev_api_gpu.py -m gpu -q > gpurun_out/${T}_pytest_knn.log 2>&1; echo "pytest knn rc=$?"; tail -4 gpurun_out/${T}_pytest_knn.log
for dbg in 0 2; do
B200VO_KNN_DBG=$dbg ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:knn_gemm --csv --log-file gpurun_out/${T}_knn_dbg$dbg.csv python benchmarks/bench_components.py --only knn --no-cv2 --reps 1 > /dev/null 2>&1
echo "dbg $dbg"; grep "m256" gpurun_out/${T}_knn_dbg$dbg.csv | awk -F'","' '{print $(NF-2), $NF}' | tail -3
done
