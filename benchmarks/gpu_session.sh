set -x
mkdir -p gpurun_out
T=r2i
B200VO_KNN=flat timeout 600 python -m pytest tests/test_knn_gpu.py tests/test_dev_api_gpu.py -m gpu -q > gpurun_out/${T}_pytest_knnflat.log 2>&1; echo "pytest flat rc=$?"; tail -5 gpurun_out/${T}_pytest_knnflat.log
timeout 600 python -m pytest tests/test_batch_gpu.py -m gpu -q -x > gpurun_out/${T}_pytest_batch.log 2>&1; echo "pytest batch rc=$?"; tail -3 gpurun_out/${T}_pytest_batch.log
for dbg in 0 1 2; do
  B200VO_KNN=flat B200VO_KNN_DBG=$dbg ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_knn_flat_dbg${dbg}.csv python benchmarks/bench_components.py --only knn --no-cv2 --reps 2 > /dev/null 2>&1
  grep -c knn_gemm gpurun_out/${T}_knn_flat_dbg${dbg}.csv; grep knn_gemm gpurun_out/${T}_knn_flat_dbg${dbg}.csv | tail -3 | awk -F'","' '{print $5, $NF}'
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_knn_v1.csv python benchmarks/bench_components.py --only knn --no-cv2 --reps 2 > /dev/null 2>&1
grep knn_gemm gpurun_out/${T}_knn_v1.csv | tail -3 | awk -F'","' '{print $5, $NF}'
B200VO_KNN=flat python benchmarks/bench_components.py --only knn --no-cv2 > gpurun_out/${T}_knn_flat.jsonl 2>&1; cat gpurun_out/${T}_knn_flat.jsonl
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_20.json 2> gpurun_out/${T}_bench_20.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench_20.err
python -c "
import json;d=json.load(open('gpurun_out/${T}_bench_20.json'));print(d['value'],d['e2e']['value'],d['single_sequence'])"
