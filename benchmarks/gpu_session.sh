set -x
mkdir -p gpurun_out
T=r2m
timeout 900 python -m pytest tests/test_sift_gpu.py -m gpu -q -x > gpurun_out/${T}_pytest_sift.log 2>&1; echo "pytest sift rc=$?"; tail -30 gpurun_out/${T}_pytest_sift.log
python benchmarks/bench_components.py --only sift --reps 10 > gpurun_out/${T}_sift.jsonl 2>&1; cat gpurun_out/${T}_sift.jsonl
