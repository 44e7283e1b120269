set -x
mkdir -p gpurun_out
T=r2z
timeout 900 python -m pytest tests/test_sift_gpu.py -m gpu -q -x > gpurun_out/${T}_pytest_sift.log 2>&1; echo "pytest sift rc=$?"; tail -3 gpurun_out/${T}_pytest_sift.log
python benchmarks/bench_components.py --only sift --reps 10 --no-cv2 > gpurun_out/${T}_sift.jsonl 2>&1; cat gpurun_out/${T}_sift.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sift_desc --csv --log-file gpurun_out/${T}_sift_launches.csv python benchmarks/bench_components.py --only sift --reps 1 --no-cv2 > /dev/null 2>&1
python profiles/summarize.py launches gpurun_out/${T}_sift_launches.csv 2>&1 | head -4
