set -x
mkdir -p gpurun_out
T=r2n
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/${T}_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_sift_launches.csv python benchmarks/bench_components.py --only sift --reps 1 --no-cv2 > /dev/null 2>&1
python profiles/summarize.py launches gpurun_out/${T}_sift_launches.csv > gpurun_out/${T}_sift_launches_summary.txt 2>&1; head -20 gpurun_out/${T}_sift_launches_summary.txt
