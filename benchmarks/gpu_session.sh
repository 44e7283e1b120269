set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2h_pytest.log
B200VO_KNN=flat timeout 600 python -m pytest tests/test_knn_gpu.py -m gpu -q > gpurun_out/r2h_pytest_knnflat.log 2>&1; echo "pytest flat rc=$?"; tail -5 gpurun_out/r2h_pytest_knnflat.log
python benchmarks/pose_phases.py > gpurun_out/r2h_pose_phases.txt 2>&1; cat gpurun_out/r2h_pose_phases.txt
for b in 64 8 1; do python benchmarks/prof_step.py --batch $b --steps 100 --warmup 5; done > gpurun_out/r2h_shards.log 2>&1; cat gpurun_out/r2h_shards.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_20.json 2> gpurun_out/r2h_bench_20.err; echo "bench rc=$?"; tail -3 gpurun_out/r2h_bench_20.err
