# Round-end session on one B200 (run through gpurun): tests, plain bench runs, then the profiler passes the files under
# profiles/ come from.  Numbers are only ever taken from the plain runs.
set -x
mkdir -p gpurun_out
T=${TAG:-r2end}
B="python bench.py --steps 2 --warmup 3 --no-single --no-cpu-baseline --no-parity"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/${T}_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_20.json 2> gpurun_out/${T}_bench_20.err; echo "bench20 rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err; echo "ref rc=$?"
python benchmarks/bench_components.py > gpurun_out/${T}_components.jsonl 2> gpurun_out/${T}_components.err; echo "components rc=$?"
if [ -n "$PROFILE" ]; then
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench200 rc=$?"
python benchmarks/pose_phases.py > gpurun_out/${T}_pose_phases.txt 2>&1
for b in 64 32 16 8 1; do python benchmarks/prof_step.py --batch $b --steps 100 --warmup 5; done > gpurun_out/${T}_shard_sizes.txt 2>&1
$B > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv $B > gpurun_out/${T}_ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:klt_kernel_v3 -c 1 -s 6 -o gpurun_out/${T}_klt $B > gpurun_out/${T}_ncu_klt.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:pnp_fused -c 1 -s 3 -o gpurun_out/${T}_pose $B > gpurun_out/${T}_ncu_pose.log 2>&1
ncu --set full --clock-control none -k regex:pyr_down_tma -c 1 -s 8 -o gpurun_out/${T}_pyr $B > gpurun_out/${T}_ncu_pyr.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:knn_gemm -c 1 -s 3 -o gpurun_out/${T}_knn python benchmarks/bench_components.py --only knn --no-cv2 --reps 2 > gpurun_out/${T}_ncu_knn.log 2>&1
ncu --set full --clock-control none -k regex:sift_blur_col -c 1 -s 3 -o gpurun_out/${T}_siftblur python benchmarks/bench_components.py --only sift --no-cv2 --reps 1 > gpurun_out/${T}_ncu_siftblur.log 2>&1
fi
