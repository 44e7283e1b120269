mkdir -p gpurun_out
for sd in 21 22; do FUZZ_SEED=$sd timeout 400 python tests/fuzz_parity.py 240 > gpurun_out/r3_fuzz_$sd.log 2>&1; echo "seed $sd rc=$?"; tail -5 gpurun_out/r3_fuzz_$sd.log; done
