mkdir -p gpurun_out
for sd in 31 32; do FUZZ_SEED=$sd timeout 400 python tests/fuzz_parity.py 240 > gpurun_out/r3_fuzzb_$sd.log 2>&1; echo "seed $sd rc=$?"; tail -5 gpurun_out/r3_fuzzb_$sd.log; done
