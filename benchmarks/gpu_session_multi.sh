# bench.py under torchrun on N GPUs of one box, as the driver launches it (run through gpurun --gpus N); LONG=1 adds a 200-step run
set -x
mkdir -p gpurun_out
T=${TAG:-r3}
N=${NGPU:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_n${N}.json 2> gpurun_out/${T}_bench_n${N}.err; echo "n$N rc=$?"; tail -2 gpurun_out/${T}_bench_n${N}.err
if [ -n "$LONG" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/${T}_bench_n${N}_200.json 2> gpurun_out/${T}_bench_n${N}_200.err; echo "n$N 200 rc=$?"
fi
