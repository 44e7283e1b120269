set -x
mkdir -p gpurun_out
T=r2l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "n2 rc=$?"; tail -3 gpurun_out/${T}_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/${T}_bench_n2_ref.json 2> gpurun_out/${T}_bench_n2_ref.err; echo "n2 ref rc=$?"; tail -3 gpurun_out/${T}_bench_n2_ref.err
