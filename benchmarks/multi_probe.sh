# Why is a shard slower under torchrun than alone?  (1) two independent one-GPU loops side by side, (2) torchrun without the clock sampler
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 python benchmarks/prof_step.py --batch 32 --distinct 16 --steps 100 --warmup 5 > gpurun_out/mp_a.txt 2>&1 &
CUDA_VISIBLE_DEVICES=1 python benchmarks/prof_step.py --batch 32 --first 32 --distinct 16 --steps 100 --warmup 5 > gpurun_out/mp_b.txt 2>&1 &
wait
grep prof_step gpurun_out/mp_a.txt gpurun_out/mp_b.txt
B200VO_NO_CLOCK_SAMPLER=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/mp_c.json 2> gpurun_out/mp_c.err; echo "rc=$?"
