#!/bin/bash
# the driver's multi-GPU bench commands, both arms, default flags
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/final_bench_${N}gpu_ref.json 2> gpurun_out/final_bench_${N}gpu_ref.err; cut -c1-300 gpurun_out/final_bench_${N}gpu_ref.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/final_bench_${N}gpu.json 2> gpurun_out/final_bench_${N}gpu.err; tail -2 gpurun_out/final_bench_${N}gpu.err; cut -c1-400 gpurun_out/final_bench_${N}gpu.json
