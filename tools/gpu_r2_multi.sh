#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_multi_gpus.txt
timeout 1200 python -m pytest tests/test_multi_gpu.py -q -x > gpurun_out/r2_multi_gpu_static.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2_multi_gpu_static.txt
tail -15 gpurun_out/r2_multi_gpu_static.txt
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_${N}gpu_strong.json 2> gpurun_out/r2_bench_${N}gpu_strong.err; tail -2 gpurun_out/r2_bench_${N}gpu_strong.err; cut -c1-700 gpurun_out/r2_bench_${N}gpu_strong.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --scaling weak > gpurun_out/r2_bench_${N}gpu_weak.json 2> gpurun_out/r2_bench_${N}gpu_weak.err; cut -c1-300 gpurun_out/r2_bench_${N}gpu_weak.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_bench_${N}gpu_ref.json 2> gpurun_out/r2_bench_${N}gpu_ref.err; cut -c1-300 gpurun_out/r2_bench_${N}gpu_ref.json
