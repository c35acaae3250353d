#!/bin/bash
# multi-GPU evidence of the final round-2 build: static() identity test, strong-scaling bench, C5 at size
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x > gpurun_out/r2b_multi_gpu_static.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_multi_gpu_static.txt
tail -3 gpurun_out/r2b_multi_gpu_static.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-static --no-extras > gpurun_out/r2b_bench_${N}gpu_strong.json 2> gpurun_out/r2b_bench_${N}gpu_strong.err; tail -2 gpurun_out/r2b_bench_${N}gpu_strong.err; cut -c1-500 gpurun_out/r2b_bench_${N}gpu_strong.json
bash tools/gpu_c5.sh $N
