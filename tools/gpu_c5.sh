#!/bin/bash
# C5 at size on N GPUs: gpu_c5.sh N
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  timeout 900 python tools/c5_scale.py > gpurun_out/c5_scale_${N}gpu.txt 2> gpurun_out/c5_scale_${N}gpu.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/c5_scale.py > gpurun_out/c5_scale_${N}gpu.txt 2> gpurun_out/c5_scale_${N}gpu.err
fi
cat gpurun_out/c5_scale_${N}gpu.txt; tail -3 gpurun_out/c5_scale_${N}gpu.err
