#!/bin/bash
# A/B of two builds of the library on the bench shape, interleaved so that clock / thermal state is shared
# usage: gpu_ab.sh NAME_A NAME_B   ("" or "cur" = the in-tree library)
mkdir -p gpurun_out
lib() { if [ -z "$1" ] || [ "$1" = cur ]; then echo ""; else echo /root/repo/cpflow_b200/lib/libcpflow_b200_$1.so; fi; }
{
for rep in 1 2; do
  for v in "$1" "$2"; do
    echo "== $v (rep $rep)"
    CPF_LIB_PATH=$(lib $v) python tools/prof_engine.py --B 100000 --T 2000 --reps 2 2>&1 | tail -2
    nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv,noheader
  done
done
} > gpurun_out/ab_$1_$2.txt 2>&1
cat gpurun_out/ab_$1_$2.txt
