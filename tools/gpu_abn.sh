#!/bin/bash
# interleaved comparison of library variants on the 4-wave C3 shape: gpu_abn.sh NAME... ("cur" = in-tree build)
mkdir -p gpurun_out
{
for rep in 1 2; do
  for v in "$@"; do
    L=""; [ "$v" != cur ] && L=/root/repo/cpflow_b200/lib/libcpflow_b200_$v.so
    echo -n "$v: "
    CPF_LIB_PATH=$L python tools/prof_engine.py --B 37888 --T 500 --reps 3 2>&1 | tail -1
  done
done
} > gpurun_out/abn.txt 2>&1
cat gpurun_out/abn.txt
