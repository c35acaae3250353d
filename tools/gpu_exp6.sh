#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_quick.sh pair6
{
for v in "" f64t384 f64t256; do
  L=""; [ -n "$v" ] && L=/root/repo/cpflow_b200/lib/libcpflow_b200_$v.so
  echo "== f64 variant ${v:-cur}"
  CPF_LIB_PATH=$L python tools/prof_engine.py --dtype f64 --B 20000 --T 200 --reps 2 2>&1 | tail -1
  CPF_LIB_PATH=$L python tools/prof_engine.py --dtype f64 --layer connected --K 61 --B 20000 --T 200 --reps 2 2>&1 | tail -1
done
} > gpurun_out/exp6_f64.txt 2>&1
cat gpurun_out/exp6_f64.txt
