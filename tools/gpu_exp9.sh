#!/bin/bash
mkdir -p gpurun_out
{
for v in cur sb256 sb512; do
 for sy in 0 3; do
  L=""; [ "$v" != cur ] && L=/root/repo/cpflow_b200/lib/libcpflow_b200_$v.so
  for n in 4 5; do
  echo -n "$v sync=$sy n=$n: "
  CPF_LIB_PATH=$L CPF_ENGINE_SYNC=$sy python tools/prof_engine.py --n $n --K $((n*12)) --loss state --B 200000 --T 100 --reps 2 2>&1 | tail -1
  done
 done
done
echo -n "relphase sync=3: "; CPF_ENGINE_SYNC=3 python tools/prof_engine.py --loss relphase --B 20000 --T 100 --reps 2 | tail -1
echo -n "relphase sync=0: "; CPF_ENGINE_SYNC=0 python tools/prof_engine.py --loss relphase --B 20000 --T 100 --reps 2 | tail -1
} > gpurun_out/exp9_single_block.txt 2>&1
cat gpurun_out/exp9_single_block.txt
