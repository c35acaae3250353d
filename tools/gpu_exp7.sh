#!/bin/bash
mkdir -p gpurun_out
{
for cfg in "4 40 2" "4 40 1" "4 40 0" "5 60 1" "5 60 0" "7 60 4" "7 60 3" "7 60 2"; do
  set -- $cfg
  echo -n "n=$1 K=$2 rb=$3: "
  CPF_SINGLE_RB=$3 python tools/prof_engine.py --n $1 --K $2 --loss state --B 200000 --T 100 --reps 2 2>&1 | tail -1
done
} > gpurun_out/exp7b_single_rb.txt 2>&1
cat gpurun_out/exp7b_single_rb.txt
