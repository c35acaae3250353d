#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for cfg in "4 40" "5 60" "6 60" "7 60"; do
  set -- $cfg
  for nl in 0 1; do
  echo -n "n=$1 K=$2 no_layered=$nl: "
  CPF_NO_LAYERED=$nl python tools/prof_engine.py --n $1 --K $2 --loss state --B 200000 --T 100 --reps 2 2>&1 | tail -1
  done
done
} > gpurun_out/exp8_single_layered.txt 2>&1
cat gpurun_out/exp8_single_layered.txt
