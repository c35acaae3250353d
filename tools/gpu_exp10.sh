#!/bin/bash
mkdir -p gpurun_out
{
for s in "" 1 2 4 5 8 10 20; do
  echo -n "slices=${s:-auto}: "
  CPF_HEIS_SLICES=$s python tools/prof_engine.py --B 100000 --T 2000 --reps 2 2>&1 | tail -1
done
for s in "" 1 4 8; do
  echo -n "B=12500 slices=${s:-auto}: "
  CPF_HEIS_SLICES=$s python tools/prof_engine.py --B 12500 --T 2000 --reps 2 2>&1 | tail -1
done
} > gpurun_out/exp10_slices.txt 2>&1
cat gpurun_out/exp10_slices.txt
