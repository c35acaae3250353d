#!/bin/bash
mkdir -p gpurun_out
{
for c in 2 4 3; do
  echo "== ctas=$c"
  CPF_HEIS_CTAS=$c CPF_HEIS_SLICES=1 python tools/prof_engine.py --B 37888 --T 500 --reps 3 2>&1 | tail -2
done
echo "== sync variants (ctas auto)"
for sy in 0 1 2; do echo "sync=$sy"; CPF_HEIS_SYNC=$sy CPF_HEIS_SLICES=1 python tools/prof_engine.py --B 37888 --T 500 --reps 2 2>&1 | tail -1; done
} > gpurun_out/exp3.txt 2>&1
cat gpurun_out/exp3.txt
