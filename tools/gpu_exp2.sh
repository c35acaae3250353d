#!/bin/bash
mkdir -p gpurun_out
{
for s in "" 1 2 4 5 10; do
  echo "== slices=${s:-auto}"
  CPF_HEIS_SLICES=$s python tools/prof_engine.py --B 100000 --T 2000 --reps 1 2>&1 | tail -1
done
echo "== B=94720 (10 exact waves)"
python tools/prof_engine.py --B 94720 --T 1000 --reps 1 2>&1 | tail -1
echo "== B=9472 T=4000"
python tools/prof_engine.py --B 9472 --T 4000 --reps 1 2>&1 | tail -1
} > gpurun_out/exp2.txt 2>&1
cat gpurun_out/exp2.txt
