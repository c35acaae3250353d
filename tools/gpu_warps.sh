#!/bin/bash
# does throughput follow the number of resident warps?  K=27 leaves shared memory for 16 warps
for w in 8 11 13 15 16; do
  echo "== K=27 CPF_HEIS_WARPS=$w"; CPF_HEIS_SKEW=0 CPF_HEIS_WARPS=$w timeout 300 python tools/prof_engine.py --K 27 --B 18944 --T 300 --reps 2 2>&1 | tail -1
done
for w in 11 12; do
  echo "== K=40 CPF_HEIS_WARPS=$w"; CPF_HEIS_SKEW=0 CPF_HEIS_WARPS=$w timeout 300 python tools/prof_engine.py --K 40 --T 300 --reps 2 2>&1 | tail -1
done
echo "== K=27 skew 36 w16"; CPF_HEIS_SKEW=36 CPF_HEIS_WARPS=16 timeout 300 python tools/prof_engine.py --K 27 --B 18944 --T 300 --reps 2 2>&1 | tail -1
echo "== K=27 skew 50 w16"; CPF_HEIS_SKEW=50 CPF_HEIS_WARPS=16 timeout 300 python tools/prof_engine.py --K 27 --B 18944 --T 300 --reps 2 2>&1 | tail -1
