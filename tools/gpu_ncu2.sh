#!/bin/bash
# ncu full capture of heis_kernel on a short C3 run (extra env passed through), plus repeated event timings
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^heis_kernel -c 1 -f -o gpurun_out/prof_heis \
  python tools/prof_c3.py 12500 40 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
for i in 1 2; do timeout 300 python tools/prof_engine.py --T 400 --reps 4 2>&1 | tail -4; done
CPF_HEIS_SKEW=36 timeout 300 python tools/prof_engine.py --T 400 --reps 4 2>&1 | tail -4
