#!/bin/bash
# ncu full capture of the engine kernel on a short C3 run + quick bench
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err
python -c "import json; d=json.load(open('gpurun_out/bench_quick.json')); print('evals/s', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:engine_kernel -c 1 -f -o gpurun_out/prof_engine \
  python tools/prof_c3.py 12500 40 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
