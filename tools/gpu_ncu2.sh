#!/bin/bash
# ncu full capture of heis_kernel on a short C3 run with full residency (57 samples per SM), then GPU tests + bench
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^heis_kernel -c 1 -f -o gpurun_out/prof_heis \
  python tools/prof_c3.py 9472 40 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-static > gpurun_out/bench_quick.json; python -c "
import json
d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])"
