#!/bin/bash
# Round-2 ncu evidence: launch list of the bench command, full capture of heis_kernel, DRAM traffic of a bench-shaped run.
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --iters 100 --no-cpu-baseline --no-static --no-extras > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_launches.csv | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^heis_kernel -c 1 -f -o gpurun_out/${TAG}_heis \
  python tools/prof_c3.py 9472 40 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full.log
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:^heis_kernel -c 64 --csv \
  --log-file gpurun_out/${TAG}_traffic.csv python tools/prof_engine.py --B 100000 --T 2000 --reps 1 > gpurun_out/${TAG}_traffic.log 2>&1
tail -2 gpurun_out/${TAG}_traffic.csv | cut -c1-300
