#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (both arms), ncu launch list, full capture and DRAM traffic of heis_kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
# launch list of the bench command (short run: same launches, fewer Adam iterations and samples)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --iters 100 --no-cpu-baseline --no-static > gpurun_out/bench_under_ncu.log 2>&1
# full capture of the engine kernel on a short C3 run (2 CTAs x 32 resident samples per SM like the bench launch)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^heis_kernel -c 1 -f -o gpurun_out/prof_heis \
  python tools/prof_c3.py 9472 40 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
# DRAM traffic of one bench launch (B = 10^5, T = 2000)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:^heis_kernel -c 1 --csv \
  --log-file gpurun_out/traffic.csv python tools/prof_engine.py --B 100000 --T 2000 --reps 1 > gpurun_out/traffic.log 2>&1
tail -2 gpurun_out/traffic.csv
