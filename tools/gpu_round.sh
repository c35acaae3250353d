#!/bin/bash
# One gpurun call: GPU tests, smoke, FP32 peak, bench (both arms), ncu launch list + full capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
ls baseline/_ref 2>&1 | head -3; python -c "import jax" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 120 ./tools/fp32_peak > gpurun_out/fp32_peak.jsonl 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
# launch list of the bench command (short run)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --iters 200 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# full capture of the engine kernel on a short C3 run
timeout 900 ncu --set full --clock-control none --import-source on -k regex:engine_kernel -c 1 -o gpurun_out/prof_engine \
  python tools/prof_c3.py 12500 40 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
