#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu4.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu4.log
tail -40 gpurun_out/r2_pytest_gpu4.log
timeout 900 python tools/grad_accuracy.py > gpurun_out/grad_accuracy_r2.txt 2> gpurun_out/grad_accuracy_r2.err; head -14 gpurun_out/grad_accuracy_r2.txt; tail -3 gpurun_out/grad_accuracy_r2.err
{
python tools/prof_engine.py --B 100000 --T 2000 --reps 2
python tools/prof_engine.py --B 12500 --T 2000 --reps 2
python tools/prof_engine.py --B 25000 --T 2000 --reps 2
python tools/prof_engine.py --B 50000 --T 2000 --reps 2
} > gpurun_out/perf_r2b.log 2>&1
cat gpurun_out/perf_r2b.log
( time timeout 1200 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err ) 2>&1 | tail -3; tail -3 gpurun_out/bench_r2a.err; cut -c1-1500 gpurun_out/bench_r2a.json
