#!/bin/bash
# A/B: base build (before the update-phase prologue hoist) vs current
mkdir -p gpurun_out
{
for rep in 1 2; do
echo "== base"; CPF_LIB_PATH=$PWD/cpflow_b200/lib/libcpflow_b200_base.so python tools/prof_engine.py --B 100000 --T 2000 --reps 2
echo "== new";  python tools/prof_engine.py --B 100000 --T 2000 --reps 2
done
echo "== new 12500 / 25000 / 50000"
python tools/prof_engine.py --B 12500 --T 2000 --reps 2
python tools/prof_engine.py --B 25000 --T 2000 --reps 2
python tools/prof_engine.py --B 50000 --T 2000 --reps 2
echo "== new star / 5q / f64"
python tools/prof_engine.py --layer star --B 100000 --T 1000 --reps 2
python tools/prof_engine.py --n 5 --K 60 --B 40000 --T 200 --reps 2
python tools/prof_engine.py --dtype f64 --B 20000 --T 200 --reps 2
} > gpurun_out/perf_r2c.log 2>&1
cat gpurun_out/perf_r2c.log
timeout 900 python -m pytest tests -m gpu -q -x -k "parity or geometry or sliced or any_layer" 2>&1 | tail -3
