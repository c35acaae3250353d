#!/bin/bash
# round 2, first call: GPU tests at the tightened tolerances, measured parity maxima, Toffoli-4 best counts, stat parity
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu.log
tail -30 gpurun_out/r2_pytest_gpu.log
timeout 600 python tools/parity_report.py > gpurun_out/parity_r2.txt 2> gpurun_out/parity_r2.err; tail -12 gpurun_out/parity_r2.txt; tail -3 gpurun_out/parity_r2.err
timeout 900 python tools/toff4_best.py 20000 > gpurun_out/toff4_best_r2.txt 2> gpurun_out/toff4_best_r2.err; cat gpurun_out/toff4_best_r2.txt; tail -3 gpurun_out/toff4_best_r2.err
timeout 900 python tools/stat_parity.py > gpurun_out/stat_parity_r2.txt 2>&1; tail -40 gpurun_out/stat_parity_r2.txt
