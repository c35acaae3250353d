#!/bin/bash
# GPU tests (default engine and the gate-list interpreter) and a quick bench line: one gpurun call.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
CPF_NO_LAYERED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2>gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python -c "import json; d=json.load(open('gpurun_out/bench_quick.json')); print('evals/s', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])"
