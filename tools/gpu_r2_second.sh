#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_full_size_properties_c3 --deselect tests/test_gpu_parity.py::test_adam_loop_parity_c3_shape_f32 > gpurun_out/r2_pytest_gpu2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu2.log
tail -60 gpurun_out/r2_pytest_gpu2.log
timeout 900 python tools/grad_conditioning.py > gpurun_out/grad_conditioning_r2.txt 2> gpurun_out/grad_conditioning_r2.err; cat gpurun_out/grad_conditioning_r2.txt; tail -3 gpurun_out/grad_conditioning_r2.err
{
echo "== baseline perf (r2 first build: generic kernel + slicing) =="
python tools/prof_engine.py --B 100000 --T 2000 --reps 2
CPF_HEIS_SLICES=1 python tools/prof_engine.py --B 100000 --T 2000 --reps 2
python tools/prof_engine.py --B 12500 --T 2000 --reps 2
CPF_HEIS_SLICES=1 python tools/prof_engine.py --B 12500 --T 2000 --reps 2
CPF_HEIS_ANY=1 python tools/prof_engine.py --B 100000 --T 2000 --reps 2
python tools/prof_engine.py --layer kite --K 25 --B 100000 --T 2000 --reps 2
python tools/prof_engine.py --layer square --K 24 --B 100000 --T 2000 --reps 2
python tools/prof_engine.py --layer star --n 5 --K 40 --B 50000 --T 500 --reps 2
} > gpurun_out/perf_r2a.log 2>&1
cat gpurun_out/perf_r2a.log
