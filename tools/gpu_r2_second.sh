#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_full_size_properties_c3 --deselect tests/test_gpu_parity.py::test_adam_loop_parity_c3_shape_f32 > gpurun_out/r2_pytest_gpu2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu2.log
tail -40 gpurun_out/r2_pytest_gpu2.log
timeout 900 python tools/grad_conditioning.py > gpurun_out/grad_conditioning_r2.txt 2> gpurun_out/grad_conditioning_r2.err; cat gpurun_out/grad_conditioning_r2.txt; tail -3 gpurun_out/grad_conditioning_r2.err
