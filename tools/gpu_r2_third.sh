#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_slicing.py > gpurun_out/diag_slicing.txt 2>&1; cat gpurun_out/diag_slicing.txt
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_full_size_properties_c3 --deselect tests/test_gpu_parity.py::test_adam_loop_parity_c3_shape_f32 > gpurun_out/r2_pytest_gpu3.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu3.log
tail -30 gpurun_out/r2_pytest_gpu3.log
