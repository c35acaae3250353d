#!/bin/bash
mkdir -p gpurun_out
{
export CPF_LIB_PATH=/root/repo/cpflow_b200/lib/libcpflow_b200_cpt1.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "adam_loop or full_size or unitary_loss_grad or time_sliced" 2>&1 | tail -3
for c in 2 1 4; do
echo "== cpt1 ctas=$c"
CPF_HEIS_CTAS=$c CPF_HEIS_SLICES=1 python tools/prof_engine.py --B 37888 --T 500 --reps 2 2>&1 | tail -1
done
unset CPF_LIB_PATH
echo "== cur"
CPF_HEIS_SLICES=1 python tools/prof_engine.py --B 37888 --T 500 --reps 2 2>&1 | tail -1
} > gpurun_out/exp4.txt 2>&1
cat gpurun_out/exp4.txt
