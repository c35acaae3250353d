#!/bin/bash
# quick check of a kernel change: GPU parity tests, then C3-shape timings (full kernel and parameter phase only)
TAG=${1:-quick}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|Error|assert" gpurun_out/${TAG}_pytest.log | head -20
{
python tools/prof_engine.py --B 37888 --T 500 --reps 2 | tail -1
[ -f cpflow_b200/lib/libcpflow_b200_nosweep.so ] && CPF_LIB_PATH=/root/repo/cpflow_b200/lib/libcpflow_b200_nosweep.so python tools/prof_engine.py --B 37888 --T 500 --reps 2 | tail -1
python tools/prof_engine.py --B 100000 --T 2000 --reps 1 | tail -1
} > gpurun_out/${TAG}_perf.txt 2>&1
cat gpurun_out/${TAG}_perf.txt
