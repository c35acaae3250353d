#!/bin/bash
# tests + ncu evidence + both bench arms for the current build
TAG=${1:-r2_v2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -25 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
bash tools/gpu_ncu_r2.sh ${TAG}
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench.json
{
python tools/prof_engine.py --B 12500 --T 2000 --reps 2
python tools/prof_engine.py --B 25000 --T 2000 --reps 2
python tools/prof_engine.py --B 50000 --T 2000 --reps 2
} > gpurun_out/${TAG}_perf_strong_shards.log 2>&1; cat gpurun_out/${TAG}_perf_strong_shards.log
