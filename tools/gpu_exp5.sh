#!/bin/bash
mkdir -p gpurun_out
{
echo "== solo CTA of 8 warps (32 samples / SM), B = 18944 = 4 waves"
CPF_HEIS_SOLO=1 CPF_HEIS_CTAS=1 CPF_HEIS_WARPS=8 CPF_HEIS_SLICES=1 python tools/prof_engine.py --B 18944 --T 500 --reps 2 2>&1 | tail -1
echo "== two CTAs of 8 warps, B = 37888"
CPF_HEIS_SLICES=1 python tools/prof_engine.py --B 37888 --T 500 --reps 2 2>&1 | tail -1
echo "== solo CTA of 16 warps (64 samples / SM)"
CPF_HEIS_SOLO=1 CPF_HEIS_CTAS=1 CPF_HEIS_SLICES=1 python tools/prof_engine.py --B 37888 --T 500 --reps 2 2>&1 | tail -1
} > gpurun_out/exp5.txt 2>&1
cat gpurun_out/exp5.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^heis_kernel -c 1 -f -o gpurun_out/r2_v6_heis python tools/prof_c3.py 9472 40 > gpurun_out/r2_v6_ncu_full.log 2>&1; tail -1 gpurun_out/r2_v6_ncu_full.log
