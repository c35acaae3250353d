#!/bin/bash
# timing experiment: sweeps only / parameter phase only / both, with two CTAs and one CTA per SM (C3 shape)
mkdir -p gpurun_out
{
for v in "" noupd nosweep; do
  for c in 2 1; do
    for w in 8 16; do
      if [ "$c" = 2 ] && [ "$w" = 16 ]; then continue; fi
      echo "== variant=${v:-full} ctas=$c warps=$w"
      L=""; [ -n "$v" ] && L=/root/repo/cpflow_b200/lib/libcpflow_b200_$v.so
      CPF_LIB_PATH=$L CPF_HEIS_CTAS=$c CPF_HEIS_WARPS=$w CPF_HEIS_SLICES=1 python tools/prof_engine.py --B 37888 --T 500 --reps 2 2>&1 | tail -1
    done
  done
done
} > gpurun_out/exp_phases.txt 2>&1
cat gpurun_out/exp_phases.txt
