#!/bin/bash
# phase-skew experiment: throughput of the C3 launch for several group splits, then the GPU parity tests
mkdir -p gpurun_out
for s in 0 50 36 45 55 64; do
  echo "== CPF_HEIS_SKEW=$s"; CPF_HEIS_SKEW=$s timeout 300 python tools/prof_engine.py --T 400 --reps 2 2>&1 | tail -1
done
for s in 0 50; do
  echo "== star CPF_HEIS_SKEW=$s"; CPF_HEIS_SKEW=$s timeout 300 python tools/prof_engine.py --layer star --T 400 --reps 2 2>&1 | tail -1
  echo "== n5 CPF_HEIS_SKEW=$s"; CPF_HEIS_SKEW=$s timeout 300 python tools/prof_engine.py --n 5 --K 60 --T 100 --reps 2 2>&1 | tail -1
  echo "== n3 CPF_HEIS_SKEW=$s"; CPF_HEIS_SKEW=$s timeout 300 python tools/prof_engine.py --n 3 --K 12 --T 400 --reps 2 2>&1 | tail -1
  echo "== f64 CPF_HEIS_SKEW=$s"; CPF_HEIS_SKEW=$s timeout 300 python tools/prof_engine.py --dtype f64 --T 100 --reps 2 2>&1 | tail -1
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
