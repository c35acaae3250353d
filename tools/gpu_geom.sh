#!/bin/bash
# throughput of heis_kernel on C3 against the number of resident samples per SM (one wave each), best of 5 reps
run() { echo -n "CTAS=$1 WARPS=$2 B=$3: "; CPF_HEIS_CTAS=$1 CPF_HEIS_WARPS=$2 timeout 300 python tools/prof_engine.py --B $3 --T 400 --reps 5 2>&1 | awk '{print $10}' | sort -g | tail -1; }
run 1 16 2368    # 16 samples/SM
run 1 16 4144    # 28
run 1 16 6364    # 43
run 1 16 7400    # 50
run 1 16 8436    # 57
run 1 16 8880    # 60
run 2 8 8288     # 2 x 28
run 2 8 6364     # 2 x 21.5
run 1 11 12500
run 1 16 100000
run 1 11 100000
run 2 7 100000
run 2 5 100000
