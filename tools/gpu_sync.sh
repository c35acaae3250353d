#!/bin/bash
run() { echo -n "$*: "; env $1 $2 timeout 300 python tools/prof_engine.py --B $3 --T $4 --reps 4 2>&1 | awk '{print $10}' | sort -g | tail -1; }
run A=1 B=1 100000 200
run CPF_HEIS_CTAS=2 CPF_HEIS_WARPS=8 100000 200
run CPF_HEIS_CTAS=2 CPF_HEIS_WARPS=8 94720 200
run A=1 B=1 94720 200
run CPF_HEIS_SYNC_EVERY=7 B=1 100000 200
run CPF_HEIS_SYNC_EVERY=100 CPF_HEIS_SYNC_BWD=5 100000 200
run CPF_HEIS_WARPS=12 B=1 100000 200
run CPF_HEIS_WARPS=14 B=1 100000 200
