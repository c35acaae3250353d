#!/bin/bash
# best-of-4 event-timed throughput of the default engine on the BASELINE configs (tools/prof_engine.py)
run() { echo -n "$*: "; timeout 300 python tools/prof_engine.py "$@" --reps 4 2>&1 | awk '{print $10}' | sort -g | tail -1; }
run --n 3 --layer connected --K 7 --B 10000 --T 400
run --n 3 --layer chain --K 12 --B 10000 --T 400
run --n 4 --layer chain --K 40 --B 12500 --T 400
run --n 4 --layer chain --K 40 --B 100000 --T 100
run --n 4 --layer star --K 40 --B 12500 --T 400
run --n 4 --layer star --K 40 --B 100000 --T 100
run --n 4 --layer connected --K 61 --B 12500 --T 100 --dtype f64
run --n 4 --layer chain --K 40 --B 12500 --T 100 --dtype f64
run --n 4 --layer connected --K 70 --B 12500 --T 200
run --n 5 --layer chain --K 60 --B 12500 --T 100
run --n 5 --layer connected --K 60 --B 12500 --T 100
run --n 4 --layer kite --K 25 --B 100000 --T 200
run --n 4 --layer square --K 24 --B 100000 --T 200
run --n 5 --layer chain --K 60 --B 12500 --T 100 --dtype f64
run --n 4 --layer chain --K 48 --B 200000 --T 100 --loss state
run --n 5 --layer chain --K 60 --B 200000 --T 100 --loss state
run --n 6 --layer chain --K 60 --B 200000 --T 100 --loss state
run --n 7 --layer chain --K 60 --B 100000 --T 100 --loss state
run --n 4 --layer chain --K 40 --B 20000 --T 100 --loss relphase
