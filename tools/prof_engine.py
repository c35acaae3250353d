"""Short single-kernel workload for ncu: C3 (4q Toffoli, K=40) fused Adam engine.
usage: python tools/prof_engine.py [--layer chain|star] [--n 4] [--K 40] [--B 12500] [--T 40] [--reps 3] [--dtype f32]
Prints the event-timed throughput of the last launch (not a bench value when run under ncu)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpflow_b200.ansatz import Ansatz
from cpflow_b200.engine import Loss, Penalty
from cpflow_b200.penalty import RegularizationOptions, make_regularization_function
from cpflow_b200.topology import chain_layer, connected_layer, fill_layers

ap = argparse.ArgumentParser()
ap.add_argument("--layer", default="chain")
ap.add_argument("--n", type=int, default=4)
ap.add_argument("--K", type=int, default=40)
ap.add_argument("--B", type=int, default=12500)
ap.add_argument("--T", type=int, default=40)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--loss", default="hs")
a = ap.parse_args()

n = a.n
layer = {"chain": chain_layer(n), "connected": connected_layer(n), "star": [[0, q] for q in range(1, n)],
         "kite": [[0, 1], [1, 2], [2, 3], [1, 3]], "square": [[0, 1], [1, 2], [2, 3], [3, 0]]}[a.layer]
anz = Ansatz(n, "cp", fill_layers(layer, a.K))
prog = anz.program
N = 1 << n
tgt = torch.eye(N, dtype=torch.complex128)
tgt[[N - 2, N - 1]] = tgt[[N - 1, N - 2]]
if a.loss == "state":
    tgt = torch.zeros(N, dtype=torch.complex128); tgt[0] = tgt[-1] = 2 ** -0.5
pf = make_regularization_function(RegularizationOptions)
pen = Penalty("piecewise", 0.001476, pf.segments, pf.period)
loss = Loss(a.loss, tgt.numpy())
dt = torch.float32 if a.dtype == "f32" else torch.float64
a0 = prog.initial_angles(0, a.B).to(dt)
flops, _ = prog.eval_cost()
for r in range(a.reps):
    st = prog.adam_state(a0.clone())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prog.adam_run(st, loss, pen, 0.1, a.T); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ev = a.B * a.T / (ms * 1e-3)
    print(f"{a.layer} n={n} K={a.K} B={a.B} T={a.T} {a.dtype} {a.loss}: {ms:.2f} ms  {ev:.4e} evals/s  "
          f"{ev * flops / 1e12:.2f} TFLOP/s(alg)")
