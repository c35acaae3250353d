"""BASELINE configs[4] ("C5") at size: 5- and 6-qubit state preparation and the 5-qubit Toffoli (C4X) template with
10^6 samples split over the GPUs of one node (one process per GPU, samples sharded by global index, no data-path
collective; device-timed, max over ranks).  Run it alone or under torchrun:

    python tools/c5_scale.py [--samples 1000000] [--iters 100]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/c5_scale.py

Rank 0 prints one JSON line per case."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpflow_b200 import _lib as L
from cpflow_b200.ansatz import Ansatz
from cpflow_b200.engine import Loss, Penalty
from cpflow_b200.parallel import shard_range
from cpflow_b200.penalty import RegularizationOptions, make_regularization_function
from cpflow_b200.topology import chain_layer, fill_layers

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=1000000)
ap.add_argument("--iters", type=int, default=100)
a = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pf = make_regularization_function(RegularizationOptions)
pen = Penalty("piecewise", 0.001, pf.segments, pf.period)


def target(n, kind):
    N = 1 << n
    if kind == "state":
        t = np.zeros(N, dtype=complex)
        t[0] = t[-1] = 2 ** -0.5                      # GHZ-n
        return t
    t = np.eye(N, dtype=complex)                      # C^{n-1}X
    t[[N - 2, N - 1]] = t[[N - 1, N - 2]]
    return t


cases = [("stateprep_5q_chain_K60", 5, 60, "state"), ("stateprep_6q_chain_K60", 6, 60, "state"),
         ("toffoli5_chain_K60_hs", 5, 60, "hs"), ("toffoli5_chain_K100_hs", 5, 100, "hs")]
for name, n, K, kind in cases:
    anz = Ansatz(n, "cp", fill_layers(chain_layer(n), K))
    prog = anz.program
    first, count = shard_range(a.samples, rank, world)
    loss = Loss(kind, target(n, kind))
    chunk = 250000                                     # samples per launch (bounds the scratch: ~0.5 GB for n = 5, K = 100)
    best = []
    ms_rank = 0.0
    for rep in range(2):                               # second pass is the timed one
        ms_rank = 0.0
        best = []
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for c0 in range(0, count, chunk):
            nb = min(chunk, count - c0)
            st = prog.adam_state(prog.initial_angles(0, a.samples, first=first + c0, count=nb, device=torch.device('cuda', local)))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            prog.adam_run(st, loss, pen, 0.1, a.iters)
            e1.record()
            torch.cuda.synchronize()
            ms_rank += e0.elapsed_time(e1)
            best.append(st.best_regloss.min())
            del st
    t = torch.tensor([ms_rank], device="cuda")
    b = torch.stack(best).min().reshape(1)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.MIN)
    if rank == 0:
        lk = {"hs": L.LOSS_HS, "state": L.LOSS_STATE}[kind]
        print(json.dumps({"case": name, "n_gpus": world, "samples": a.samples, "adam_iterations": a.iters,
                          "ms": float(t), "evals_per_s": a.samples * a.iters / (float(t) * 1e-3),
                          "engine": "heis" if prog.launch_plan(min(count, chunk), lk, torch.float32)["engine"] == 1 else "state-adjoint",
                          "best_regloss": float(b), "scaling": "strong (10^6 samples over the ranks)"}), flush=True)
if world > 1:
    dist.destroy_process_group()
