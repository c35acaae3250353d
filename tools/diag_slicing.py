"""Diagnostic: which samples of a time-sliced Adam run differ from the single-launch run."""
import os
import sys

import numpy as np
import torch
from scipy.stats import unitary_group

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_lib as P  # noqa: E402
from cpflow_b200.ansatz import Ansatz  # noqa: E402
from cpflow_b200.engine import Loss  # noqa: E402
from cpflow_b200.topology import chain_layer, fill_layers  # noqa: E402

anz = Ansatz(4, "cp", fill_layers(chain_layer(4), 12))
V = unitary_group.rvs(16, random_state=1)
n_sm = torch.cuda.get_device_properties(0).multi_processor_count
plan = anz.program.launch_plan(10 ** 6, n_sm=n_sm)
slots = plan["samples_per_cta"] * plan["ctas_per_sm"] * n_sm
B = slots + slots // 3 + 5
a = anz.program.initial_angles(2, B)
print("slots", slots, "B", B)


def run(T, k, step_chunks=None):
    os.environ["CPF_HEIS_SLICES"] = str(k)
    st = anz.program.adam_state(a.clone())
    for t in (step_chunks or [T]):
        anz.program.adam_run(st, Loss("hs", V), P.pen(), 0.1, t)
    torch.cuda.synchronize()
    return st


def ranges(idx):
    idx = np.asarray(idx)
    if len(idx) == 0:
        return "none"
    cuts = np.flatnonzero(np.diff(idx) > 1)
    starts = np.concatenate([[idx[0]], idx[cuts + 1]])
    ends = np.concatenate([idx[cuts], [idx[-1]]])
    return " ".join(f"[{s}..{e}]" for s, e in list(zip(starts, ends))[:12]) + (" ..." if len(starts) > 12 else "")


for T, k in ((60, 2), (40, 2), (60, 3), (60, 20), (60, 30), (2000, 25)):
    ref = run(T, 1)
    user = run(T, 1, [T // k] * k)         # the same chunks through the public step0 mechanism
    st = run(T, k)
    for name in ("angles", "best_regloss", "best_params", "m"):
        x, y, z = getattr(st, name), getattr(ref, name), getattr(user, name)
        bad = (x != y).reshape(B, -1).any(1).nonzero().flatten().cpu().numpy()
        badu = (z != y).reshape(B, -1).any(1).nonzero().flatten().cpu().numpy()
        print(f"T={T} k={k} {name:13s} sliced-vs-single mismatching samples: {len(bad):6d} {ranges(bad)} | "
              f"user-chunked-vs-single: {len(badu)} {ranges(badu)}")
