"""Complex64 gradient accuracy of the HS loss on the C3 shape over thousands of samples (run on the GPU box; output kept
in profiles/grad_accuracy_r2.txt): per-sample norm-wise error |g32 - g64| / |g64| of both float32 engines against the
float64 engine, its quantiles, the batch-level Frobenius error, and the summation condition number of the trace
t = Tr(V^dag U) = sum_i y_i (cond = sum |y_i| / |sum y_i|) next to the worst samples — the first r2 capture showed
no correlation with it (0.08): the tail is the kernel's own rounding, not conditioning.  Also the error of the
complex64 Adam loop against the float32 oracle as a function of the horizon."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_lib as P  # noqa: E402
from cpflow_b200.ansatz import Ansatz  # noqa: E402
from cpflow_b200.engine import Loss  # noqa: E402
from cpflow_b200.gates import u_toff4  # noqa: E402
from cpflow_b200.topology import chain_layer, fill_layers  # noqa: E402


def run(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)


for lname, layer in (("star", P.STAR4), ("chain", chain_layer(4))):
    anz = Ansatz(4, "cp", fill_layers(layer, 40))
    prog = anz.program
    B = 4096
    a = prog.initial_angles(0, 100000, first=25000, count=B)
    loss = Loss("hs", u_toff4)
    pen = P.pen(0.001476)
    _, _, g64 = prog.loss_grad(a.double(), loss, pen)
    _, _, gh = prog.loss_grad(a, loss, pen)
    _, _, ga = run({"CPF_ENGINE": "adjoint"}, lambda: prog.loss_grad(a, loss, pen))
    U = prog.unitary(a.double())
    V = torch.tensor(u_toff4, dtype=U.dtype, device=U.device)
    y = torch.einsum("ij,bij->bi", V.conj(), U) if False else (V.conj()[None] * U).sum(1)   # column sums of conj(V) o U
    cond = (y.abs().sum(1) / y.sum(1).abs()).cpu().numpy()
    eh = ((gh.double() - g64).norm(dim=1) / g64.norm(dim=1)).cpu().numpy()
    ea = ((ga.double() - g64).norm(dim=1) / g64.norm(dim=1)).cpu().numpy()
    print(f"# {lname}: {B} samples, K=40, r=0.001476; error = |g32 - g64| / |g64| per sample")
    for name, e in (("heis", eh), ("adjoint", ea)):
        q = np.quantile(e, [0.5, 0.9, 0.99, 1.0])
        print(f"{name:8s} median {q[0]:.2e}  p90 {q[1]:.2e}  p99 {q[2]:.2e}  max {q[3]:.2e}   "
              f"batch-level |G32-G64|_F/|G64|_F {float((({'heis': gh, 'adjoint': ga}[name]).double() - g64).norm() / g64.norm()):.2e}   "
              f"max of error/cond {np.max(e / cond):.2e}  corr(log e, log cond) {np.corrcoef(np.log(e), np.log(cond))[0, 1]:.2f}")
    top = np.argsort(-eh)[:8]
    print("worst heis samples: " + "  ".join(f"(err {eh[i]:.1e} adj {ea[i]:.1e} cond {cond[i]:.0f})" for i in top))
    print(f"cond quantiles: median {np.median(cond):.1f} p99 {np.quantile(cond, 0.99):.1f} max {cond.max():.1f}")

print("#\n# complex64 Adam loop vs the (float32-input, float64-arithmetic) oracle, C3 chain, 32 samples: error vs horizon")
for T in (1, 2, 3, 4, 6, 10):
    m = P.measure_adam_loop(4, chain_layer(4), 40, u_toff4, B=32, T=T, dt=torch.float32)
    print(f"T={T:2d} " + " ".join(f"{k}={v:.2e}" for k, v in m.items() if isinstance(v, float)))
