"""Shared measurement code of the GPU parity tests and of tools/parity_report.py: the CUDA engine
(through the C ABI) against the CPU oracle on identical seeded angle batches.  Error measures are the
ones of SURVEY.md 8(d): loss / reg relative to the batch maximum, gradients norm-wise per sample."""
import numpy as np
import torch
from scipy.stats import unitary_group

from oracle import cpflow_oracle as O
from cpflow_b200.ansatz import Ansatz
from cpflow_b200.engine import Loss, Penalty
from cpflow_b200.penalty import RegularizationOptions, make_regularization_function
from cpflow_b200.topology import chain_layer, connected_layer, fill_layers

PF = make_regularization_function(RegularizationOptions)
# BASELINE.json north_star: 1e-5 relative in complex64, 1e-12 in complex128
TOL = {torch.float64: 1e-12, torch.float32: 1e-5}
STAR4 = [[0, 1], [0, 2], [0, 3]]
KITE4 = [[0, 1], [1, 2], [2, 3], [1, 3]]        # paper/results/toff4_kite_xyz
SQUARE4 = [[0, 1], [1, 2], [2, 3], [3, 0]]      # paper/results/toff4_square_xyz
CONFIGS = [(3, chain_layer(3), 5, "xyz"), (4, STAR4, 10, "xyz"), (2, [[0, 1]], 3, "xz"),
           (5, connected_layer(5), 12, "xyz"), (4, [[3, 1], [2, 0]], 7, "zyx"), (4, chain_layer(4), 40, "xyz"),
           (3, connected_layer(3), 7, "xyz"), (5, chain_layer(5), 9, "xz"), (4, STAR4, 40, "xyz"),
           (4, KITE4, 25, "xyz"), (4, SQUARE4, 21, "xyz"), (5, [[0, 1], [0, 2], [0, 3], [0, 4]], 13, "xyz")]


def pen(r=0.01):
    return Penalty("piecewise", r, PF.segments, PF.period)


def rel(a, b):
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


def setup(n, layer, K, rg):
    anz = Ansatz(n, "cp", fill_layers(layer, K), rg)
    oanz = O.cp_ansatz(layer, K, rg)
    return anz, oanz, O.ansatz_program(oanz)


def measure_loss_grad(n, layer, K, rg, dt, B=37, dev="cuda", kinds=("hs", "relphase", "state")):
    """Worst errors of unitary / loss / reg / gradient over a batch of B generic angle vectors.
    Returns {kind: {'loss':, 'reg':, 'grad':}, 'unitary': max abs}.  The oracle sees the dtype-rounded
    inputs, so input rounding is not part of the error."""
    anz, oanz, ops = setup(n, layer, K, rg)
    N = 2 ** n
    a64 = np.random.default_rng(n * 100 + K).uniform(0, 2 * np.pi, (B, anz.num_angles))
    a = torch.tensor(a64, dtype=dt, device=dev)
    a_o = torch.tensor(a.cpu().numpy().astype(np.float64))
    out = {}
    u = anz.program.unitary(a).cpu().numpy()
    uo = O.program_unitary_batched(n, ops, a_o).numpy()
    out["unitary"] = float(np.abs(u - uo).max())
    V = unitary_group.rvs(N, random_state=1)
    for kind in kinds:
        tgt = V[:, 0].copy() if kind == "state" else V
        lo, rg_, gr = anz.program.loss_grad(a, Loss(kind, tgt), pen())
        ol, orr, og = O.loss_and_grad_batched(n, ops, a_o, kind, torch.tensor(tgt), oanz.cp_mask, 0.01,
                                              O.make_regularization_function())
        g, og = gr.cpu().numpy().astype(np.float64), og.numpy()
        gn = np.linalg.norm(g - og, axis=1) / np.linalg.norm(og, axis=1)
        lo2, _, none = anz.program.loss_grad(a, Loss(kind, tgt), pen(), want_grad=False)
        out[kind] = {"loss": rel(lo.cpu().numpy(), ol.numpy()), "reg": rel(rg_.cpu().numpy(), orr.numpy()),
                     "grad": float(gn.max()), "loss_only_same_bits": bool(none is None and torch.equal(lo, lo2))}
    return out


def oracle_run(n, ops, target, a0, lr, T, cp_mask, r, **kw):
    return O.mynimize_repeated(n, ops, "hs", torch.tensor(target), a0, lr, T, cp_mask, r,
                               O.make_regularization_function(), **kw)


def adam_case_inputs(n, layer, K, B, freeze, seed=0):
    """Initial angles (float64 copies of the threefry float32 draws) of an Adam-loop parity case; freeze=True is the
    verification variant (cp_utils.py:205-247): a third of the CP angles pulled near 0 / pi, then projected and frozen
    with the oracle's project_cp_angles (cp_utils.py:70-77, 111-141)."""
    oanz = O.cp_ansatz(layer, K, "xyz")
    a0 = O.generate_initial_angles(seed, oanz.num_angles, oanz.cp_mask, batch_size=B).astype(np.float64)
    if not freeze:
        return a0, None
    rng = np.random.default_rng(seed + 1)
    cp_idx = np.flatnonzero(oanz.cp_mask)
    fm = np.zeros(a0.shape, dtype=bool)
    for b in range(B):
        k = rng.choice(cp_idx, len(cp_idx) // 3, replace=False)
        a0[b, k] = rng.choice([0.0, np.pi], len(k)) + rng.uniform(-0.15, 0.15, len(k))
        po, fo = O.project_cp_angles(a0[b].astype(np.float32), oanz.cp_mask, 0.2)
        a0[b], fm[b] = po.astype(np.float64), fo
    return a0, fm


def oracle_adam_case(n, layer, K, target, B, T, freeze, r=0.001476, lr=0.1, dt=torch.float64):
    """The oracle's loop (optimization.py:28-94, 362) on an Adam-loop parity case -> dict of arrays."""
    oanz = O.cp_ansatz(layer, K, "xyz")
    ops = O.ansatz_program(oanz)
    a0, fm = adam_case_inputs(n, layer, K, B, freeze)
    a0t = torch.tensor(a0).to(dt)
    if freeze:
        res = O.mynimize_repeated(n, ops, "hs", torch.tensor(target), a0t, lr, T, freeze_mask=torch.tensor(fm))
    else:
        res = oracle_run(n, ops, target, a0t, lr, T, oanz.cp_mask, r)
    return {"init_regloss": np.array([x["regloss"][0].item() for x in res]),
            "best_regloss": np.array([x["regloss"][1].item() for x in res]),
            "best_reg": np.array([x["reg"][1].item() for x in res]),
            "best_params": np.stack([x["params"][1].numpy() for x in res])}


def measure_adam_loop(n, layer, K, target, B, T, dt=torch.float64, r=0.001476, lr=0.1, freeze=False, dev="cuda",
                      expect=None):
    """The fused Adam loop (cpf_adam_run) against the oracle loop from the same initial angles: max abs errors of
    init regloss, best regloss, best reg, best params.  `expect`: stored oracle outputs (tests/golden/adam_c3.npz,
    written by tests/golden/make_adam_c3.py); None runs the oracle now."""
    anz, oanz, ops = setup(n, layer, K, "xyz")
    a0, fm = adam_case_inputs(n, layer, K, B, freeze)
    if expect is None:
        expect = oracle_adam_case(n, layer, K, target, B, T, freeze, r, lr, dt)
    a0t = torch.tensor(a0).to(dt).to(dev)
    fmt = torch.tensor(fm).to(torch.uint8).to(dev).contiguous() if freeze else None
    st = anz.program.adam_state(a0t.clone(), freeze=fmt)
    anz.program.adam_run(st, Loss("hs", target), None if freeze else pen(r), lr, T)
    torch.cuda.synchronize()
    out = {k: float(np.abs(getattr(st, k).cpu().numpy() - expect[k]).max())
           for k in ("init_regloss", "best_regloss", "best_reg", "best_params")}
    out["engine"] = anz.program.launch_plan(B, dtype=dt)["engine"]
    if freeze:
        out["frozen_moved"] = float((st.angles[fmt.bool()] - a0t[fmt.bool()]).abs().max()) if fm.any() else 0.0
    return out
