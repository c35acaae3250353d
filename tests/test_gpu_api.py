"""GPU tests of the reference's open API on top of the engine: arbitrary
`unitary_loss_func(U)` callables (main.py:528-529), arbitrary `cp_regularization_func` callables (main.py:536-539),
cp_distribution='normal' (cp_utils.py:38-40) and `Synthesize.adaptive` (main.py:695-864)."""
import math

import numpy as np
import pytest
import torch

import cpflow_b200 as cp
from oracle import cpflow_oracle as O
from conftest import hst
from cpflow_b200.engine import Loss, Penalty, TorchLoss
from cpflow_b200.gates import u_toff3
from cpflow_b200.hyper import next_seed
from cpflow_b200.optimization import ProgramLoss, mynimize_repeated, run_adam_batch
from cpflow_b200.penalty import RegularizationOptions, make_regularization_function, tabulate_penalty
from cpflow_b200.topology import chain_layer, connected_layer, fill_layers

pytestmark = pytest.mark.gpu
PF = make_regularization_function(RegularizationOptions)
CCZ = np.diag([1, 1, 1, 1, 1, 1, 1, -1]).astype(complex)


def _hs_callable(target, dtype):
    V = torch.as_tensor(target).to("cuda", dtype)
    n = V.shape[0]
    return lambda u: 1 - torch.abs((u * V.conj()).sum()) ** 2 / n ** 2      # cost_HST, matrix_utils.py:35-42


@pytest.mark.parametrize("dt,tol", [(torch.float64, 1e-9), (torch.float32, 2e-4)])
def test_callable_loss_equals_the_fused_path(dt, tol):
    """The HS loss written as a torch callable and run through the host-driven loop (cpf_unitary -> autograd
    cotangent -> cpf_adjoint_from_cotangent -> cpf_adam_step) gives the results of the fused kernel, with penalty,
    with a freeze mask, and with history."""
    n, layer, K = 3, chain_layer(3), 6
    anz = cp.Ansatz(n, "cp", fill_layers(layer, K))
    prog = anz.program
    cdt = torch.complex128 if dt == torch.float64 else torch.complex64
    pen = Penalty("piecewise", 0.002, PF.segments, PF.period)
    a0 = prog.initial_angles(1, 9).to(dt)
    T = 40 if dt == torch.float64 else 4      # float32 trajectories separate quickly (Adam's first steps)
    fused = run_adam_batch(prog, Loss("hs", u_toff3), pen, a0, 0.1, T)
    call = run_adam_batch(prog, TorchLoss(_hs_callable(u_toff3, cdt)), pen, a0, 0.1, T)
    assert float((fused.regloss - call.regloss).abs().max()) < tol
    assert float((fused.reg - call.reg).abs().max()) < tol
    assert float((fused.params - call.params).abs().max()) < tol * 50
    if dt == torch.float64:
        fm = torch.zeros(9, anz.num_angles, dtype=torch.uint8, device="cuda")
        fm[:, ::5] = 1
        f2 = run_adam_batch(prog, Loss("hs", u_toff3), None, a0, 0.01, T, freeze=fm)
        c2 = run_adam_batch(prog, TorchLoss(_hs_callable(u_toff3, cdt)), None, a0, 0.01, T, freeze=fm)
        assert float((f2.regloss - c2.regloss).abs().max()) < tol
        assert torch.equal(c2.params[:, 1][fm.bool()], a0[fm.bool()]) or \
            float((c2.params[:, 1] - f2.params[:, 1]).abs().max()) < 1e-7
        h1 = run_adam_batch(prog, Loss("hs", u_toff3), pen, a0, 0.1, 12, keep_history=True)
        h2 = run_adam_batch(prog, TorchLoss(_hs_callable(u_toff3, cdt)), pen, a0, 0.1, 12, keep_history=True)
        assert float((h1.params - h2.params).abs().max()) < 1e-9 and float((h1.regloss - h2.regloss).abs().max()) < 1e-10
        # against the oracle's loop directly
        ores = O.mynimize_repeated(n, O.ansatz_program(O.cp_ansatz(layer, K)), "hs", torch.tensor(u_toff3), a0.cpu(),
                                   0.1, T, anz.cp_mask, 0.002, O.make_regularization_function())
        obr = np.array([r["regloss"][1].item() for r in ores])
        assert np.abs(call.regloss[:, 1].cpu().numpy() - obr).max() < 1e-9


def test_synthesize_with_a_callable_loss_and_penalty(tmp_path):
    """README example posed the reference's way: `unitary_loss_func=<callable>` and a callable penalty; the result
    objects carry the callable's loss and survive save / load (dill)."""
    V = torch.as_tensor(CCZ).to("cuda", torch.complex64)

    def my_loss(u):
        return 1 - torch.abs((u * V.conj()).sum()) ** 2 / 64

    syn = cp.Synthesize(chain_layer(3), unitary_loss_func=my_loss, label="ccz_callable",
                        cp_regularization_func=lambda a: PF(a))
    assert isinstance(syn.unitary_loss_func, TorchLoss) and len(syn.cp_regularization_func.segments) == 9
    opts = cp.StaticOptions(num_cp_gates=12, accepted_num_cz_gates=10, num_samples=10, num_gd_iterations=600,
                            num_gd_iterations_at_verification=1500)
    res = syn.static(opts, save_to=str(tmp_path / "cc"))
    ref = cp.Synthesize(chain_layer(3), target_unitary=CCZ).static(opts, save_results=False)
    assert len(res.decompositions) >= 1
    # same samples, same loss: the same prospective set (complex64 trajectories drift, counts agree within one)
    assert abs(len(res.decompositions) - len(ref.decompositions)) <= 2
    for d in res.decompositions:
        assert d.cz_count <= 10 and d.loss <= 1e-5 and hst(d.unitary, CCZ) < 1e-5
    back = cp.Results.load(str(tmp_path / "cc"))
    assert [d.cz_count for d in back.decompositions] == [d.cz_count for d in res.decompositions]
    # a loss the declarative specs cannot express: HS distance modulo a diagonal on the first qubit
    with pytest.raises(ValueError, match="periodic|piecewise"):
        cp.Synthesize(chain_layer(3), target_unitary=CCZ, cp_regularization_func=lambda a: np.sin(np.asarray(a)) ** 2)


def test_cp_distribution_normal():
    """cp_dist='normal': CP angles 1.5 * N(0,1) from a second split of the sample key, other angles as 'uniform'
    (cp_utils.py:31-40).  Non-CP angles are bit-exact against the oracle; the normal draws (XLA's erf_inv polynomial,
    restated on both sides; libm log vs CUDA logf) agree to a few ulp; shard-independent."""
    anz = cp.Ansatz(3, "cp", fill_layers(chain_layer(3), 12))
    P, mask = anz.num_angles, anz.cp_mask.astype(bool)
    ref = O.generate_initial_angles(7, P, anz.cp_mask, cp_dist="normal", batch_size=40)
    got = anz.program.initial_angles(7, 40, cp_dist="normal").cpu().numpy()
    uni = anz.program.initial_angles(7, 40).cpu().numpy()
    assert np.array_equal(got[:, ~mask], ref[:, ~mask]) and np.array_equal(got[:, ~mask], uni[:, ~mask])
    assert np.abs(got[:, mask] - ref[:, mask]).max() < 2e-6 * max(1.0, np.abs(ref[:, mask]).max())
    part = anz.program.initial_angles(7, 40, first=11, count=9, cp_dist="normal").cpu().numpy()
    assert np.array_equal(part, got[11:20])
    big = anz.program.initial_angles(3, 20000, cp_dist="normal").cpu().numpy()[:, mask]
    assert abs(big.mean()) < 0.02 and abs(big.std() - 1.5) < 0.02
    syn = cp.Synthesize(chain_layer(3), target_unitary=CCZ)
    res = syn.static(cp.StaticOptions(num_cp_gates=12, accepted_num_cz_gates=10, num_samples=16,
                                      cp_distribution="normal"), save_results=False)
    assert all(d.cz_count <= 10 for d in res.decompositions)


def _score(cz_counts, n):
    return float(-np.log2((2.0 ** (-np.array(cz_counts, dtype=np.float32))).sum() / n)) if cz_counts else math.inf


def test_adaptive_seed_chain_score_resume_and_logs(tmp_path, trials):
    """Synthesize.adaptive (main.py:695-864) on Toffoli-3, chain: the seed chain (main.py:798-799), the score
    formula (main.py:735-737, also against a stored reference trial), saving after every evaluation, resuming from
    saved trials (main.py:773-781), verification of improvements only, and keep_logs (main.py:751-755)."""
    path = str(tmp_path / "t3_adaptive")
    syn = cp.Synthesize(chain_layer(3), target_unitary=u_toff3, label="t3_adaptive")
    opts = cp.AdaptiveOptions(min_num_cp_gates=8, max_num_cp_gates=16, max_evals=4, num_samples=60,
                              num_gd_iterations_at_verification=2500, keep_logs=True)
    res = syn.adaptive(opts, save_to=path)
    tr = res.trials.results
    assert len(tr) == 4
    seed = opts.random_seed
    for t in tr:                                   # seed chain: seed_{i+1} = int(split(PRNGKey(seed_i))[1][1])
        seed = next_seed(seed)
        assert t["random_seed"] == seed
        assert 8 <= t["num_cp_gates"] <= 16 and t["r"] > 0 and t["layer"] == chain_layer(3)
        assert t["cz_counts"] == sorted(t["cz_counts"])
        assert t["loss"] == pytest.approx(_score(t["cz_counts"], 60), abs=1e-5)
        assert len(t["prospective_decompositions"]) == len(t["cz_counts"])        # keep_logs
        assert set(t["attachments"]) == {"prospective_decompositions", "static_options", "unitary_loss_func"}
    chain = [opts.random_seed]
    for _ in range(4):
        chain.append(O.next_adaptive_seed(chain[-1]))
    assert [t["random_seed"] for t in tr] == chain[1:]
    # each trial's prospective set is what static stages 1-2 give for the same (k, r, seed)
    t = tr[-1]
    so = opts.get_static(t["num_cp_gates"], t["r"])
    so.random_seed, so.accepted_num_cz_gates = t["random_seed"], 10 ** 6
    _, cand = syn._prospective(so)
    assert [int(c) for c in cand[:, 1].tolist()] == t["cz_counts"]
    # decompositions: strictly improving CZ counts below the theoretical lower bound start (main.py:783-788, 824-855)
    counts = [d.cz_count for d in res.decompositions]
    assert counts and all(b < a for a, b in zip(counts, counts[1:])) and counts[0] < 14
    for d in res.decompositions:
        assert hst(d.unitary, u_toff3) < 1e-5 and d._adaptive_options is opts
    # resume: two more evaluations continue the seed chain from the saved file and keep earlier results
    opts2 = cp.AdaptiveOptions(min_num_cp_gates=8, max_num_cp_gates=16, max_evals=6, num_samples=60,
                               num_gd_iterations_at_verification=2500)
    res2 = cp.Synthesize(chain_layer(3), target_unitary=u_toff3, label="t3_adaptive").adaptive(opts2, save_to=path)
    tr2 = res2.trials.results
    assert len(tr2) == 6 and [t["random_seed"] for t in tr2[:4]] == chain[1:]
    assert tr2[4]["random_seed"] == O.next_adaptive_seed(chain[-1])
    assert "prospective_decompositions" not in tr2[5]
    assert len(res2.decompositions) >= len(res.decompositions)
    assert [d.cz_count for d in res2.decompositions[:len(counts)]] == counts
    assert res2.best_hyperparameters()[0] == [min(tr2, key=lambda t: t["loss"])["num_cp_gates"],
                                              min(tr2, key=lambda t: t["loss"])["r"]]
    # the score formula against a trial stored by the reference itself
    rec = trials["paper/results/toff3_chain_xyz"]
    ref_t = next(t for t in rec["trials"] if isinstance(t["cz_counts"], list) and t["cz_counts"])
    assert _score(ref_t["cz_counts"], 200) == pytest.approx(ref_t["score"], abs=1e-4)
    # a finished run does nothing more
    res3 = cp.Synthesize(chain_layer(3), target_unitary=u_toff3, label="t3_adaptive").adaptive(opts2, save_to=path)
    assert len(res3.trials.results) == 6
