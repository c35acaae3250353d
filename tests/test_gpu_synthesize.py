"""GPU tests of the host API on top of the engine: mynimize_repeated's return contract, selection,
batched verification, Synthesize.static end to end, and statistical parity with the reference's
stored hyperopt trials (same template, r, sample count: prospective fraction, minimum CZ count and
score must agree within sampling error — SURVEY.md §4.3 "statistical pins")."""
import math

import numpy as np
import pytest
import torch

import cpflow_b200 as cp
from oracle import cpflow_oracle as O
from conftest import hst
from cpflow_b200 import cp_utils as CU
from cpflow_b200.engine import Loss, Penalty
from cpflow_b200.gates import u_toff3, u_toff4
from cpflow_b200.optimization import ProgramLoss, RawResults, mynimize, mynimize_repeated, unitary_learn
from cpflow_b200.penalty import RegularizationOptions, make_regularization_function
from cpflow_b200.topology import chain_layer, connected_layer, fill_layers

pytestmark = pytest.mark.gpu
PF = make_regularization_function(RegularizationOptions)
CCZ = np.diag([1, 1, 1, 1, 1, 1, 1, -1]).astype(complex)


def test_mynimize_repeated_return_contract():
    anz = cp.Ansatz(3, "cp", fill_layers(chain_layer(3), 6))
    pl = ProgramLoss(anz.program, Loss("hs", u_toff3))
    pen = Penalty("piecewise", 0.002, PF.segments, PF.period)
    a0 = O.generate_initial_angles(0, anz.num_angles, anz.cp_mask, batch_size=5)
    res = mynimize_repeated(pl, anz.num_angles, learning_rate=0.1, num_iterations=50, initial_params_batch=a0,
                            regularization_func=pen, keep_history=False)
    assert len(res) == 5
    r = res[2]
    assert set(r) == {"params", "loss", "reg", "regloss"}
    assert r["params"].shape == (2, anz.num_angles) and r["regloss"].shape == (2,)
    assert isinstance(r["params"], np.ndarray)
    assert np.array_equal(r["params"][0], a0[2])
    assert np.allclose(r["loss"] + r["reg"], r["regloss"])
    assert r["regloss"][1] <= r["regloss"][0]
    # oracle agreement of the whole returned structure (f32, short horizon)
    ores = O.mynimize_repeated(3, O.ansatz_program(O.cp_ansatz(chain_layer(3), 6)), "hs",
                               torch.tensor(u_toff3), torch.tensor(a0), 0.1, 4,
                               anz.cp_mask, 0.002, O.make_regularization_function())
    res12 = mynimize_repeated(pl, learning_rate=0.1, num_iterations=4, initial_params_batch=a0,
                              regularization_func=pen, keep_history=False)
    for a, b in zip(res12, ores):
        assert np.abs(a["regloss"] - b["regloss"].numpy()).max() < 2e-4
        assert np.abs(a["reg"] - b["reg"].numpy()).max() < 2e-4
    # history mode and single-vector input
    h = mynimize_repeated(pl, learning_rate=0.1, num_iterations=20, initial_params_batch=a0[0],
                          regularization_func=pen, keep_history=True)
    assert isinstance(h, dict) and h["params"].shape == (20, anz.num_angles) and h["loss"].shape == (20,)
    assert np.array_equal(h["params"][0], a0[0])
    ph, lh = mynimize(pl, learning_rate=0.1, num_iterations=20, initial_params=a0[0], keep_history=True)
    h0 = mynimize_repeated(pl, learning_rate=0.1, num_iterations=20, initial_params_batch=a0[0], keep_history=True)
    assert set(h0) == {"params", "loss"}
    assert ph.shape == (20, anz.num_angles) and np.array_equal(lh, h0["loss"]) and np.array_equal(ph, h0["params"])
    # num_repeats without initial conditions; unitary_learn keys
    rr = mynimize_repeated(pl, num_repeats=3, num_iterations=5, keep_history=False)
    assert len(rr) == 3 and set(rr[0]) == {"params", "loss"}
    ul = unitary_learn(anz, u_toff3, num_repeats=2, num_iterations=10, keep_history=True)
    assert len(ul) == 2 and set(ul[0]) == {"params", "loss", "reg", "regloss"} and np.all(ul[0]["reg"] == 0)
    with pytest.raises(NotImplementedError):
        mynimize_repeated(pl, method="natural adam")
    with pytest.raises(TypeError):
        mynimize_repeated(lambda a: 0.0, 3)


def test_cz_value_count_and_filter_paths_agree():
    a = np.array([0.1, 3.2, 1.0, 6.2, 3.0, -0.05, 6.30, 9.5], dtype=np.float32)
    assert list(CU.cz_value(a, 0.2)) == list(O.cz_value(a, 0.2))
    assert CU.count_cz(a, 0.2) == O.count_cz(a, 0.2)
    assert CU.project_cp_angle(3.2) == math.pi and CU.project_cp_angle(0.1) == 0 and CU.project_cp_angle(6.2) == 0
    assert np.array_equal(CU.insert_params(np.array([0., 1, 2, 3]), np.array([-1., -2, -4]), [0, 2, 4]),
                          [-1, 0, -2, 1, -4, 2, 3])
    anz = cp.Ansatz(3, "cp", fill_layers(connected_layer(3), 7))
    pl = ProgramLoss(anz.program, Loss("hs", u_toff3))
    pen = Penalty("piecewise", 0.0013, PF.segments, PF.period)
    raw = mynimize_repeated(pl, learning_rate=0.1, num_iterations=600, regularization_func=pen, keep_history=False,
                            initial_params_batch=anz.program.initial_angles(3, 64), return_device=True)
    assert isinstance(raw, RawResults)
    fast = CU.filter_cp_results(raw, anz.cp_mask, 13, 1e-2, program=anz.program)
    slow = CU.filter_cp_results(list(raw.numpy()), anz.cp_mask, 13, 1e-2)
    assert [c for c, _ in fast] == [c for c, _ in slow] and len(fast) > 0
    for (c1, r1), (c2, r2) in zip(fast, slow):
        assert np.array_equal(r1["params"].cpu().numpy(), r2["params"])
    assert [c for c, _ in fast] == sorted(c for c, _ in fast)
    # oracle's selection on the same histories
    osel = O.filter_cp_results(list(raw.numpy()), anz.cp_mask, 13, 1e-2)
    assert [c for c, _ in osel] == [c for c, _ in slow]


def test_verify_and_convert():
    anz = cp.Ansatz(3, "cp", fill_layers(connected_layer(3), 7))
    opts = cp.StaticOptions(num_cp_gates=7, accepted_num_cz_gates=8, num_gd_iterations_at_verification=1500)
    pl = ProgramLoss(anz.program, Loss("hs", u_toff3))
    pen = Penalty("piecewise", 0.0013, PF.segments, PF.period)
    raw = mynimize_repeated(pl, learning_rate=0.1, num_iterations=1500, regularization_func=pen,
                            keep_history=False, initial_params_batch=anz.program.initial_angles(1, 48))
    sel = CU.filter_cp_results(raw, anz.cp_mask, 8, 1e-3)
    assert sel, "no prospective result among 48 samples (expected ~90%)"
    cz, res = sel[0]
    success, num_cz, circ, u, best = CU.verify_cp_result(res, anz, Loss("hs", u_toff3), opts)
    assert num_cz == cz
    circ_f, u_f, free = CU.convert_cp_to_cz(anz, res["params"][1], 0.2)
    assert len(free) == len(best) and len(free) < anz.num_angles
    if success:
        assert hst(u(best).astype(complex), u_toff3) < 5e-6
        qc = circ(best)
        assert qc.count_ops().get("cp", 0) == 7
    batch = CU.verify_cp_results([r for _, r in sel[:5]], anz, Loss("hs", u_toff3), opts)
    assert batch[0][0] == success and batch[0][1] == num_cz
    assert np.allclose(batch[0][4], best, atol=1e-6)


def test_static_ccz_readme_example(tmp_path):
    """BASELINE configs[0]: CCZ on the chain 0-1-2, 12 CP gates, 10 samples (README.md:30-45)."""
    syn = cp.Synthesize(chain_layer(3), target_unitary=CCZ, label="ccz_chain")
    opts = cp.StaticOptions(num_cp_gates=12, accepted_num_cz_gates=10, num_samples=10)
    res = syn.static(opts, save_to=str(tmp_path / "ccz"))
    assert len(res.decompositions) >= 1
    for d in res.decompositions:
        assert d.cz_count <= 10 and d.loss <= 1e-5
        assert hst(d.unitary, CCZ) < 1e-5
        assert d.cz_count == d.circuit.count_ops()["cz"] and 1 <= d.cz_depth <= d.cz_count
        assert set(d.circuit.count_ops()) <= {"rz", "rx", "cz"}
        assert "CZ count" in repr(d) and d.type == "Approximate" and d._static_options is opts
        # batched construction (one cpf_unitary for all verified results, lazy gate list) == the gate-list route
        from cpflow_b200.circuit import gates_depth
        assert d.cz_depth == gates_depth(["cz"], d.circuit)
        assert hst(d.circuit.unitary(), d.unitary) < 1e-12
        assert abs(d.loss - Loss("hs", CCZ)(d.circuit.unitary())) < 1e-9      # 1 - |t|^2 / 64 cancels to ~1e-7
        u_func, circ_func, free = d._cp_data
        assert hst(u_func(free).astype(complex), d.unitary) < 1e-5 and circ_func(free).count_ops()["cp"] == 12
    import pickle
    back = pickle.loads(pickle.dumps(res))           # plain pickle is enough (no closures in _cp_data)
    assert [x.cz_count for x in back.decompositions] == [x.cz_count for x in res.decompositions]
    assert min(d.cz_count for d in res.decompositions) <= 8   # README shows an 8-CZ result
    back = cp.Results.load(str(tmp_path / "ccz"))
    assert len(back.decompositions) == len(res.decompositions)
    assert back.decompositions[0].cz_count == res.decompositions[0].cz_count
    # a second call appends to the saved results (main.py:614-618, 683)
    res2 = syn.static(cp.StaticOptions(num_cp_gates=12, accepted_num_cz_gates=10, num_samples=4, random_seed=1),
                      save_to=str(tmp_path / "ccz"))
    assert len(res2.decompositions) >= len(res.decompositions)


def _score(cz_counts, n):
    return -math.log2(sum(2.0 ** (-c) for c in cz_counts) / n)


# Files written by the released reference (the pre-release `toff4_*_xyz` files used an older template
# layout and converge less often at large K).
@pytest.mark.parametrize("file,k,B_ref", [("paper/results/toff3_conn_xyz", 7, 200),
                                          ("paper/results/toff3_chain_xyz", 14, 200),
                                          ("tutorial/results/toff4_star", 30, 500),
                                          ("tutorial/results/toff4_star", 29, 500)])
def test_statistical_parity_with_stored_trials(trials, file, k, B_ref):
    rec = trials[file]
    cand = [t for t in rec["trials"] if isinstance(t["cz_counts"], list) and (k is None or t["num_cp_gates"] == k)]
    t = max(cand, key=lambda t: len(t["cz_counts"]))
    from cpflow_b200.topology import num_qubits_from_layer
    n = num_qubits_from_layer(rec["layer"])
    target = u_toff3 if n == 3 else u_toff4
    syn = cp.Synthesize(rec["layer"], target_unitary=target)
    B = 4 * B_ref
    opts = cp.StaticOptions(num_cp_gates=t["num_cp_gates"], r=t["r"], accepted_num_cz_gates=10 ** 6,
                            num_samples=B, random_seed=t["random_seed"])
    anz, candidates = syn._prospective(opts)
    cz = [int(c) for c in candidates[:, 1].tolist()]
    p_ref, p_new = len(t["cz_counts"]) / B_ref, len(cz) / B
    sigma = math.sqrt(p_ref * (1 - p_ref) / B_ref + p_new * (1 - p_new) / B + 1e-6)
    assert abs(p_ref - p_new) < 4 * sigma + 0.02, (p_ref, p_new)
    assert cz == sorted(cz)
    # our 4x larger sample must reach the reference's minimum CZ count (or better by at most 2)
    assert min(t["cz_counts"]) - 2 <= min(cz) <= min(t["cz_counts"]) + (1 if n == 4 else 0)
    assert abs(_score(cz, B) - t["score"]) < 0.8, (_score(cz, B), t["score"])


def test_static_toffoli3_finds_known_optimum():
    """Toffoli-3 on all-to-all connectivity: best known count 6 CZ (paper/CPFlow.tex:413), hit rate
    ~28% of samples at K=7, r=0.00131 (paper/CPFlow.tex:419)."""
    syn = cp.Synthesize(connected_layer(3), target_unitary=u_toff3, label="t3")
    opts = cp.StaticOptions(num_cp_gates=7, r=0.00131, accepted_num_cz_gates=6, num_samples=400)
    res = syn.static(opts, save_results=False)
    counts = [d.cz_count for d in res.decompositions]
    assert counts and min(counts) == 6
    assert 0.12 < len(counts) / 400 < 0.45
    d = res.decompositions[0]
    assert hst(d.unitary, u_toff3) < 1e-5


# Best known CZ counts of the 4-qubit Toffoli per topology (paper/CPFlow.tex:480) at hyper-parameter points the
# reference's own stored trials found productive (tests/golden/trials.json); profiles/toff4_best_r2.txt keeps a full run.
@pytest.mark.parametrize("name,layer,K,r,best,B", [
    ("star", [[0, 1], [0, 2], [0, 3]], 26, 0.000793, 16, 20000),
    ("chain", chain_layer(4), 26, 0.000256, 18, 60000),
    ("kite", [[0, 1], [1, 2], [2, 3], [1, 3]], 25, 0.000648, 14, 20000),
    ("square", [[0, 1], [1, 2], [2, 3], [3, 0]], 24, 0.000528, 16, 20000),
    ("connected", connected_layer(4), 23, 0.000528, 14, 20000)])
def test_static_toffoli4_reaches_best_known_counts(name, layer, K, r, best, B):
    """Full Synthesize.static() runs reproduce the reference's minimal CZ counts on the benchmark target of the
    headline metric (north_star: 'identical best CZ counts'): 16 (star) / 18 (chain) / 14 (kite) / 16 (square) /
    14 (connected).  Kite and square run on the run-time-pair Heisenberg kernel."""
    syn = cp.Synthesize(layer, target_unitary=u_toff4, label=f"toff4_{name}")
    opts = cp.StaticOptions(num_cp_gates=K, r=r, accepted_num_cz_gates=best, num_samples=B)
    res = syn.static(opts, save_results=False)
    counts = sorted(d.cz_count for d in res.decompositions)
    assert counts and counts[0] <= best, (name, counts[:5], len(syn.last_prospective_cz_counts))
    assert counts[0] >= best - 1          # a count below the best known one would be news: look at it before trusting it
    for d in res.decompositions:
        assert hst(d.unitary, u_toff4) < 1e-5 and d.cz_count == d.circuit.count_ops()["cz"]
