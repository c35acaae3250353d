"""CPU tests of the gate-list circuit IR, the CP -> CZ rewriting, the ZXZ merge, the hyper-parameter
sampler and the options / results objects.  Circuit unitaries are checked with the ORACLE's numpy
simulator (tests may use the oracle; the product evaluates circuits on the CUDA engine)."""
import math
import os
import pickle

import numpy as np
import pytest

from oracle import cpflow_oracle as O
from conftest import hst
from cpflow_b200 import circuit as CI
from cpflow_b200.ansatz import Ansatz
from cpflow_b200.topology import chain_layer, connected_layer, fill_layers


def sim(qc):
    return O.program_unitary_np(qc.num_qubits, [tuple(o) for o in qc.to_ops()], np.zeros(0))


def test_ansatz_circuit_matches_program():
    anz = Ansatz(3, "cp", fill_layers(connected_layer(3), 5), "xyz")
    a = np.random.default_rng(0).uniform(0, 6.28, anz.num_angles)
    qc = anz.circuit(a)
    assert qc.count_ops() == {"rz": 6 + 10, "rx": 3 + 10, "ry": 10, "cp": 5}
    u = O.program_unitary_np(3, [tuple(o) for o in anz.ops], a)
    assert np.abs(sim(qc) - u).max() < 1e-13


@pytest.mark.parametrize("a", [0.0, 5e-7, math.pi, math.pi - 5e-7, 0.3, 1.7, -2.2, 5.9, 2 * math.pi])
def test_cp_to_cz_equivalence(a):
    """exact_decompositions.py:42-74 semantics with cp_threshold=1e-6 (main.py:284)."""
    qc = CI.Circuit(3).rx(0.3, 0).ry(1.1, 2).cp(a, 2, 0).rz(0.7, 2)
    new = CI.cp_to_cz_circuit(qc, cp_threshold=1e-6)
    assert "cp" not in new.count_ops()
    ncz = new.count_ops().get("cz", 0)
    assert ncz == (0 if abs(a) <= 1e-6 else (1 if abs(a - math.pi) <= 1e-6 else 2))
    assert hst(sim(new), sim(qc)) < 1e-10


def test_convert_to_zxz_merges_and_preserves():
    rng = np.random.default_rng(1)
    qc = CI.Circuit(3)
    for _ in range(30):
        k = rng.integers(0, 5)
        if k < 3:
            qc.append(["rx", "ry", "rz"][k], [int(rng.integers(3))], [rng.uniform(-7, 7)])
        elif k == 3:
            q = rng.choice(3, 2, replace=False)
            qc.cz(int(q[0]), int(q[1]))
        else:
            qc.h(int(rng.integers(3)))
    new = CI.convert_to_ZXZ(qc)
    assert set(new.count_ops()) <= {"rz", "rx", "cz"}
    assert hst(sim(new), sim(qc)) < 1e-10
    # between two entanglers every wire carries at most rz rx rz
    run = [0] * 3
    for g in new.data:
        if len(g.qubits) == 1:
            run[g.qubits[0]] += 1
            assert run[g.qubits[0]] <= 3
        else:
            for q in g.qubits:
                run[q] = 0
    # special cases drop trivial rotations
    only_x = CI.convert_to_ZXZ(CI.Circuit(1).rx(0.4, 0).rx(0.5, 0))
    assert [g.name for g in only_x.data] == ["rx"] and abs(only_x.data[0].params[0] - 0.9) < 1e-12
    assert len(CI.convert_to_ZXZ(CI.Circuit(1).rz(0.4, 0).rz(-0.4, 0)).data) == 0
    for u in (np.array([[0, 1], [1, 0]], complex), np.diag([1, 1j]), np.eye(2)):
        z1, x, z2 = CI.zxz_angles(u)
        got = sim(CI.Circuit(1).rz(z1, 0).rx(x, 0).rz(z2, 0))
        assert hst(got, u) < 1e-12


def test_counts_and_depth_against_stored_circuits(gatelist_kats):
    """cz_count / cz_depth computed on the stored gate lists equal the stored values
    (main.py:268-269 via exact_decompositions.py:280-290)."""
    meta, arrs = gatelist_kats
    for m in meta[::2]:
        key = m["key"]
        qc = CI.Circuit(m["n"])
        for k, q0, q1, p in zip(m["kinds"], arrs["q0_" + key], arrs["q1_" + key], arrs["p_" + key]):
            if k == "cz":
                qc.cz(int(q0), int(q1))
            elif k in ("rx", "ry", "rz"):
                qc.append(k, [int(q0)], [float(p)])
            else:
                qc.append(k, [int(q0)] + ([int(q1)] if q1 >= 0 else []))
        assert CI.gates_count(["cz"], qc) == m["cz_count"]
        assert CI.gates_depth(["cz"], qc) == m["cz_depth"], key


def test_full_pipeline_reproduces_stored_cz_count(ansatz_kats):
    """template at the stored angles -> cp_to_cz (1e-6) -> ZXZ has the stored cz_count and the
    stored unitary (Decomposition._from_cp_circuit, main.py:281-291)."""
    meta, arrs = ansatz_kats
    for m in meta[::9]:
        anz = Ansatz(m["n"], "cp", fill_layers(m["layer"], m["num_cp_gates"]), m["rotation_gates"])
        ang = arrs["angles_" + m["key"]]
        qc = CI.convert_to_ZXZ(CI.cp_to_cz_circuit(anz.circuit(ang), 1e-6))
        assert CI.gates_count(["cz"], qc) == m["cz_count"]
        assert hst(sim(qc), arrs["u_" + m["key"]]) < 1e-10


def test_qasm_export():
    s = CI.Circuit(2).rz(0.5, 0).cz(0, 1).cp(0.25, 1, 0).qasm()
    assert "qreg q[2];" in s and "cz q[0],q[1];" in s and "cu1(0.25) q[1],q[0];" in s


def test_options_and_results_objects(tmp_path):
    import cpflow_b200 as cp
    with pytest.raises(TypeError):
        cp.StaticOptions(num_cp_gates=5)
    with pytest.raises(TypeError):
        cp.AdaptiveOptions(min_num_cp_gates=3)
    o = cp.StaticOptions(num_cp_gates=12, accepted_num_cz_gates=10)
    assert (o.num_samples, o.learning_rate, o.num_gd_iterations, o.r, o.entry_loss, o.target_loss,
            o.threshold_cp, o.learning_rate_at_verification, o.num_gd_iterations_at_verification,
            o.random_seed, o.rotation_gates, o.method, o.cp_distribution) == \
        (100, 0.1, 2000, 0.00055, 1e-3, 1e-6, 0.2, 0.01, 5000, 0, "xyz", "adam", "uniform")
    a = cp.AdaptiveOptions(min_num_cp_gates=3, max_num_cp_gates=20, num_samples=7)
    s = a.get_static(9, 0.002)
    assert (s.num_cp_gates, s.r, s.accepted_num_cz_gates, s.num_samples) == (9, 0.002, None, 7)
    ro = cp.RegularizationOptions
    assert (ro.function, ro.ymax, ro.plato_0) == ("linear", 2, 0.05)
    # results round-trip through the pickler, save_to default, loss spec picklable
    syn = cp.Synthesize(chain_layer(3), target_unitary=np.eye(8), label="idtest")
    assert syn.num_qubits == 3 and syn.unitary_loss_func.kind == "hs"
    res = cp.Results(syn.unitary_loss_func, syn.layer, label="idtest")
    assert res.save_to == "results/idtest"
    res.save_to = str(tmp_path / "sub" / "r")
    res.save()
    back = cp.Results.load(res.save_to)
    assert back.label == "idtest" and back.layer == syn.layer
    assert abs(back.loss_function(np.eye(8))) < 1e-15
    with pytest.raises(AssertionError):
        cp.Synthesize(chain_layer(3), target_unitary=np.eye(4))
    # any callable is accepted as unitary_loss_func, as in the reference (main.py:528-529); non-callables are not
    from cpflow_b200.engine import TorchLoss
    assert isinstance(cp.Synthesize(chain_layer(3), unitary_loss_func=lambda u: 0.0).unitary_loss_func, TorchLoss)
    with pytest.raises(TypeError):
        cp.Synthesize(chain_layer(3), unitary_loss_func=3.0)
    # plain pickle (no dill) round trip of the picklable constrained callables stored in Decomposition._cp_data
    import pickle
    from cpflow_b200.cp_utils import _constrained_funcs, insert_params
    from cpflow_b200.ansatz import Ansatz
    from cpflow_b200.topology import fill_layers
    anz = Ansatz(3, "cp", fill_layers(chain_layer(3), 4))
    circ, u = _constrained_funcs(anz, np.array([0.0, np.pi], dtype=np.float32), [15, 22])
    circ2, u2 = pickle.loads(pickle.dumps((circ, u)))
    free = np.arange(anz.num_angles - 2, dtype=np.float32)
    assert circ2(free).count_ops() == circ(free).count_ops() and u2.indices == [15, 22]
    assert np.array_equal(circ2.full(free), insert_params(free, [0.0, np.float32(np.pi)], [15, 22]))


def test_seed_chain_and_tpe(trials):
    from cpflow_b200.hyper import TPESampler, next_seed
    rec = trials["paper/results/toff3_conn_xyz"]["trials"]
    seeds = [t["random_seed"] for t in rec]
    assert seeds[0] == next_seed(0)
    for a, b in zip(seeds[:60], seeds[1:61]):
        assert next_seed(a) == b
    sp = TPESampler(3, 30, 0.00055, 0.5, n_startup=10)
    rng = np.random.default_rng(0)
    res = []
    for _ in range(60):
        k, r = sp.suggest(res, rng)
        assert 3 <= k <= 30 and r > 0
        res.append({"num_cp_gates": k, "r": r, "loss": abs(k - 14) + abs(math.log(r / 0.001))})
    early = np.mean([t["loss"] for t in res[:10]])
    late = np.mean([t["loss"] for t in res[-20:]])
    assert late < early   # proposals concentrate on the good region


def test_batched_cz_count_and_depth_match_the_built_circuits():
    """`cp_template_cz_count_depth` (used by Synthesize.static() for all verified results at once) against the
    circuits built gate by gate (exact_decompositions.py:42-74, 280-290)."""
    from cpflow_b200.ansatz import Ansatz
    from cpflow_b200.topology import fill_layers
    rng = np.random.default_rng(3)
    for n, layer, K in [(3, [[0, 1], [1, 2]], 9), (4, [[0, 1], [0, 2], [0, 3]], 14), (4, [[3, 1], [2, 0], [1, 2]], 11)]:
        anz = Ansatz(n, "cp", fill_layers(layer, K))
        cp_idx = np.flatnonzero(anz.cp_mask)
        A = rng.uniform(0, 2 * np.pi, (25, anz.num_angles)).astype(np.float32)
        for b in range(25):
            k = rng.choice(cp_idx, len(cp_idx) // 2, replace=False)
            A[b, k] = rng.choice(np.array([0.0, np.pi], dtype=np.float32), len(k))
        A[0, cp_idx[0]] = 5e-7          # inside the 1e-6 window
        A[1, cp_idx[1]] = np.float32(2 * np.pi)   # NOT folded: the rule acts on the raw angle
        cnt, dep = CI.cp_template_cz_count_depth(anz.all_placements, A[:, cp_idx], n)
        for b in range(25):
            qc = CI.convert_to_ZXZ(CI.cp_to_cz_circuit(anz.circuit(A[b]), 1e-6))
            assert cnt[b] == CI.gates_count(["cz"], qc) and dep[b] == CI.gates_depth(["cz"], qc)
