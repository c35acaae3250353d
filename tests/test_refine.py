"""Refine path: the batched engine version of reduce_all_1q_angles against the
sequential oracle restatement of exact_decompositions.py:77-113, and Decomposition.refine end to end."""
import math

import numpy as np
import pytest
import torch

from oracle import cpflow_oracle as O


def _circuit_ops(qc):
    kind = {"rx": O.RX, "ry": O.RY, "rz": O.RZ}
    ops, angles, wires = [], [], []
    for g in qc.data:
        if g.name in kind:
            ops.append((kind[g.name], g.qubits[0], -1, len(angles), 0.0))
            angles.append(g.params[0]); wires.append(g.qubits[0])
        else:
            ops.append((O.CZ, g.qubits[0], g.qubits[1], -1, 0.0))
    return ops, np.array(angles), wires


def test_oracle_reduce_merges_and_zeroes_redundant_rotations():
    """CPU: rz(a) rz(b) on one wire merge, an identity rotation pair cancels, a needed angle stays."""
    n = 2
    ops = [(O.RZ, 0, -1, 0, 0.0), (O.RZ, 0, -1, 1, 0.0), (O.RX, 1, -1, 2, 0.0), (O.CZ, 0, 1, -1, 0.0),
           (O.RX, 1, -1, 3, 0.0), (O.RX, 1, -1, 4, 0.0)]
    a = np.array([0.3, 0.5, 0.7, 1.1, -1.1])
    target = O.program_unitary_np(n, ops, a)
    loss = lambda x: 1 - abs(np.sum(O.program_unitary_np(n, ops, x) * target.conj())) ** 2 / 16
    r = O.reduce_all_1q_angles(loss, a, [0, 0, 1, 1, 1], 1e-9)
    assert np.allclose(r, [0.0, 0.8, 0.7, 0.0, 0.0])
    assert loss(r) < 1e-12


@pytest.mark.gpu
def test_batched_reduce_equals_sequential_oracle():
    import cpflow_b200 as cp
    from cpflow_b200 import exact_decompositions as ED
    from cpflow_b200.circuit import Circuit, convert_to_ZXZ, cp_to_cz_circuit
    from cpflow_b200.engine import Loss
    rng = np.random.default_rng(0)
    # a CP template with some CP angles at 0 / pi and redundant rotations, lowered like reduce_angles does
    qc = Circuit(3)
    for q in range(3):
        qc.rz(rng.uniform(-3, 3), q).rx(rng.uniform(-3, 3), q).rz(rng.uniform(-3, 3), q)
    for (p0, p1), a in zip([(0, 1), (1, 2), (0, 2), (0, 1)], [math.pi, 0.004, math.pi - 0.003, 1.3]):
        qc.cp(a, p0, p1)
        qc.rx(rng.uniform(-3, 3), p0).rz(rng.uniform(-3, 3), p0).rz(0.4, p1).rz(-0.4, p1)
    target = qc.unitary()
    loss = Loss("hs", target)
    low = convert_to_ZXZ(cp_to_cz_circuit(qc, cp_threshold=0.01))
    prog, angles, wires = ED.circuit_angle_program(low)
    got = ED.reduce_all_1q_angles(prog, loss, angles, wires, threshold=1e-5)
    ops, a0, w0 = _circuit_ops(low)
    assert w0 == wires and np.array_equal(a0, angles)
    oloss = lambda x: 1 - abs(np.sum(O.program_unitary_np(3, ops, x) * target.conj())) ** 2 / 64
    want = O.reduce_all_1q_angles(oloss, angles, wires, 1e-5)
    assert np.allclose(got, want, atol=1e-12)
    assert (np.abs(want) < 1e-12).sum() > (np.abs(angles) < 1e-12).sum()      # something was reduced


@pytest.mark.gpu
def test_decomposition_refine_end_to_end():
    """Toffoli-3 on all-to-all connectivity: a 6-CZ static() result refines to rational angles
    (multiples of pi/4, paper/CPFlow.tex Fig. 1) and still implements the target."""
    import cpflow_b200 as cp
    from cpflow_b200.gates import u_toff3
    from cpflow_b200.topology import connected_layer
    from conftest import hst
    syn = cp.Synthesize(connected_layer(3), target_unitary=u_toff3, label="t3")
    opts = cp.StaticOptions(num_cp_gates=7, r=0.00131, accepted_num_cz_gates=6, num_samples=200)
    res = syn.static(opts, save_results=False)
    assert res.decompositions
    kinds = set()
    for d in res.decompositions[:8]:
        n_rot = sum(g.name in ("rx", "ry", "rz") for g in d.circuit.data)
        msg = d.refine()
        kinds.add(d.type)
        assert msg == f"Refined to {d.type}" and d.type in ("Approximate", "Rational")
        assert hst(d.circuit.unitary(), u_toff3) < 1e-5
        assert d.circuit.count_ops().get("cz", 0) == 6
        assert sum(g.name in ("rx", "ry", "rz") for g in d.circuit.data) <= n_rot
        if d.type == "Rational":
            for g in d.circuit.data:
                if g.name in ("rx", "rz"):
                    k = g.params[0] / (math.pi / 32)
                    assert abs(k - round(k)) < 1e-9
    assert "Rational" in kinds
