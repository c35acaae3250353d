"""Minimal gate-list circuit IR (replaces the qiskit objects the reference attaches to results).

The reference returns qiskit `QuantumCircuit`s (main.py:193-222, 261-291) and post-processes them
with qiskit passes (exact_decompositions.py:42-74, 142-190, 280-290).  qiskit is not a dependency
here; this module keeps what the results API needs: an ordered gate list over {rz, rx, ry, cp, cz,
cx, h, ...}, gate counting, CZ depth, the CP -> CZ rewriting and the ZXZ merge of single-qubit
gates, OpenQASM export, and evaluation of the circuit's unitary on the CUDA engine.

Conventions follow the rest of the package (and the stored reference results): qubit 0 is the most
significant bit, R_s(a) = exp(-i a s / 2), CP(a) = diag(1, 1, 1, e^{ia}).
"""
import cmath
import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib as L

_KIND = {'rx': L.RX, 'ry': L.RY, 'rz': L.RZ, 'cp': L.CP, 'cz': L.CZ, 'cx': L.CX}
TWO_PI = 2 * math.pi


@dataclass
class Gate:
    name: str
    qubits: tuple
    params: tuple = ()

    def __repr__(self):
        p = f"({', '.join(f'{x:.6g}' for x in self.params)})" if self.params else ''
        return f"{self.name}{p} q{list(self.qubits)}"


def _su2(name, a):
    c, s = math.cos(a / 2), math.sin(a / 2)
    if name == 'rx':
        return np.array([[c, -1j * s], [-1j * s, c]])
    if name == 'ry':
        return np.array([[c, -s], [s, c]], dtype=complex)
    if name == 'rz':
        return np.array([[c - 1j * s, 0], [0, c + 1j * s]])
    if name == 'h':
        return np.array([[1, 1], [1, -1]], dtype=complex) / math.sqrt(2)
    raise ValueError(name)


def _wrap(a):
    """angle -> (-pi, pi]"""
    a = math.fmod(a, TWO_PI)
    if a > math.pi:
        a -= TWO_PI
    elif a <= -math.pi:
        a += TWO_PI
    return a


def zxz_angles(u, tol=1e-12):
    """(z1, x, z2) with u ~ Rz(z2) Rx(x) Rz(z1) up to a global phase (time order rz(z1), rx, rz(z2));
    the job of OneQubitEulerDecomposer('ZXZ') in exact_decompositions.py:142-156."""
    u = np.asarray(u, dtype=complex)
    u = u / cmath.sqrt(np.linalg.det(u))     # SU(2): [[al, -conj(be)], [be, conj(al)]]
    al, be = u[0, 0], u[1, 0]
    x = 2 * math.atan2(abs(be), abs(al))
    if abs(be) < tol:                         # diagonal: a single rz
        return _wrap(-2 * cmath.phase(al)), 0.0, 0.0
    if abs(al) < tol:                         # anti-diagonal: rx(pi) and one rz
        return _wrap(-2 * (cmath.phase(be) + math.pi / 2)), math.pi, 0.0
    s = -2 * cmath.phase(al)                  # z1 + z2
    d = -2 * (cmath.phase(be) + math.pi / 2)  # z1 - z2
    return _wrap((s + d) / 2), x, _wrap((s - d) / 2)


class Circuit:
    """Ordered gate list on `num_qubits` qubits."""

    def __init__(self, num_qubits, data=None, global_phase=0.0):
        self.num_qubits = int(num_qubits)
        self.data = list(data or [])
        self.global_phase = float(global_phase)

    # ---- construction (qiskit-like spelling so reference user code reads the same) ----
    def append(self, name, qubits, params=()):
        qubits = tuple(int(q) for q in (qubits if isinstance(qubits, (list, tuple)) else [qubits]))
        if any(q < 0 or q >= self.num_qubits for q in qubits):
            raise ValueError(f"qubit out of range in {name}{qubits}")
        self.data.append(Gate(name, qubits, tuple(float(p) for p in params)))
        return self

    def rx(self, a, q): return self.append('rx', [q], [a])
    def ry(self, a, q): return self.append('ry', [q], [a])
    def rz(self, a, q): return self.append('rz', [q], [a])
    def h(self, q): return self.append('h', [q])
    def cp(self, a, q0, q1): return self.append('cp', [q0, q1], [a])
    def cz(self, q0, q1): return self.append('cz', [q0, q1])
    def cx(self, q0, q1): return self.append('cx', [q0, q1])

    def copy(self):
        return Circuit(self.num_qubits, [Gate(g.name, g.qubits, g.params) for g in self.data], self.global_phase)

    def __len__(self):
        return len(self.data)

    def __repr__(self):
        ops = self.count_ops()
        return f"<Circuit {self.num_qubits}q, {len(self.data)} gates {ops}>"

    # ---- inspection (exact_decompositions.py:273-290) ----
    def count_ops(self):
        out = {}
        for g in self.data:
            out[g.name] = out.get(g.name, 0) + 1
        return out

    def depth(self, filter_function=None):
        """Circuit depth counting only gates for which filter_function(gate) is true."""
        level = [0] * self.num_qubits
        for g in self.data:
            if filter_function is not None and not filter_function(g):
                continue
            lv = max(level[q] for q in g.qubits) + 1
            for q in g.qubits:
                level[q] = lv
        return max(level) if level else 0

    # ---- lowering to the engine ----
    def to_ops(self):
        """Constant-angle primitive ops for cpf_program_create ('h' is lowered to rz rx rz)."""
        ops = []
        for g in self.data:
            if g.name in ('rx', 'ry', 'rz'):
                ops.append((_KIND[g.name], g.qubits[0], -1, -1, g.params[0]))
            elif g.name == 'cp':
                ops.append((L.CP, g.qubits[0], g.qubits[1], -1, g.params[0]))
            elif g.name in ('cz', 'cx'):
                ops.append((_KIND[g.name], g.qubits[0], g.qubits[1], -1, 0.0))
            elif g.name == 'h':
                for k in (L.RZ, L.RX, L.RZ):
                    ops.append((k, g.qubits[0], -1, -1, math.pi / 2))
            else:
                raise ValueError(f"gate {g.name!r} cannot be lowered to the engine")
        return ops

    def program(self):
        from .engine import Program
        return Program(self.num_qubits, self.to_ops(), 0)

    def unitary(self, dtype=None, device='cuda'):
        """The circuit's unitary (numpy complex128, modulo global phase) evaluated by cpf_unitary."""
        import torch
        dtype = dtype or torch.float64
        u = self.program().unitary(torch.zeros(1, 0, dtype=dtype, device=device))
        return u[0].cpu().numpy()

    # ---- export ----
    def qasm(self):
        lines = ['OPENQASM 2.0;', 'include "qelib1.inc";', f'qreg q[{self.num_qubits}];']
        for g in self.data:
            p = f"({','.join(repr(x) for x in g.params)})" if g.params else ''
            name = 'cu1' if g.name == 'cp' else g.name
            lines.append(f"{name}{p} {','.join(f'q[{q}]' for q in g.qubits)};")
        return '\n'.join(lines) + '\n'


def gate_filter(gate_names, gate):
    return gate.name in gate_names


def gates_count(gate_names, circuit):
    """exact_decompositions.py:280-286."""
    ops = circuit.count_ops()
    return sum(ops.get(n, 0) for n in gate_names)


def gates_depth(gate_names, circuit):
    """exact_decompositions.py:289-290."""
    return circuit.depth(filter_function=lambda g: gate_filter(gate_names, g))


def cp_to_cz_circuit(circuit, cp_threshold=0.2):
    """exact_decompositions.py:42-74: CP(a) with |a| <= thr -> nothing, |a - pi| <= thr -> CZ, else
    the two-CZ form  CP(a) ~ Rz_0(a/2) Rz_1(a/2) . H_1 CZ Rx_1(-a/2) CZ H_1  (global phase dropped;
    the reference gets an equivalent form from qiskit's transpiler)."""
    out = Circuit(circuit.num_qubits, global_phase=circuit.global_phase)
    for g in circuit.data:
        if g.name != 'cp':
            out.data.append(Gate(g.name, g.qubits, g.params))
            continue
        a = g.params[0]
        q0, q1 = g.qubits
        if abs(a) <= cp_threshold:
            continue
        if abs(a - math.pi) <= cp_threshold:
            out.cz(q0, q1)
            continue
        out.h(q1).cz(q0, q1).rx(-a / 2, q1).cz(q0, q1).h(q1)
        out.rz(a / 2, q0).rz(a / 2, q1)
        out.global_phase += a / 4
    return out


def cp_template_cz_count_depth(all_placements, cp_angles, num_qubits, cp_threshold=1e-6):
    """CZ count and CZ depth of `cp_to_cz_circuit(template)` for a batch of CP-angle vectors [B,K] without building
    the circuits: CP(a) contributes 0 / 1 / 2 CZ gates by the rule above (on the raw angle), and the two CZ of the
    two-CZ form sit back to back on the same pair, so the depth is a per-qubit level scan over the blocks
    (`Circuit.depth` restricted to CZ gates)."""
    a = np.asarray(cp_angles, dtype=np.float64)
    mult = np.where(np.abs(a) <= cp_threshold, 0, np.where(np.abs(a - math.pi) <= cp_threshold, 1, 2))
    level = np.zeros((a.shape[0], num_qubits), dtype=np.int64)
    for k, (q0, q1) in enumerate(all_placements):
        m = mult[:, k]
        lv = np.maximum(level[:, q0], level[:, q1]) + m
        on = m > 0
        level[on, q0] = lv[on]
        level[on, q1] = lv[on]
    depth = level.max(1) if num_qubits else np.zeros(a.shape[0], dtype=np.int64)
    return mult.sum(1), depth


def convert_to_ZXZ(circuit, drop_tol=1e-9):
    """exact_decompositions.py:176-190: merge every run of single-qubit gates into one SU(2) and
    re-express it as rz rx rz, dropping rotations whose angle is 0 mod 2 pi."""
    n = circuit.num_qubits
    pending = [None] * n
    out = Circuit(n, global_phase=circuit.global_phase)

    def flush(q):
        u = pending[q]
        pending[q] = None
        if u is None:
            return
        z1, x, z2 = zxz_angles(u)
        for name, a in (('rz', z1), ('rx', x), ('rz', z2)):
            if abs(_wrap(a)) > drop_tol:
                out.append(name, [q], [a])

    for g in circuit.data:
        if len(g.qubits) == 1:
            m = _su2(g.name, g.params[0] if g.params else 0.0)
            q = g.qubits[0]
            pending[q] = m if pending[q] is None else m @ pending[q]
        else:
            for q in g.qubits:
                flush(q)
            out.data.append(Gate(g.name, g.qubits, g.params))
    for q in range(n):
        flush(q)
    return out
