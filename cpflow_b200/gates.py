"""Gate vocabulary and benchmark targets (mirror of reference cpflow/gates.py).

Gate matrices live in the CUDA engine (csrc/engine.cuh); here are only the names, the kind codes of
the C ABI and the closed-form Toffoli targets that the reference builds with qiskit
(gates.py:95-106: `mct` + `reverse_bits` = identity with the last two basis states swapped).
"""
import numpy as np

from ._lib import RX, RY, RZ, CP, CZ, CX  # noqa: F401

GATE_KINDS = {'rx': RX, 'ry': RY, 'rz': RZ, 'cp': CP, 'cz': CZ, 'cx': CX}
GATE_QUBITS = {'rx': 1, 'ry': 1, 'rz': 1, 'cp': 2, 'cz': 2, 'cx': 2}


def gate_kind(name):
    if name not in GATE_KINDS:
        raise TypeError(f"Gate '{name}' not implemented.")  # same error as reference gates.py:82-83
    return GATE_KINDS[name]


def toffoli(num_qubits):
    n = 2 ** num_qubits
    u = np.eye(n, dtype=np.complex128)
    u[[n - 2, n - 1]] = u[[n - 1, n - 2]]
    return u


u_toff3 = toffoli(3)
u_toff4 = toffoli(4)
u_toff5 = toffoli(5)
