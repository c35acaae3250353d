"""Template circuits (mirror of reference cpflow/main.py:23-239: EntanglingBlock, split_angles,
build_unitary, Ansatz) compiled to a declarative gate program for the CUDA engine."""
import numpy as np
import torch

from . import _lib as L
from .engine import Program
from .gates import gate_kind

_ROT = {'x': L.RX, 'y': L.RY, 'z': L.RZ}


def block_num_angles(entangling_gate_name, rotation_gates):
    """EntanglingBlock.get_num_angles (reference main.py:32-34)."""
    return 2 * len(rotation_gates) + (entangling_gate_name == 'cp')


def split_angles(angles, num_qubits, num_block_angles, layer_len=0, num_layers=0):
    """Reference main.py:85-103 (index bookkeeping only)."""
    angles = np.asarray(angles)
    surface_angles = angles[:3 * num_qubits].reshape(num_qubits, 3)
    block_angles = angles[3 * num_qubits:].reshape(-1, num_block_angles)
    if num_layers is None:
        layers_angles = []
    else:
        layers_angles = block_angles[:layer_len * num_layers].reshape(num_layers, layer_len, num_block_angles)
    free_block_angles = block_angles[layer_len * num_layers:]
    cp_angles = [b[-1] for b in block_angles] if num_block_angles % 2 == 1 else []
    return {'surface angles': surface_angles, 'block angles': block_angles, 'layers angles': layers_angles,
            'free block angles': free_block_angles, 'cp angles': cp_angles}


def ansatz_ops(num_qubits, entangling_gate_name, rotation_gates, all_placements):
    """Time-ordered primitive gates of build_unitary (reference main.py:106-146): surface round
    Rz(a0) Rx(a1) Rz(a2) per qubit (main.py:122-124), then per block the entangler followed by, for
    each letter, R(a[2j]) on placement[0] and R(a[2j+1]) on placement[1] (main.py:43-46, 69-82)."""
    n = num_qubits
    ek = gate_kind(entangling_gate_name)
    nb = block_num_angles(entangling_gate_name, rotation_gates)
    ops = []
    for q in range(n):
        ops += [(L.RZ, q, -1, 3 * q, 0.0), (L.RX, q, -1, 3 * q + 1, 0.0), (L.RZ, q, -1, 3 * q + 2, 0.0)]
    for b, (p0, p1) in enumerate(all_placements):
        base = 3 * n + nb * b
        ops.append((ek, p0, p1, base + nb - 1 if ek == L.CP else -1, 0.0))
        for j, letter in enumerate(rotation_gates):
            ops.append((_ROT[letter], p0, -1, base + 2 * j, 0.0))
            ops.append((_ROT[letter], p1, -1, base + 2 * j + 1, 0.0))
    return ops


class Ansatz:
    """Reference main.py:149-239.  `unitary(angles)` runs on the GPU through cpf_unitary."""

    def __init__(self, num_qubits, entangling_gate_name, placements, rotation_gates='xyz'):
        self.num_qubits = num_qubits
        self.entangling_gate_name = entangling_gate_name
        self.rotation_gates = rotation_gates
        placements = dict(placements)
        placements.setdefault('layers', [[], 0])
        placements.setdefault('free', [])
        self.placements = placements
        self.layer, self.num_layers = placements['layers']
        self.free_placements = placements['free']
        self.all_placements = [list(p) for p in list(self.layer) * self.num_layers + list(self.free_placements)]
        self.num_blocks = len(self.all_placements)
        nb = block_num_angles(entangling_gate_name, rotation_gates)
        self.num_block_angles = nb
        self.num_angles = 3 * num_qubits + nb * self.num_blocks
        if entangling_gate_name == 'cp':
            mask = np.zeros(self.num_angles, dtype=np.int32)
            mask[3 * num_qubits + nb - 1::nb] = 1  # main.py:181-184
            self.cp_mask = mask
        self.ops = ansatz_ops(num_qubits, entangling_gate_name, rotation_gates, self.all_placements)
        self._program = None

    @property
    def program(self):
        if self._program is None:
            self._program = Program(self.num_qubits, self.ops, self.num_angles)
        return self._program

    def unitary(self, angles, dtype=None, device='cuda'):
        """2^n x 2^n unitary at `angles` ([P] or [B,P]; numpy or torch)."""
        is_torch = isinstance(angles, torch.Tensor)
        a = angles if is_torch else torch.as_tensor(np.asarray(angles))
        if dtype is None:
            dtype = a.dtype if a.dtype in (torch.float32, torch.float64) else torch.float32
        single = a.dim() == 1
        a = a.reshape(-1, self.num_angles).to(device=device, dtype=dtype).contiguous()
        u = self.program.unitary(a)
        if single:
            u = u[0]
        return u if is_torch else u.cpu().numpy()

    def __getstate__(self):
        st = dict(self.__dict__)
        st['_program'] = None
        return st

    def circuit(self, angles):
        """Gate-list circuit of the template at `angles` (reference main.py:193-222, numeric angles
        only): rz rx rz per qubit, then per block cp and the rotation letters on both qubits."""
        from .circuit import Circuit
        angles = np.asarray(angles, dtype=np.float64)
        if angles.shape != (self.num_angles,):
            raise ValueError(f"expected {self.num_angles} angles, got shape {angles.shape}")
        name = {L.RX: 'rx', L.RY: 'ry', L.RZ: 'rz', L.CP: 'cp', L.CZ: 'cz', L.CX: 'cx'}
        qc = Circuit(self.num_qubits)
        for kind, q0, q1, p, c in self.ops:
            a = float(angles[p]) if p >= 0 else c
            if kind in (L.RX, L.RY, L.RZ):
                qc.append(name[kind], [q0], [a])
            elif kind == L.CP:
                qc.append('cp', [q0, q1], [a])
            else:
                qc.append(name[kind], [q0, q1])
        return qc

    def constrained(self, fixed_params, indices):
        """Program with parameters `indices` frozen at `fixed_params` — the device-side form of
        constrained_function(anz.unitary, ...) (reference cp_utils.py:100-108): returns
        (Program over the remaining free angles, free index list)."""
        fixed = dict(zip([int(i) for i in indices], [float(x) for x in fixed_params]))
        free_idx = [i for i in range(self.num_angles) if i not in fixed]
        remap = {old: new for new, old in enumerate(free_idx)}
        ops = []
        for kind, q0, q1, p, c in self.ops:
            if p >= 0 and p in fixed:
                ops.append((kind, q0, q1, -1, fixed[p]))
            elif p >= 0:
                ops.append((kind, q0, q1, remap[p], c))
            else:
                ops.append((kind, q0, q1, p, c))
        return Program(self.num_qubits, ops, len(free_idx)), free_idx

    def learn(self, u_target, method='adam', learning_rate=0.1, target_loss=1e-7, keep_history=True, **kwargs):
        """Reference main.py:224-239."""
        from .optimization import unitary_learn
        return unitary_learn(self, u_target, self.num_angles, method=method, learning_rate=learning_rate,
                             target_loss=target_loss, keep_history=keep_history, **kwargs)
