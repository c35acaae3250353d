"""CPU ORACLE — test infrastructure, NOT product code.

A plain numpy / torch-CPU restatement of the variational-synthesis hot path of idnm/cpflow
(the reference, read-only at /root/reference during the build).  Every function cites the
reference file:line it restates.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
CPU-baseline / `--impl reference` legs may import this module; `cpflow_b200/` never does.

PARITY PINNING.  The reference is pure Python on JAX 0.3.x + optax 0.1.1 and cannot be
imported here (no jax/optax/qiskit in the image), so the oracle is pinned as follows:

* forward math (gates, gate placement, ansatz layout, HS loss): pinned by the reference's own
  stored result files — 148 (angles -> unitary) and 170 gate-list known answers extracted
  into tests/golden/ by tests/golden/make_golden.py (see tests/test_oracle_golden.py);
* threefry PRNG (`jax.random.split/uniform`, jax 0.3.4): pinned by the Random123 known-answer
  vectors, by the JAX documentation values and by the chain of `random_seed`s stored in the
  reference's hyperopt trials (main.py:798-799);
* gradients: the reference stores none ("parity unpinned" at the value level); the oracle uses
  torch complex autograd over the restated forward, cross-checked by finite differences and
  by an independent hand-written adjoint sweep (`hand_adjoint_grad`);
* optax 0.1.1 Adam (third-party, pinned version setup.py:25): restated from its published
  update rule; "parity unpinned" at the step level (no stored trajectories exist).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

PI = math.pi

# --------------------------------------------------------------------------------------
# gates.py:10-58
# --------------------------------------------------------------------------------------


def _cdtype(dtype):
    return {torch.float32: torch.complex64, torch.float64: torch.complex128}[dtype]


def pauli(name, cdtype=torch.complex128):
    """gates.py:10-17."""
    m = {"x": [[0, 1], [1, 0]], "y": [[0, -1j], [1j, 0]], "z": [[1, 0], [0, -1]]}[name]
    return torch.tensor(m, dtype=cdtype)


def rotation_matrix(name, a):
    """gates.py:22-35: cos(a/2) I - i sin(a/2) sigma.  `a` is a 0-d real tensor."""
    cd = _cdtype(a.dtype)
    c = torch.cos(a / 2).to(cd)
    s = torch.sin(a / 2).to(cd)
    return c * torch.eye(2, dtype=cd) - 1j * pauli(name, cd) * s


def cp_mat(a):
    """gates.py:51-58: diag(1, 1, 1, exp(i a))."""
    cd = _cdtype(a.dtype)
    ph = torch.exp(1j * a.to(cd))
    d = torch.stack([torch.ones((), dtype=cd), torch.ones((), dtype=cd), torch.ones((), dtype=cd), ph])
    return torch.diag(d)


def cz_mat(cd=torch.complex128):
    """gates.py:45-48."""
    return torch.diag(torch.tensor([1, 1, 1, -1], dtype=cd))


def cx_mat(cd=torch.complex128):
    """gates.py:40-43."""
    return torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=cd)


def toffoli_target(n, cd=torch.complex128):
    """gates.py:95-106 (qiskit `mct` + reverse_bits): identity with the last two basis
    states swapped, big-endian (confirmed by the stored circuits)."""
    N = 2 ** n
    u = torch.eye(N, dtype=cd)
    u[N - 2, N - 2] = 0
    u[N - 1, N - 1] = 0
    u[N - 2, N - 1] = 1
    u[N - 1, N - 2] = 1
    return u


# --------------------------------------------------------------------------------------
# circuit_assembly.py:7-45
# --------------------------------------------------------------------------------------


def gate_transposition(placement):
    """circuit_assembly.py:7-13."""
    position_index = [(placement[i], i) for i in range(len(placement))]
    position_index.sort()
    return [i for _, i in position_index]


def transposition(n_qubits, placement):
    """circuit_assembly.py:16-28."""
    gate_width = len(placement)
    t = list(range(gate_width, n_qubits))
    for position, insertion in zip(sorted(placement), gate_transposition(placement)):
        t.insert(position, insertion)
    return t


def apply_gate_to_tensor(gate, tensor, placement):
    """circuit_assembly.py:31-45 (tensordot over the gate's input axes, transpose back)."""
    gate_width = gate.dim() // 2
    tensor_width = tensor.dim() // 2
    gate_contraction_axes = list(range(gate_width, 2 * gate_width))
    contraction = torch.tensordot(gate, tensor, dims=(gate_contraction_axes, list(placement)))
    t = transposition(tensor_width, placement) + list(range(tensor_width, 2 * tensor_width))
    return contraction.permute(t)


# --------------------------------------------------------------------------------------
# topology.py:7-20, 36-38
# --------------------------------------------------------------------------------------


def connected_layer(num_qubits):
    return [[i, j] for i in range(num_qubits) for j in range(i + 1, num_qubits)]


def chain_layer(num_qubits):
    return [[i, i + 1] for i in range(num_qubits - 1)]


def fill_layers(layer, depth):
    num_complete_layers = depth // len(layer)
    return {"layers": [layer, num_complete_layers], "free": layer[: depth % len(layer)]}


def num_qubits_from_layer(layer):
    return max(item for sub in layer for item in sub) + 1


# --------------------------------------------------------------------------------------
# main.py:23-146, 149-191
# --------------------------------------------------------------------------------------


def block_num_angles(entangling_gate_name, rotation_gates):
    """main.py:32-34."""
    return 2 * len(rotation_gates) + (entangling_gate_name == "cp")


def block_unitary(entangling_gate_name, rotation_gates, angles):
    """main.py:69-82: entangler first, then kron(R(up), R(down)) per letter."""
    cd = _cdtype(angles.dtype)
    if entangling_gate_name == "cp":
        u = cp_mat(angles[-1])
    elif entangling_gate_name == "cz":
        u = cz_mat(cd)
    elif entangling_gate_name == "cx":
        u = cx_mat(cd)
    else:
        raise TypeError(entangling_gate_name)
    up = angles[::2]
    down = angles[1::2][: len(up)]
    for xyz, a0, a1 in zip(rotation_gates, up, down):
        u = torch.kron(rotation_matrix(xyz, a0), rotation_matrix(xyz, a1)) @ u
    return u


def build_unitary(num_qubits, entangling_gate_name, rotation_gates, placements, angles):
    """main.py:106-146 (the fori_loop over layers is unrolled; same gate order)."""
    layer, num_layers = placements["layers"]
    free_placements = placements["free"]
    nb = block_num_angles(entangling_gate_name, rotation_gates)
    n = num_qubits
    cd = _cdtype(angles.dtype)
    surface = angles[: 3 * n].reshape(n, 3)
    blocks = angles[3 * n:].reshape(-1, nb)
    u = torch.eye(2 ** n, dtype=cd).reshape([2] * (2 * n))
    for i in range(n):
        a = surface[i]
        gate = rotation_matrix("z", a[2]) @ rotation_matrix("x", a[1]) @ rotation_matrix("z", a[0])
        u = apply_gate_to_tensor(gate, u, [i])
    all_placements = list(layer) * num_layers + list(free_placements)
    for a, p in zip(blocks, all_placements):
        gate = block_unitary(entangling_gate_name, rotation_gates, a).reshape(2, 2, 2, 2)
        u = apply_gate_to_tensor(gate, u, p)
    return u.reshape(2 ** n, 2 ** n)


class Ansatz:
    """main.py:149-191 (num_angles, cp_mask, unitary)."""

    def __init__(self, num_qubits, entangling_gate_name, placements, rotation_gates="xyz"):
        self.num_qubits = num_qubits
        self.entangling_gate_name = entangling_gate_name
        self.rotation_gates = rotation_gates
        placements = dict(placements)
        placements.setdefault("layers", [[], 0])
        placements.setdefault("free", [])
        self.placements = placements
        self.layer, self.num_layers = placements["layers"]
        self.free_placements = placements["free"]
        self.all_placements = list(self.layer) * self.num_layers + list(self.free_placements)
        self.num_blocks = len(self.all_placements)
        nb = block_num_angles(entangling_gate_name, rotation_gates)
        self.num_block_angles = nb
        self.num_angles = 3 * num_qubits + nb * self.num_blocks
        if entangling_gate_name == "cp":
            m = np.zeros(self.num_angles, dtype=np.int64)
            m[3 * num_qubits + nb - 1:: nb] = 1
            self.cp_mask = m

    def unitary(self, angles):
        return build_unitary(self.num_qubits, self.entangling_gate_name, self.rotation_gates,
                             self.placements, angles)


def cp_ansatz(layer, num_cp_gates, rotation_gates="xyz"):
    """main.py:560: Ansatz(n, 'cp', fill_layers(layer, K), rg)."""
    return Ansatz(num_qubits_from_layer(layer), "cp", fill_layers(layer, num_cp_gates), rotation_gates)


# --------------------------------------------------------------------------------------
# gate program IR (shared vocabulary with the C ABI: include/cpflow_b200.h)
# --------------------------------------------------------------------------------------

RX, RY, RZ, CP, CZ, CX = 0, 1, 2, 3, 4, 5
_ROT = {"x": RX, "y": RY, "z": RZ}


def ansatz_program(anz):
    """Flatten an Ansatz into time-ordered primitive ops (kind, q0, q1, param_index, const).

    Order follows main.py:119-146 and main.py:69-82: surface Rz(a0) Rx(a1) Rz(a2) per qubit,
    then per block: entangler, then per letter R(up) on placement[0], R(down) on placement[1].
    """
    n = anz.num_qubits
    ops = []
    for q in range(n):
        ops.append((RZ, q, -1, 3 * q + 0, 0.0))
        ops.append((RX, q, -1, 3 * q + 1, 0.0))
        ops.append((RZ, q, -1, 3 * q + 2, 0.0))
    nb = anz.num_block_angles
    for b, (p0, p1) in enumerate(anz.all_placements):
        base = 3 * n + nb * b
        if anz.entangling_gate_name == "cp":
            ops.append((CP, p0, p1, base + nb - 1, 0.0))
        elif anz.entangling_gate_name == "cz":
            ops.append((CZ, p0, p1, -1, 0.0))
        else:
            ops.append((CX, p0, p1, -1, 0.0))
        for j, letter in enumerate(anz.rotation_gates):
            ops.append((_ROT[letter], p0, -1, base + 2 * j, 0.0))
            ops.append((_ROT[letter], p1, -1, base + 2 * j + 1, 0.0))
    return ops


def program_unitary_np(n, ops, angles):
    """Straightforward numpy state-matrix simulator of a primitive-op program (big-endian:
    qubit 0 is the most significant bit of the row index, circuit_assembly.py:31-45).
    Restates qiskit_circ_to_jax_unitary's `u(angles)` (circuit_assembly.py:48-81)."""
    N = 2 ** n
    angles = np.asarray(angles, dtype=np.float64)
    u = np.eye(N, dtype=np.complex128)
    idx = np.arange(N)
    for kind, q0, q1, pi, const in ops:
        a = angles[pi] if pi >= 0 else const
        if kind in (RX, RY, RZ):
            c, s = math.cos(a / 2), math.sin(a / 2)
            sig = {RX: np.array([[0, 1], [1, 0]]), RY: np.array([[0, -1j], [1j, 0]]),
                   RZ: np.array([[1, 0], [0, -1]])}[kind]
            g = c * np.eye(2) - 1j * s * sig
            m = 1 << (n - 1 - q0)
            lo = idx[(idx & m) == 0]
            hi = lo | m
            a0, a1 = u[lo].copy(), u[hi].copy()
            u[lo] = g[0, 0] * a0 + g[0, 1] * a1
            u[hi] = g[1, 0] * a0 + g[1, 1] * a1
        elif kind in (CP, CZ):
            m0, m1 = 1 << (n - 1 - q0), 1 << (n - 1 - q1)
            sel = (idx & m0 != 0) & (idx & m1 != 0)
            ph = np.exp(1j * a) if kind == CP else -1.0
            u[sel] *= ph
        elif kind == CX:
            m0, m1 = 1 << (n - 1 - q0), 1 << (n - 1 - q1)
            ctrl = idx[(idx & m0 != 0) & (idx & m1 == 0)]
            tmp = u[ctrl].copy()
            u[ctrl] = u[ctrl | m1]
            u[ctrl | m1] = tmp
        else:
            raise ValueError(kind)
    return u


# --------------------------------------------------------------------------------------
# losses — matrix_utils.py:35-42 and the notebook losses
# --------------------------------------------------------------------------------------


def cost_HST(u, u_target):
    """matrix_utils.py:35-42."""
    n = u_target.shape[0]
    return 1 - torch.abs((u * u_target.conj()).sum()) ** 2 / n ** 2


def cost_state_prep(u, psi):
    """tutorial/CPFlow_tutorial.ipynb:1318: 1 - |<psi| U |0>|^2."""
    return 1 - torch.abs((psi.conj() * u[:, 0]).sum()) ** 2


def cost_relative_phase(u, u_target):
    """tutorial/CPFlow_tutorial.ipynb:1507: 1 - sum |conj(V_ij) U_ij|^2 / N."""
    n = u_target.shape[0]
    return 1 - (torch.abs(u_target.conj() * u) ** 2).sum() / n


def theoretical_lower_bound(n):
    """matrix_utils.py:11-14."""
    return int((4 ** n - 3 * n - 1) / 4 + 1)


# --------------------------------------------------------------------------------------
# penalty.py:14-15, 44-97; main.py:328-335
# --------------------------------------------------------------------------------------


@dataclass
class RegularizationOptions:
    function: str = "linear"
    ymax: float = 2
    xmax: float = PI / 2
    plato_0: float = 0.05
    plato_1: float = 0.05
    plato_2: float = 0.05


def line_coeffs(x0, y0, x1, y1):
    """penalty.py:14-15 evaluated in Python doubles (the reference's `line` receives Python
    floats for everything but `x`): returns (slope, intercept)."""
    return (y1 - y0) / (x1 - x0), (x0 * y1 - x1 * y0) / (x0 - x1)


def penalty_segments(xmax=PI / 2, ymax=2, plato_0=0.05, plato_1=0.05, plato_2=0.05):
    """penalty.py:44-71 as a first-true-wins table [(lo, hi, slope, intercept)], meaning
    `lo < a <= hi` (lo = -inf for the first row).  No match -> 0."""
    pi = PI
    pts = [
        (-math.inf, plato_0, (0, 0, plato_0, 0)),
        (plato_0, xmax - plato_2, (plato_0, 0, xmax - plato_2, ymax)),
        (xmax - plato_2, xmax + plato_2, (xmax - plato_2, ymax, xmax + plato_2, ymax)),
        (xmax + plato_2, pi - plato_1, (xmax + plato_2, ymax, pi - plato_1, 1)),
        (pi - plato_1, pi + plato_1, (pi - plato_1, 1, pi + plato_1, 1)),
        (pi + plato_1, pi + xmax - plato_2, (pi + plato_1, 1, pi + xmax - plato_2, ymax)),
        (pi + xmax - plato_2, pi + xmax + plato_2, (pi + xmax - plato_2, ymax, pi + xmax + plato_2, ymax)),
        (pi + xmax + plato_2, 2 * pi - plato_0, (pi + xmax + plato_2, ymax, 2 * pi - plato_0, 0)),
        (2 * pi - plato_0, 2 * pi, (2 * pi - plato_0, 0, 2 * pi, 0)),
    ]
    out = [(lo, hi) + line_coeffs(*ln) for lo, hi, ln in pts]
    out.append((2 * pi, 3 * pi, 0.0, 1.0))  # penalty.py:56, :67 ("workaround" constant 1)
    return out


def cp_penalty_linear(a, xmax=PI / 2, ymax=2, plato_0=0.05, plato_1=0.05, plato_2=0.05):
    """penalty.py:44-71 on a real tensor (elementwise; differentiable w.r.t. `a`).
    `a % 2pi` is jnp.mod == torch.remainder; jnp.piecewise is first-true-wins."""
    dt = a.dtype
    a = torch.remainder(a, torch.tensor(2 * PI, dtype=dt))
    res = torch.zeros_like(a)
    done = torch.zeros_like(a, dtype=torch.bool)
    for lo, hi, slope, icpt in penalty_segments(xmax, ymax, plato_0, plato_1, plato_2):
        cond = (a <= torch.tensor(hi, dtype=dt))
        if lo != -math.inf:
            cond = cond & (torch.tensor(lo, dtype=dt) < a)
        val = torch.tensor(slope, dtype=dt) * a + torch.tensor(icpt, dtype=dt)
        take = cond & ~done
        res = torch.where(take, val, res)
        done = done | cond
    return res


def cp_penalty_L1(a):
    """penalty.py:74-76."""
    return torch.abs(a)


def make_regularization_function(options=RegularizationOptions):
    """penalty.py:79-97 (the reference passes the CLASS, main.py:539; attributes are read the
    same way from class or instance)."""
    if options.function == "linear":
        o = options
        return lambda a: cp_penalty_linear(a, o.xmax, o.ymax, o.plato_0, o.plato_1, o.plato_2)
    if options.function == "L1":
        return cp_penalty_L1
    raise ValueError("penalty function not supported")


def regularization_value(angles, cp_mask, r, penalty_func):
    """main.py:563-564: r * sum_i R(angs_i * cp_mask_i).  `angles` is [..., P]."""
    mask = torch.as_tensor(cp_mask, dtype=angles.dtype)
    return r * penalty_func(angles * mask).sum(-1)


# --------------------------------------------------------------------------------------
# batched forward (same math as build_unitary, vectorised over samples for speed)
# --------------------------------------------------------------------------------------


def _rot_batched(kind, a):
    """[B] angles -> [B,2,2] rotation matrices (gates.py:22-35)."""
    cd = _cdtype(a.dtype)
    c = torch.cos(a / 2).to(cd)
    s = torch.sin(a / 2).to(cd)
    z = torch.zeros_like(c)
    if kind == RX:
        rows = [[c, -1j * s], [-1j * s, c]]
    elif kind == RY:
        rows = [[c, -s], [s, c]]
    else:
        rows = [[c - 1j * s, z], [z, c + 1j * s]]
    return torch.stack([torch.stack(r, -1) for r in rows], -2)


def program_unitary_batched(n, ops, angles):
    """angles [B,P] real tensor -> U [B,N,N]; differentiable.  Big-endian placement."""
    B = angles.shape[0]
    N = 2 ** n
    cd = _cdtype(angles.dtype)
    u = torch.eye(N, dtype=cd).expand(B, N, N).clone()
    idx = torch.arange(N)
    for kind, q0, q1, pi, const in ops:
        a = angles[:, pi] if pi >= 0 else torch.full((B,), const, dtype=angles.dtype)
        if kind in (RX, RY, RZ):
            g = _rot_batched(kind, a)
            v = u.reshape(B, 2 ** q0, 2, 2 ** (n - 1 - q0) * N)
            u = torch.einsum("bij,bajc->baic", g, v).reshape(B, N, N)
        elif kind in (CP, CZ):
            m0, m1 = 1 << (n - 1 - q0), 1 << (n - 1 - q1)
            sel = ((idx & m0) != 0) & ((idx & m1) != 0)
            if kind == CP:
                ph = torch.exp(1j * a.to(cd))
            else:
                ph = -torch.ones(B, dtype=cd)
            d = torch.where(sel[None, :], ph[:, None], torch.ones((), dtype=cd))
            u = u * d[:, :, None]
        elif kind == CX:
            m0, m1 = 1 << (n - 1 - q0), 1 << (n - 1 - q1)
            perm = torch.where((idx & m0) != 0, idx ^ m1, idx)
            u = u[:, perm, :]
        else:
            raise ValueError(kind)
    return u


def loss_batched(kind, u, target):
    """kind in {'hs','state','relphase'}; u [B,N,N]."""
    N = u.shape[-1]
    if kind == "hs":
        t = (u * target.conj()[None]).sum((-1, -2))
        return 1 - torch.abs(t) ** 2 / N ** 2
    if kind == "state":
        t = (target.conj()[None] * u[:, :, 0]).sum(-1)
        return 1 - torch.abs(t) ** 2
    if kind == "relphase":
        return 1 - (torch.abs(target.conj()[None] * u) ** 2).sum((-1, -2)) / N
    raise ValueError(kind)


def loss_and_grad_batched(n, ops, angles, loss_kind, target, cp_mask=None, r=0.0, penalty_func=None):
    """value_and_grad(regloss) of optimization.py:331-340 for a batch: returns
    (loss [B], reg [B], grad [B,P]) with grad = d(loss+reg)/d(angles)."""
    a = angles.clone().requires_grad_(True)
    u = program_unitary_batched(n, ops, a)
    loss = loss_batched(loss_kind, u, target)
    if penalty_func is not None and cp_mask is not None:
        reg = regularization_value(a, cp_mask, r, penalty_func)
    else:
        reg = torch.zeros_like(loss)
    (loss + reg).sum().backward()
    return loss.detach(), reg.detach(), a.grad.detach()


# --------------------------------------------------------------------------------------
# hand adjoint sweep — independent gradient cross-check in numpy
# --------------------------------------------------------------------------------------


def hand_adjoint_grad(n, ops, angles, loss_kind, target):
    """O(1)-memory adjoint: forward, seed lambda, then walk the program backwards applying
    inverse gates to both phi and lambda.  Returns (loss, grad[P]) in float64."""
    N = 2 ** n
    angles = np.asarray(angles, dtype=np.float64)
    target = np.asarray(target)
    phi = program_unitary_np(n, ops, angles)
    if loss_kind == "hs":
        t = np.sum(np.conj(target) * phi)
        loss = 1 - abs(t) ** 2 / N ** 2
        lam = -(t / N ** 2) * target
    elif loss_kind == "state":
        t = np.sum(np.conj(target) * phi[:, 0])
        loss = 1 - abs(t) ** 2
        lam = np.zeros_like(phi)
        lam[:, 0] = -t * target
        phi = phi.copy()
    elif loss_kind == "relphase":
        loss = 1 - np.sum(np.abs(np.conj(target) * phi) ** 2) / N
        lam = -(1.0 / N) * np.abs(target) ** 2 * phi
    else:
        raise ValueError(loss_kind)
    lam = lam.astype(np.complex128)
    grad = np.zeros_like(angles)
    idx = np.arange(N)
    sig = {RX: np.array([[0, 1], [1, 0]], dtype=complex), RY: np.array([[0, -1j], [1j, 0]]),
           RZ: np.array([[1, 0], [0, -1]], dtype=complex)}

    def apply1(mat, q, x):
        m = 1 << (n - 1 - q)
        lo = idx[(idx & m) == 0]
        hi = lo | m
        a0, a1 = x[lo].copy(), x[hi].copy()
        x[lo] = mat[0, 0] * a0 + mat[0, 1] * a1
        x[hi] = mat[1, 0] * a0 + mat[1, 1] * a1

    for kind, q0, q1, pi, const in reversed(ops):
        a = angles[pi] if pi >= 0 else const
        if kind in (RX, RY, RZ):
            if pi >= 0:
                sp = phi.copy()
                apply1(sig[kind], q0, sp)
                grad[pi] += np.imag(np.sum(np.conj(lam) * sp))
            c, s = math.cos(a / 2), math.sin(a / 2)
            ginv = c * np.eye(2) + 1j * s * sig[kind]
            apply1(ginv, q0, phi)
            apply1(ginv, q0, lam)
        elif kind in (CP, CZ):
            m0, m1 = 1 << (n - 1 - q0), 1 << (n - 1 - q1)
            sel = (idx & m0 != 0) & (idx & m1 != 0)
            if kind == CP and pi >= 0:
                grad[pi] += -2.0 * np.imag(np.sum(np.conj(lam[sel]) * phi[sel]))
            ph = np.exp(-1j * a) if kind == CP else -1.0
            phi[sel] *= ph
            lam[sel] *= ph
        elif kind == CX:
            m0, m1 = 1 << (n - 1 - q0), 1 << (n - 1 - q1)
            ctrl = idx[(idx & m0 != 0) & (idx & m1 == 0)]
            for x in (phi, lam):
                tmp = x[ctrl].copy()
                x[ctrl] = x[ctrl | m1]
                x[ctrl | m1] = tmp
    return loss, grad


# --------------------------------------------------------------------------------------
# optax 0.1.1 adam (third-party; setup.py:25) + optimization.py:14-94, 209-382
# --------------------------------------------------------------------------------------


class AdamState:
    def __init__(self, params):
        self.count = 0
        self.mu = torch.zeros_like(params)
        self.nu = torch.zeros_like(params)


def adam_update(grads, state, lr, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    """optax.adam(lr) == chain(scale_by_adam(b1,b2,eps,eps_root), scale(-lr)) in optax 0.1.1:
    mu = (1-b1) g + b1 mu ; nu = (1-b2) g^2 + b2 nu ; count += 1 ;
    mu_hat = mu / (1 - b1^count) ; nu_hat = nu / (1 - b2^count) ;
    update = -lr * mu_hat / (sqrt(nu_hat + eps_root) + eps)."""
    dt = grads.dtype
    state.mu = (1 - b1) * grads + b1 * state.mu
    state.nu = (1 - b2) * grads * grads + b2 * state.nu
    state.count += 1
    c1 = torch.tensor(1.0, dtype=dt) - torch.tensor(b1, dtype=dt) ** state.count
    c2 = torch.tensor(1.0, dtype=dt) - torch.tensor(b2, dtype=dt) ** state.count
    mu_hat = state.mu / c1
    nu_hat = state.nu / c2
    upd = mu_hat / (torch.sqrt(nu_hat + eps_root) + eps)
    return -lr * upd


def adam_minimize_batched(loss_and_grad, params0, lr, num_iterations, keep_history=False,
                          freeze_mask=None):
    """optax_minimize (optimization.py:28-94) vmapped over the batch (optimization.py:362).

    `loss_and_grad(params[B,P]) -> (regloss[B], grad[B,P])`.
    No-history: returns params [B,2,P] = [initial, best] and regloss [B,2]; strict `<`,
    pre-update params (optimization.py:61-75).  History: params [B,T,P], regloss [B,T]
    (optimization.py:52-59, 77-86).  `freeze_mask[B,P]` (True = frozen) emulates the reduced
    parameter vector of constrained_function (cp_utils.py:100-108): frozen entries never move.
    """
    params = params0.clone()
    B, P = params.shape
    state = AdamState(params)
    init_loss, _ = loss_and_grad(params)
    best_loss = init_loss.clone()
    best_params = params.clone()
    if keep_history:
        ph = torch.zeros(B, num_iterations, P, dtype=params.dtype)
        lh = torch.zeros(B, num_iterations, dtype=params.dtype)
        ph[:, 0] = params
        lh[:, 0] = init_loss
    for i in range(num_iterations):
        loss, g = loss_and_grad(params)
        if freeze_mask is not None:
            g = torch.where(freeze_mask, torch.zeros_like(g), g)
        upd = adam_update(g, state, lr)
        if freeze_mask is not None:
            upd = torch.where(freeze_mask, torch.zeros_like(upd), upd)
        new_params = params + upd
        if keep_history:
            if i + 1 < num_iterations:
                ph[:, i + 1] = new_params
            lh[:, i] = loss
        else:
            better = loss < best_loss
            best_loss = torch.where(better, loss, best_loss)
            best_params = torch.where(better[:, None], params, best_params)
        params = new_params
    if keep_history:
        return ph, lh
    return torch.stack([params0, best_params], 1), torch.stack([init_loss, best_loss], 1)


def mynimize_repeated(n, ops, loss_kind, target, initial_params_batch, learning_rate=0.1,
                      num_iterations=2000, cp_mask=None, r=0.0, penalty_func=None,
                      keep_history=False, freeze_mask=None):
    """optimization.py:269-382 restated on the declarative spec (the reference takes closures).
    Returns the reference's list of dicts {'params','loss','reg','regloss'}."""
    x0 = torch.as_tensor(initial_params_batch)
    single = x0.dim() == 1
    if single:
        x0 = x0[None]

    def lg(p):
        loss, reg, g = loss_and_grad_batched(n, ops, p, loss_kind, target, cp_mask, r, penalty_func)
        return loss + reg, g

    params_h, regloss_h = adam_minimize_batched(lg, x0, learning_rate, num_iterations,
                                                keep_history=keep_history, freeze_mask=freeze_mask)
    if penalty_func is not None and cp_mask is not None:
        reg_h = regularization_value(params_h, cp_mask, r, penalty_func)
    else:
        reg_h = torch.zeros_like(regloss_h)
    loss_h = regloss_h - reg_h
    results = [{"params": p, "loss": l, "reg": rg, "regloss": rl}
               for p, l, rg, rl in zip(params_h, loss_h, reg_h, regloss_h)]
    return results[0] if single else results


# --------------------------------------------------------------------------------------
# cp_utils.py:45-77, 111-141, 144-202
# --------------------------------------------------------------------------------------


def cz_value(a, threshold=0.2):
    """cp_utils.py:45-57 on a float32 numpy array."""
    a = np.asarray(a, dtype=np.float32)
    t = np.float32(threshold)
    a = np.mod(a, np.float32(2 * PI))
    out = np.full(a.shape, 2, dtype=np.int64)
    out[np.abs(a - np.float32(PI)) < t] = 1
    out[np.abs(a - np.float32(2 * PI)) < t] = 0
    out[a < t] = 0
    return out


def count_cz(angles, threshold=0.2):
    """cp_utils.py:59-67."""
    return int(cz_value(angles, threshold).sum())


def project_cp_angles(angles, cp_mask, threshold=0.2):
    """cp_utils.py:70-77 + 111-141: returns (projected angles [P] f32, frozen mask [P]).
    CP angles within `threshold` of 0/2pi -> 0, of pi -> float32(pi); those are frozen."""
    a = np.asarray(angles, dtype=np.float32).copy()
    mask = np.asarray(cp_mask) == 1
    am = np.mod(a, np.float32(2 * PI))
    near_pi = mask & (np.abs(am - np.float32(PI)) < np.float32(threshold))
    near_0 = mask & ~near_pi & ((np.abs(am) < np.float32(threshold)) |
                                (np.abs(am - np.float32(2 * PI)) < np.float32(threshold)))
    out = a.copy()  # non-projected angles keep their original value (cp_utils.py:135)
    out[near_pi] = np.float32(PI)
    out[near_0] = 0.0
    return out, (near_pi | near_0)


def evaluate_cp_result(res, cp_mask, threshold=0.2):
    """cp_utils.py:144-164."""
    regloss = np.asarray(res["regloss"])
    best_i = int(np.argmin(regloss))
    loss = float(np.asarray(res["loss"])[best_i])
    angles = np.asarray(res["params"])[best_i]
    cz = count_cz(angles * np.asarray(cp_mask), threshold=threshold)
    return cz, loss, angles


def filter_cp_results(res_list, cp_mask, threshold_cz_count, threshold_loss, threshold_cp=0.2):
    """cp_utils.py:167-202."""
    selected = []
    for res in res_list:
        cz, loss, _ = evaluate_cp_result(res, cp_mask, threshold=threshold_cp)
        if cz <= threshold_cz_count and loss <= threshold_loss:
            selected.append([cz, res])
    selected.sort(key=lambda x: x[0])
    return selected


# --------------------------------------------------------------------------------------
# jax 0.3.4 threefry PRNG (third-party; main.py:543, 566; cp_utils.py:31-32;
# trigonometric_utils.py:35-38) — restated from the published Threefry-2x32-20 algorithm
# --------------------------------------------------------------------------------------

_ROTATIONS = ((13, 15, 26, 6), (17, 29, 16, 24))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds (Salmon et al. 2011), as used by jax._src.prng."""
    with np.errstate(over="ignore"):
        x0 = np.asarray(x0, dtype=np.uint32).copy()
        x1 = np.asarray(x1, dtype=np.uint32).copy()
        k0 = np.uint32(k0)
        k1 = np.uint32(k1)
        ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for g in range(5):
            for rot in _ROTATIONS[g % 2]:
                x0 = x0 + x1
                x1 = (x1 << np.uint32(rot)) | (x1 >> np.uint32(32 - rot))
                x1 = x1 ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + np.uint32(g + 1)
    return x0, x1


def prng_key(seed):
    """jax.random.PRNGKey for 0 <= seed < 2^32 (x32 mode): [0, seed]."""
    return np.array([(int(seed) >> 32) & 0xFFFFFFFF, int(seed) & 0xFFFFFFFF], dtype=np.uint32)


def _threefry_2x32_counts(key, count):
    count = np.asarray(count, dtype=np.uint32).ravel()
    odd = count.size % 2
    if odd:
        count = np.concatenate([count, np.zeros(1, np.uint32)])
    h = count.size // 2
    o0, o1 = threefry2x32(key[0], key[1], count[:h], count[h:])
    out = np.concatenate([o0, o1])
    return out[:-1] if odd else out


def prng_split(key, num=2):
    """jax.random.split: threefry over iota(2*num), reshaped (num, 2)."""
    return _threefry_2x32_counts(key, np.arange(2 * num, dtype=np.uint32)).reshape(num, 2)


def prng_uniform(key, size, minval=0.0, maxval=1.0):
    """jax.random.uniform(key, (size,), float32, minval, maxval)."""
    bits = _threefry_2x32_counts(key, np.arange(size, dtype=np.uint32))
    fb = (bits >> np.uint32(9)) | np.float32(1.0).view(np.uint32)
    floats = fb.view(np.float32) - np.float32(1.0)
    lo, hi = np.float32(minval), np.float32(maxval)
    return np.maximum(lo, floats * (hi - lo) + lo).astype(np.float32)


def random_angles(num_angles, key):
    """trigonometric_utils.py:35-38."""
    return prng_uniform(key, num_angles, 0.0, 2 * PI)


def erf_inv_f32(x):
    """lax.erf_inv in float32 as XLA computes it (xla/client/lib/math.cc, ErfInv32: Giles' polynomial
    approximation, jaxlib 0.3.x: w = -log((1-x)(1+x)); two degree-8 polynomials in w - 2.5 / sqrt(w) - 3).
    Third-party arithmetic that is not under /root/reference: restated from the published algorithm; no reference
    artefact stores normal draws ("parity unpinned" for cp_dist='normal', which the reference never uses by
    default, main.py:360)."""
    x = np.asarray(x, dtype=np.float32)
    one = np.float32(1)
    w = -np.log((one - x) * (one + x)).astype(np.float32)
    lt = w < np.float32(5)
    w = np.where(lt, w - np.float32(2.5), np.sqrt(np.maximum(w, 0)).astype(np.float32) - np.float32(3)).astype(np.float32)
    c_lt = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
            -0.00417768164, 0.246640727, 1.50140941]
    c_gt = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
            0.00943887047, 1.00167406, 2.83297682]
    p = np.where(lt, np.float32(c_lt[0]), np.float32(c_gt[0])).astype(np.float32)
    for a, b in zip(c_lt[1:], c_gt[1:]):
        p = (np.where(lt, np.float32(a), np.float32(b)) + p * w).astype(np.float32)
    return (p * x).astype(np.float32)


def prng_normal(key, size):
    """jax.random.normal(key, (size,), float32) of jax 0.3.x (_normal_real): u = uniform(key, minval=nextafter(-1, 0),
    maxval=1); sqrt(2) * erf_inv(u)."""
    lo = np.nextafter(np.float32(-1), np.float32(0))
    u = prng_uniform(key, size, lo, 1.0)
    return (np.float32(np.sqrt(2)) * erf_inv_f32(u)).astype(np.float32)


def generate_initial_angles(seed, num_angles, cp_mask, cp_dist="uniform", batch_size=1):
    """main.py:541-548 + cp_utils.py:13-42: 'uniform', '0' (CP angles zeroed) and 'normal' (CP angles
    1.5 * random.normal from a second split of the sample's key, cp_utils.py:38-40)."""
    key = prng_key(seed)
    keys = prng_split(key, batch_size + 1)
    out = np.zeros((batch_size, num_angles), dtype=np.float32)
    mask = np.asarray(cp_mask, dtype=np.float32)
    for b in range(batch_size):
        k1, sub = prng_split(keys[b + 1], 2)          # key, subkey = random.split(key)   cp_utils.py:31
        out[b] = random_angles(num_angles, sub)
        if cp_dist == "normal":
            sub2 = prng_split(k1, 2)[1]               # key, subkey = random.split(key)   cp_utils.py:39
            out[b] = out[b] * (1 - mask) + np.float32(1.5) * prng_normal(sub2, num_angles) * mask
    if cp_dist == "0":
        out = out * (1 - mask)[None]
    elif cp_dist not in ("uniform", "normal"):
        raise ValueError(cp_dist)
    return out


def next_adaptive_seed(seed):
    """main.py:798-799: `_, subkey = split(PRNGKey(seed)); seed = int(subkey[1])`."""
    return int(prng_split(prng_key(seed), 2)[1][1])


# --------------------------------------------------------------------------------------
# refine path: exact_decompositions.py:77-113, sequential restatement
# --------------------------------------------------------------------------------------


def reduce_all_1q_angles(loss_func, initial_angles, wires, threshold=1e-5):
    """exact_decompositions.py:77-113 unrolled into a loop: angle k (earlier ones already final) is set to zero
    if the loss stays below `threshold`, else merged into the first later angle on the same wire for which
    a_i -/+ a_k passes (sign -1 tried first)."""
    angles = np.array(initial_angles, dtype=np.float64)
    G = len(angles)
    for k in range(G):
        trial = angles.copy()
        trial[k] = 0.0
        if loss_func(trial) < threshold:                       # reduce_first_1q_angle, :90-92
            angles = trial
            continue
        done = False
        for i in range(k + 1, G):                               # :95-98
            if wires[i] != wires[k]:                            # can_reduce_two_angles, :104-105
                continue
            for sign in (-1, 1):                                # :107-113
                trial = angles.copy()
                trial[i] = angles[i] + sign * angles[k]
                trial[k] = 0.0
                if loss_func(trial) < threshold:
                    angles, done = trial, True
                    break
            if done:
                break
    return angles
