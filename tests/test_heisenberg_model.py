"""The math of the Heisenberg-picture kernel (csrc/heis_impl.cuh), restated in numpy by
tools/heisenberg_model.py, against the oracle: Pauli-basis pivot, SO(3) and CP maps, and the full
loss + gradient of the HS loss (matrix_utils.py:35-42) for CP / CZ templates.  CPU only."""
import math
import os
import sys

import numpy as np
import pytest
import torch
from scipy.stats import unitary_group

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import heisenberg_model as H  # noqa: E402
from oracle import cpflow_oracle as O  # noqa: E402


def _embed(g, bit, n):
    G = np.array([[1.0 + 0j]])
    for b in reversed(range(n)):
        G = np.kron(G, g if b == bit else np.eye(2))
    return G


def test_pivot_and_gate_maps_against_brute_force():
    rng = np.random.default_rng(1)
    n, N = 3, 8
    y = rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))
    t, h = H.to_pauli(y)
    A = 1j * np.conj(t) / N ** 2 * y
    Hm = (A + A.conj().T) / 2
    hb = H.pauli_brute(Hm)
    assert np.abs(h - hb).max() < 1e-14
    g = H.rot_mat(H.RZ, 0.7) @ H.rot_mat(H.RY, -1.1) @ H.rot_mat(H.RX, 2.2)
    for bit in range(n):
        G = _embed(g, bit, n)
        assert np.abs(H.conj_su2(hb, H.so3_of(g).T, bit) - H.pauli_brute(G.conj().T @ Hm @ G)).max() < 1e-14
    idx = np.arange(N)
    for b1, b2 in [(0, 1), (1, 0), (0, 2), (2, 1)]:
        for a in (0.83, math.pi, -2.1):
            d = np.ones(N, dtype=complex)
            d[((idx >> b1) & 1 == 1) & ((idx >> b2) & 1 == 1)] = np.exp(1j * a)
            Cm = np.diag(d)
            assert np.abs(H.conj_cp(hb, a, b1, b2) - H.pauli_brute(Cm.conj().T @ Hm @ Cm)).max() < 1e-14


def test_so3_from_quaternion_formula():
    """heis_su2_to_so3: w = Re alpha, z = -Im alpha, y = Re beta, x = -Im beta."""
    rng = np.random.default_rng(3)
    for _ in range(20):
        g = H.rot_mat(H.RZ, rng.uniform(-7, 7)) @ H.rot_mat(H.RY, rng.uniform(-7, 7)) @ H.rot_mat(H.RX, rng.uniform(-7, 7))
        al, be = g[0, 0], g[1, 0]
        w, z, y, x = al.real, -al.imag, be.real, -be.imag
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        assert np.abs(R - H.so3_of(g)).max() < 1e-14


@pytest.mark.parametrize("n,layer,K,rg,ent", [(3, O.chain_layer(3), 5, "xyz", "cp"), (4, O.chain_layer(4), 7, "xyz", "cp"),
                                              (4, [[0, 1], [0, 2], [0, 3]], 4, "xz", "cp"),
                                              (3, O.connected_layer(3), 6, "zyx", "cp"), (2, [[0, 1]], 3, "xyz", "cp")])
def test_heisenberg_gradient_equals_oracle(n, layer, K, rg, ent):
    rng = np.random.default_rng(K)
    anz = O.cp_ansatz(layer, K, rg)
    ops = O.ansatz_program(anz)
    tgt = unitary_group.rvs(1 << n, random_state=3)
    a = rng.uniform(0, 2 * np.pi, anz.num_angles)
    l, g = H.grad_hs(n, ops, a, tgt)
    ol, _, og = O.loss_and_grad_batched(n, ops, torch.tensor(a)[None], "hs", torch.tensor(tgt))
    assert abs(l - float(ol[0])) < 1e-13
    assert np.abs(g - og[0].numpy()).max() < 1e-13
    l2, g2 = O.hand_adjoint_grad(n, ops, a, "hs", tgt)
    assert abs(l - l2) < 1e-13 and np.abs(g - g2).max() < 1e-13


def test_heisenberg_gradient_with_cz_and_constants():
    """CZ entanglers (is_cz) and constant angles (projected / frozen CP gates, cp_utils.py:70-108)."""
    n = 3
    ops = [(H.RZ, 0, -1, 0, 0.0), (H.RX, 1, -1, 1, 0.0), (H.CZ, 0, 1, -1, 0.0), (H.RY, 0, -1, 2, 0.0),
           (H.CP, 1, 2, -1, math.pi), (H.RX, 2, -1, 3, 0.0), (H.CP, 0, 2, 4, 0.0), (H.RZ, 1, -1, -1, 0.37),
           (H.RY, 2, -1, 5, 0.0)]
    a = np.random.default_rng(0).uniform(0, 6.28, 6)
    tgt = unitary_group.rvs(8, random_state=5)
    l, g = H.grad_hs(n, ops, a, tgt)
    l2, g2 = O.hand_adjoint_grad(n, ops, a, "hs", tgt)
    assert abs(l - l2) < 1e-13 and np.abs(g - g2).max() < 1e-13


@pytest.mark.parametrize("n,layer,K,rg", [(3, O.chain_layer(3), 5, "xyz"), (4, [[0, 1], [0, 2], [0, 3]], 7, "xz"),
                                          (4, O.connected_layer(4), 9, "zyx"), (2, [[0, 1]], 3, "xyz")])
def test_merged_diagonal_forward(n, layer, K, rg):
    """Forward sweep in ZYZ form with merged diagonals and pending phases (heis_impl.cuh: forward) equals the
    gate-by-gate forward up to a global phase; the Pauli vector that seeds the backward sweep is identical."""
    rng = np.random.default_rng(2)
    anz = O.cp_ansatz(layer, K, rg)
    ops = O.ansatz_program(anz)
    tgt = unitary_group.rvs(1 << n, random_state=3)
    a = rng.uniform(0, 2 * np.pi, anz.num_angles)
    y, y2 = H.forward(n, ops, a, tgt), H.forward_merged(n, ops, a, tgt)
    ph = y2.flat[0] / y.flat[0]
    assert abs(abs(ph) - 1) < 1e-13 and np.abs(y * ph - y2).max() < 1e-13
    (t1, h1), (t2, h2) = H.to_pauli(y), H.to_pauli(y2)
    assert abs(abs(t1) - abs(t2)) < 1e-13 and np.abs(h1 - h2).max() < 1e-14


def test_zz_split_of_the_cp_gate():
    """CP(a) ~ Rz Rz exp(i a/4 ZZ): the pair rotation of the ZZ part against brute force, and the full gradient
    with the Rz halves folded into the block's fused gates (the kernel's backward sweep) against the oracle."""
    rng = np.random.default_rng(2)
    N = 8
    y = rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))
    Hm = (y + y.conj().T) / 2
    hb = H.pauli_brute(Hm)
    idx = np.arange(N)
    for b1, b2 in [(0, 1), (1, 0), (2, 0), (1, 2)]:
        zz = np.array([(1 - 2 * ((i >> b1) & 1)) * (1 - 2 * ((i >> b2) & 1)) for i in idx])
        U = np.diag(np.exp(1j * (0.77 / 4) * zz))
        assert np.abs(H.conj_zz(hb, 0.77, b1, b2) - H.pauli_brute(U.conj().T @ Hm @ U)).max() < 1e-14
    for n, layer, K, rg in [(3, O.chain_layer(3), 5, "xyz"), (4, [[0, 1], [0, 2], [0, 3]], 7, "xz"),
                            (4, O.connected_layer(4), 9, "zyx")]:
        anz = O.cp_ansatz(layer, K, rg)
        ops = O.ansatz_program(anz)
        tgt = unitary_group.rvs(1 << n, random_state=3)
        a = rng.uniform(0, 2 * np.pi, anz.num_angles)
        l, g = H.grad_hs_zz(n, ops, a, tgt)
        l2, g2 = O.hand_adjoint_grad(n, ops, a, "hs", tgt)
        assert abs(l - l2) < 1e-13 and np.abs(g - g2).max() < 1e-13


def test_lifting_form_of_the_rotations():
    """heis_impl.cuh / engine.cuh ry_lift, zz_quad: three shears with t = -s/(1+c) are the rotation, for every
    angle the kernel can meet (phi in [0, pi/2] for Ry, |a/2| <= pi/2 for the canonical entangler angle)."""
    rng = np.random.default_rng(5)
    for phi in list(rng.uniform(0, math.pi / 2, 20)) + [0.0, math.pi / 2]:
        c, s = math.cos(phi), math.sin(phi)
        t, s2 = H.lift_coeffs(c, s)
        assert -1.0 - 1e-15 <= t <= 0.0
        a, b = rng.normal(size=2)
        la, lb = H.lift_rotate(a, b, t, s2)
        assert abs(la - (c * a - s * b)) < 1e-14 and abs(lb - (s * a + c * b)) < 1e-14
        assert abs((1.0 + t * s) - c) < 1e-15          # the staging recovers cos as 1 + t s
    for a in list(rng.uniform(-20, 20, 50)) + [0.0, math.pi, 2 * math.pi, -math.pi]:
        c, s, t = H.canonical_half_angle(a)
        assert c >= 0 and abs(t) <= 1.0 + 1e-12
        # same gate: e^{ia} from the half angle
        assert abs(complex(c * c - s * s, 2 * c * s) - np.exp(1j * a)) < 1e-13
        x, y = rng.normal(size=2)
        lx, ly = H.lift_rotate(x, y, t, s)
        assert abs(lx - (c * x - s * y)) < 1e-13 and abs(ly - (s * x + c * y)) < 1e-13


def test_so3_from_zyz_data_and_staged_rows():
    """stage_layer / zyz_to_so3: the SO(3) matrix of G Rz(a/2) from (ty, sy, u_out, u_in) equals so3_of of the
    2x2 matrix, whichever sign the canonical half angle took, and the staged lane rows reproduce M (X, Y, Z)."""
    rng = np.random.default_rng(6)
    for _ in range(20):
        g = H.rot_mat(H.RZ, rng.uniform(0, 7)) @ H.rot_mat(H.RY, rng.uniform(0, 7)) @ H.rot_mat(H.RX, rng.uniform(0, 7))
        cy, sy, u_in, u_out = H.zyz(g)
        ty, _ = H.lift_coeffs(cy, sy)
        a = rng.uniform(-10, 10)
        c, s, _ = H.canonical_half_angle(a)
        m, row_iz, row_xy = H.so3_from_zyz(ty, sy, u_out, u_in * complex(c, s))
        # reference: SO(3) of G Rz(a'/2), a' = a or a - 2 pi (the canonical representative)
        ah = 2 * math.atan2(s, c)
        ref = H.so3_of(g @ H.rot_mat(H.RZ, ah / 2)).T
        assert np.abs(m - ref).max() < 1e-13
        X, Y, Z, I = rng.normal(size=4)
        # lane holding (I, Z): e0 = I, e1 = Z; lane holding (X, Y): e0 = X, e1 = Y
        send_iz = row_iz[0] * I + row_iz[1] * Z
        send_xy = row_xy[0] * X + row_xy[1] * Y
        i2 = row_iz[2] * I + row_iz[3] * Z + row_iz[4] * send_xy
        z2 = row_iz[5] * I + row_iz[6] * Z + row_iz[7] * send_xy
        x2 = row_xy[2] * X + row_xy[3] * Y + row_xy[4] * send_iz
        y2 = row_xy[5] * X + row_xy[6] * Y + row_xy[7] * send_iz
        want = ref @ np.array([X, Y, Z])
        assert abs(i2 - I) < 1e-14 and np.abs(np.array([x2, y2, z2]) - want).max() < 1e-13


@pytest.mark.parametrize("n,layer,K,rg,ent", [(3, O.chain_layer(3), 5, "xyz", "cp"), (4, O.chain_layer(4), 7, "xyz", "cp"),
                                              (4, [[0, 1], [0, 2], [0, 3]], 4, "xz", "cp"),
                                              (3, O.connected_layer(3), 6, "zyx", "cp"), (2, [[0, 1]], 3, "xyz", "cp"),
                                              (3, O.chain_layer(3), 4, "xyz", "cz")])
def test_merged_zxz_backward_sweep(n, layer, K, rg, ent):
    """The kernel's backward sweep: Rx / merged-Rz undo per gate, gradient sums in the rotated frame, frame change
    and chain rule of the parameter phase: equals the oracle gradient."""
    rng = np.random.default_rng(K + 11)
    anz = O.cp_ansatz(layer, K, rg)
    ops = O.ansatz_program(anz)
    if ent == "cz":      # CZ entanglers (projected CP gates): no parameter, a = pi exactly
        ops = [(H.CZ, q0, q1, -1, 0.0) if kind == H.CP else (kind, q0, q1, pi, c) for kind, q0, q1, pi, c in ops]
    tgt = unitary_group.rvs(1 << n, random_state=4)
    a = rng.uniform(-2 * np.pi, 4 * np.pi, anz.num_angles)
    l, g = H.grad_hs_zxz(n, ops, a, tgt)
    l2, g2 = O.hand_adjoint_grad(n, ops, a, "hs", tgt)
    assert abs(l - l2) < 1e-13 and np.abs(g - g2).max() < 1e-12
