"""Numpy model of the engine's Heisenberg-picture gradient (the math of csrc/heis_impl.cuh).

HS loss L = 1 - |t|^2/N^2, t = Tr(V^dag U), U = G_M ... G_1.  With Y = U V^dag (forward sweep started
from V^dag instead of the identity) t = Tr(Y), and for a gate G_k = exp(-i theta sigma/2):
    dL/dtheta_k = Tr(H_k sigma),   H_k = Herm(s Z_k),  s = i conj(t)/N^2,
    Z_k = G_k..G_1 V^dag G_M..G_{k+1},  Z_M = Y,  Z_{k-1} = G_k^dag Z_k G_k.
H_k is Hermitian, so in the Pauli basis it is a REAL vector h[x, z] (x = bit-flip mask, z = sign mask):
    h[x, z] = Re( i^{|x&z|} * s * W[x, z] ),  W[x, z] = sum_r (-1)^{|z&r|} Y[r, r^x]   (WHT over r)
and every gate acts on h by a REAL linear map: a fused 1q gate rotates (h_X, h_Y, h_Z) of its qubit by
the transpose of its SO(3) matrix, a CP gate rotates two "difference pairs" per (z1, z2) quad.
The gradient sums are single entries of h: no reductions.

This file is test infrastructure (tests/test_heisenberg_model.py checks it against the oracle)."""
import math

import numpy as np

RX, RY, RZ, CP, CZ, CX = 0, 1, 2, 3, 4, 5
SIG = {RX: np.array([[0, 1], [1, 0]], dtype=complex), RY: np.array([[0, -1j], [1j, 0]]),
       RZ: np.array([[1, 0], [0, -1]], dtype=complex)}


def popcount(v):
    return bin(int(v)).count("1")


def rot_mat(kind, a):
    return math.cos(a / 2) * np.eye(2) - 1j * math.sin(a / 2) * SIG[kind]


def so3_of(g):
    """R[c, b] = Tr(sigma_c g sigma_b g^dag)/2  (g sigma_b g^dag = sum_c R[c,b] sigma_c)."""
    s = [SIG[RX], SIG[RY], SIG[RZ]]
    return np.array([[np.real(np.trace(s[c] @ g @ s[b] @ g.conj().T)) / 2 for b in range(3)] for c in range(3)])


def apply_row(y, g, bit):
    """y <- (g on row-index bit `bit`) y."""
    N = y.shape[0]
    idx = np.arange(N)
    lo = idx[(idx >> bit) & 1 == 0]
    hi = lo | (1 << bit)
    a0, a1 = y[lo].copy(), y[hi].copy()
    y[lo] = g[0, 0] * a0 + g[0, 1] * a1
    y[hi] = g[1, 0] * a0 + g[1, 1] * a1


def forward(n, ops, angles, target):
    """Y = U V^dag, up to nothing (exact)."""
    N = 1 << n
    y = np.conj(np.asarray(target, dtype=complex)).T.copy()
    idx = np.arange(N)
    for kind, q0, q1, pi, const in ops:
        a = angles[pi] if pi >= 0 else const
        if kind in (RX, RY, RZ):
            apply_row(y, rot_mat(kind, a), n - 1 - q0)
        elif kind in (CP, CZ):
            sel = ((idx >> (n - 1 - q0)) & 1 == 1) & ((idx >> (n - 1 - q1)) & 1 == 1)
            y[sel] *= np.exp(1j * a) if kind == CP else -1.0
        else:
            raise ValueError("model covers rotations and diagonal two-qubit gates")
    return y


def wht_diagonals(y):
    """W[x, z] = sum_r (-1)^{|z&r|} Y[r, r^x]: gather the x-diagonals, then a WHT over r."""
    N = y.shape[0]
    r = np.arange(N)
    w = np.empty((N, N), dtype=complex)
    for x in range(N):
        a = y[r, r ^ x].copy()
        h = 1
        while h < N:                      # in-place butterflies, bit by bit
            for i in range(N):
                if i & h == 0:
                    a[i], a[i | h] = a[i] + a[i | h], a[i] - a[i | h]
            h <<= 1
        w[x] = a
    return w


def to_pauli(y):
    """Returns (t, h) with h[x, z] = Re(i^{|x&z|} s W[x,z]), s = i conj(t)/N^2."""
    N = y.shape[0]
    w = wht_diagonals(y)
    t = w[0, 0]
    s = 1j * np.conj(t) / N ** 2
    h = np.empty((N, N))
    for x in range(N):
        for z in range(N):
            h[x, z] = np.real((1j) ** (popcount(x & z) % 4) * s * w[x, z])
    return t, h


def conj_su2(h, m3, bit):
    """h <- coefficients of G^dag H G for a 1q gate with (hX,hY,hZ)' = m3 @ (hX,hY,hZ), m3 = so3_of(G).T.
    Components of qubit `bit`: X=(x=1,z=0), Y=(1,1), Z=(0,1), I=(0,0) untouched."""
    N = h.shape[0]
    b = 1 << bit
    out = h.copy()
    for x in range(N):
        if x & b:
            continue
        for z in range(N):
            if z & b:
                continue
            v = np.array([h[x | b, z], h[x | b, z | b], h[x, z | b]])
            nv = m3 @ v
            out[x | b, z], out[x | b, z | b], out[x, z | b] = nv
    return out


def conj_cp(h, a, bit1, bit2):
    """h <- coefficients of CP^dag H CP.  Per (z1, z2) quad e[z1][z2] at fixed x and other z bits:
      group A (x1=1,x2=0): pairs (e00,e01), (e10,e11)
      group B (x1=0,x2=1): pairs (e00,e10), (e01,e11)
      group C (x1=1,x2=1): pairs (e00,e11), (e01,e10) with the second pair SUMMED
    v1 = (p0 - p1)/2, v2 = (p2 -/+ p3)/2 rotate by the CP angle."""
    N = h.shape[0]
    b1, b2 = 1 << bit1, 1 << bit2
    c, s = math.cos(a), math.sin(a)
    out = h.copy()
    for x in range(N):
        x1, x2 = (x & b1) != 0, (x & b2) != 0
        if not (x1 or x2):
            continue
        for z in range(N):
            if z & (b1 | b2):
                continue
            e = {(i, j): h[x, z | (b1 if i else 0) | (b2 if j else 0)] for i in (0, 1) for j in (0, 1)}
            if x1 and not x2:
                p = [(0, 0), (0, 1), (1, 0), (1, 1)]; sg = 1.0
            elif x2 and not x1:
                p = [(0, 0), (1, 0), (0, 1), (1, 1)]; sg = 1.0
            else:
                p = [(0, 0), (1, 1), (0, 1), (1, 0)]; sg = -1.0
            v1 = (e[p[0]] - e[p[1]]) / 2
            v2 = (e[p[2]] - sg * e[p[3]]) / 2
            d1 = (c - 1) * v1 + SGN_CP * s * v2
            d2 = (c - 1) * v2 - SGN_CP * s * v1
            e[p[0]] += d1; e[p[1]] -= d1
            e[p[2]] += d2; e[p[3]] -= sg * d2
            for (i, j), val in e.items():
                out[x, z | (b1 if i else 0) | (b2 if j else 0)] = val
    return out


SGN_CP = 1.0   # fixed by tests against brute force (see _selfcheck)


def grad_hs(n, ops, angles, target):
    """(loss, grad[P]) of the HS loss through the Heisenberg sweep."""
    N = 1 << n
    y = forward(n, ops, angles, target)
    t, h = to_pauli(y)
    loss = 1 - abs(t) ** 2 / N ** 2
    grad = np.zeros(len(angles))
    for kind, q0, q1, pi, const in reversed(ops):
        a = angles[pi] if pi >= 0 else const
        if kind in (RX, RY, RZ):
            b = 1 << (n - 1 - q0)
            if pi >= 0:
                grad[pi] += {RX: h[b, 0], RY: h[b, b], RZ: h[0, b]}[kind]
            h = conj_su2(h, so3_of(rot_mat(kind, a)).T, n - 1 - q0)
        else:
            b1, b2 = 1 << (n - 1 - q0), 1 << (n - 1 - q1)
            if kind == CP and pi >= 0:
                grad[pi] += -0.5 * (h[0, 0] - h[0, b1] - h[0, b2] + h[0, b1 | b2])
            h = conj_cp(h, a if kind == CP else math.pi, n - 1 - q0, n - 1 - q1)
    return loss, grad


def pauli_brute(hm):
    """h[x,z] = Re Tr(P_{x,z} Hm) by explicit Pauli strings (slow; for checks)."""
    N = hm.shape[0]
    n = N.bit_length() - 1
    out = np.empty((N, N))
    for x in range(N):
        for z in range(N):
            P = np.array([[1.0 + 0j]])
            for bit in reversed(range(n)):       # most significant bit first in the kron
                xb, zb = (x >> bit) & 1, (z >> bit) & 1
                s = {(0, 0): np.eye(2), (1, 0): SIG[RX], (1, 1): SIG[RY], (0, 1): SIG[RZ]}[(xb, zb)]
                P = np.kron(P, s)
            out[x, z] = np.real(np.trace(P @ hm))
    return out


# ---------------------------------------------------------------------------------------------
# Forward sweep with merged diagonals (heis_impl.cuh: forward): every fused one-qubit gate is written
# G ~ diag(1, u_out) Ry diag(1, u_in) (ZYZ form, global phase dropped: the HS loss and its gradient do
# not see it); u_in merges with the block's CP phase and with the pending u_out of the previous gate on
# the same qubit into one two-qubit diagonal (1, B, A, A B e^{ia}); u_out stays pending.
# ---------------------------------------------------------------------------------------------
def zyz(g):
    al, be = g[0, 0], g[1, 0]
    cy, sy = abs(al), abs(be)
    pa = al / cy if cy > 1e-150 else 1.0
    pb = be / sy if sy > 1e-150 else 1.0
    return cy, sy, np.conj(pa * pb), pb * np.conj(pa)      # cy, sy, u_in, u_out


def layered_structure(n, ops, angles):
    """Fused gates of a layered template: surface[q] = 2x2, blocks = [(lo, hi, cp_phase, g_lo, g_hi)]."""
    def fuse(seq):
        g = np.eye(2, dtype=complex)
        for kind, a in seq:
            g = rot_mat(kind, a) @ g
        return g
    pend = {q: [] for q in range(n)}
    surface, blocks = {}, []
    cur = None
    for kind, q0, q1, pi, const in ops:
        a = angles[pi] if pi >= 0 else const
        if kind in (RX, RY, RZ):
            pend[q0].append((kind, a))
        else:
            if cur is None:
                surface = {q: fuse(pend[q]) for q in range(n)}
            else:
                blocks.append((cur[0], cur[1], cur[2], fuse(pend[cur[0]]), fuse(pend[cur[1]])))
            for q in (q0, q1) if cur is None else cur[:2]:
                pend[q] = []
            if cur is None:
                pend = {q: [] for q in range(n)}
            lo, hi = min(q0, q1), max(q0, q1)
            cur = (lo, hi, np.exp(1j * a) if kind == CP else -1.0)
    if cur is None:
        surface = {q: fuse(pend[q]) for q in range(n)}
    else:
        blocks.append((cur[0], cur[1], cur[2], fuse(pend[cur[0]]), fuse(pend[cur[1]])))
    return surface, blocks


def forward_merged(n, ops, angles, target):
    """Y' = e^{i gamma} U V^dag through Ry rotations and merged diagonals only."""
    N = 1 << n
    y = np.conj(np.asarray(target, dtype=complex)).T.copy()
    idx = np.arange(N)
    bit = lambda q: (idx >> (n - 1 - q)) & 1
    surface, blocks = layered_structure(n, ops, angles)
    pending = {}
    for q in range(n):
        cy, sy, u_in, u_out = zyz(surface[q])
        y[bit(q) == 1] *= u_in
        apply_row(y, np.array([[cy, -sy], [sy, cy]]), n - 1 - q)
        pending[q] = u_out
    for lo, hi, cp, g_lo, g_hi in blocks:
        cl, sl, uil, uol = zyz(g_lo)
        ch, sh, uih, uoh = zyz(g_hi)
        A, Bv = pending[lo] * uil, pending[hi] * uih
        y[(bit(lo) == 1) & (bit(hi) == 0)] *= A
        y[(bit(lo) == 0) & (bit(hi) == 1)] *= Bv
        y[(bit(lo) == 1) & (bit(hi) == 1)] *= A * Bv * cp
        apply_row(y, np.array([[cl, -sl], [sl, cl]]), n - 1 - lo)
        apply_row(y, np.array([[ch, -sh], [sh, ch]]), n - 1 - hi)
        pending[lo], pending[hi] = uol, uoh
    for q in range(n):
        y[bit(q) == 1] *= pending[q]
    return y


# ---------------------------------------------------------------------------------------------
# Backward sweep with the CP gate split as CP(a) ~ Rz_lo(a/2) Rz_hi(a/2) exp(i (a/4) Z Z): the two Rz are
# absorbed into the SO(3) matrices of the block's fused gates (M' = Rz3(-a/2) M), what is left of the CP gate
# is a plain pair rotation by a/2 (no group-dependent pairing): heis_impl.cuh, zz_bwd.
# ---------------------------------------------------------------------------------------------
def conj_zz(h, a, bit1, bit2):
    """h <- coefficients of ZZ'^dag H ZZ', ZZ' = exp(i (a/4) Z1 Z2)."""
    N = h.shape[0]
    b1, b2 = 1 << bit1, 1 << bit2
    c, s = math.cos(a / 2), math.sin(a / 2)
    out = h.copy()
    for x in range(N):
        x1, x2 = (x & b1) != 0, (x & b2) != 0
        if x1 == x2:
            continue
        sg = 1.0 if x1 else -1.0
        for z in range(N):
            if z & (b1 | b2):
                continue
            e00, e01, e10, e11 = h[x, z], h[x, z | b2], h[x, z | b1], h[x, z | b1 | b2]
            out[x, z] = c * e00 - s * e11
            out[x, z | b1 | b2] = c * e11 + s * e00
            out[x, z | b2] = c * e01 - sg * s * e10
            out[x, z | b1] = c * e10 + sg * s * e01
    return out


def grad_hs_zz(n, ops, angles, target):
    """Same gradient as grad_hs for a layered template, with the Rz halves of every CP gate fused into the
    block's one-qubit gates (checks the bookkeeping of the kernel's backward sweep)."""
    N = 1 << n
    y = forward_merged(n, ops, angles, target)
    t, h = to_pauli(y)
    loss = 1 - abs(t) ** 2 / N ** 2
    grad = np.zeros(len(angles))
    # group the op list into blocks (entangler + following rotations) walking backwards
    i = len(ops) - 1
    while i >= 0:
        j = i
        while j >= 0 and ops[j][0] in (RX, RY, RZ):
            j -= 1
        rots = ops[j + 1:i + 1]
        ent = ops[j] if j >= 0 else None
        # undo the rotations one by one (gradients of the individual angles), except that the SO(3)
        # matrix of the FIRST rotation in time on each block qubit also carries Rz(a/2)
        for kind, q0, q1, pi, const in reversed(rots):
            a = angles[pi] if pi >= 0 else const
            b = 1 << (n - 1 - q0)
            if pi >= 0:
                grad[pi] += {RX: h[b, 0], RY: h[b, b], RZ: h[0, b]}[kind]
            h = conj_su2(h, so3_of(rot_mat(kind, a)).T, n - 1 - q0)
        if ent is not None:
            kind, q0, q1, pi, const = ent
            a = (angles[pi] if pi >= 0 else const) if kind == CP else math.pi
            for q in (q0, q1):
                h = conj_su2(h, so3_of(rot_mat(RZ, a / 2)).T, n - 1 - q)
            b1, b2 = 1 << (n - 1 - q0), 1 << (n - 1 - q1)
            if kind == CP and pi >= 0:
                grad[pi] += -0.5 * (h[0, 0] - h[0, b1] - h[0, b2] + h[0, b1 | b2])
            h = conj_zz(h, a, n - 1 - q0, n - 1 - q1)
        i = j - 1
    return loss, grad


# ---------------------------------------------------------------------------------------------
# Arithmetic forms of the current kernel (heis_impl.cuh): rotations as three lifting shears, the SO(3)
# matrix of a fused gate from its ZYZ data, and the canonical half angle of the entangler.
# ---------------------------------------------------------------------------------------------
def lift_coeffs(c, s):
    """(t, s) of the rotation [[c, -s], [s, c]] with c >= 0: a += t b; b += s a; a += t b."""
    return -s / (1.0 + c), s


def lift_rotate(a, b, t, s):
    a = a + t * b
    b = b + s * a
    a = a + t * b
    return a, b


def so3_from_zyz(ty, sy, u_out, u_in_eff):
    """M = R(G')^T for G' = diag(1, u_out) Ry diag(1, u_in_eff) from the forward data (ty, sy) of the lifting
    form (cos = 1 + ty sy); u_in_eff = u_in e^{i a/2} carries the gate's share of its CP gate.  Returned in the
    staged layout: row for (I, Z) lanes, row for (X, Y) lanes (ka, kb, k00, k01, k02, k10, k11, k12)."""
    cy = 1.0 + ty * sy
    ct, st = cy * cy - sy * sy, 2.0 * cy * sy
    co, so = u_out.real, u_out.imag
    ci, si = u_in_eff.real, u_in_eff.imag
    m = np.array([[co * ct * ci - so * si, so * ct * ci + co * si, -st * ci],
                  [-(co * ct * si + so * ci), co * ci - so * ct * si, st * si],
                  [co * st, so * st, ct]])
    row_iz = np.array([0.0, 1.0, 1.0, 0.0, 0.0, 0.0, m[2, 2], 1.0])
    row_xy = np.array([m[2, 0], m[2, 1], m[0, 0], m[0, 1], m[0, 2], m[1, 0], m[1, 1], m[1, 2]])
    return m, row_iz, row_xy


def canonical_half_angle(a):
    """(c, s, t) of the entangler angle a as the kernel keeps them: c = cos(a/2) >= 0 (CP(a) = CP(a - 2 pi))."""
    c, s = math.cos(a / 2), math.sin(a / 2)
    if c < 0:
        c, s = -c, -s
    return c, s, -s / (1.0 + c)


# ---------------------------------------------------------------------------------------------
# Backward sweep in Z-X-Z form with the Rz factors merged across the entanglers (heis_impl.cuh: backward).
# G ~ Rz(phi_out) Ry(theta) Rz(phi_in) = Rz(phi_out + pi/2) Rx(theta) Rz(phi_in - pi/2).  In the Pauli basis Rx
# mixes (Y, Z): for a qubit whose x bit is a lane bit both sit in the SAME register slot of the two paired lanes, so
# the exchange needs no send computation (2 FMA + 1 SHFL per pair); Rz mixes (X, Y), which one lane holds (4 FMA).
# Everything between the Rx of consecutive gates on a qubit is Z-type (Rz factors, CP halves, ZZ rotations) and
# commutes, so the outgoing Rz of a gate is undone together with the incoming Rz of the NEXT gate on that qubit:
#   per gate: undo Rx(theta), then undo Rz(zeta), e^{i zeta} = pending u_out(prev) * u_in * e^{i a/2}
# (= the forward sweep's merged diagonal A or B times the entangler's half angle; the +-pi/2 cancel).  The gradient
# sums are read in the frame where the gate's own outgoing Rz(phi_out + pi/2) is already undone; the parameter
# phase rotates (S_X, S_Y) back with u_out.
# ---------------------------------------------------------------------------------------------
def layered_gates(n, ops, angles):
    """Fused gates in kernel slot order (surface gate of qubit q: slot q; block k: lower-qubit gate slot n + 2k,
    higher-qubit gate n + 2k + 1), each a list of (kind, pidx, angle) in time order, and the blocks
    [(lo, hi, kind, pidx, angle)]."""
    gates = {q: [] for q in range(n)}
    blocks = []
    for kind, q0, q1, pi, const in ops:
        a = angles[pi] if pi >= 0 else const
        if kind in (RX, RY, RZ):
            if not blocks:
                gates[q0].append((kind, pi, a))
            else:
                lo, hi = blocks[-1][0], blocks[-1][1]
                k = len(blocks) - 1
                gates.setdefault(n + 2 * k + (0 if q0 == lo else 1), []).append((kind, pi, a))
        else:
            k = len(blocks)
            blocks.append((min(q0, q1), max(q0, q1), kind, pi, a if kind == CP else math.pi))
            gates[n + 2 * k] = []
            gates[n + 2 * k + 1] = []
    return [gates[s] for s in range(n + 2 * len(blocks))], blocks


def rz_undo(h, u, bit):
    """h <- coefficients of Rz(zeta)^dag H Rz(zeta), e^{i zeta} = u: (X, Y) -> (c X + s Y, -s X + c Y)."""
    N = h.shape[0]
    b = 1 << bit
    c, s = u.real, u.imag
    out = h.copy()
    for x in range(N):
        if x & b:
            continue
        for z in range(N):
            if z & b:
                continue
            X, Y = h[x | b, z], h[x | b, z | b]
            out[x | b, z], out[x | b, z | b] = c * X + s * Y, -s * X + c * Y
    return out


def rx_undo(h, ct, st, bit):
    """h <- coefficients of Rx(theta)^dag H Rx(theta): (Y, Z) -> (ct Y + st Z, -st Y + ct Z)."""
    N = h.shape[0]
    b = 1 << bit
    out = h.copy()
    for x in range(N):
        if x & b:
            continue
        for z in range(N):
            if z & b:
                continue
            Y, Z = h[x | b, z | b], h[x, z | b]
            out[x | b, z | b], out[x, z | b] = ct * Y + st * Z, -st * Y + ct * Z
    return out


def grad_hs_zxz(n, ops, angles, target):
    """(loss, grad[P]) through the merged Z-X-Z backward sweep, including the parameter phase's frame change of the
    gradient sums and its chain rule through the fused rotations."""
    N = 1 << n
    gates, blocks = layered_gates(n, ops, angles)
    y = forward_merged(n, ops, angles, target)
    t, h = to_pauli(y)
    loss = 1 - abs(t) ** 2 / N ** 2

    def fuse(seq):
        g = np.eye(2, dtype=complex)
        for kind, _, a in seq:
            g = rot_mat(kind, a) @ g
        return g
    data = [zyz(fuse(g)) for g in gates]               # cy, sy, u_in, u_out
    qubit_of = list(range(n)) + [q for lo, hi, *_ in blocks for q in (lo, hi)]
    prev, last = {}, {q: q for q in range(n)}
    for s in range(n, len(gates)):
        prev[s] = last[qubit_of[s]]
        last[qubit_of[s]] = s
    bit = lambda q: n - 1 - q
    for q in range(n):                                   # tail: the outgoing Rz of the last gate on every qubit
        h = rz_undo(h, data[last[q]][3] * 1j, bit(q))
    S = {}
    grad = np.zeros(len(angles))

    def undo_gate(h, s, u_in_eff):
        b = 1 << bit(qubit_of[s])
        S[s] = np.array([h[b, 0], h[b, b], h[0, b]])
        cy, sy = data[s][0], data[s][1]
        h = rx_undo(h, cy * cy - sy * sy, 2 * cy * sy, bit(qubit_of[s]))
        return rz_undo(h, u_in_eff, bit(qubit_of[s]))
    for k in reversed(range(len(blocks))):
        lo, hi, kind, pi, a = blocks[k]
        c, s_, _ = canonical_half_angle(a)
        for s in (n + 2 * k + 1, n + 2 * k):
            h = undo_gate(h, s, data[prev[s]][3] * data[s][2] * complex(c, s_))
        b1, b2 = 1 << bit(lo), 1 << bit(hi)
        if kind == CP and pi >= 0:
            grad[pi] += -0.5 * (h[0, 0] - h[0, b1] - h[0, b2] + h[0, b1 | b2])
        h = conj_zz(h, 2 * math.atan2(s_, c), bit(lo), bit(hi))
    for q in reversed(range(n)):
        h = undo_gate(h, q, data[q][2] * (-1j))
    # parameter phase: back to the gate's output frame, then the chain rule through G = R_2 R_1 R_0
    for s, seq in enumerate(gates):
        w = data[s][3] * 1j
        sx, sy, sz = S[s]
        vec = np.array([w.real * sx - w.imag * sy, w.imag * sx + w.real * sy, sz])
        for kind, pi, a in reversed(seq):
            if pi >= 0:
                grad[pi] += vec[{RX: 0, RY: 1, RZ: 2}[kind]]
            vec = so3_of(rot_mat(kind, a)).T @ vec
    return loss, grad
