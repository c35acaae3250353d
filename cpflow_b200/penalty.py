"""Penalty function and regularization (mirror of reference cpflow/penalty.py).

The reference passes Python closures to the optimizer; the CUDA engine needs a declarative table, so
`make_regularization_function` returns a `PenaltyFunction`: a callable (numpy, for inspection and
plotting) that also carries the first-true-wins segment table consumed by the kernels.
"""
import math
from dataclasses import dataclass

import numpy as np


@dataclass
class RegularizationOptions:
    """Reference main.py:328-335 (same names and defaults)."""
    function: str = 'linear'
    ymax: float = 2
    xmax: float = math.pi / 2
    plato_0: float = 0.05
    plato_1: float = 0.05
    plato_2: float = 0.05


def line(x, x0, y0, x1, y1):
    """Reference penalty.py:14-15."""
    return (y1 - y0) / (x1 - x0) * x + (x0 * y1 - x1 * y0) / (x0 - x1)


def _coeffs(x0, y0, x1, y1):
    return (y1 - y0) / (x1 - x0), (x0 * y1 - x1 * y0) / (x0 - x1)


class PenaltyFunction:
    """R(a) for one CP angle.  kind 'piecewise': a <- a mod period, first segment with
    lo < a <= hi gives slope*a + intercept, no match gives 0 (jnp.piecewise, penalty.py:71).
    kind 'l1': |a| (penalty.py:74-76)."""

    def __init__(self, kind, segments=(), period=2 * math.pi):
        self.kind = kind
        self.segments = [tuple(map(float, s)) for s in segments]
        self.period = float(period)

    def __call__(self, a):
        a = np.asarray(a, dtype=np.float64)
        if self.kind == 'l1':
            return np.abs(a)
        am = np.mod(a, self.period)
        out = np.zeros_like(am)
        done = np.zeros(am.shape, dtype=bool)
        for lo, hi, slope, icpt in self.segments:
            cond = (lo < am) & (am <= hi)
            take = cond & ~done
            out = np.where(take, slope * am + icpt, out)
            done |= cond
        return out


def cp_penalty_linear(xmax, ymax, plato_0, plato_1, plato_2):
    """Segment table of the reference's effective cp_penalty_linear (penalty.py:44-71)."""
    pi = math.pi
    rows = [
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
    segs = [(lo, hi) + _coeffs(*ln) for lo, hi, ln in rows]
    segs.append((2 * pi, 3 * pi, 0.0, 1.0))  # penalty.py:56, 67
    return PenaltyFunction('piecewise', segs, 2 * pi)


def cp_penalty_L1():
    """Reference penalty.py:74-76."""
    return PenaltyFunction('l1')


def make_regularization_function(options):
    """Reference penalty.py:79-97; `options` may be the RegularizationOptions class itself (main.py:539)."""
    if options.function == 'linear':
        return cp_penalty_linear(options.xmax, options.ymax, options.plato_0, options.plato_1, options.plato_2)
    elif options.function == 'L1':
        return cp_penalty_L1()
    raise ValueError(f"penalty function {options.function!r} not supported")


def tabulate_penalty(func, period=2 * math.pi, max_segments=16, tol=1e-6, grid=1 << 16):
    """A user `cp_regularization_func` (reference main.py:536-539 accepts any callable R(a)) as the segment table
    the kernels consume.  The function must be `period`-periodic, continuous and piecewise linear with at most
    `max_segments` pieces on [0, period] (the reference's own 'linear' penalty has 9); it is sampled on a fine grid,
    kinks are located from second differences, each piece is fitted through two interior points and the table is checked
    against the samples.  Raises ValueError with the fit error when the function does not fit."""
    xs = np.linspace(0.0, period, grid + 1)
    try:
        ys = np.asarray(func(xs), dtype=np.float64)
        if ys.shape != xs.shape:
            raise TypeError
    except Exception:
        ys = np.array([float(func(float(x))) for x in xs])
    if not np.all(np.isfinite(ys)):
        raise ValueError("cp_regularization_func returned non-finite values on [0, period]")
    scale = max(1.0, float(np.abs(ys).max()))
    probe = np.array([0.3, 1.1, 2.9, 4.4]) * period / (2 * math.pi)
    wrapped = np.array([float(np.asarray(func(float(x + period)))) for x in probe])
    here = np.array([float(np.asarray(func(float(x)))) for x in probe])
    if np.abs(wrapped - here).max() > tol * scale:
        raise ValueError(f"cp_regularization_func is not periodic with period {period:.6g} (the engine evaluates "
                         f"R(a mod period)); use PenaltyFunction('l1') for |a|")
    h = xs[1] - xs[0]
    yi = ys[:-1]                                 # [0, period): the sample AT the period belongs to the next wrap
    d2 = np.abs(yi[2:] - 2 * yi[1:-1] + yi[:-2])
    kink = np.flatnonzero(d2 > 1e-9 * scale + 1e-3 * h * h) + 1           # grid points next to a slope change
    # merge neighbouring flagged points (a kink between two grid points flags both) into one breakpoint
    breaks = [0]
    i = 0
    while i < len(kink):
        j = i
        while j + 1 < len(kink) and kink[j + 1] == kink[j] + 1:
            j += 1
        breaks.append(int(kink[i]) if j == i else None)
        if breaks[-1] is None:
            # the true kink lies between the flagged points: intersect the two neighbouring lines
            a, b = int(kink[i]) - 1, int(kink[j]) + 1
            if a < 1 or b > grid - 1:
                breaks[-1] = int(kink[i])
            else:
                s0 = (ys[a] - ys[a - 1]) / h
                s1 = (ys[b + 1] - ys[b]) / h if b + 1 <= grid else s0
                x_star = (ys[b] - ys[a] + s0 * xs[a] - s1 * xs[b]) / (s0 - s1) if s0 != s1 else xs[kink[i]]
                breaks[-1] = float(min(max(x_star, xs[a]), xs[b]))
        i = j + 1
    pts = [0.0] + [xs[b] if isinstance(b, int) else b for b in breaks[1:]] + [float(period)]
    pts = sorted(set(round(p, 15) for p in pts))
    if len(pts) - 1 > max_segments:
        raise ValueError(f"cp_regularization_func needs {len(pts) - 1} linear pieces on [0, period]; the engine's "
                         f"penalty table holds {max_segments}")
    segs = []
    for k in range(len(pts) - 1):
        lo, hi = pts[k], pts[k + 1]
        m0, m1 = lo + (hi - lo) * 0.25, lo + (hi - lo) * 0.75      # interior points: slopes unaffected by the kinks
        y0, y1 = float(np.asarray(func(m0))), float(np.asarray(func(m1)))
        slope = (y1 - y0) / (m1 - m0)
        icpt = y0 - slope * m0
        segs.append((-math.inf if k == 0 else lo, hi, slope, icpt))
    for (_, hi, s0, i0), (_, _, s1, i1) in zip(segs[:-1], segs[1:]):
        if abs((s0 * hi + i0) - (s1 * hi + i1)) > 1e-4 * scale:
            raise ValueError(f"cp_regularization_func jumps at a = {hi:.6g}: only continuous piecewise-linear "
                             f"functions can be tabulated exactly")
    pf = PenaltyFunction('piecewise', segs, period)
    err = float(np.abs(pf(xs[:-1]) - ys[:-1]).max())
    if err > max(tol * scale, 4 * h * max(abs(s[2]) for s in segs)):
        raise ValueError(f"cp_regularization_func is not piecewise linear with <= {max_segments} pieces: the fitted "
                         f"table deviates by {err:.3e} (allowed {tol * scale:.1e})")
    pf.fit_error = err
    return pf
