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
