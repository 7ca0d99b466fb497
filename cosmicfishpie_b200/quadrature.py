"""Quadrature weight vectors that reproduce the reference's scipy integrators.

The Fisher reduction runs on the GPU as a weighted sum over lattice cells, so the reference's
`scipy.integrate.simpson(y, x=...)` calls (cosmicfish.py:510-511, spectro_like.py:178-181,53) are
turned into weight vectors once on the host: simpson(y, x) == sum(w * y) to rounding.
"""
from __future__ import annotations

import numpy as np


_SIMPSON_MEMO = {}


def simpson_weights(x) -> np.ndarray:
    """Weights of composite Simpson on the sample points `x` as scipy >= 1.11 computes it:
    per-pair rule with the actual (possibly unequal) spacings, and for an even number of samples
    Cartwright's correction on the last interval.  Remembered per grid (every forecast of a lattice asks again);
    the result is read-only."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    key = x.tobytes()
    hit = _SIMPSON_MEMO.get(key)
    if hit is None:
        if len(_SIMPSON_MEMO) >= 32:
            _SIMPSON_MEMO.clear()
        hit = _SIMPSON_MEMO[key] = _simpson_weights(x)
        hit.setflags(write=False)
    return hit


def _simpson_weights(x) -> np.ndarray:
    n = x.size
    w = np.zeros(n)
    if n == 1:
        return w
    if n == 2:
        w[:] = 0.5 * (x[1] - x[0])
        return w
    h = np.diff(x)

    def basic(stop):  # pairs (0,1,2), (2,3,4), ... up to index `stop`
        i = np.arange(0, stop - 1, 2)
        h0, h1 = h[i], h[i + 1]
        hs, hp, r = h0 + h1, h0 * h1, h0 / h1
        np.add.at(w, i, hs / 6.0 * (2.0 - 1.0 / r))
        np.add.at(w, i + 1, hs / 6.0 * (hs * hs / hp))
        np.add.at(w, i + 2, hs / 6.0 * (2.0 - r))

    if n % 2 == 1:
        basic(n - 1)
    else:
        basic(n - 2)
        h0, h1 = h[-2], h[-1]
        w[-1] += (2.0 * h1**2 + 3.0 * h0 * h1) / (6.0 * (h0 + h1))
        w[-2] += (h1**2 + 3.0 * h1 * h0) / (6.0 * h0)
        w[-3] -= h1**3 / (6.0 * h0 * (h0 + h1))
    return w


def trapezoid_weights(x) -> np.ndarray:
    """sum(w*y) == scipy.integrate.trapezoid(y, x)."""
    x = np.asarray(x, dtype=np.float64)
    w = np.zeros(x.size)
    d = np.diff(x)
    w[:-1] += d / 2.0
    w[1:] += d / 2.0
    return w
