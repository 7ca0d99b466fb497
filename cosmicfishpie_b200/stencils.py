"""Finite-difference stencils (reference: cosmicfishpie/fishermatrix/derivatives.py).

A stencil is a list of (m, w) -- evaluate the observable at theta + m*h with integer weight w -- and a
denominator factor D: derivative = (sum_m w f(theta + m h)) / (D h).  Keeping the weights integral and
dividing once reproduces the reference's `(fwd - bwd) / (2 * step)` bit for bit, including exact zeros
for parameters an observable does not depend on.

  3PT      derivatives.py:93-111     (f(+h) - f(-h)) / (2h)                      -- what every production
                                       caller of the reference actually uses
  4PT_FWD  derivatives.py:209-225    (-11 f(0) + 18 f(h) - 9 f(2h) + 2 f(3h)) / (6h)
  5PT      (north_star; not in the reference)  (-f(2h) + 8 f(h) - 8 f(-h) + f(-2h)) / (12h)
"""
from __future__ import annotations

STENCILS = {
    "3PT": ([(+1, +1.0), (-1, -1.0)], 2.0),
    "4PT_FWD": ([(0, -11.0), (1, 18.0), (2, -9.0), (3, 2.0)], 6.0),
    "5PT": ([(+2, -1.0), (+1, +8.0), (-1, -8.0), (-2, +1.0)], 12.0),
}


def stepsize(fiducial_value: float, eps: float) -> float:
    """Relative step, absolute when the fiducial is zero (derivatives.py:140-143)."""
    return fiducial_value * eps if fiducial_value != 0.0 else eps


def get(name: str):
    try:
        return STENCILS[name]
    except KeyError:
        raise ValueError(f"unknown stencil {name!r}; available: {sorted(STENCILS)}") from None
