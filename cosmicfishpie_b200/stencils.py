"""Finite-difference stencils (reference: cosmicfishpie/fishermatrix/derivatives.py).

A stencil is a list of (m, w) -- evaluate the observable at theta + m*h with integer weight w -- and a
denominator factor D: derivative = (sum_m w f(theta + m h)) / (D h).  Keeping the weights integral and
dividing once reproduces the reference's `(fwd - bwd) / (2 * step)` bit for bit, including exact zeros
for parameters an observable does not depend on.

  3PT      derivatives.py:93-111     (f(+h) - f(-h)) / (2h)                      -- what every production
                                       caller of the reference actually uses
  4PT_FWD  derivatives.py:209-225    (-11 f(0) + 18 f(h) - 9 f(2h) + 2 f(3h)) / (6h)
  5PT      (north_star; not in the reference)  (-f(2h) + 8 f(h) - 8 f(-h) + f(-2h)) / (12h)
  STEM     derivatives.py:348-441    slope of the least-squares line through f at fid*(1 +- eps) for every tabulated
                                       eps (external tables), else at 11 points on fid*(1 + [-5 eps, 5 eps]); photometric
                                       probes only.  A least-squares slope is a fixed linear stencil,
                                       sum_i (x_i - xbar) f_i / sum_i (x_i - xbar)^2, so it runs through the same device
                                       path (:func:`stem_offsets`).  The reference drops the outermost points while the
                                       fit residual exceeds an absolute 1e-3; angular spectra are <= 1e-4, the caller
                                       checks the residual and refuses the (unreachable) trimmed case.
"""
from __future__ import annotations

import numpy as np

STEM_THRESHOLD = 1.0e-3  # derivatives.py:373

STENCILS = {
    "3PT": ([(+1, +1.0), (-1, -1.0)], 2.0),
    "4PT_FWD": ([(0, -11.0), (1, 18.0), (2, -9.0), (3, 2.0)], 6.0),
    "5PT": ([(+2, -1.0), (+1, +8.0), (-1, -8.0), (-2, +1.0)], 12.0),
}


def stepsize(fiducial_value: float, eps: float) -> float:
    """Relative step, absolute when the fiducial is zero (derivatives.py:140-143)."""
    return fiducial_value * eps if fiducial_value != 0.0 else eps


def stem_offsets(fiducial_value: float, eps: float, external_eps=None, numstem: int = 11, mult_eps_factor: int = 5):
    """Abscissae of the SteM fit as absolute offsets from the fiducial value (derivatives.py:356-379): with external
    tables +-eps for every eps of `extfiles["eps_values"]` (the parameter's own eps is ignored), else `numstem` points
    on [-5 eps, 5 eps]; relative to the fiducial unless it is zero."""
    if external_eps is not None:
        e = np.asarray(external_eps, dtype=float)
        d = np.concatenate([-e[::-1], e])
    else:
        d = np.linspace(-eps * mult_eps_factor, eps * mult_eps_factor, numstem)
    return fiducial_value * d if fiducial_value != 0.0 else d


def slope_weights(x):
    """(w, den): least-squares slope of a line through (x_i, f_i) = sum_i w_i f_i / den."""
    x = np.asarray(x, dtype=float)
    w = x - np.mean(x)
    return w, float(np.sum(w * w))


def offsets_and_weights(name: str, fiducial_value: float, eps: float, external_eps=None):
    """[(offset, weight)], denominator of stencil `name` for one parameter: derivative =
    sum_i weight_i f(theta + offset_i) / denominator."""
    if name == "STEM":
        x = stem_offsets(fiducial_value, eps, external_eps)
        w, den = slope_weights(x)
        return list(zip(x.tolist(), w.tolist())), den
    pts, den = get(name)
    h = stepsize(fiducial_value, eps)
    return [(m * h, w) for m, w in pts], den * h


def get(name: str):
    try:
        return STENCILS[name]
    except KeyError:
        raise ValueError(f"unknown stencil {name!r}; available: {sorted(STENCILS)}") from None
