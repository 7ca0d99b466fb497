"""The derivative engine as a callable seam (reference: fishermatrix/derivatives.py:18-91, the de-facto plugin API
`derivatives(observable, fiducial, special_deriv_function, freeparams).result`).

Two routes, same result structure {parameter: {key: array}} (keys "ells" / "z_bins" pass through, as in
derivatives.py:167-187):

* `observable` is the bound `get_obs` of a `SpectroDerivs` or `getcls` of a `PhotoCov` of this package -- what
  `FisherMatrix.compute()` itself uses: the whole stencil is ONE fused device job (every stencil point's lattice or
  C_ell on the GPU, the stencil applied there), reached through the object's own `compute_derivs`;
* any other callable (a user's own observable: a dict -> dict of host arrays): it is evaluated at the stencil points
  in the reference's order and the stencil -- a fixed linear combination with integral weights and ONE division,
  `stencils.offsets_and_weights` -- is applied to what it returns.  Those arrays live on the host because the callable
  does; nothing here falls back from a device path.

Stencils: "3PT" (derivatives.py:113-207), "4PT_FWD" (:209-346), "STEM" (:348-441; with external tables the abscissae
are +-eps for every tabulated eps; the reference's trimming of points whose fit residual exceeds 1e-3 is checked and
refused), and "5PT".  "POLY" does not exist in the reference either (it raises AttributeError there)."""
from __future__ import annotations

import copy

import numpy as np

from . import stencils

_PASS_THROUGH = {"ells", "z_bins"}


class derivatives:
    def __init__(self, observable, fiducial, special_deriv_function=None, derivatives_type="3PT", freeparams=dict(),
                 observables_type=None, external_settings=None, feed_lvl=None):
        from . import config as cfg

        cur = cfg._CURRENT
        self.freeparams = freeparams if freeparams != dict() else cur.freeparams
        self.observable = observable
        self.fiducial = fiducial
        self.special = special_deriv_function
        self.feed_lvl = feed_lvl if feed_lvl is not None else (cur.settings["feedback"] if cur else 0)
        self.observables_type = observables_type if observables_type is not None else (cur.obs if cur else ["plain"])
        self.external_settings = external_settings if external_settings is not None else (cur.external if cur else None)
        self.derivatives_type = derivatives_type if derivatives_type is not None else cur.settings["derivatives"]
        if self.derivatives_type not in ("3PT", "4PT_FWD", "5PT", "STEM"):
            raise ValueError("ERROR: I don't know this derivative type!!!")
        self.result = self._device_route() if self._own_observable() else self._generic_route()

    # --- this package's own observables: one fused device job ---------------------------------------------------
    def _own_observable(self):
        from .photo import PhotoCov
        from .spectro import SpectroDerivs

        owner = getattr(self.observable, "__self__", None)
        name = getattr(self.observable, "__name__", "")
        return (isinstance(owner, SpectroDerivs) and name == "get_obs") or (isinstance(owner, PhotoCov) and name == "getcls")

    def _device_route(self):
        from .spectro import SpectroDerivs

        owner = self.observable.__self__
        if any(self.fiducial.get(k) != v for k, v in getattr(owner, "fiducial_allpars", getattr(owner, "allparsfid", {})).items()
               if k in self.fiducial and not isinstance(v, str)):
            return self._generic_route()  # derivatives about another point than the object's own fiducial
        prev = owner.stencil
        owner.stencil = self.derivatives_type
        try:
            if isinstance(owner, SpectroDerivs):
                owner.job = None
                return owner.compute_derivs(freeparams=self.freeparams)
            owner._job = owner._cls_all = None
            owner.build_job(self.freeparams)
            owner._job.run()
            return owner.compute_derivs()
        finally:
            owner.stencil = prev

    # --- any other callable -------------------------------------------------------------------------------------
    def _combine(self, outs, weights, den):
        """sum_i w_i f_i / den on dicts of arrays (or plain arrays for observables_type 'plain')."""
        first = outs[0]
        if not isinstance(first, dict):
            acc = sum(w * np.asarray(o, dtype=float) for w, o in zip(weights, outs))
            return acc / den
        out = {}
        for key in first:
            if key in _PASS_THROUGH:
                out[key] = first[key]
            else:
                acc = sum(w * np.asarray(o[key], dtype=float) for w, o in zip(weights, outs))
                out[key] = acc / den
        return out

    def _generic_route(self):
        ext_eps = (self.external_settings or {}).get("eps_values") if self.derivatives_type == "STEM" else None
        result = {}
        for par in self.freeparams:
            if self.special is not None:
                special = self.special(par)
                if special is not None:  # an exact derivative supplied by the caller (e.g. shot-noise parameters)
                    result[par] = special
                    continue
            offs, den = stencils.offsets_and_weights(self.derivatives_type, self.fiducial[par], self.freeparams[par], ext_eps)
            outs, weights = [], []
            for off, w in offs:
                point = copy.deepcopy(self.fiducial)
                point[par] = point[par] + off
                outs.append(self.observable(point))
                weights.append(w)
            if self.derivatives_type == "STEM":
                self._check_stem(par, offs, outs)
            result[par] = self._combine(outs, weights, den)
        return result

    @staticmethod
    def _check_stem(par, offs, outs):
        """derivatives.py:402-419 drops the outermost abscissae while the straight-line fit leaves a squared residual
        above an absolute 1e-3 in any element; a fixed stencil cannot do that, so the case is detected and refused."""
        x = np.array([o for o, _ in offs])
        keys = [k for k in outs[0] if k not in _PASS_THROUGH] if isinstance(outs[0], dict) else [None]
        w, den = stencils.slope_weights(x)
        for key in keys:
            y = np.array([np.asarray(o[key] if key is not None else o, dtype=float) for o in outs])
            slope = np.tensordot(w, y, axes=(0, 0)) / den
            fit = np.mean(y, axis=0) + (x - np.mean(x)).reshape((-1,) + (1,) * (y.ndim - 1)) * slope
            if np.max(np.sum((y - fit) ** 2, axis=0)) > stencils.STEM_THRESHOLD:
                raise NotImplementedError(f"SteM derivative of {par}: fit residual above {stencils.STEM_THRESHOLD}; the "
                                          "reference would trim the stencil element by element, which is not supported")
