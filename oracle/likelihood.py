"""Oracle: photometric and spectroscopic likelihoods (test infrastructure only; see oracle/__init__.py).

Restates:
  likelihood/photo_like.py:28-97     _cells_from_cls, _chi2_per_obs
  likelihood/photo_like.py:180-256   PhotometricLikelihood.compute_chi2 (incl. the duplicated LL term)
  likelihood/spectro_like.py:25-56   observable_Pgg, legendre_Pgg
  likelihood/spectro_like.py:69-140  compute_covariance_legendre, compute_chi2_legendre
  likelihood/spectro_like.py:143-233 compute_wedge_chi2, compute_theory_spectro
  utilities/legendre_tools.py:56-123 Wigner-3j sum matrices m00..m04
"""
from __future__ import annotations

from copy import copy

import numpy as np
from scipy.integrate import simpson
from scipy.interpolate import make_interp_spline
from scipy.special import eval_legendre

from .cosmo import DEFAULT_SETTINGS
from .photo import Cls
from .spectro import GalSpectro, SpectroCov


# ------------------------------------------------------------------------------ photometric
def cells_from_cls(result, specs, observables):
    """photo_like.py:28-75."""
    ells = np.array(result["ells"], copy=True)
    data = {"ells": ells}
    wl = list(specs["binrange_WL"]) if "WL" in observables else []
    gc = list(specs["binrange_GCph"]) if "GCph" in observables else []
    if wl:
        c = np.empty((len(ells), len(wl), len(wl)))
        for i in wl:
            for j in wl:
                noise = (specs["ellipt_error"] ** 2.0) / specs["ngalbin_WL"][i - 1] if i == j else 0.0
                c[:, i - 1, j - 1] = result[f"WL {i}xWL {j}"] + noise
        data["Cell_LL"] = c
    if gc:
        c = np.empty((len(ells), len(gc), len(gc)))
        for i in gc:
            for j in gc:
                noise = (1.0 / specs["ngalbin_GCph"][i - 1]) if i == j else 0.0
                c[:, i - 1, j - 1] = result[f"GCph {i}xGCph {j}"] + noise
        data["Cell_GG"] = c
    if wl and gc:
        c = np.empty((len(ells), len(gc), len(wl)))
        for i in gc:
            for j in wl:
                c[:, i - 1, j - 1] = result[f"GCph {i}xWL {j}"]
        data["Cell_GL"] = c
    return data


def chi2_per_obs(cell_fid, cell_th, ells, dells):
    """photo_like.py:78-97 (determinant-ratio form)."""
    dfid = np.linalg.det(cell_fid)
    dth = np.linalg.det(cell_th)
    dmix = 0.0
    for idx in range(cell_fid.shape[-1]):
        mix = copy(cell_th)
        mix[:, idx, :] = cell_fid[:, idx, :]
        dmix += np.linalg.det(mix)
    integrand = (2 * ells + 1) * (dmix / dth + np.log(dth / dfid) - cell_fid.shape[-1])
    return float(np.sum(np.concatenate([((integrand[1:] + integrand[:-1]) / 2) * dells[:-1], integrand[-1:] * dells[-1:]])))


def ell_subset(ells, limit):
    """photo_like.py:189-204."""
    working = np.array(ells, copy=True)
    if limit is not None:
        idx = np.searchsorted(working, limit)
        working = np.insert(working, idx, limit)
    else:
        idx = len(working)
    deltas = np.diff(working)
    if deltas.size < idx:
        deltas = np.concatenate([deltas, [deltas[-1] if deltas.size else 1.0]])
    else:
        deltas = deltas[:idx]
    return idx, working[:idx], deltas


def photo_chi2(data, theory, specs):
    """photo_like.py:180-256."""
    ells = data["ells"]
    chi2 = 0.0
    fw, fg = specs.get("fsky_WL", 1.0), specs.get("fsky_GCph", 1.0)
    lw, lg = specs.get("lmax_WL"), specs.get("lmax_GCph")
    lx = min(lw, lg) if (lw is not None and lg is not None) else None
    if "Cell_LL" in data and "Cell_LL" in theory:
        n, e, d = ell_subset(ells, lw)
        chi2 += fw * chi2_per_obs(data["Cell_LL"][:n], theory["Cell_LL"][:n], e, d)
    if "Cell_GG" in data and "Cell_GG" in theory:
        n, e, d = ell_subset(ells, lg)
        chi2 += fg * chi2_per_obs(data["Cell_GG"][:n], theory["Cell_GG"][:n], e, d)
    if "Cell_GL" in data and "Cell_GL" in theory and "Cell_GG" in theory and "Cell_LL" in theory:
        n, e, d = ell_subset(ells, lx)

        def big(c):
            return np.block([[c["Cell_LL"][:n], np.transpose(c["Cell_GL"], (0, 2, 1))[:n]],
                             [c["Cell_GL"][:n], c["Cell_GG"][:n]]])

        chi2 += np.sqrt(fw * fg) * chi2_per_obs(big(data), big(theory), e, d)
        chi2 += fw * chi2_per_obs(data["Cell_LL"][:n], theory["Cell_LL"][:n], e, d)
    return float(chi2)


class PhotoLikelihood:
    """PhotometricLikelihood (photo_like.py:100-256) on the oracle's Cls."""

    def __init__(self, bank, specs, settings, observables, fiducial_cosmopars, photopars, IApars, biaspars, lum):
        self.bank, self.specs, self.observables, self.lum = bank, specs, list(observables), lum
        s = dict(DEFAULT_SETTINGS)
        s.update(settings or {})
        self.settings = s
        self.fid = (dict(fiducial_cosmopars), dict(photopars), dict(IApars), dict(biaspars))
        self.data = self.cells({})

    def cells(self, params):
        p = dict(params)
        groups = []
        for tmpl in self.fid:
            g = dict(tmpl)
            for k in tmpl:
                if k in p:
                    g[k] = p.pop(k)
            groups.append(g)
        cp, pp, ia, bp = groups
        c = Cls(self.bank.cosmo(cp), self.specs, self.settings, self.observables, pp, ia, bp, self.lum)
        return cells_from_cls(c.compute_all(), self.specs, self.observables)

    def loglike(self, params):
        return -0.5 * photo_chi2(self.data, self.cells(params), self.specs)


# ------------------------------------------------------------------------------ spectroscopic
_M = {
    "m00": [1.0, 0, 0, 0, 0.2, 0, 0, 0, 0.111111111111111],
    "m22": [0.2, 2 / 35, 2 / 35, 2 / 35, 0.0857142857142857, 12 / 385, 2 / 35, 12 / 385, 0.0397158397158397],
    "m44": [0.111111111111111, 20 / 693, 18 / 1001, 20 / 693, 0.0397158397158397, 20 / 1001, 18 / 1001, 20 / 1001,
            0.0310865604983252],
    "m02": [0, 0.2, 0, 0.2, 2 / 35, 0.0571428571428571, 0, 0.0571428571428571, 20 / 693],
    "m24": [0, 0.0571428571428571, 20 / 693, 0.0571428571428571, 12 / 385, 0.0397158397158397, 20 / 693,
            0.0397158397158397, 20 / 1001],
    "m04": [0, 0, 0.111111111111111, 0, 2 / 35, 20 / 693, 0.111111111111111, 20 / 693, 18 / 1001],
}
M = {k: np.array(v, dtype=float).reshape(3, 3) for k, v in _M.items()}
LEG_ELLS = np.arange(0, 5, 2, dtype="int")


class SpectroLikelihoodOracle:
    def __init__(self, bank, specs, settings, fiducial_cosmopars, spectrobiaspars, pshotpars, nonlinpars):
        s = dict(DEFAULT_SETTINGS)
        s.update(settings or {})
        self.bank, self.specs, self.s = bank, specs, s
        self.fid = (dict(fiducial_cosmopars), dict(spectrobiaspars), dict(pshotpars), dict(nonlinpars))
        self.fidcosmo = bank.cosmo(fiducial_cosmopars)
        fo = GalSpectro(self.fidcosmo, self.fidcosmo, specs, s, spectrobiaspars, nonlinpars)
        self.cov = SpectroCov(fo)
        self.zmids = self.cov.zmids
        h = fiducial_cosmopars["h"]
        self.kgrid = np.linspace(specs["kmin_GCsp"] * h, specs["kmax_GCsp"] * h, s["spectro_Pk_k_samples"])
        self.mugrid = np.linspace(0.0, 1.0, s["spectro_Pk_mu_samples"])
        self.kmesh, self.mumesh = np.meshgrid(self.kgrid, self.mugrid)
        self.volume = np.array([self.cov.volume_survey(i) for i in range(len(self.zmids))])
        self.data = self.observable(self.cov, None)

    def observable(self, cov, shot):
        """spectro_like.py:25-41."""
        shot = np.zeros_like(self.zmids) if shot is None else shot
        out = np.zeros((len(self.zmids), len(self.mugrid), len(self.kgrid)))
        for i, z in enumerate(self.zmids):
            out[i] = cov.noisy_Pgg(z, self.kmesh, self.mumesh) + shot[i]
        return out

    def theory(self, params):
        """spectro_like.py:192-233."""
        p = dict(params)
        cp0, bp0, ps0, nl0 = self.fid
        shot = np.zeros(len(self.zmids))
        for i, key in enumerate(ps0):
            shot[i] = p.pop(key, ps0[key])

        def upd(t):
            g = dict(t)
            for k in t:
                if k in p:
                    g[k] = p.pop(k)
            return g

        bp, nl, cp = upd(bp0), upd(nl0), upd(cp0)
        cosmo = self.fidcosmo if cp == cp0 else self.bank.cosmo(cp)
        ob = GalSpectro(cosmo, self.fidcosmo, self.specs, self.s, bp, nl)
        return self.observable(SpectroCov(ob), shot)

    def wedge_chi2(self, theory):
        """spectro_like.py:143-189."""
        delta = self.data - theory
        cov = (8 * np.pi**2 / (self.kgrid[None, None, :] ** 2 * self.volume[:, None, None])) * self.data**2
        mu_int = 2 * simpson(delta**2 / cov, x=self.kgrid, axis=2)
        return np.sum(simpson(mu_int, x=self.mugrid, axis=1))

    def legendre(self, obs):
        """spectro_like.py:48-56 -> (Nk, Nz, 3)."""
        lv = eval_legendre(LEG_ELLS[None, :], self.mugrid[:, None])[None, :, None, :]
        pl = (2.0 * LEG_ELLS[None, None, :] + 1) * simpson(lv * obs[:, :, :, None], x=self.mugrid, axis=1)
        return pl.transpose(1, 0, 2)

    def legendre_cov(self, P_ell):
        """spectro_like.py:69-122."""
        k = self.kgrid
        Pb = P_ell[:, :, :, None, None]
        nk, nz, nl = P_ell.shape
        P00 = Pb[:, :, 0] ** 2 * M["m00"][None, None]
        P22 = Pb[:, :, 1] ** 2 * M["m22"][None, None]
        P44 = Pb[:, :, 2] ** 2 * M["m44"][None, None]
        P02 = Pb[:, :, 0] * Pb[:, :, 1] * M["m02"][None, None]
        P24 = Pb[:, :, 1] * Pb[:, :, 2] * M["m24"][None, None]
        P04 = Pb[:, :, 0] * Pb[:, :, 2] * M["m04"][None, None]
        mbm = (2 * (2 * LEG_ELLS[None, None, :, None] + 1) * (2 * LEG_ELLS[None, None, None, :] + 1)
               / self.volume[None, :, None, None] * (P00 + P22 + P44 + 2 * P02 + 2 * P24 + 2 * P04))
        lnk = np.log(k)
        edges = np.zeros(nk + 1)
        edges[1:-1] = (lnk[1:] + lnk[:-1]) / 2.0
        edges[0] = edges[1] - (lnk[1] - lnk[0])
        edges[-1] = edges[-2] + (lnk[-1] - lnk[-2])
        kvol = 4 * np.pi / 3.0 * np.diff(np.exp(3 * edges))
        spl = make_interp_spline(lnk, mbm * (k[:, None, None, None]) ** 3, k=1, axis=0)
        cov = np.zeros((nk, nz, nl, nl))
        for i in range(nk):
            cov[i] = spl.integrate(edges[i], edges[i + 1])
        cov *= 2 * (2 * np.pi) ** 4 / (kvol[:, None, None, None] ** 2)
        return cov, np.linalg.inv(cov)

    @staticmethod
    def legendre_chi2(pd, pt, icov):
        return np.sum((pd[..., None] - pt[..., None]) * icov * (pd[..., None, :] - pt[..., None, :]))
