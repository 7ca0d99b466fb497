"""Oracle: photometric 3x2pt -- n_i(z) windows, Limber C_ell, covariance, Fisher.

Restates (test infrastructure only; see oracle/__init__.py):
  photo_window.py:60-242   GalaxyPhotoDist (dNdz, erf-convolved n_i, norm by 1000-pt trapezoid)
  nuisance.py:89-230       IA eNLA window, galaxy bias models, luminosity ratio
  photo_obs.py:113-160     much_faster_integral_efficiency
  photo_obs.py:164-324     ComputeCls.__init__ (ell / z grids)
  photo_obs.py:457-512     sqrtP_limber
  photo_obs.py:514-722     kernels / genwindow
  photo_obs.py:798-928     computecls_vectorized / clsintegral
  photo_cov.py:116-245     getcls / getclsnoise / get_covmat
  cosmicfish.py:643-744    photo_LSS_fishermatrix_einsum
"""
from __future__ import annotations

import copy
from itertools import product

import numpy as np
from scipy import integrate
from scipy.interpolate import InterpolatedUnivariateSpline, interp1d
from scipy.special import erf

from .cosmo import DEFAULT_SETTINGS

SR = np.power(180 / np.pi, 2)


def finish_photo_specs(specs):
    """Derived spec entries of config.py:376-445 (ngalbin_*, binrange_*, z0)."""
    sp = dict(specs)

    def ngal_per_bin(ngal_sqarmin, zbins):
        nb = len(zbins[:-1])
        return ((ngal_sqarmin * 3600) / nb) * SR * np.ones_like(np.asarray(zbins[:-1], float))

    sp["z_bins_WL"] = np.array(sp.get("z_bins_WL", sp.get("z_bins_ph")))
    sp["ngal_sqarmin_WL"] = sp.get("ngal_sqarmin_WL", sp.get("ngal_sqarmin"))
    sp["ngalbin_WL"] = ngal_per_bin(sp["ngal_sqarmin_WL"], sp["z_bins_WL"])
    sp["binrange_WL"] = range(1, len(sp["z_bins_WL"]))
    sp["z_bins_GCph"] = sp.get("z_bins_GCph", sp.get("z_bins_ph"))
    sp["ngal_sqarmin_GCph"] = sp.get("ngal_sqarmin_GCph", sp.get("ngal_sqarmin"))
    sp["ngalbin_GCph"] = ngal_per_bin(sp["ngal_sqarmin_GCph"], sp["z_bins_GCph"])
    sp["binrange_GCph"] = range(1, len(sp["z_bins_GCph"]))
    sp["z0"] = sp["zm"] / np.sqrt(2)
    sp["z0_p"] = sp["z0"]
    return sp


def default_photo_bias(specs):
    """config.py:673-692 for the binned models."""
    model = specs["ph_bias_model"]
    prtz = specs["ph_bias_parametrization"]
    out = {"bias_model": model}
    if model in ("binned", "binned_constant"):
        if prtz[model]["generate_bias_keys"]:
            zb = specs.get("z_bins_GCph", specs.get("z_bins_ph"))
            for ind in range(1, len(zb)):
                out[prtz[model]["keystr"] + str(ind)] = np.sqrt(1 + 0.5 * (zb[ind] + zb[ind - 1]))
        else:
            for k in prtz[model]:
                if k not in ("generate_bias_keys", "keystr"):
                    out[k] = prtz[model][k]
    else:
        out.update(prtz[model])
    return out


_PHOTODIST_CACHE = {}


def photo_dist(specs, photopars):
    """GalaxyPhotoDist depends only on (bin edges, n(z) shape, photo-z parameters): build once."""
    key = (tuple(specs["z_bins_WL"]), tuple(specs["z_bins_GCph"]), specs["z0"], specs["ngamma"],
           tuple(sorted(photopars.items())))
    if key not in _PHOTODIST_CACHE:
        _PHOTODIST_CACHE[key] = PhotoDist(specs, photopars)
    return _PHOTODIST_CACHE[key]


class PhotoDist:
    """photo_window.py GalaxyPhotoDist."""

    def __init__(self, specs, photopars):
        self._cache = {}
        self.zb = {"WL": specs["z_bins_WL"], "GCph": specs["z_bins_GCph"]}
        self.z0, self.z0_p, self.ngamma = specs["z0"], specs["z0_p"], specs["ngamma"]
        self.photo = photopars
        self.z_max = np.max([specs["z_bins_GCph"][-1], specs["z_bins_WL"][-1]])
        self.normalization = {"GCph": self.norm("GCph"), "WL": self.norm("WL")}

    def dNdz(self, z):
        return (z / self.z0_p) ** 2 * np.exp(-((z / self.z0) ** self.ngamma))

    def ngal_photoz(self, z, i, obs):
        zb, p = self.zb[obs], self.photo
        t1 = p["cb"] * p["fout"] * erf((np.sqrt(1 / 2) * (z - p["zo"] - p["co"] * zb[i - 1])) / (p["sigma_o"] * (1 + z)))
        t2 = -p["cb"] * p["fout"] * erf((np.sqrt(1 / 2) * (z - p["zo"] - p["co"] * zb[i])) / (p["sigma_o"] * (1 + z)))
        t3 = p["co"] * (1 - p["fout"]) * erf((np.sqrt(1 / 2) * (z - p["zb"] - p["cb"] * zb[i - 1])) / (p["sigma_b"] * (1 + z)))
        t4 = -p["co"] * (1 - p["fout"]) * erf((np.sqrt(1 / 2) * (z - p["zb"] - p["cb"] * zb[i])) / (p["sigma_b"] * (1 + z)))
        return self.dNdz(z) * (t1 + t2 + t3 + t4) / (2 * p["co"] * p["cb"])

    def norm(self, obs):
        zint = np.linspace(0.0, self.z_max, 1000)
        dz = self.z_max / 1000  # sic (photo_window.py:210)
        # scalar calls on purpose: numpy's scalar pow/exp and its SIMD array loops differ in the
        # last ulp, and the reference evaluates point by point (photo_window.py:212-215, :242)
        out = [
            integrate.trapezoid([self.ngal_photoz(z, i, obs) for z in zint], dx=dz)
            for i in range(1, len(self.zb[obs]))
        ]
        out.insert(0, None)
        return out

    def norm_ngal_photoz(self, z, i, obs):
        key = (i, obs, z.tobytes())
        if key not in self._cache:  # pure function of (photopars, z-grid): memoised, values unchanged
            self._cache[key] = np.array([self.ngal_photoz(zi, i, obs) for zi in z]) / self.normalization[obs][i]
        return self._cache[key]


class PhotoNuisance:
    """nuisance.py:89-230 (photometric part)."""

    def __init__(self, specs, settings, lumratio_table, observables):
        self.specs, self.settings = specs, settings
        self.z = np.linspace(
            min(specs["z_bins_WL"][0], specs["z_bins_GCph"][0]),
            max(specs["z_bins_WL"][-1], specs["z_bins_GCph"][-1]) + 1,
            50 * settings["accuracy"],
        )
        if "WL" in observables:
            self.lumratio = interp1d(lumratio_table[:, 0], lumratio_table[:, 1], kind="linear")

    def gcph_bias(self, biaspars, ibin=1):
        z = self.z
        model = biaspars["bias_model"]
        if model == "sqrt":
            return interp1d(z, biaspars["b0"] * np.sqrt(1 + z), kind="linear")
        if model == "binned":
            edges = np.asarray(self.specs["z_bins_GCph"])
            brang = self.specs["binrange_GCph"]
            last = brang[-1]
            bb = np.array([biaspars[f"b{k}"] for k in brang], dtype=float)
            return lambda za: bb[np.clip(np.searchsorted(edges, np.asarray(za), side="right") - 1, 0, last - 1)]
        if model == "binned_constant":
            return np.vectorize(lambda zz: biaspars["b" + str(ibin)])
        if model == "flagship":
            b = biaspars["A"] + biaspars["B"] / (1 + np.exp(-biaspars["C"] * (z - biaspars["D"])))
            return interp1d(z, b, kind="linear")
        raise ValueError(model)

    def IA(self, IApars, cosmo):
        Om = cosmo.Omegam_of_z(0.0)
        piv = self.settings["pivot_z_IA"]
        z = self.z
        CIA = 0.0134 * (1 + piv)
        fac = -IApars["AIA"] * CIA * Om
        z_dep = (1 + z) / (1 + piv)
        win = fac * z_dep ** IApars["etaIA"] * ((self.lumratio(z) ** IApars["betaIA"]) / (cosmo.growth(z).flatten()))
        return InterpolatedUnivariateSpline(z, win, k=1)


def _efficiency(i, ngal_func, chi, z):
    """photo_obs.py:113-160."""
    dz = np.mean(np.diff(z))
    ngal = ngal_func(z, i, "WL")
    noc = ngal / chi

    def back(y, dx):
        yr = y[::-1]
        seg = 0.5 * (yr[:-1] + yr[1:]) * dx
        return np.concatenate([[0.0], np.cumsum(seg)])[::-1]

    eff = back(ngal, dz) - chi * back(noc, dz)
    return interp1d(z, eff, kind="cubic")


class Cls:
    """photo_obs.py ComputeCls (FAST paths)."""

    def __init__(self, cosmo, specs, settings, observables, photopars, IApars, biaspars, lumratio_table):
        s = dict(DEFAULT_SETTINGS)
        s.update(settings or {})
        self.s, self.specs, self.cosmo = s, specs, cosmo
        self.observables = [o for o in observables if o in ("GCph", "WL")]
        self.binrange = {"GCph": specs["binrange_GCph"], "WL": specs["binrange_WL"]}
        self.biaspars, self.IApars = biaspars, IApars
        self.nuis = PhotoNuisance(specs, s, lumratio_table, self.observables)
        if "WL" in self.observables:
            self.IAvalue = self.nuis.IA(IApars, cosmo)
        self.window = photo_dist(specs, photopars)
        if "GCph" in self.observables and "WL" in self.observables:
            self.ellmax = max(specs["lmax_GCph"], specs["lmax_WL"])
            self.ellmin = min(specs["lmin_GCph"], specs["lmin_WL"])
        elif "GCph" in self.observables:
            self.ellmax, self.ellmin = specs["lmax_GCph"], specs["lmin_GCph"]
        else:
            self.ellmax, self.ellmin = specs["lmax_WL"], specs["lmin_WL"]
        self.zsamp = int(round(200 * s["accuracy"]))
        self.ellsamp = int(round(100 * s["accuracy"])) if s["ell_sampling"] == "accuracy" else s["ell_sampling"]
        self.ell = np.logspace(np.log10(self.ellmin), np.log10(self.ellmax + 10), num=self.ellsamp)
        self.z_min = np.min([specs["z_bins_GCph"][0], specs["z_bins_WL"][0]])
        self.z_max = np.max([specs["z_bins_GCph"][-1], specs["z_bins_WL"][-1]])
        self.z = np.linspace(self.z_min, self.z_max, self.zsamp)
        self.dz = np.mean(np.diff(self.z))
        self.chi = cosmo.comoving(self.z)

    def sqrtP_limber(self):
        shp = (self.zsamp, self.ellsamp)
        self.sqrtPell = {k: np.zeros(shp) for k in ("WL", "WL_IA", "GCph", "GCph_IA")}
        kappa = (self.ell[:, None] + 1 / 2) / self.chi  # (ell, z)
        idx = np.where(kappa < self.specs["kmax"])
        el, zi = idx
        p = np.sqrt(self.cosmo.matpow(self.z[zi], kappa[el, zi], nonlinear=self.s["nonlinear"])) / self.chi[zi]
        for k in ("WL", "WL_IA", "GCph"):  # Sigma_MG == 1 on this path
            self.sqrtPell[k][zi, el] = p
        if self.s["GCph_Tracer"] == "clustering":  # photo_obs.py:499-509: GCph on the cdm+baryon spectrum
            self.sqrtPell["GCph"][zi, el] = np.sqrt(self.cosmo.matpow(
                self.z[zi], kappa[el, zi], nonlinear=self.s["nonlinear"], tracer="clustering")) / self.chi[zi]

    def compute_kernels(self):
        z = self.z
        hub = self.cosmo.Hubble(z)
        if "GCph" in self.observables:
            self.GC = [None] + [
                interp1d(z, self.window.norm_ngal_photoz(z, i, "GCph") * hub, kind="cubic") for i in self.binrange["GCph"]
            ]
        if "WL" in self.observables:
            eff = [None] + [_efficiency(i, self.window.norm_ngal_photoz, self.chi, z) for i in self.binrange["WL"]]
            pref = (3.0 / 2.0) * self.cosmo.Hubble(0.0) ** 2.0 * self.cosmo.Omegam_of_z(0.0)
            self.WL, self.WL_IA = [None], [None]
            for i in self.binrange["WL"]:
                self.WL.append(interp1d(z, pref * (1.0 + z) * self.chi * eff[i](z), kind="cubic"))
                self.WL_IA.append(interp1d(z, self.window.norm_ngal_photoz(z, i, "WL") * self.IAvalue(z) * hub, kind="cubic"))

    def genwindow(self, obs, i):
        z = self.z
        if obs == "GCph":
            bz = self.nuis.gcph_bias(self.biaspars, i)(z)
            return self.GC[i](z) * bz, self.GC[i](z) * 0.0
        return self.WL[i](z), self.WL_IA[i](z)

    def compute_all(self):
        self.sqrtP_limber()
        self.compute_kernels()
        sp = self.specs
        hub = self.cosmo.Hubble(self.z)
        cls = {"ells": self.ell}
        for o1, o2 in product(self.observables, self.observables):
            m1 = (self.ell >= sp["lmin_" + o1]) & (self.ell <= sp["lmax_" + o1])
            m2 = (self.ell >= sp["lmin_" + o2]) & (self.ell <= sp["lmax_" + o2])
            for b1, b2 in product(self.binrange[o1], self.binrange[o2]):
                w1, w1ia = self.genwindow(o1, b1)
                w2, w2ia = self.genwindow(o2, b2)
                intgn = (
                    self.sqrtPell[o1] * w1[:, None] / np.sqrt(hub)[:, None]
                    + self.sqrtPell[o1 + "_IA"] * w1ia[:, None] / np.sqrt(hub)[:, None]
                ) * (
                    self.sqrtPell[o2] * w2[:, None] / np.sqrt(hub)[:, None]
                    + self.sqrtPell[o2 + "_IA"] * w2ia[:, None] / np.sqrt(hub)[:, None]
                )
                clint = integrate.trapezoid(intgn, dx=self.dz, axis=0)
                clint[~m1] = 0
                clint[~m2] = 0
                clinterp = interp1d(self.ell, clint, kind="linear")
                mask = m1 & m2
                fin = np.zeros(len(self.ell))
                fin[mask] = clinterp(self.ell[mask])
                cls[f"{o1} {b1}x{o2} {b2}"] = fin
        self.result = cls
        return cls


class PhotoCov:
    """photo_cov.py getcls / noise / covariance (dense arrays instead of DataFrames)."""

    def __init__(self, bank, specs, settings, observables, fiducial_cosmopars, photopars, IApars, biaspars, lumratio_table):
        self.bank, self.specs, self.settings = bank, specs, settings
        self.observables = [o for o in observables if o in ("GCph", "WL")]
        self.cosmopars, self.photopars, self.IApars, self.biaspars = fiducial_cosmopars, photopars, IApars, biaspars
        self.lum = lumratio_table
        self.allparsfid = {**fiducial_cosmopars, **IApars, **biaspars, **photopars}
        self.binrange = {"GCph": specs["binrange_GCph"], "WL": specs["binrange_WL"]}
        self.fsky = {"WL": specs.get("fsky_WL"), "GCph": specs.get("fsky_GCph")}
        self.noise = {
            "WL": (specs["ellipt_error"] ** 2.0) / np.array(specs["ngalbin_WL"]),
            "GCph": 1.0 / np.array(specs["ngalbin_GCph"]),
        }
        self.cols = [f"{o} {i}" for o in self.observables for i in self.binrange[o]]

    def getcls(self, allpars):
        cp = {k: allpars[k] for k in self.cosmopars}
        pp = {k: allpars[k] for k in self.photopars}
        ia = {k: allpars[k] for k in self.IApars}
        bp = {k: allpars[k] for k in self.biaspars}
        c = Cls(self.bank.cosmo(cp), self.specs, self.settings, self.observables, pp, ia, bp, self.lum)
        self.last_cls = c
        return c.compute_all()

    def noisy(self, cls):
        out = dict(cls)
        for o in self.observables:
            for i in self.binrange[o]:
                k = f"{o} {i}x{o} {i}"
                out[k] = out[k] + self.noise[o][i - 1]
        return out

    def covmat(self, noisy):
        n = len(self.cols)
        cov = np.zeros((len(noisy["ells"]), n, n))
        for a, ca in enumerate(self.cols):
            f = np.sqrt(self.fsky[ca.split(" ")[0]])
            for b, cb in enumerate(self.cols):
                cov[:, a, b] = noisy[ca + "x" + cb] / f
        return cov


def _stem_derivative(pc, par, eps_values, threshold=1.0e-3, numstem=11):
    """derivatives.derivative_stem with external tables (derivatives.py:348-441): straight-line fit through the
    observable at fid*(1 +- eps) for every tabulated eps; the outermost points are dropped while the fit residual
    exceeds an ABSOLUTE 1e-3 (never, for angular spectra of order 1e-5 and below).  Restated as written, including
    the trimmed fit pairing the shortened abscissae with the FIRST ordinates (:410-413)."""
    fid = pc.allparsfid
    eps_v = np.array(eps_values)
    d_eps = np.concatenate([-eps_v[::-1], eps_v])
    stepsize = fid[par] * d_eps if fid[par] != 0.0 else d_eps
    obs_mod = []
    for step in stepsize:
        ap = copy.deepcopy(fid)
        ap[par] = ap[par] + step
        obs_mod.append(pc.getcls(ap))
    dpar = {"ells": obs_mod[0]["ells"]}
    for key in obs_mod[0]:
        if key == "ells":
            continue
        temp = []
        for ind in range(len(dpar["ells"])):
            residuals, counter, tempstep = 1000, 0, stepsize
            while residuals > threshold:
                fit = np.polyfit(tempstep, [obs_mod[i][key][ind] for i in range(len(tempstep))], 1, full=True)
                residuals = fit[1]
                if residuals > threshold:
                    tempstep = tempstep[1:-1]
                counter += 1
                if numstem - counter < 3:
                    raise RuntimeError(f"{par} derivative could not converge")
            temp.append(fit[0][0])
        dpar[key] = np.array(temp)
    return dpar


def photo_fisher(bank, specs, settings, observables, fiducial_cosmopars, photopars, IApars, biaspars,
                 freeparams, lumratio_table, return_all=False, stencil="3PT"):
    """FisherMatrix.compute() photo branch (cosmicfish.py:225-251, 643-744)."""
    s = dict(DEFAULT_SETTINGS)
    s.update(settings or {})
    pc = PhotoCov(bank, specs, s, observables, fiducial_cosmopars, photopars, IApars, biaspars, lumratio_table)
    cls = pc.getcls(pc.allparsfid)
    noisy = pc.noisy(cls)
    cov = pc.covmat(noisy)
    derivs = {}
    for par in freeparams:  # derivatives.py:131-207
        fid = pc.allparsfid
        step = fid[par] * freeparams[par] if fid[par] != 0.0 else freeparams[par]
        if stencil == "STEM":
            derivs[par] = _stem_derivative(pc, par, bank.eps_values)
            continue
        from .spectro import STENCILS

        pts, den = STENCILS[stencil]
        acc = None
        for m, w in pts:
            ap = copy.deepcopy(fid)
            ap[par] = ap[par] + m * step
            o = pc.getcls(ap)
            if acc is None:
                acc = {k: (o[k] if k == "ells" else w * o[k]) for k in o}
            else:
                acc = {k: (acc[k] if k == "ells" else acc[k] + w * o[k]) for k in o}
        derivs[par] = {k: (acc[k] if k == "ells" else acc[k] / (den * step)) for k in acc}
    lvec = noisy["ells"]
    lave = np.convolve(lvec, np.ones(2) / 2, mode="valid")
    dl = np.diff(lvec)
    cols = pc.cols
    inv = np.linalg.pinv(cov)
    der = np.array([[derivs[p][f"{a}x{b}"] for a in cols for b in cols] for p in freeparams])
    der = der.reshape(len(freeparams), len(cols), len(cols), -1)
    factor = (lave + 0.5) * dl
    FV = np.zeros((len(lave), len(freeparams), len(freeparams)))
    for li in range(len(lave)):
        d = der[:, :, :, li]
        m1 = np.einsum("pij,jk->pik", d, inv[li])
        m2 = np.einsum("ij,pjk->pik", inv[li], m1)
        FV[li] = np.einsum("pij,qji->pq", d, m2) * factor[li]
    out = {"fisher": np.sum(FV, axis=0), "param_names": list(freeparams)}
    if return_all:
        out.update(cls=cls, noisy=noisy, cov=cov, derivs=derivs, cols=cols, photocov=pc)
    return out
