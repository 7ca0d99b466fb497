"""Oracle: spectroscopic galaxy clustering P_obs(k,mu,z), V_eff, derivatives and Fisher.

Restates (test infrastructure only; see oracle/__init__.py):
  spectro_obs.py:342-933   ComputeGalSpectro  (AP, Kaiser, FoG, dewiggling, spec-z error)
  spectro_cov.py:134-228   SpectroCov         (volumes, number density, V_eff)
  spectro_cov.py:456-561   SpectroDerivs      (get_obs, exact Ps derivative with the stale-pobs quirk)
  derivatives.py:113-207   derivative_3pt
  cosmicfish.py:366-542    set_pk_settings, pk_LSS_Fisher ... fish_integrand
  nuisance.py:232-366      z-bins, bias lookup, sigma_p/sigma_v rescale
  utilities/utils.py:131-155 bisection
"""
from __future__ import annotations

import copy

import numpy as np
from scipy import integrate
from scipy.integrate import simpson

from .cosmo import DEFAULT_SETTINGS

SR = np.power(180 / np.pi, 2)
AREASKY = 4 * np.pi * SR


def moving_average(x, periods=2):
    return np.convolve(x, np.ones(periods) / periods, mode="valid")


def bisection(array, value):
    """utilities/utils.py:131-155."""
    array = np.asarray(array)
    n = len(array)
    if np.isclose(value, array[0]) or value < array[0]:
        return 0
    if np.isclose(value, array[-1]) or value > array[-1]:
        return n - 2
    left, right = 0, n - 1
    while left < right:
        mid = (left + right) // 2
        if np.isclose(array[mid], value):
            return mid
        elif array[mid] < value:
            left = mid + 1
        else:
            right = mid
    return left - 1


class SpectroNuisance:
    """nuisance.py:232-366 (GCsp part)."""

    def __init__(self, specs, spectrobiaspars, nonlinpars):
        self.specs = specs
        zb, inds = [], []
        for key, val in specs["z_bins_sp"].items():
            zb.append(val)
            inds.append(key)
        self.zbins = np.unique(np.concatenate(zb))
        self.inds = inds
        self.zmids = moving_average(self.zbins)
        self.dndz = np.array([specs["dndOmegadz"][ii] for ii in inds])
        self.model = specs["sp_bias_model"]
        self.root = specs["sp_bias_root"]
        self.sample = specs["sp_bias_sample"]
        self.bias = spectrobiaspars
        self.nonlin = nonlinpars

    def bias_at_zm(self):
        b = np.ones(len(self.zmids))
        if "linear" in self.model:
            for ii, zi in enumerate(self.inds):
                b[ii] = self.bias[self.root + self.sample + "_" + str(zi)]
            if self.model == "linear_log":
                b = np.exp(b)
        return b

    def zindex(self, z):
        n = bisection(self.zbins, z) + 1
        return min(max(n, self.inds[0]), self.inds[-1])

    def bias_at_z(self, z):
        return self.bias_at_zm()[self.zindex(z) - 1]

    def kscale(self, k):
        if k is not None and self.model == "linear_Qbias":
            return (1 + k**2 * self.bias.get("A2", 0.0)) / (1 + k * self.bias.get("A1", 0.0))
        return 1

    def rescale(self, z, key):
        return self.nonlin.get(key + "_" + str(self.zindex(z)), 1.0)


class GalSpectro:
    """spectro_obs.py ComputeGalSpectro for the gg spectrum."""

    def __init__(self, cosmo, fidcosmo, specs, settings, spectrobiaspars, nonlinpars, use_bias_funcs=False):
        self.use_bias_funcs = use_bias_funcs  # spectro_obs.py:568-572
        self.cosmo, self.fid = cosmo, fidcosmo
        self.specs = specs
        s = dict(DEFAULT_SETTINGS)
        s.update(settings or {})
        self.s = s
        self.nuis = SpectroNuisance(specs, spectrobiaspars, nonlinpars)
        self.k_grid = np.logspace(np.log10(0.001), np.log10(5), 1024)  # :261-263
        self.linear = s["GCsp_linear"]
        self.tracer = s["GCsp_Tracer"]  # spectro_obs.py:164
        par = specs.get("nonlinear_parametrization", {"default": ""}) or {}
        self.rescale_sigmap = par.get("rescale_sigmap", False)
        self.rescale_sigmav = par.get("rescale_sigmav", False)
        self.dz_err = specs["spec_sigma_dz"]
        self.dz_type = specs["spec_sigma_dz_type"]

    def qpar(self, z):
        return self.fid.Hubble(z) / self.cosmo.Hubble(z)

    def qper(self, z):
        return self.cosmo.angdist(z) / self.fid.angdist(z)

    def kpar(self, z, k, mu):
        return k * mu * (1 / self.qpar(z))

    def kper(self, z, k, mu):
        return k * np.sqrt(1 - mu**2) * (1 / self.qper(z))

    def k_units_change(self, k):
        if self.s["kh_rescaling_bug"]:
            return k * (self.cosmo.cosmopars["h"] / self.fid.cosmopars["h"])
        return k

    def kmu_ap(self, z, k, mu):
        if not self.s["AP_effect"]:
            return k, mu
        kap = np.sqrt(self.kpar(z, k, mu) ** 2 + self.kper(z, k, mu) ** 2)
        return kap, self.kpar(z, k, mu) / kap

    def spec_err_z(self, z, k, mu):
        dz = self.dz_err if self.dz_type == "constant" else self.dz_err * (1 + z)
        err = dz * (1 / self.cosmo.Hubble(z)) * self.kpar(z, k, mu)
        return np.exp(-(1 / 2) * err**2)

    def bao(self, z):
        if not self.s["AP_effect"]:
            return 1
        return 1 / (self.qper(z) ** 2 * self.qpar(z))

    def kaiser(self, z, k, mu, b_i=None):
        """spectro_obs.py:585-634; b_i replaces the whole bias term, use_bias_funcs interpolates the per-bin biases
        linearly in z (nuisance.py:327-344)."""
        if b_i is not None:
            b = b_i
        elif self.use_bias_funcs:
            from scipy.interpolate import InterpolatedUnivariateSpline

            b = InterpolatedUnivariateSpline(self.nuis.zmids, self.nuis.bias_at_zm(), k=1)(z) * self.nuis.kscale(k)
        else:
            b = self.nuis.bias_at_z(z) * self.nuis.kscale(k)
        if self.s["bfs8terms"]:
            f = self.cosmo.fsigma8_of_z(z, k, tracer=self.tracer)  # spectro_obs.py:628-630
        else:
            f = self.cosmo.f_growthrate(z, k, tracer=self.tracer)
        return b + f * mu**2

    def moments(self, z, m):
        c = self.fid if self.s["fix_cosmo_nl_terms"] else self.cosmo
        ff = (c.f_growthrate(z, self.k_grid) ** m).flatten()
        pp = c.matpow(z, self.k_grid).flatten()
        return (1 / (6 * np.pi**2)) * integrate.trapezoid(pp * ff, x=self.k_grid)

    def sigmap(self, z):
        if self.linear:
            return 0
        sp = np.sqrt(self.moments(z, 2))
        if self.rescale_sigmap:
            sp *= self.nuis.rescale(z, "sigmap")
        return sp

    def sigmav(self, z, mu):
        if self.linear:
            return 0
        f0, f1, f2 = self.moments(z, 0), self.moments(z, 1), self.moments(z, 2)
        sv = np.sqrt(f0 + 2 * mu**2 * f1 + mu**2 * f2)
        if self.rescale_sigmav:
            sv *= self.nuis.rescale(z, "sigmav")
        return sv

    def fog(self, z, k, mu):
        if (self.s["FoG_switch"] is False) or self.linear:
            return 1
        return 1 / (1 + (k * mu * self.sigmap(z)) ** 2)

    def dewiggled(self, z, k, mu):
        g = 0 if self.linear else self.sigmav(z, mu) ** 2
        den = self.cosmo.sigma8_of_z(z, tracer=self.tracer) ** 2 if self.s["bfs8terms"] else 1  # spectro_obs.py:778-807
        pdd = self.cosmo.matpow(z, k, tracer=self.tracer) / den
        pnw = self.cosmo.nonwiggle_pow(z, k, tracer=self.tracer) / den
        return pdd * np.exp(-g * k**2) + pnw * (1 - np.exp(-g * k**2))

    def observed_Pgg(self, z, k, mu, b_i=None):
        """spectro_obs.py:845-911."""
        if self.s["kh_rescaling_beforespecerr_bug"]:
            k = self.k_units_change(k)
            err = self.spec_err_z(z, k, mu)
        else:
            err = self.spec_err_z(z, k, mu)
            k = self.k_units_change(k)
        k, mu = self.kmu_ap(z, k, mu)
        return self.bao(z) * (self.kaiser(z, k, mu, b_i) ** 2) * self.dewiggled(z, k, mu) * self.fog(z, k, mu) * (err**2) + 0

    def lnpobs_gg(self, z, k, mu):
        return np.log(self.observed_Pgg(z, k, mu))

    def observed_P_gg_ij(self, z, k, mu):
        """spectro_obs.py:985-1019 `observed_P_ij` for si = sj = "g" (the form the likelihood uses; same
        physics as observed_Pgg, different order of the floating-point products)."""
        err = self.spec_err_z(z, k, mu)
        k = self.k_units_change(k)
        k, mu = self.kmu_ap(z, k, mu)
        bao = self.bao(z)
        ka = self.kaiser(z, k, mu)
        fog = self.fog(z, k, mu)
        pdw = self.dewiggled(z, k, mu)
        fac = ka * 1 * err * 1 + np.sqrt(0.0)
        return bao * fog * pdw * fac * fac


class SpectroCov:
    """spectro_cov.py:134-228 on the fiducial observable."""

    def __init__(self, fid_obs: GalSpectro):
        self.obs = fid_obs
        sp = fid_obs.specs
        if "fsky_spectro" in sp:
            self.fsky = sp["fsky_spectro"]
        else:  # config.py:581-583
            self.fsky = sp["area_survey_spectro"] / AREASKY
        n = fid_obs.nuis
        self.dnz, self.zbins, self.zmids = n.dndz, n.zbins, n.zmids
        self.dz_bins = np.diff(self.zbins)

    def volume_bin(self, zi, zj):
        d1 = self.obs.fid.angdist(zi)
        d2 = self.obs.fid.angdist(zj)
        return (4 * np.pi / 3) * (pow((1 + zj) * d2, 3) - pow((1 + zi) * d1, 3))

    def d_volume(self, i):
        return self.volume_bin(self.zbins[i], self.zbins[i + 1])

    def volume_survey(self, i):
        return self.fsky * self.d_volume(i)

    def n_density(self, zi):
        w = np.where(np.isclose(self.zmids, zi, rtol=1e-2))[0]
        if len(w) == 0:
            return 0
        i = w[0]
        return AREASKY * (self.dnz[i] * self.dz_bins[i] / self.d_volume(i))

    def veff(self, zi, k, mu):
        npobs = self.n_density(zi) * self.obs.observed_Pgg(zi, k, mu)
        cov = (1 / (8 * (np.pi**2))) * (npobs / (1 + npobs)) ** 2
        if zi < self.zmids[0] or zi > self.zmids[-1]:
            cov = np.zeros_like(cov)
        return cov

    def noisy_Pgg(self, z, k, mu):
        """spectro_cov.py:361-370 for si=sj='g' (observed_P_ij == observed_Pgg for gg with zero extra shot)."""
        return self.obs.observed_P_gg_ij(z, k, mu) + 1 / self.n_density(z)


STENCILS = {  # (offset multiple of h, weight) and denominator factor; derivatives.py:93-111, 209-225
    "3PT": ([(+1, +1.0), (-1, -1.0)], 2.0),
    "4PT_FWD": ([(0, -11.0), (1, 18.0), (2, -9.0), (3, 2.0)], 6.0),
    "5PT": ([(+2, -1.0), (+1, +8.0), (-1, -8.0), (-2, +1.0)], 12.0),  # not in the reference (north_star)
}


def spectro_fisher(bank, specs, settings, fiducial_cosmopars, spectrobiaspars, pshotpars, nonlinpars,
                   freeparams, max_z_bins=None, return_all=False, stencil="3PT"):
    """FisherMatrix.compute() spectro branch (cosmicfish.py:253-330)."""
    s = dict(DEFAULT_SETTINGS)
    s.update(settings or {})
    fidcosmo = bank.cosmo(fiducial_cosmopars)
    fid_obs = GalSpectro(fidcosmo, fidcosmo, specs, s, spectrobiaspars, nonlinpars)
    cov = SpectroCov(fid_obs)
    zmids = cov.zmids
    freeparams = dict(freeparams)
    if max_z_bins is not None:
        zmids = zmids[:max_z_bins]
        # cosmicfish.py eliminate_zbinned_freepars: drop nuisance parameters of removed bins
        for key in list(freeparams):  # cosmicfish.py:813-839
            if "_" in key:
                try:
                    ibin = int(key.split("_")[1])
                except ValueError:
                    continue
                if ibin > max_z_bins:
                    freeparams.pop(key)
    h = fiducial_cosmopars["h"]
    kgrid = np.linspace(specs["kmin_GCsp"] * h, specs["kmax_GCsp"] * h, s["spectro_Pk_k_samples"])
    mugrid = np.linspace(0.0, 1.0, s["spectro_Pk_mu_samples"])
    kmesh, mumesh = np.meshgrid(kgrid, mugrid)
    veff = [cov.veff(z, kmesh, mumesh) for z in zmids]

    fid_all = {**fiducial_cosmopars, **spectrobiaspars, **pshotpars, **nonlinpars}
    state = {"pobs": None}

    def get_obs(allpars):  # spectro_cov.py:456-508
        cp = {k: allpars[k] for k in fiducial_cosmopars}
        bp = {k: allpars[k] for k in spectrobiaspars}
        nl = {k: allpars[k] for k in nonlinpars}
        cosmo = fidcosmo if cp == fiducial_cosmopars else bank.cosmo(cp)
        ob = GalSpectro(cosmo, fidcosmo, specs, s, bp, nl)
        state["pobs"] = ob
        return {i: ob.lnpobs_gg(z, kmesh, mumesh) for i, z in enumerate(zmids)}

    derivs = {}
    for par in freeparams:  # derivatives.py:131-207
        if "Ps" in par:  # spectro_cov.py:510-532 (uses the LAST evaluated stencil point)
            d = {}
            for i, z in enumerate(zmids):
                pgg = state["pobs"].observed_Pgg(z, kmesh, mumesh)
                ii = np.where(np.isclose(zmids, z))[0][0] + 1
                kron = 1 if ii == int(par.split("_")[-1]) else 0
                d[i] = kron * (1 / pgg)
            derivs[par] = d
            continue
        step = fid_all[par] * freeparams[par] if fid_all[par] != 0.0 else freeparams[par]
        pts, den = STENCILS[stencil]
        acc = None
        for m, w in pts:
            ap = copy.deepcopy(fid_all)
            ap[par] = ap[par] + m * step
            o = get_obs(ap)
            acc = {i: w * o[i] for i in o} if acc is None else {i: acc[i] + w * o[i] for i in o}
        derivs[par] = {i: acc[i] / (den * step) for i in acc}

    plist = list(freeparams)
    F = np.zeros((len(zmids), len(plist), len(plist)))
    for b in range(len(zmids)):
        vol = cov.volume_survey(b)
        for i, pi in enumerate(plist):
            for j in range(i + 1):
                intg = kmesh**2 * 2 * vol * veff[b] * derivs[pi][b] * derivs[plist[j]][b]
                val = simpson(simpson(intg, x=mumesh[:, 0], axis=0), x=kmesh[0, :])
                F[b, i, j] = val
                F[b, j, i] = val
    out = {"fisher": np.sum(F, axis=0), "fisher_per_bin": F, "param_names": plist}
    if return_all:
        out.update(derivs=derivs, veff=veff, kgrid=kgrid, mugrid=mugrid, zmids=zmids, fid_obs=fid_obs, cov=cov,
                   lnp_fid=[fid_obs.lnpobs_gg(z, kmesh, mumesh) for z in zmids])
    return out
