"""Oracle: cosmology table facade (restates cosmology.py:1188-1543 `external_input` and
cosmology.py:1546-2010 `cosmo_functions` for `code="external"`).  Test infrastructure only."""
from __future__ import annotations

import math

import numpy as np
from scipy import integrate
from scipy.interpolate import InterpolatedUnivariateSpline, RectBivariateSpline
from scipy.signal import savgol_filter

C_KMS = 299792.458

DEFAULT_SETTINGS = {
    # config.py:225-275 -- only the keys the hot path reads
    "nonlinear": True,
    "nonlinear_photo": True,
    "bfs8terms": False,
    "AP_effect": True,
    "FoG_switch": True,
    "GCsp_linear": False,
    "fix_cosmo_nl_terms": True,
    "pivot_z_IA": 0.0,
    "accuracy": 1,
    "spectro_Pk_k_samples": 1025,
    "spectro_Pk_mu_samples": 17,
    "savgol_polyorder": 3,
    "savgol_width": 1.358528901113328,
    "savgol_internalsamples": 800,
    "savgol_internalkmin": 0.001,
    "eps_cosmopars": 0.01,
    "eps_gal_nuispars": 0.0001,
    "eps_gal_nonlinpars": 0.01,
    "GCsp_Tracer": "matter",
    "GCph_Tracer": "matter",
    "ell_sampling": "accuracy",
    "kh_rescaling_bug": False,
    "kh_rescaling_beforespecerr_bug": False,
    "activateMG": False,
    "external_activateMG": False,
}


def _dcom_trapz(zi, hfunc):
    """cosmology.py:36-55: comoving distance by a 100-point trapezoid of 1/H."""
    zt = np.linspace(0.0, zi, 100)
    return integrate.trapezoid(1.0 / hfunc(zt), zt)


class Cosmo:
    """One cosmology's spline facade built from a table dict
    {z,k,H,s8,D,f,Pl,Pnl} (cosmology.py:1423-1540)."""

    def __init__(self, tables, cosmopars, settings=None, k_units="1/Mpc", r_units="Mpc"):
        self.settings = dict(DEFAULT_SETTINGS)
        if settings:
            self.settings.update(settings)
        self.cosmopars = dict(cosmopars)
        pkf = rf = kf = 1.0
        if k_units == "h/Mpc":  # cosmology.py:1438-1440
            pkf = (1.0 / cosmopars["h"]) ** 3
            kf = cosmopars["h"]
        if r_units == "Mpc/h":
            rf = 1.0 / cosmopars["h"]
        elif r_units == "km/s/Mpc":
            rf = C_KMS
        self.zgrid = np.asarray(tables["z"], float).flatten()
        self.kgrid = kf * np.asarray(tables["k"], float).flatten()
        self.h_of_z = InterpolatedUnivariateSpline(self.zgrid, (1.0 / rf) * np.asarray(tables["H"]).flatten())
        dcom = np.array([_dcom_trapz(zi, self.h_of_z) for zi in self.zgrid])
        self.com_dist = InterpolatedUnivariateSpline(self.zgrid, dcom)
        self.ang_dist = InterpolatedUnivariateSpline(self.zgrid, dcom / (1.0 + self.zgrid))
        self.D_growth_zk = RectBivariateSpline(self.zgrid, self.kgrid, tables["D"], kx=3, ky=3)
        self.f_growthrate_zk = RectBivariateSpline(self.zgrid, self.kgrid, tables["f"], kx=3, ky=3)
        self.s8_of_z = InterpolatedUnivariateSpline(self.zgrid, np.asarray(tables["s8"]).flatten())
        self.Pk_l = RectBivariateSpline(self.zgrid, self.kgrid, pkf * np.asarray(tables["Pl"]), kx=3, ky=3)
        self.Pk_nl = RectBivariateSpline(self.zgrid, self.kgrid, pkf * np.asarray(tables["Pnl"]), kx=3, ky=3)
        if "Plcb" in tables:  # cdm+baryon ("clustering" tracer) tables, cosmology.py:1508-1531
            self.f_growthrate_cb_zk = RectBivariateSpline(self.zgrid, self.kgrid, tables["fcb"], kx=3, ky=3)
            self.s8_cb_of_z = InterpolatedUnivariateSpline(self.zgrid, np.asarray(tables["s8cb"]).flatten())
            self.Pk_cb_l = RectBivariateSpline(self.zgrid, self.kgrid, pkf * np.asarray(tables["Plcb"]), kx=3, ky=3)
            self.Pk_cb_nl = RectBivariateSpline(self.zgrid, self.kgrid, pkf * np.asarray(tables["Pnlcb"]), kx=3, ky=3)

    # accessors, cosmology.py:1616-1977 -------------------------------------------------
    def Hubble(self, z):
        return self.h_of_z(z)

    def E_hubble(self, z):
        return self.Hubble(z) / self.Hubble(0.0)

    def angdist(self, z):
        return self.ang_dist(z)

    def comoving(self, z):
        return self.com_dist(z)

    def matpow(self, z, k, nonlinear=False, tracer="matter"):
        if tracer == "clustering":  # cosmology.py:1715-1716
            return (self.Pk_cb_nl if nonlinear else self.Pk_cb_l)(z, k, grid=False)
        return (self.Pk_nl if nonlinear else self.Pk_l)(z, k, grid=False)

    def f_growthrate(self, z, k=None, tracer="matter"):
        if tracer == "clustering":  # cosmology.py:1937-1939
            return self.f_growthrate_cb_zk(z, 0.0001 if k is None else k, grid=False)
        return self.f_growthrate_zk(z, 0.0001 if k is None else k, grid=False)

    def growth(self, z, k=None):
        return self.D_growth_zk(z, 0.0001 if k is None else k, grid=False)

    def sigma8_of_z(self, z, tracer="matter"):
        return self.s8_cb_of_z(z) if tracer == "clustering" else self.s8_of_z(z)  # cosmology.py:1851-1857

    def fsigma8_of_z(self, z, k=None, tracer="matter"):
        return self.f_growthrate(z, k, tracer=tracer) * self.sigma8_of_z(z, tracer=tracer)

    def Omegam_of_z(self, z):
        # cosmology.py:1904-1905 (external branch)
        return (self.cosmopars["Omegam"] * (1 + z) ** 3) / self.E_hubble(z) ** 2

    def nonwiggle_table(self, z, nonlinear=False, tracer="matter"):
        """(k_samples, P_nw samples) of the degree-1 spline of cosmology.py:1790-1812."""
        s = self.settings
        unitsf = self.cosmopars["h"]
        kmin_loc = unitsf * s["savgol_internalkmin"]
        kmax_loc = unitsf * np.max(self.kgrid)
        logk = np.linspace(np.log(kmin_loc), np.log(kmax_loc), s["savgol_internalsamples"])
        dlnk = np.mean(logk[1:] - logk[0:-1])
        n_savgol = int(np.round(s["savgol_width"] / np.log(1 + dlnk)))
        intp_p = InterpolatedUnivariateSpline(
            logk, np.log(self.matpow(z, np.exp(logk), nonlinear=nonlinear, tracer=tracer).flatten()), k=1
        )
        pow_sg = savgol_filter(intp_p(logk), n_savgol, s["savgol_polyorder"])
        return np.exp(logk), np.exp(pow_sg)

    def nonwiggle_pow(self, z, k, nonlinear=False, tracer="matter"):
        kk, pp = self.nonwiggle_table(z, nonlinear, tracer)
        return InterpolatedUnivariateSpline(kk, pp, k=1)(k)


# ---------------------------------------------------------------------------------------
def closest(lst, K):
    lst = np.asarray(lst)
    return lst[(np.abs(lst - K)).argmin()]


def round_decimals_up(number):
    """utilities/utils.py numerics.round_decimals_up (new branch)."""
    exponent = math.floor(math.log10(number))
    mantissa = number / (10**exponent)
    return math.ceil(mantissa * 10) / 10 * (10**exponent)


def folder_for(cosmopars, fiducial, param_names, eps_values, folder_names=None, e00=True):
    """Folder selection by snapping to the nearest tabulated epsilon (cosmology.py:1375-1421)."""
    folder_names = folder_names or param_names
    name = "fiducial_eps_0"
    for par, fpar in zip(param_names, folder_names):
        if not np.isclose(cosmopars[par], fiducial[par], rtol=1e-5):
            if fiducial[par] != 0:
                d = (cosmopars[par] / fiducial[par]) - 1.0
            else:
                d = cosmopars[par] - fiducial[par]
            sign = {-1.0: "mn", 1.0: "pl"}[float(np.sign(d))]
            ev = np.array(eps_values)
            d = closest(np.concatenate((ev, -ev, np.array([0]))), d)
            d = round_decimals_up(abs(d))
            tok = "{:.1E}".format(abs(d)).replace(".", "p")
            if not e00:
                tok = tok.replace("E-0", "E-")
            name = fpar + "_" + sign + "_eps_" + tok
            break
    return name


class Bank:
    """In-memory table bank {folder: tables} + the `extfiles` metadata; caches Cosmo objects
    per folder (the reference re-reads and re-fits per call; results are identical)."""

    def __init__(self, tables_by_folder, fiducial, param_names, eps_values=(0.01,), settings=None,
                 k_units="1/Mpc", r_units="Mpc"):
        self.tables = tables_by_folder
        self.fiducial = dict(fiducial)
        self.param_names = list(param_names)
        self.eps_values = list(eps_values)
        self.settings = settings
        self.k_units, self.r_units = k_units, r_units
        self._cache = {}

    def cosmo(self, cosmopars, cache=True):
        folder = folder_for(cosmopars, self.fiducial, self.param_names, self.eps_values)
        key = (folder, tuple(sorted((k, v) for k, v in cosmopars.items() if not isinstance(v, str))))
        if cache and key in self._cache:
            return self._cache[key]
        c = Cosmo(self.tables[folder], cosmopars, self.settings, self.k_units, self.r_units)
        if cache:
            self._cache[key] = c
        return c
