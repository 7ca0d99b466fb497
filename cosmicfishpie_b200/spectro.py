"""Spectroscopic galaxy clustering (GCsp): host-side mirror of the reference's operator interface.

Same class names, constructor arguments and public methods as the reference
(cosmicfishpie/LSSsurvey/spectro_obs.py `ComputeGalSpectro`, spectro_cov.py `SpectroCov`,
`SpectroDerivs`), but every lattice evaluation, the derivative stencil and the Fisher reduction run in
the CUDA kernels of csrc/cfp_spectro.cu through the C ABI (include/cfp_b200.h).  The host only
prepares O(100) scalars per (stencil point, z-bin) from 1-D splines plus the no-wiggle tables.

There is no CPU implementation of the lattice math in this package.
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional

import numpy as np
from scipy.integrate import trapezoid

from . import _lib, stencils
from .cosmology import CosmoSet, cosmo_functions
from .quadrature import simpson_weights

SR = np.power(180 / np.pi, 2)
AREASKY = 4 * np.pi * SR


def moving_average(data, periods=2):
    return np.convolve(data, np.ones(periods) / periods, mode="valid")


def _isclose(a, b):
    """np.isclose(a, b) for finite scalars with numpy's default tolerances (asymmetric in b)."""
    return abs(a - b) <= 1e-8 + 1e-5 * abs(b)


def bisection(array, value):
    """Bin lookup with np.isclose edge rules (reference: utilities/utils.py:131-155)."""
    array = [float(a) for a in array]
    value = float(value)
    n = len(array)
    if _isclose(value, array[0]) or value < array[0]:
        return 0
    if _isclose(value, array[-1]) or value > array[-1]:
        return n - 2
    lo, hi = 0, n - 1
    while lo < hi:
        mid = (lo + hi) // 2
        if _isclose(array[mid], value):
            return mid
        if array[mid] < value:
            lo = mid + 1
        else:
            hi = mid
    return lo - 1


class SpectroNuisance:
    """z-bins, dN/dOmega/dz, bias and sigma_p/sigma_v rescale lookups (nuisance.py:232-366)."""

    def __init__(self, specs, spectrobiasparams, spectrononlinearpars):
        self.specs = specs
        zb, inds = [], []
        for key, val in specs["z_bins_sp"].items():
            zb.append(val)
            inds.append(key)
        self.sp_zbins = np.unique(np.concatenate(zb))
        self.sp_zbins_inds = inds
        self.sp_zbins_mids = moving_average(self.sp_zbins)
        self.sp_dndz = np.array([specs["dndOmegadz"][i] for i in inds])
        self.sp_bias_model = specs["sp_bias_model"]
        self.sp_bias_root = specs["sp_bias_root"]
        self.sp_bias_sample = specs["sp_bias_sample"]
        self.Spectrobiasparams = spectrobiasparams
        self.spectrononlinearpars = spectrononlinearpars

    def gcsp_zbins(self):
        return self.sp_zbins

    def gcsp_zbins_mids(self):
        return self.sp_zbins_mids

    def gcsp_dndz(self):
        return self.sp_dndz

    def gcsp_bias_at_zm(self):
        b = np.ones(len(self.sp_zbins_mids))
        if "linear" in self.sp_bias_model:
            for i, zi in enumerate(self.sp_zbins_inds):
                b[i] = self.Spectrobiasparams[self.sp_bias_root + self.sp_bias_sample + "_" + str(zi)]
            if self.sp_bias_model == "linear_log":
                b = np.exp(b)
        return b

    def gcsp_zvalue_to_zindex(self, z):
        n = bisection(self.sp_zbins, z) + 1
        return min(max(n, self.sp_zbins_inds[0]), self.sp_zbins_inds[-1])

    def gcsp_bias_at_z(self, z):
        return self.gcsp_bias_at_zm()[self.gcsp_zvalue_to_zindex(z) - 1]

    def gcsp_bias_interp(self):
        """Degree-1 spline through the per-bin biases at the bin centres (nuisance.py:327-344)."""
        from scipy.interpolate import InterpolatedUnivariateSpline

        return InterpolatedUnivariateSpline(self.gcsp_zbins_mids(), self.gcsp_bias_at_zm(), k=1)

    def vectorized_gcsp_bias_at_z(self, z):
        return np.vectorize(self.gcsp_bias_at_z)(z)

    def gcsp_bias_at_zi(self, zi):
        """Bias of the bin with (1-based) index zi (nuisance.py:268-283)."""
        return self.gcsp_bias_at_zm()[zi - 1]

    def gcsp_bias_kscale(self, k, z=None):
        """Scale dependence of the bias: (1 + k^2 A2) / (1 + k A1) for the Q-bias model, else 1 (nuisance.py:306-325)."""
        if k is not None and self.sp_bias_model == "linear_Qbias":
            a1, a2 = self.qbias_coeffs()
            return (1 + k**2 * a2) / (1 + k * a1)
        return 1

    def vectorized_gcsp_rescale_sigmapv_at_z(self, z, sigma_key="sigmap"):
        return np.vectorize(lambda zz: self.gcsp_rescale_sigmapv_at_z(zz, sigma_key=sigma_key))(z)

    def qbias_coeffs(self):
        if self.sp_bias_model == "linear_Qbias":
            return self.Spectrobiasparams.get("A1", 0.0), self.Spectrobiasparams.get("A2", 0.0)
        return 0.0, 0.0

    def gcsp_rescale_sigmapv_at_z(self, z, sigma_key="sigmap"):
        return self.spectrononlinearpars.get(sigma_key + "_" + str(self.gcsp_zvalue_to_zindex(z)), 1.0)

    def extra_Pshot_noise(self):
        return 0  # nuisance.py:364-366


_K_MOMENTS = np.logspace(np.log10(0.001), np.log10(5), 1024)  # spectro_obs.py:261-263


def _theta_moments(cosmo: cosmo_functions, z: float):
    """(I0, I1, I2): (1/6pi^2) int P_lin f^m dk on the fixed 1024-point grid (spectro_obs.py:724-755)."""
    cache = cosmo.__dict__.setdefault("_moment_cache", {})
    key = float(z)
    if key not in cache:
        pp = cosmo.matpow(z, _K_MOMENTS).flatten()
        f = cosmo.f_growthrate(z, _K_MOMENTS)
        cache[key] = tuple(
            (1 / (6 * np.pi**2)) * trapezoid(pp * (f**m).flatten(), x=_K_MOMENTS) for m in (0, 1, 2)
        )
    return cache[key]


class ComputeGalSpectro:
    """Observed galaxy power spectrum of one parameter point (reference: spectro_obs.py:20-933).

    `observed_Pgg` / `lnpobs_gg` evaluate on the GPU.  Intensity-mapping members are not provided
    (IM is out of scope)."""

    def __init__(self, cosmopars, fiducial_cosmopars=None, spectrobiaspars=None, IMbiaspars=None,
                 spectrononlinearpars=None, PShotpars=None, fiducial_cosmo=None, bias_samples=None,
                 use_bias_funcs=False, configuration=None, cosmo_set: Optional[CosmoSet] = None):
        if configuration is None:
            from . import config as config_module

            configuration = config_module.current()
        self.config = configuration
        cfg = getattr(configuration, "config", configuration)  # a FisherMatrix duck-types as a config
        self._cfg = cfg
        self.feed_lvl = configuration.settings["feedback"]
        # private top-level copies: only top-level keys ("kmax", "kmin") are rebound below, nested
        # specification entries are read-only here, so one level is enough
        self.observables = list(configuration.obs)
        self.specs = dict(configuration.specs)
        self.settings = configuration.settings
        self.s8terms = self.settings["bfs8terms"]
        self.tracer = self.settings["GCsp_Tracer"]  # "matter" | "clustering" (cdm+baryon tables), spectro_obs.py:164
        if self.tracer not in ("matter", "clustering"):
            raise ValueError("GCsp_Tracer must be 'matter' or 'clustering'")
        self.fiducial_cosmopars = dict(cfg.fiducialparams if fiducial_cosmopars is None else fiducial_cosmopars)
        if self.fiducial_cosmopars != cfg.fiducialparams:
            raise AttributeError("fiducial_cosmopars not equal to the configuration's fiducialparams")
        self.fiducialcosmo = cfg.fiducialcosmo
        self.cosmopars = cosmopars
        self.cosmo_set = cosmo_set if cosmo_set is not None else _cosmo_set_of(cfg)
        self.cosmo_index = self.cosmo_set.add(cosmopars)
        self.cosmo = self.cosmo_set.cosmos[self.cosmo_index]
        self.fiducial_spectrobiaspars = dict(cfg.Spectrobiasparams)
        self.spectrobiaspars = self.fiducial_spectrobiaspars if spectrobiaspars is None else spectrobiaspars
        self.fiducial_spectrononlinearpars = dict(cfg.Spectrononlinearparams)
        self.spectrononlinearpars = (
            self.fiducial_spectrononlinearpars if spectrononlinearpars is None else spectrononlinearpars
        )
        self.fiducial_IMbiaspars = dict(cfg.IMbiasparams)
        self.IMbiaspars = self.fiducial_IMbiaspars if IMbiaspars is None else IMbiaspars
        self.nuisance = SpectroNuisance(self.specs, self.spectrobiaspars, self.spectrononlinearpars)
        self.extraPshot = self.nuisance.extra_Pshot_noise()
        self.gcsp_z_bin_mids = self.nuisance.gcsp_zbins_mids()
        self.fiducial_PShotpars = dict(cfg.PShotparams)
        self.PShotpars = self.fiducial_PShotpars if PShotpars is None else PShotpars
        self.allpars = {**self.cosmopars, **self.spectrobiaspars, **self.IMbiaspars, **self.PShotpars,
                        **self.spectrononlinearpars}
        self.fiducial_allpars = {**self.fiducial_cosmopars, **self.fiducial_spectrobiaspars,
                                 **self.fiducial_IMbiaspars, **self.fiducial_PShotpars,
                                 **self.fiducial_spectrononlinearpars}
        self.obs_spectrum = ["g", "g"]
        self.specs["kmax"] = self.specs["kmax_GCsp"]
        self.specs["kmin"] = self.specs["kmin_GCsp"]
        self.k_grid = _K_MOMENTS
        self.linear_switch = self.settings["GCsp_linear"]
        self.FoG_switch = self.settings["FoG_switch"]
        self.APbool = self.settings["AP_effect"]
        self.fix_cosmo_nl_terms = self.settings["fix_cosmo_nl_terms"]
        prtz = self.specs.get("nonlinear_parametrization", {"default": ""}) or {}
        self.rescale_sigmap = prtz.get("rescale_sigmap", False)
        self.rescale_sigmav = prtz.get("rescale_sigmav", False)
        self.dz_err = self.specs["spec_sigma_dz"]
        self.dz_type = self.specs["spec_sigma_dz_type"]
        self.kh_rescaling_bug = self.settings["kh_rescaling_bug"]
        self.kh_rescaling_beforespecerr_bug = self.settings["kh_rescaling_beforespecerr_bug"]
        self.bias_samples = bias_samples
        self.use_bias_funcs = use_bias_funcs  # per-bin biases interpolated linearly in z (spectro_obs.py:568-572)

    # --- cheap host scalars (same formulas as the reference) ----------------------------------
    def qparallel(self, z):
        return self.fiducialcosmo.scalar("Hubble", z) / self.cosmo.scalar("Hubble", z)

    def qperpendicular(self, z):
        return self.cosmo.scalar("angdist", z) / self.fiducialcosmo.scalar("angdist", z)

    def BAO_term(self, z):
        return 1 / (self.qperpendicular(z) ** 2 * self.qparallel(z)) if self.APbool else 1

    # coordinate maps of the lattice (spectro_obs.py:386-506): closed-form host helpers for inspection -- the lattice
    # kernel folds them into per-(point, mu row) constants (RowC in csrc/cfp_spectro.cu) and never calls these
    def kpar(self, z, k, mu):
        return k * mu * (1 / self.qparallel(z))

    def kper(self, z, k, mu):
        return k * np.sqrt(1 - mu**2) * (1 / self.qperpendicular(z))

    def k_units_change(self, k):
        if self.kh_rescaling_bug:
            return k * (self.cosmo.cosmopars["h"] / self.fiducialcosmo.cosmopars["h"])
        return k

    def kmu_alc_pac(self, z, k, mu):
        if not self.APbool:
            return k, mu
        kap = np.sqrt(self.kpar(z, k, mu) ** 2 + self.kper(z, k, mu) ** 2)
        return kap, self.kpar(z, k, mu) / kap

    def spec_err_z(self, z, k, mu):
        dz = self.dz_err if self.dz_type == "constant" else self.dz_err * (1 + z)
        err = dz * (1 / self.cosmo.scalar("Hubble", z)) * self.kpar(z, k, mu)
        return np.exp(-(1 / 2) * err**2)

    def sigmavNL(self, zz, mu):
        """Damping scale of the BAO wiggles (spectro_obs.py:699-722)."""
        if self.linear_switch:
            return 0
        f0, f1, f2 = (self.P_ThetaTheta_Moments(zz, m) for m in (0, 1, 2))
        sv = np.sqrt(f0 + 2 * mu**2 * f1 + mu**2 * f2)
        if self.rescale_sigmav:
            sv = sv * self.nuisance.gcsp_rescale_sigmapv_at_z(zz, sigma_key="sigmav")
        return sv

    def observed_P_ij(self, z, k, mu, bsi="g", bsj="g"):
        """Galaxy auto-spectrum through the lattice kernel (spectro_obs.py:913-933 for bias samples g, g)."""
        if (bsi, bsj) != ("g", "g"):
            raise NotImplementedError("only the galaxy auto-spectrum is in scope")
        return self.observed_Pgg(z, k, mu)

    def lnpobs_ij(self, z, k, mu, bsi="g", bsj="g"):
        if (bsi, bsj) != ("g", "g"):
            raise NotImplementedError("only the galaxy auto-spectrum is in scope")
        return self.lnpobs_gg(z, k, mu)

    def P_ThetaTheta_Moments(self, zz, moment=0):
        c = self.fiducialcosmo if self.fix_cosmo_nl_terms else self.cosmo
        return _theta_moments(c, zz)[moment]

    def sigmapNL(self, zz):
        if self.linear_switch:
            return 0
        sp = np.sqrt(self.P_ThetaTheta_Moments(zz, 2))
        if self.rescale_sigmap:
            sp *= self.nuisance.gcsp_rescale_sigmapv_at_z(zz, sigma_key="sigmap")
        return sp

    # --- the factors of P_obs on a lattice, evaluated by the device (cfp_spectro_plan_terms) ----------------------
    def _terms(self, z, k, mu, b_i=None):
        """[5][Nmu][Nk] at redshift z for a separable (k, mu) mesh: Kaiser, FoG, de-wiggled, spec-z error, P_obs."""
        k, mu = np.asarray(k, dtype=float), np.asarray(mu, dtype=float)
        kb, mb = np.broadcast_arrays(k, mu)
        if not (kb.ndim == 2 and np.all(kb == kb[0:1, :]) and np.all(mb == mb[:, 0:1])):
            raise NotImplementedError("the term accessors take a separable mesh (k, mu from np.meshgrid)")
        return _evaluate_terms(self, float(z), kb[0, :], mb[:, 0], b_i=b_i)

    def kaiserTerm(self, z, k, mu, b_i=None, just_rsd=False, bias_sample="g"):
        """b(z) b(k) + f(z, k) mu^2 at the (k, mu) given -- like the reference's accessor, which `observed_Pgg` calls
        with the already AP-distorted coordinates (spectro_obs.py:584-637); `just_rsd`: 1 + (f / b) mu^2."""
        if bias_sample != "g":
            raise NotImplementedError("only the galaxy sample is in scope")
        kaiser = self._terms(z, k, mu, b_i=b_i)[0]
        if just_rsd:
            rsd = self._terms(z, k, mu, b_i=0.0)[0]  # f mu^2
            return 1 + rsd / (kaiser - rsd)
        return kaiser

    def FingersOfGod(self, z, k, mu, mode="Lorentz"):
        """1 / (1 + (k mu sigma_p)^2) (spectro_obs.py:639-676; 1 when FoG is switched off or the model is linear)."""
        if mode != "Lorentz":
            raise NotImplementedError("only the Lorentzian damping exists in the Fisher path")
        return self._terms(z, k, mu)[1]

    def dewiggled_pdd(self, z, k, mu):
        """P_lin e^{-g} + P_nw (1 - e^{-g}), g = sigma_v^2(mu) k^2, normalised like the reference (spectro_obs.py:811-843)."""
        return self._terms(z, k, mu)[2]

    def spec_err_lattice(self, z, k, mu):
        """exp(-err^2 / 2) on the lattice: `spec_err_z` of the reference evaluated by the device."""
        return self._terms(z, k, mu)[3]

    def flags(self) -> int:
        f = 0
        if self.APbool:
            f |= _lib.SPF_AP
        if self.FoG_switch:
            f |= _lib.SPF_FOG
        if self.linear_switch:
            f |= _lib.SPF_LINEAR
        if self.kh_rescaling_beforespecerr_bug:
            f |= _lib.SPF_KH_BEFORE_ERR
        return f

    def scalar_block(self, z: float) -> np.ndarray:
        """The CFP_SP_* block of this parameter point at redshift z (include/cfp_b200.h)."""
        b = np.zeros(_lib.SP_NSCAL)
        b[_lib.SP_Z] = z
        b[_lib.SP_COSMO] = self.cosmo_index
        b[_lib.SP_QPAR] = self.qparallel(z)
        b[_lib.SP_QPER] = self.qperpendicular(z)
        b[_lib.SP_HUBBLE] = self.cosmo.scalar("Hubble", z)
        b[_lib.SP_DZ] = self.dz_err if self.dz_type == "constant" else self.dz_err * (1 + z)
        if self.use_bias_funcs:
            b[_lib.SP_BIAS] = float(self.nuisance.gcsp_bias_interp()(z))
        else:
            b[_lib.SP_BIAS] = self.nuisance.gcsp_bias_at_z(z)
        b[_lib.SP_A1], b[_lib.SP_A2] = self.nuisance.qbias_coeffs()
        if not self.linear_switch:
            m0, m1, m2 = (self.P_ThetaTheta_Moments(z, m) for m in (0, 1, 2))
            b[_lib.SP_SIGMAP] = self.sigmapNL(z)
            b[_lib.SP_I0], b[_lib.SP_I1], b[_lib.SP_I2] = m0, m1, m2
        b[_lib.SP_SVRESCALE] = (
            self.nuisance.gcsp_rescale_sigmapv_at_z(z, sigma_key="sigmav") if self.rescale_sigmav else 1.0
        )
        if self.s8terms:
            s8 = float(self.cosmo.scalar("sigma8_cb_of_z" if self.tracer == "clustering" else "sigma8_of_z", z))
            b[_lib.SP_FS8], b[_lib.SP_INVS8SQ] = s8, 1.0 / s8**2
        else:
            b[_lib.SP_FS8], b[_lib.SP_INVS8SQ] = 1.0, 1.0
        b[_lib.SP_KHFAC] = (
            self.cosmo.cosmopars["h"] / self.fiducialcosmo.cosmopars["h"] if self.kh_rescaling_bug else 1.0
        )
        b[_lib.SP_PSHOT] = self.extraPshot
        return b

    # --- lattice evaluation on the device --------------------------------------------------------
    def observed_Pgg(self, z, k, mu, b_i=None):
        """P_obs(z; k, mu) for a separable mesh (k, mu from np.meshgrid) or any broadcastable pair of
        arrays; evaluated by the CUDA lattice kernel (reference: spectro_obs.py:845-911)."""
        return np.exp(self.lnpobs_gg(z, k, mu, b_i=b_i))

    def lnpobs_gg(self, z, k, mu, b_i=None):
        """ln P_obs; `b_i` (a number) replaces the whole bias term b(z) b(k) of the Kaiser factor
        (spectro_obs.py:617-618).  Arrays of b_i are not supported."""
        if b_i is not None and np.ndim(b_i) != 0:
            raise NotImplementedError("b_i must be a number")
        k = np.asarray(k, dtype=float)
        mu = np.asarray(mu, dtype=float)
        kb, mb = np.broadcast_arrays(k, mu)
        shape = kb.shape
        if kb.ndim == 2 and np.all(kb == kb[0:1, :]) and np.all(mb == mb[:, 0:1]):
            kg, mg = kb[0, :], mb[:, 0]  # meshgrid layout (Nmu, Nk)
            out = _evaluate_lattice([self], [float(z)], kg, mg, b_i=b_i)[0, 0]
            return out.reshape(shape)
        flat_k, flat_m = kb.ravel(), mb.ravel()
        out = np.empty(flat_k.size)
        # generic point list: one lattice row per distinct mu keeps the kernel's separable layout
        for m in np.unique(flat_m):
            sel = flat_m == m
            out[sel] = _evaluate_lattice([self], [float(z)], flat_k[sel], np.array([m]), b_i=b_i)[0, 0, 0]
        return out.reshape(shape)


def table_slots(settings):
    """Bank slots of P_lin(z,k) and f(z,k) of the GCsp tracer (spectro_obs.py:628-630, 780)."""
    if settings["GCsp_Tracer"] == "clustering":
        return _lib.TAB_PLIN_CB, _lib.TAB_F_CB
    return _lib.TAB_PLIN, _lib.TAB_F


def _cosmo_set_of(cfg) -> CosmoSet:
    cs = cfg.__dict__.get("_cosmo_set")
    if cs is None:
        cs = CosmoSet(cfg, slots=table_slots(cfg.settings))
        cfg.__dict__["_cosmo_set"] = cs
    return cs



def prepare_cosmologies(cs: CosmoSet, zs, obs: "ComputeGalSpectro"):
    """Everything the scalar blocks and no-wiggle tables of the set's cosmologies need from the Boltzmann tables, fitted
    and integrated on the device for all of them at once (CosmoSet.prepare / prepare_spectro, csrc/cfp_prep.cu)."""
    cs.prepare(scalars=zs)
    if not obs.linear_switch:
        cs.prepare_spectro(zs, tracer=obs.tracer)


def _base_blocks(fid_obs: "ComputeGalSpectro", cs: CosmoSet, c: int, zs) -> np.ndarray:
    """[Z][NSCAL] scalar blocks of cosmology `c` of the set with the configuration's FIDUCIAL nuisance values.
    Everything in them that is not a nuisance value derives from the cached Boltzmann tables (H, D_A, sigma8, the
    theta-theta moments at the bin centres), so the array is cached on the cosmology facade, keyed by what else enters
    (bin centres, switches, redshift-error model, fiducial nuisance values)."""
    cosmo = cs.cosmos[c]
    key = (tuple(float(z) for z in zs), fid_obs.flags(), fid_obs.s8terms, fid_obs.tracer, fid_obs.fix_cosmo_nl_terms,
           fid_obs.kh_rescaling_bug, fid_obs.dz_err, fid_obs.dz_type, fid_obs.rescale_sigmap, fid_obs.rescale_sigmav,
           tuple(fid_obs.fiducial_spectrobiaspars.items()), tuple(fid_obs.fiducial_spectrononlinearpars.items()),
           tuple(np.asarray(fid_obs.nuisance.sp_zbins).tolist()), fid_obs.nuisance.sp_bias_model)
    cache = cosmo.__dict__.setdefault("_sp_base_cache", {})
    hit = cache.get(key)
    if hit is None or hit[0] is not fid_obs.fiducialcosmo:  # identity check: the fiducial facade enters q_par, q_perp
        # the parameter point "cosmology c, fiducial nuisance values": a shallow copy of one template object per
        # fiducial observable (the constructor costs more than the four blocks when a batch brings thousands of
        # cosmologies), with the three attributes that name the cosmology rebound
        tmpl = fid_obs.__dict__.get("_base_template")
        if tmpl is None or tmpl.cosmo_set is not cs:
            tmpl = fid_obs.__dict__["_base_template"] = ComputeGalSpectro(
                fid_obs.fiducial_cosmopars, spectrobiaspars=fid_obs.fiducial_spectrobiaspars,
                spectrononlinearpars=fid_obs.fiducial_spectrononlinearpars, PShotpars=fid_obs.fiducial_PShotpars,
                configuration=fid_obs.config, cosmo_set=cs)
        obs = copy.copy(tmpl)
        obs.cosmopars, obs.cosmo, obs.cosmo_index = cosmo.cosmopars, cosmo, c
        arr = np.array([obs.scalar_block(z) for z in zs])
        arr.setflags(write=False)
        hit = (fid_obs.fiducialcosmo, arr)
        cache[key] = hit
    out = hit[1].copy()
    out[:, _lib.SP_COSMO] = c
    return out


def scalar_blocks_from_matrix(fm, fid_obs: "ComputeGalSpectro", zs, names, matrix):
    """Scalar blocks [B][Z][NSCAL], additive shot noise [B][Z] and bank index [B] of a batch of parameter points given
    as a matrix [B, len(names)] -- stencil points of a Fisher job or points of a likelihood batch.  One cached block per
    (cosmology, bin) plus the nuisance columns: no Python object per point.  Unlisted parameters stay fiducial."""
    cs = fid_obs.cosmo_set
    matrix = np.atleast_2d(np.asarray(matrix, dtype=float))
    B, Z = matrix.shape[0], len(zs)
    col = {n: i for i, n in enumerate(names)}

    def column(key, fid):
        return matrix[:, col[key]] if key in col else np.full(B, float(fid))

    ckeys = [k for k in fm.fiducialcosmopars if k in col]
    cosmo_of_s = np.zeros(B, dtype=np.int64)
    if ckeys:
        uniq, first, inv = np.unique(matrix[:, [col[k] for k in ckeys]], axis=0, return_index=True, return_inverse=True)
        idx = np.empty(len(uniq), dtype=np.int64)
        for u in np.argsort(first):  # cosmologies enter the set in the order of their first appearance
            cp = dict(fm.fiducialcosmopars)
            cp.update({k: float(v) for k, v in zip(ckeys, uniq[u])})
            idx[u] = cs.add(cp)
        cosmo_of_s = idx[np.asarray(inv).ravel()]
    prepare_cosmologies(cs, zs, fid_obs)
    cu, cinv = np.unique(cosmo_of_s, return_inverse=True)
    scal = np.stack([_base_blocks(fid_obs, cs, int(c), zs) for c in cu])[np.asarray(cinv).ravel()]  # [B, Z, NSCAL]
    nu = fid_obs.nuisance
    for zb, z in enumerate(zs):
        b = nu.gcsp_zvalue_to_zindex(z)
        if "linear" in nu.sp_bias_model:
            key = nu.sp_bias_root + nu.sp_bias_sample + "_" + str(b)
            vals = column(key, fm.Spectrobiaspars[key])
            scal[:, zb, _lib.SP_BIAS] = np.exp(vals) if nu.sp_bias_model == "linear_log" else vals
        if fid_obs.rescale_sigmap and not fid_obs.linear_switch:
            k2 = "sigmap_" + str(b)
            scal[:, zb, _lib.SP_SIGMAP] = np.sqrt(scal[:, zb, _lib.SP_I2]) * column(k2, fm.Spectrononlinpars.get(k2, 1.0))
        if fid_obs.rescale_sigmav:
            k2 = "sigmav_" + str(b)
            scal[:, zb, _lib.SP_SVRESCALE] = column(k2, fm.Spectrononlinpars.get(k2, 1.0))
    if nu.sp_bias_model == "linear_Qbias":
        scal[:, :, _lib.SP_A1] = column("A1", fm.Spectrobiaspars.get("A1", 0.0))[:, None]
        scal[:, :, _lib.SP_A2] = column("A2", fm.Spectrobiaspars.get("A2", 0.0))[:, None]
    shot = np.zeros((B, Z))
    for i, key in enumerate(fm.PShotpars.keys()):
        if i < Z:
            shot[:, i] = column(key, fm.PShotpars[key])
    return scal, shot, cosmo_of_s


_PNW_MEMO = {}


def _pnw_arrays(cosmo_set: CosmoSet, zs, used):
    """No-wiggle knots/values [NC][nnw], [NC][Z][nnw]; rows of unused cosmologies are left as ones.  Remembered per
    (cosmologies, redshifts): every forecast of a fiducial asks for the same arrays (read-only)."""
    tracer = cosmo_set.config.settings["GCsp_Tracer"]
    nnw = cosmo_set.config.settings["savgol_internalsamples"]
    cosmos = list(cosmo_set.cosmos)
    key = (tuple(id(c) for c in cosmos), tuple(float(z) for z in zs), tuple(sorted(used)), tracer, nnw)
    hit = _PNW_MEMO.get(key)
    if hit is not None and all(a is b for a, b in zip(hit[0], cosmos)):
        return hit[1], hit[2]
    NC, Z = len(cosmos), len(zs)
    pk = np.ones((NC, nnw))
    pk[:] = np.arange(1, nnw + 1)[None, :]
    pp = np.ones((NC, Z, nnw))
    for c in sorted(used):
        for j, z in enumerate(zs):
            kk, vv, _ = cosmos[c].nonwiggle_table(z, tracer=tracer)
            pk[c] = kk
            pp[c, j] = vv
    pk.setflags(write=False)
    pp.setflags(write=False)
    if len(_PNW_MEMO) >= 16:
        _PNW_MEMO.pop(next(iter(_PNW_MEMO)))
    _PNW_MEMO[key] = (cosmos, pk, pp)
    return pk, pp


def _evaluate_lattice(points: List[ComputeGalSpectro], zs, kgrid, mugrid, b_i=None) -> np.ndarray:
    """ln P_obs [S][Z][Nmu][Nk] of several parameter points (no derivatives); b_i: a number replacing the bias term."""
    cs = points[0].cosmo_set
    S, Z = len(points), len(zs)
    prepare_cosmologies(cs, zs, points[0])
    scal = np.array([[p.scalar_block(z) for z in zs] for p in points])
    if b_i is not None:
        scal[:, :, _lib.SP_BIAS] = float(b_i)
        scal[:, :, _lib.SP_A1] = scal[:, :, _lib.SP_A2] = 0.0
    alias = np.tile(np.arange(S, dtype=np.int32)[:, None], (1, Z))
    linear = points[0].linear_switch
    used = {p.cosmo_index for p in points}
    pk, pp = _pnw_arrays(cs, zs, set() if linear else used)
    plan = _lib.SpectroPlan(
        cs.context(), cs.bank(), k=kgrid, mu=mugrid, wk=np.zeros_like(kgrid),
        wmu=np.zeros_like(mugrid), scal=scal, alias=alias, pnw_k=pk, pnw_p=pp, stw=np.zeros((1, S)),
        stden=np.ones(1), ps_bin=np.array([-1], dtype=np.int32), ps_src=0, nbar=np.zeros(Z), vol=np.zeros(Z),
        flags=points[0].flags(), tab_plin=table_slots(points[0].settings)[0], tab_f=table_slots(points[0].settings)[1])
    plan.run(materialize=_lib.MAT_LNP)
    _, lnp, _, _ = plan.fetch(lnp=True)
    plan.close()
    return lnp


def _evaluate_terms(point: ComputeGalSpectro, z: float, kgrid, mugrid, b_i=None) -> np.ndarray:
    """[5][Nmu][Nk] factors of P_obs of one parameter point at redshift z (cfp_spectro_plan_terms)."""
    cs = point.cosmo_set
    prepare_cosmologies(cs, [z], point)
    scal = np.array([[point.scalar_block(z)]])
    if b_i is not None:
        scal[:, :, _lib.SP_BIAS] = float(b_i)
        scal[:, :, _lib.SP_A1] = scal[:, :, _lib.SP_A2] = 0.0
    # the reference's accessors take (k, mu) as given (observed_Pgg hands them the AP-distorted values): no remap here
    scal[:, :, _lib.SP_KHFAC] = 1.0
    flags = point.flags() & ~_lib.SPF_AP
    pk, pp = _pnw_arrays(cs, [z], set() if point.linear_switch else {point.cosmo_index})
    plan = _lib.SpectroPlan(
        cs.context(), cs.bank(), k=kgrid, mu=mugrid, wk=np.zeros_like(kgrid), wmu=np.zeros_like(mugrid), scal=scal,
        alias=np.zeros((1, 1), dtype=np.int32), pnw_k=pk, pnw_p=pp, stw=np.zeros((1, 1)), stden=np.ones(1),
        ps_bin=np.array([-1], dtype=np.int32), ps_src=0, nbar=np.zeros(1), vol=np.zeros(1), flags=flags,
        tab_plin=table_slots(point.settings)[0], tab_f=table_slots(point.settings)[1])
    out = plan.terms(0)[:, 0]
    plan.close()
    return out


class SpectroCov:
    """Survey volume, number density and effective volume (reference: spectro_cov.py:19-228)."""

    def __init__(self, fiducialpars, fiducial_specobs=None, bias_samples=["g", "g"], configuration=None):
        if configuration is None:
            from . import config as config_module

            configuration = config_module.current()
        self.config = configuration
        self.feed_lvl = configuration.settings["feedback"]
        try:
            self.fsky_spectro = configuration.specs["fsky_spectro"]
            self.area_survey = self.fsky_spectro * AREASKY
        except KeyError:
            self.area_survey = configuration.specs["area_survey_spectro"]
            self.fsky_spectro = self.area_survey / AREASKY
        if fiducial_specobs is None:
            fiducial_specobs = ComputeGalSpectro(fiducialpars, fiducial_cosmopars=fiducialpars,
                                                 bias_samples=bias_samples, configuration=configuration)
        self.pk_obs = fiducial_specobs
        n = self.pk_obs.nuisance
        self.dnz = n.gcsp_dndz()
        self.z_bins = n.gcsp_zbins()
        self.z_bin_mids = n.gcsp_zbins_mids()
        self.dz_bins = np.diff(self.z_bins)
        self.global_z_bin_mids = self.inter_z_bin_mids = self.z_bin_mids
        self.global_z_bins = self.inter_z_bins = self.z_bins

    def volume_bin(self, zi, zj):
        fid = self.pk_obs.fiducialcosmo  # (memoised per redshift on the facade: every forecast asks for the bin edges)
        d1 = fid.scalar("angdist", zi) if hasattr(fid, "scalar") else fid.angdist(zi)
        d2 = fid.scalar("angdist", zj) if hasattr(fid, "scalar") else fid.angdist(zj)
        return (4 * np.pi / 3) * (pow((1 + zj) * d2, 3) - pow((1 + zi) * d1, 3))

    def d_volume(self, ibin):
        return self.volume_bin(self.global_z_bins[ibin], self.global_z_bins[ibin + 1])

    def volume_survey(self, ibin):
        return self.fsky_spectro * self.d_volume(ibin)

    @property
    def volume_survey_array(self):
        """Missing in the reference (SURVEY.md §8a-L2'): per-bin survey volumes, used by the likelihood."""
        return np.array([self.volume_survey(i) for i in range(len(self.global_z_bin_mids))])

    def n_density(self, zi):
        memo = self.__dict__.setdefault("_n_density_memo", {})
        key = float(zi)
        if key in memo:
            return memo[key]
        memo[key] = out = self._n_density(zi)
        return out

    def _n_density(self, zi):
        # np.isclose(mids, zi, rtol=1e-2) spelled out for a handful of bins: |m - zi| <= atol + rtol |zi|
        tol = 1e-8 + 1e-2 * abs(float(zi))
        w = [i for i, m in enumerate(self.inter_z_bin_mids) if abs(float(m) - float(zi)) <= tol]
        if len(w) == 0:
            print(f"Warning: zi = {zi} not in global_z_bin_mids")
            return 0
        i = w[0]
        return AREASKY * (self.dnz[i] * self.dz_bins[i] / self.d_volume(i))

    def veff(self, zi, k, mu):
        npobs = self.n_density(zi) * self.pk_obs.observed_Pgg(zi, k, mu)
        cov = (1 / (8 * (np.pi**2))) * (npobs / (1 + npobs)) ** 2
        if zi < self.inter_z_bin_mids[0] or zi > self.inter_z_bin_mids[-1]:
            cov = np.zeros_like(cov)
        return cov

    def cov(self, ibin, k, mu):
        """Covariance of the clustering spectrum in bin ibin (spectro_cov.py:230-252), restated as written: the
        reference hands the bin INDEX to veff(), which expects a redshift."""
        zmid = self.global_z_bin_mids[ibin]
        veffS = self.veff(ibin, k, mu) * self.volume_survey(ibin)
        pobs = self.pk_obs.observed_Pgg(zmid, k, mu)
        return (2 * (2 * np.pi) ** 3 / veffS) * pobs**2 * (1 / k) ** 3

    def noisy_P_ij(self, z, k, mu, si="g", sj="g"):
        if not (si == "g" and sj == "g"):
            raise NotImplementedError("only the galaxy auto-spectrum is in scope")
        return self.pk_obs.observed_Pgg(z, k, mu) + 1 / self.n_density(z)


class SpectroJob:
    """Everything one GCsp Fisher needs on the device: stencil points, scalar blocks, aliases,
    no-wiggle tables, quadrature weights.  Built once per FisherMatrix.compute()."""

    def __init__(self, fiducial_obs: ComputeGalSpectro, cov: SpectroCov, zmids, kgrid, mugrid, freeparams: Dict,
                 stencil: str = "3PT"):
        cfg = fiducial_obs._cfg
        self.fid = fiducial_obs
        self.cov = cov
        self.zmids = np.asarray(zmids, dtype=float)
        self.kgrid, self.mugrid = np.asarray(kgrid, float), np.asarray(mugrid, float)
        self.freeparams = dict(freeparams)
        self.parnames = list(self.freeparams)
        self.fiducial_allpars = fiducial_obs.fiducial_allpars
        if stencil == "STEM":  # derivatives.py:427
            raise ValueError("STEM derivative not availabe for spectropscopic probes yet!")
        pts, den = stencils.get(stencil)
        cs = fiducial_obs.cosmo_set
        fid_all = self.fiducial_allpars
        point_pars = [dict(fid_all)]   # s = 0: the fiducial
        npar = len(self.parnames)
        stw_rows, stden, ps_bin = [], np.ones(npar), -np.ones(npar, dtype=np.int32)
        last_eval, ps_src = None, None
        for ip, par in enumerate(self.parnames):
            row = {}
            if "Ps" in par:  # analytic shot-noise derivative, spectro_cov.py:510-532
                if last_eval is None:
                    raise AttributeError("a Ps_* parameter precedes every finite-difference parameter: the reference "
                                         "has no observable to evaluate 1/P_obs on (spectro_cov.py:526)")
                if ps_src is not None and ps_src != last_eval:
                    raise NotImplementedError("Ps_* parameters interleaved with finite-difference parameters")
                ps_src = last_eval
                ps_bin[ip] = int(par.split("_")[-1]) - 1
                stw_rows.append(row)
                continue
            h = stencils.stepsize(fid_all[par], self.freeparams[par])
            stden[ip] = den * h
            for m, w in pts:
                if m == 0:
                    row[0] = row.get(0, 0.0) + w
                    continue
                ap = dict(fid_all)  # flat {name: number}
                ap[par] = ap[par] + m * h
                point_pars.append(ap)
                row[len(point_pars) - 1] = w
                last_eval = len(point_pars) - 1
            stw_rows.append(row)
        self.point_pars = point_pars
        self._cfg, self._points = cfg, None
        S, Z = len(point_pars), len(self.zmids)
        self.stw = np.zeros((npar, S))
        for ip, row in enumerate(stw_rows):
            for s, w in row.items():
                self.stw[ip, s] = w
        self.stden, self.ps_bin, self.ps_src = stden, ps_bin, (0 if ps_src is None else ps_src)
        if fiducial_obs.use_bias_funcs:
            pts = self.points
            prepare_cosmologies(cs, self.zmids, fiducial_obs)
            self.scal = np.array([[p.scalar_block(z) for z in self.zmids] for p in pts])
        else:
            # every stencil point as a row of a matrix: one cached block per (cosmology, bin) + the nuisance columns
            from types import SimpleNamespace

            names = list(fid_all)
            mat = np.array([[ap[k] for k in names] for ap in point_pars], dtype=float)
            view = SimpleNamespace(fiducialcosmopars=cfg.fiducialparams, Spectrobiaspars=cfg.Spectrobiasparams,
                                   Spectrononlinpars=cfg.Spectrononlinearparams, PShotpars={})
            self.scal, _, _ = scalar_blocks_from_matrix(view, fiducial_obs, self.zmids, names, mat)
        # a stencil point whose block equals the fiducial's in a bin reuses the fiducial lattice there
        same = np.all(self.scal == self.scal[0:1], axis=2)
        self.alias = np.where(same, 0, np.arange(S)[:, None]).astype(np.int32)
        used = {int(c) for c in self.scal[:, :, _lib.SP_COSMO].ravel()}
        self.pnw_k, self.pnw_p = _pnw_arrays(cs, self.zmids, set() if fiducial_obs.linear_switch else used)
        self.wk, self.wmu = simpson_weights(self.kgrid), simpson_weights(self.mugrid)
        self.nbar = np.array([cov.n_density(z) for z in self.zmids])
        self.vol = np.array([cov.volume_survey(i) for i in range(Z)])
        self.flags = fiducial_obs.flags()
        self.tabs = table_slots(fiducial_obs.settings)
        self.cosmo_set = cs
        self._plan = None

    @property
    def points(self):
        """ComputeGalSpectro object of every stencil point (built on demand: the job itself works on the matrix of
        their parameter values)."""
        if self._points is None:
            self._points = [self.fid] + [self._point(ap, self._cfg, self.fid.cosmo_set) for ap in self.point_pars[1:]]
        return self._points

    @staticmethod
    def _point(allpars, cfg, cs):
        """ComputeGalSpectro of one stencil point (spectro_cov.py:456-475)."""
        return ComputeGalSpectro(
            cosmopars={k: allpars[k] for k in cfg.fiducialparams},
            fiducial_cosmopars=cfg.fiducialparams,
            spectrobiaspars={k: allpars[k] for k in cfg.Spectrobiasparams},
            spectrononlinearpars={k: allpars[k] for k in cfg.Spectrononlinearparams},
            PShotpars={k: allpars[k] for k in cfg.PShotparams},
            configuration=cfg, cosmo_set=cs)

    def plan(self) -> _lib.SpectroPlan:
        if self._plan is None:
            cs = self.cosmo_set
            self._plan = _lib.SpectroPlan(
                cs.context(), cs.bank(), k=self.kgrid, mu=self.mugrid, wk=self.wk, wmu=self.wmu,
                scal=self.scal, alias=self.alias, pnw_k=self.pnw_k, pnw_p=self.pnw_p, stw=self.stw, stden=self.stden,
                ps_bin=self.ps_bin, ps_src=self.ps_src, nbar=self.nbar, vol=self.vol, flags=self.flags,
                tab_plin=self.tabs[0], tab_f=self.tabs[1])
        return self._plan

    def run(self, materialize=0):
        p = self.plan()
        p.run(materialize=materialize)
        return p

    def close(self):
        if self._plan is not None:
            self._plan.close()
            self._plan = None


class SpectroDerivs:
    """Derivatives of ln P_obs w.r.t. every free parameter (reference: spectro_cov.py:373-601).

    `compute_derivs` returns the reference's structure {par: {"z_bins": z_array, i: (Nmu, Nk)}}; the
    lattices, the stencil and (through `fisher_per_bin`) the Fisher reduction are one fused device pass."""

    def __init__(self, z_array, pk_kmesh, pk_mumesh, fiducial_spectro_obj, bias_samples=["g", "g"],
                 configuration=None, stencil="3PT"):
        self.observables = fiducial_spectro_obj.observables
        self.bias_samples = bias_samples
        self.fiducial_cosmopars = fiducial_spectro_obj.fiducial_cosmopars
        self.fiducial_spectrobiaspars = fiducial_spectro_obj.fiducial_spectrobiaspars
        self.fiducial_PShotpars = fiducial_spectro_obj.PShotpars
        self.fiducial_allpars = fiducial_spectro_obj.fiducial_allpars
        self.fiducial_spectrononlinearpars = fiducial_spectro_obj.fiducial_spectrononlinearpars
        self.fiducial_cosmo = fiducial_spectro_obj.fiducialcosmo
        self.fiducial_obj = fiducial_spectro_obj
        self.z_array = np.asarray(z_array)
        self.pk_kmesh, self.pk_mumesh = pk_kmesh, pk_mumesh
        # The device works on a separable (mu, k) lattice.  A np.meshgrid pair maps onto it directly; any other
        # broadcastable pair (the reference evaluates whatever it is given) is evaluated on the lattice of its
        # distinct k and mu values and gathered back.
        kk, mm = np.asarray(pk_kmesh, dtype=float), np.asarray(pk_mumesh, dtype=float)
        # (broadcast views, what FisherMatrix.set_pk_settings passes, are separable by construction: no 2 x Nmu x Nk scan)
        if kk.ndim == 2 and kk.shape == mm.shape and (
                (kk.strides[0] == 0 and mm.strides[1] == 0) or (np.all(kk == kk[0:1, :]) and np.all(mm == mm[:, 0:1]))):
            self._kgrid, self._mugrid, self._gather = kk[0, :], mm[:, 0], None
        else:
            kb, mb = np.broadcast_arrays(kk, mm)
            self._kgrid, self._mugrid = np.unique(kb), np.unique(mb)
            self._gather = (np.searchsorted(self._mugrid, mb), np.searchsorted(self._kgrid, kb))
        self.freeparams = None
        self.stencil = stencil
        self.config = configuration if configuration is not None else fiducial_spectro_obj.config
        self.feed_lvl = self.config.settings["feedback"]
        self.job: Optional[SpectroJob] = None
        self.pobs = None

    def get_obs(self, allpars):
        """{"z_bins": z_array, i: ln P_obs lattice} at an arbitrary parameter point (spectro_cov.py:481-508)."""
        cfg = self.fiducial_obj._cfg
        self.pobs = SpectroJob._point(allpars, cfg, self.fiducial_obj.cosmo_set)
        lat = _evaluate_lattice([self.pobs], list(self.z_array), self._kgrid, self._mugrid)
        out = {"z_bins": self.z_array}
        for i in range(len(self.z_array)):
            out[i] = lat[0, i] if self._gather is None else lat[0, i][self._gather]
        return out

    def exact_derivs(self, par):
        """Analytic derivative of ln P_obs w.r.t. a shot-noise parameter Ps_i: delta_{i, bin} / P_obs, with P_obs of the
        LAST evaluated stencil point (spectro_cov.py:510-532); None for any other parameter -- the
        `special_deriv_function` signature of the derivative engine.  Read from the device job's derivative lattices."""
        if "Ps" not in par:
            return None
        if getattr(self, "derivs", None) is None or par not in self.derivs:
            self.compute_derivs()
        if par not in self.derivs:
            return None
        return {"z_bins": self.z_array, **{i: self.derivs[par][i] for i in range(len(self.z_array))}}

    def kronecker_bins(self, par):
        """[delta_{i, bin of par}] over the bins (spectro_cov.py:534-561)."""
        ib = int(par.split("_")[-1]) - 1
        return np.array([1.0 if i == ib else 0.0 for i in range(len(self.z_array))])

    def build_job(self, cov: SpectroCov, freeparams=None) -> SpectroJob:
        if freeparams:
            self.freeparams = freeparams
        self.job = SpectroJob(self.fiducial_obj, cov, self.z_array, self._kgrid, self._mugrid, self.freeparams,
                              self.stencil)
        return self.job

    def compute_derivs(self, freeparams=dict(), cov: Optional[SpectroCov] = None):
        if freeparams != dict():
            self.freeparams = freeparams
        if self.job is None:
            if cov is None:
                cov = SpectroCov(self.fiducial_cosmopars, fiducial_specobs=self.fiducial_obj, configuration=self.config)
            self.build_job(cov)
        plan = self.job.run(materialize=_lib.MAT_DERIVS)
        _, _, D, V = plan.fetch(derivs=True)
        g = self._gather
        self.veff = V if g is None else np.array([v[g] for v in V])
        derivs = {}
        for ip, par in enumerate(self.job.parnames):
            d = {i: (D[ip, i] if g is None else D[ip, i][g]) for i in range(len(self.z_array))}
            if self.job.ps_bin[ip] < 0:
                d = {"z_bins": self.z_array, **d}
            derivs[par] = d
        self.derivs = derivs
        return derivs
