"""Photometric probes (WL, GCph, 3x2pt): host-side mirror of the reference's operator interface.

Same class names and public methods as the reference (cosmicfishpie/LSSsurvey/photo_window.py
`GalaxyPhotoDist`, photo_obs.py `ComputeCls`, photo_cov.py `PhotoCov`).  The Limber power-spectrum table,
lensing efficiencies, tomographic kernels, the C_ell contraction, the covariance Cholesky and the Fisher
trace run in the CUDA kernels of csrc/cfp_photo.cu; the host prepares cosmology-independent windows
(n_i(z), erf photo-z convolution), 1-D background splines on the z grid, and the O(50)-point IA window.

There is no CPU implementation of the C_ell / Fisher math in this package.
"""
from __future__ import annotations

import copy
import os
from time import time
from typing import Dict, List, Optional

import numpy as np
from scipy.integrate import trapezoid
from scipy.interpolate import InterpolatedUnivariateSpline, interp1d
from scipy.special import erf

from . import _lib, stencils
from .cosmology import CosmoSet

_WINDOW_CACHE: Dict[tuple, "GalaxyPhotoDist"] = {}


def _cfg_of(configuration):
    if configuration is None:
        from . import config as config_module

        configuration = config_module.current()
    return configuration, getattr(configuration, "config", configuration)


class GalaxyPhotoDist:
    """n(z) and photo-z convolved, normalised bin distributions (reference: photo_window.py:15-242).

    Vectorised over z.  The only operation numpy evaluates differently for scalars and arrays is `pow`
    (libm vs SIMD); the reference calls point by point, so `dNdz` takes its power per element and the
    windows agree with the reference bit for bit."""

    def __init__(self, photopars, configuration=None):
        configuration, _ = _cfg_of(configuration)
        sp = configuration.specs
        self.z_bins_WL = sp["z_bins_WL"]
        self.n_bins_WL = len(self.z_bins_WL)
        self.z_bins_GCph = sp["z_bins_GCph"]
        self.n_bins_GCph = len(self.z_bins_GCph)
        self.z0, self.z0_p, self.ngamma = sp["z0"], sp["z0_p"], sp["ngamma"]
        self.photo = photopars
        self.z_min = np.min([sp["z_bins_GCph"][0], sp["z_bins_WL"][0]])
        self.z_max = np.max([sp["z_bins_GCph"][-1], sp["z_bins_WL"][-1]])
        self.normalization = {"GCph": self.norm("GCph"), "WL": self.norm("WL")}

    def dNdz(self, z):
        z = np.asarray(z, dtype=float)
        expo = z / self.z0
        powv = np.array([e**self.ngamma for e in expo.ravel()]).reshape(expo.shape)
        return (z / self.z0_p) ** 2 * np.exp(-powv)

    def n_i(self, z, i, obs="GCph"):
        """dN/dz restricted to tomographic bin i, no photo-z errors (photo_window.py:85-110)."""
        zb = self.z_bins_WL if obs == "WL" else self.z_bins_GCph
        nb = self.n_bins_WL if obs == "WL" else self.n_bins_GCph
        if i == 0 or i > nb:
            return None
        z = np.atleast_1d(z)
        out = self.dNdz(z)
        out[~((z <= zb[i]) & (z >= zb[i - 1]))] = 0.0
        return out

    def ngal_photoz(self, z, i, obs):
        if obs == "GCph":
            zb = self.z_bins_GCph
        elif obs == "WL":
            zb = self.z_bins_WL
        else:
            raise ValueError("obs must be either 'GCph' or 'WL'")
        z = np.asarray(z, dtype=float)
        p = self.photo
        r = np.sqrt(1 / 2)
        t1 = p["cb"] * p["fout"] * erf((r * (z - p["zo"] - p["co"] * zb[i - 1])) / (p["sigma_o"] * (1 + z)))
        t2 = -p["cb"] * p["fout"] * erf((r * (z - p["zo"] - p["co"] * zb[i])) / (p["sigma_o"] * (1 + z)))
        t3 = p["co"] * (1 - p["fout"]) * erf((r * (z - p["zb"] - p["cb"] * zb[i - 1])) / (p["sigma_b"] * (1 + z)))
        t4 = -p["co"] * (1 - p["fout"]) * erf((r * (z - p["zb"] - p["cb"] * zb[i])) / (p["sigma_b"] * (1 + z)))
        return self.dNdz(z) * (t1 + t2 + t3 + t4) / (2 * p["co"] * p["cb"])

    def norm(self, obs):
        zb = self.z_bins_GCph if obs == "GCph" else self.z_bins_WL
        zint = np.linspace(0.0, self.z_max, 1000)
        dz = self.z_max / 1000  # (sic) photo_window.py:210
        out = [trapezoid(self.ngal_photoz(zint, i, obs), dx=dz) for i in range(1, len(zb))]
        out.insert(0, None)
        return out

    def norm_ngal_photoz(self, z, i, obs):
        return self.ngal_photoz(z, i, obs) / self.normalization[obs][i]


def galaxy_photo_dist(photopars, configuration) -> GalaxyPhotoDist:
    """Windows depend only on (bin edges, n(z) shape, photo-z parameters): computed once and cached."""
    sp = configuration.specs
    key = (tuple(sp["z_bins_WL"]), tuple(sp["z_bins_GCph"]), float(sp["z0"]), float(sp["ngamma"]),
           tuple(sorted(photopars.items())))
    if key not in _WINDOW_CACHE:
        _WINDOW_CACHE[key] = GalaxyPhotoDist(photopars, configuration)
    return _WINDOW_CACHE[key]


_LUM_CACHE = {}


def _lumratio_table(extdir):
    """<L>(z)/L*(z) table of the eNLA model: `lumratio_file.dat` of `external_data_dir` (nuisance.py:193-230); the
    package's own data directory ships the same table as .npy.  None if the directory has no such file -- the
    reference then uses a ratio of 1 (with a warning)."""
    if extdir not in _LUM_CACHE:
        here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
        txt = os.path.join(extdir or "", "lumratio_file.dat")
        if extdir and os.path.isfile(txt):
            _LUM_CACHE[extdir] = np.loadtxt(txt)
        elif extdir is None or os.path.abspath(extdir) == here:
            _LUM_CACHE[extdir] = np.load(os.path.join(here, "lumratio.npy"))
        else:
            _LUM_CACHE[extdir] = None
    return _LUM_CACHE[extdir]


class PhotoNuisance:
    """Galaxy bias models and the eNLA intrinsic-alignment window (reference: nuisance.py:89-230)."""

    def __init__(self, configuration=None):
        configuration, _ = _cfg_of(configuration)
        self.specs, self.settings = configuration.specs, configuration.settings
        self.observables = configuration.obs
        sp = self.specs
        self.z = np.linspace(min(sp["z_bins_WL"][0], sp["z_bins_GCph"][0]),
                             max(sp["z_bins_WL"][-1], sp["z_bins_GCph"][-1]) + 1, 50 * self.settings["accuracy"])
        if "WL" in self.observables:
            lum = _lumratio_table(self.settings.get("external_data_dir"))
            if lum is None:  # nuisance.py:220-228
                self.lumratio = lambda z: np.ones_like(z) if hasattr(z, "__len__") else 1.0
            else:
                self.lumratio = interp1d(lum[:, 0], lum[:, 1], kind="linear")
            self._lum_z = self.lumratio(self.z)

    def gcph_bias(self, biaspars, ibin=1):
        z = self.z
        model = biaspars["bias_model"]
        if model == "sqrt":
            return interp1d(z, biaspars["b0"] * np.sqrt(1 + z), kind="linear")
        if model == "binned":
            edges = np.asarray(self.specs["z_bins_GCph"])
            brang = self.specs["binrange_GCph"]
            last = brang[-1]
            vals = np.array([biaspars[f"b{k}"] for k in brang], dtype=float)
            return lambda za: vals[np.clip(np.searchsorted(edges, np.asarray(za), side="right") - 1, 0, last - 1)]
        if model == "binned_constant":
            return lambda za: np.full(np.shape(za), biaspars["b" + str(ibin)], dtype=float)
        if model == "flagship":
            b = biaspars["A"] + biaspars["B"] / (1 + np.exp(-biaspars["C"] * (z - biaspars["D"])))
            return interp1d(z, b, kind="linear")
        raise ValueError("unknown galaxy bias model: " + str(model))

    def IA(self, IApars, cosmo):
        if IApars["IA_model"] != "eNLA":
            raise ValueError("only the eNLA intrinsic alignment model exists")
        # memoised accessors of this package's cosmology facade; any object with the reference's
        # `Omegam_of_z` / `growth` methods works too
        Om = cosmo.scalar("Omegam_of_z", 0.0) if hasattr(cosmo, "scalar") else cosmo.Omegam_of_z(0.0)
        piv = self.settings["pivot_z_IA"]
        z = self.z
        CIA = 0.0134 * (1 + piv)
        fac = -IApars["AIA"] * CIA * Om
        z_dep = (1 + z) / (1 + piv)
        growth = cosmo.on_grid("growth", z) if hasattr(cosmo, "on_grid") else np.asarray(cosmo.growth(z))
        win = fac * z_dep ** IApars["etaIA"] * ((self._lum_z ** IApars["betaIA"]) / (growth.flatten()))
        return InterpolatedUnivariateSpline(z, win, k=1)


_GRID_MEMO = {}


def _photo_grid(configuration) -> "_PhotoGrid":
    """The `_PhotoGrid` of a configuration, remembered by everything that enters it (a forecast builds the grid three
    times; specifications may be edited between calls, so the key is the values, not the object)."""
    configuration, cfg = _cfg_of(configuration)
    sp, st = configuration.specs, configuration.settings
    try:
        key = (tuple(configuration.obs), tuple(sp["binrange_GCph"]), tuple(sp["binrange_WL"]),
               sp["lmax_GCph"], sp["lmax_WL"], sp["lmin_GCph"], sp["lmin_WL"], st["accuracy"], st["ell_sampling"],
               float(sp["z_bins_GCph"][0]), float(sp["z_bins_WL"][0]), float(sp["z_bins_GCph"][-1]),
               float(sp["z_bins_WL"][-1]), sp["ellipt_error"], np.asarray(sp["ngalbin_WL"], float).tobytes(),
               np.asarray(sp["ngalbin_GCph"], float).tobytes(), sp["fsky_GCph"], sp["fsky_WL"], float(sp["kmax"]),
               st["nonlinear"], st["GCph_Tracer"], bool(st["activateMG"]), bool(st["external_activateMG"]))
        hash(key)
    except (KeyError, TypeError):
        return _PhotoGrid(configuration)
    g = _GRID_MEMO.get(key)
    if g is None:
        if len(_GRID_MEMO) >= 32:
            _GRID_MEMO.clear()
        g = _GRID_MEMO[key] = _PhotoGrid(configuration)
        for a in (g.ell, g.z, g.kind, g.lmin, g.lmax, g.noise, g.sqrtfsky):
            a.setflags(write=False)
    return g


class _PhotoGrid:
    """ell / z sampling and tomographic row layout shared by every stencil point (photo_obs.py:287-318)."""

    def __init__(self, configuration):
        configuration, cfg = _cfg_of(configuration)
        sp, st = configuration.specs, configuration.settings
        self.observables = [o for o in configuration.obs if o in ("GCph", "WL")]
        if not self.observables:
            raise ValueError("Observables not specified correctly")
        self.binrange = {"GCph": sp["binrange_GCph"], "WL": sp["binrange_WL"]}
        if len(self.observables) == 2:
            ellmax, ellmin = max(sp["lmax_GCph"], sp["lmax_WL"]), min(sp["lmin_GCph"], sp["lmin_WL"])
        else:
            o = self.observables[0]
            ellmax, ellmin = sp["lmax_" + o], sp["lmin_" + o]
        self.ellmax, self.ellmin = ellmax, ellmin
        self.zsamp = int(round(200 * st["accuracy"]))
        if st["ell_sampling"] == "accuracy":
            self.ellsamp = int(round(100 * st["accuracy"]))
        elif isinstance(st["ell_sampling"], int):
            self.ellsamp = st["ell_sampling"]
        else:
            raise ValueError("ell_sampling should be an integer or 'accuracy'")
        self.ell = np.logspace(np.log10(ellmin), np.log10(ellmax + 10), num=self.ellsamp)
        self.z_min = np.min([sp["z_bins_GCph"][0], sp["z_bins_WL"][0]])
        self.z_max = np.max([sp["z_bins_GCph"][-1], sp["z_bins_WL"][-1]])
        self.z = np.linspace(self.z_min, self.z_max, self.zsamp)
        self.dz = np.mean(np.diff(self.z))
        # rows: all bins of observable 0, then observable 1 (cosmicfish.py:686-691)
        self.rows = [(o, i) for o in self.observables for i in self.binrange[o]]
        self.cols = [f"{o} {i}" for o, i in self.rows]
        self.kind = np.array([0 if o == "WL" else 1 for o, _ in self.rows], dtype=np.int32)
        self.lmin = np.array([sp["lmin_" + o] for o, _ in self.rows], dtype=float)
        self.lmax = np.array([sp["lmax_" + o] for o, _ in self.rows], dtype=float)
        noise_wl = (sp["ellipt_error"] ** 2.0) / np.array(sp["ngalbin_WL"])       # photo_cov.py:111
        noise_gc = 1.0 / np.array(sp["ngalbin_GCph"])                            # photo_cov.py:114
        self.noise = np.array([noise_wl[i - 1] if o == "WL" else noise_gc[i - 1] for o, i in self.rows])
        self.sqrtfsky = np.array([np.sqrt(sp["fsky_" + o]) for o, _ in self.rows])  # photo_cov.py:228
        self.kmax = float(sp["kmax"])
        self.tab_p = _lib.TAB_PNL if st["nonlinear"] else _lib.TAB_PLIN
        # sqrt(P) of the clustering rows is taken on the tracer's spectrum (photo_obs.py:483-507)
        if st["GCph_Tracer"] == "clustering":
            self.tab_pg = _lib.TAB_PNL_CB if st["nonlinear"] else _lib.TAB_PLIN_CB
        elif st["GCph_Tracer"] == "matter":
            self.tab_pg = self.tab_p
        else:
            raise ValueError("GCph_Tracer must be 'matter' or 'clustering'")
        if st["activateMG"] or st["external_activateMG"]:
            raise NotImplementedError("modified-gravity Sigma(z,k) is not supported")


def _ngal_rows(win, g: _PhotoGrid) -> np.ndarray:
    """Normalised n_i(z) of every tomographic row on the job's z grid, cached on the window object."""
    cache = win.__dict__.setdefault("_ngal_rows_cache", {})
    key = (g.zsamp, float(g.z_min), float(g.z_max), tuple(g.rows))
    if key not in cache:
        arr = np.array([win.norm_ngal_photoz(g.z, i, o) for o, i in g.rows])
        arr.setflags(write=False)
        cache[key] = arr
    return cache[key]


class PhotoJob:
    """One photometric job on the device: stencil points (parameter dictionaries) -> C_ell of each and,
    if a stencil is given, the Fisher matrix."""

    def __init__(self, configuration, points: List[dict], stw, stden, cosmo_set: Optional[CosmoSet] = None):
        configuration, cfg = _cfg_of(configuration)
        self.configuration, self.cfg = configuration, cfg
        g = self.grid = _photo_grid(configuration)
        cs = self.cosmo_set = cosmo_set if cosmo_set is not None else _cosmo_set_of(cfg)
        nuis = PhotoNuisance(configuration)
        fid_photo = cfg.photoparams
        S, Nz, Nb = len(points), g.zsamp, len(g.rows)
        self.cosmo_of_s = np.zeros(S, dtype=np.int32)
        self.bias = np.ones((S, Nb, Nz))
        self.ia = np.zeros((S, Nz))
        # most stencil points share their nuisance values with the fiducial: evaluate each distinct
        # (IA parameters, cosmology) window and each distinct bias row once
        ia_memo, bias_memo = {}, {}
        per_bin_bias = cfg.Photobiasparams.get("bias_model") == "binned_constant"
        win_sets, win_index = [fid_photo], {tuple(fid_photo.items()): 0}
        self.win_of_s = np.zeros(S, dtype=np.int32)
        for ap in points:
            cs.add({k: ap[k] for k in cfg.fiducialparams})
        cs.prepare(grids=[g.z, nuis.z], scalars=[0.0])  # H, distances, growth of all the job's cosmologies: one device call
        for s, ap in enumerate(points):
            photopars = {k: ap[k] for k in fid_photo}
            if photopars != fid_photo:  # a stencil point of a photo-z parameter: its own n_i(z) windows
                key = tuple(photopars.items())
                if key not in win_index:
                    win_index[key] = len(win_sets)
                    win_sets.append(photopars)
                self.win_of_s[s] = win_index[key]
            cp = {k: ap[k] for k in cfg.fiducialparams}
            c = cs.add(cp)
            self.cosmo_of_s[s] = c
            if "WL" in g.observables:
                key = (c,) + tuple(ap[k] for k in cfg.IAparams)
                if key not in ia_memo:
                    ia_memo[key] = nuis.IA({k: ap[k] for k in cfg.IAparams}, cs.cosmos[c])(g.z)
                self.ia[s] = ia_memo[key]
            if "GCph" in g.observables:
                bkey = tuple(ap[k] for k in cfg.Photobiasparams)
                bp = None
                for a, (o, i) in enumerate(g.rows):
                    if o == "GCph":
                        key = bkey + ((i,) if per_bin_bias else ())
                        if key not in bias_memo:
                            if bp is None:
                                bp = {k: ap[k] for k in cfg.Photobiasparams}
                            bias_memo[key] = nuis.gcph_bias(bp, i)(g.z)
                        self.bias[s, a] = bias_memo[key]
        NC = len(cs)
        self.chi = np.array([c.on_grid("comoving", g.z) for c in cs.cosmos])
        self.hub = np.array([c.on_grid("Hubble", g.z) for c in cs.cosmos])
        self.wlpref = np.array([(3.0 / 2.0) * c.scalar("Hubble", 0.0) ** 2.0 * c.scalar("Omegam_of_z", 0.0)
                                for c in cs.cosmos])
        win = galaxy_photo_dist(fid_photo, configuration)
        self.window = win
        self.ngal = _ngal_rows(win, g)
        if len(win_sets) > 1:
            self.ngal = np.stack([_ngal_rows(galaxy_photo_dist(pp, configuration), g) for pp in win_sets])
        else:
            self.win_of_s = None
        self.stw, self.stden = np.asarray(stw, float), np.asarray(stden, float)
        self.NC = NC
        self._plan = None

    @classmethod
    def from_matrix(cls, configuration, names, matrix, cosmo_set: Optional[CosmoSet] = None):
        """Batch of parameter points given as a matrix [B, len(names)] (likelihood sampling): the per-point
        windows are assembled with array operations instead of one Python object per point.  Unlisted
        parameters stay at their fiducial values.  Same numbers as the point-by-point constructor up to the
        rounding of the IA window's linear interpolation."""
        configuration, cfg = _cfg_of(configuration)
        self = cls.__new__(cls)
        self.configuration, self.cfg = configuration, cfg
        g = self.grid = _photo_grid(configuration)
        cs = self.cosmo_set = cosmo_set if cosmo_set is not None else _cosmo_set_of(cfg)
        nuis = PhotoNuisance(configuration)
        matrix = np.atleast_2d(np.asarray(matrix, dtype=float))
        B, Nz, Nb = matrix.shape[0], g.zsamp, len(g.rows)
        col = {n: i for i, n in enumerate(names)}
        for n in names:
            if n in cfg.photoparams and np.any(matrix[:, col[n]] != float(cfg.photoparams[n])):
                raise NotImplementedError("photo-z parameters cannot be varied per point by the array path")

        def column(key, fid):
            return matrix[:, col[key]] if key in col else np.full(B, float(fid))

        # cosmology of every point -> index into the bank (nearest tabulated cosmology, as the reference does)
        ckeys = [k for k in cfg.fiducialparams if k in col]
        self.cosmo_of_s = np.zeros(B, dtype=np.int32)
        if ckeys:
            sub = matrix[:, [col[k] for k in ckeys]]
            uniq, first, inv = np.unique(sub, axis=0, return_index=True, return_inverse=True)
            idx = np.empty(len(uniq), dtype=np.int32)
            for u in np.argsort(first):  # cosmologies enter the set in the order of their first appearance
                cp = dict(cfg.fiducialparams)
                cp.update({k: float(v) for k, v in zip(ckeys, uniq[u])})
                idx[u] = cs.add(cp)
            self.cosmo_of_s = idx[np.asarray(inv).ravel()]
        # H, distances, growth of all new cosmologies in one device call, evaluated on the grids this job reads
        cs.prepare(grids=[g.z, nuis.z], scalars=[0.0])
        self.zmap = None  # bias sampled on the z grid, or on fewer samples + a z -> sample map (binned models)
        self.bias = None
        self.ia = np.zeros((B, Nz))
        self.ia_model = None
        if "WL" in g.observables:
            ia = cfg.IAparams
            if ia["IA_model"] != "eNLA":
                raise ValueError("only the eNLA intrinsic alignment model exists")
            A, eta, beta = column("AIA", ia["AIA"]), column("etaIA", ia["etaIA"]), column("betaIA", ia["betaIA"])
            piv = nuis.settings["pivot_z_IA"]
            zz = nuis.z
            Om = np.array([c.scalar("Omegam_of_z", 0.0) for c in cs.cosmos])
            growth = np.array([c.on_grid("growth", zz).flatten() for c in cs.cosmos])  # [NC, 50]
            j = np.clip(np.searchsorted(zz, g.z, side="right") - 1, 0, len(zz) - 2)
            t = (g.z - zz[j]) / (zz[j + 1] - zz[j])
            # the window itself -- two powers per (point, redshift), B x Nz doubles -- is evaluated on the device from
            # three numbers per point (cfp_photo_plan_set_ia_enla); `ia_window()` gives the same array on the host
            self.ia = None
            self.ia_model = dict(zdep=(1 + zz) / (1 + piv), lum=nuis._lum_z, growth=growth,
                                 fac=-A * (0.0134 * (1 + piv)) * Om[self.cosmo_of_s], eta=eta, beta=beta, j=j, t=t)
        if "GCph" in g.observables:
            bp = cfg.Photobiasparams
            model = bp["bias_model"]
            gc_rows = [a for a, (o, _) in enumerate(g.rows) if o == "GCph"]
            if model == "binned":
                # piecewise constant in z: one sample per tomographic bin and a z -> bin map, not B*Nb*Nz doubles
                edges = np.asarray(configuration.specs["z_bins_GCph"])
                brang = list(configuration.specs["binrange_GCph"])
                vals = np.stack([column(f"b{k}", bp[f"b{k}"]) for k in brang], axis=1)            # [B, nbins]
                self.zmap = np.clip(np.searchsorted(edges, g.z, side="right") - 1, 0, brang[-1] - 1).astype(np.int32)
                self.bias = np.ascontiguousarray(vals)  # [B, nbins]: every clustering row shares the point's vector
            elif model == "binned_constant":
                self.zmap = np.zeros(Nz, dtype=np.int32)
                self.bias = np.ones((B, Nb, 1))
                for a in gc_rows:
                    i = g.rows[a][1]
                    self.bias[:, a, 0] = column(f"b{i}", bp[f"b{i}"])
            else:
                self.bias = np.ones((B, Nb, Nz))
                pts = [dict(bp, **{k: float(matrix[s, col[k]]) for k in bp if k in col}) for s in range(B)]
                for s, b in enumerate(pts):
                    for a in gc_rows:
                        self.bias[s, a] = nuis.gcph_bias(b, g.rows[a][1])(g.z)
        if self.bias is None:
            self.zmap = np.zeros(Nz, dtype=np.int32)
            self.bias = np.ones((B, Nb, 1))
        self.chi = np.array([c.on_grid("comoving", g.z) for c in cs.cosmos])
        self.hub = np.array([c.on_grid("Hubble", g.z) for c in cs.cosmos])
        self.wlpref = np.array([(3.0 / 2.0) * c.scalar("Hubble", 0.0) ** 2.0 * c.scalar("Omegam_of_z", 0.0)
                                for c in cs.cosmos])
        win_ = galaxy_photo_dist(cfg.photoparams, configuration)
        self.window = win_
        self.ngal = _ngal_rows(win_, g)
        self.stw, self.stden = np.zeros((1, B)), np.ones(1)
        self.NC = len(cs)
        self._plan = None
        return self

    def ia_window(self) -> np.ndarray:
        """Intrinsic-alignment window [S][Nz] of every point on the z grid, on the host (what the device evaluates from
        `ia_model`; nuisance.py:190-230)."""
        if self.ia is not None:
            return self.ia
        m = self.ia_model
        gr = m["growth"][self.cosmo_of_s]
        win = m["fac"][:, None] * m["zdep"][None, :] ** m["eta"][:, None] * ((m["lum"][None, :] ** m["beta"][:, None]) / gr)
        return np.take(win, m["j"], axis=1) * (1.0 - m["t"])[None, :] + np.take(win, m["j"] + 1, axis=1) * m["t"][None, :]

    def dense_bias(self) -> np.ndarray:
        """Galaxy bias [S][Nb][Nz] on the z grid, whatever compact form the job holds it in (per-bin samples and a
        z -> bin map, per point or per (point, row)); lensing rows carry 1."""
        b = self.bias
        zmap = getattr(self, "zmap", None)
        if zmap is None:
            return b
        if b.ndim == 2:  # one vector per point, shared by the clustering rows
            out = np.ones((b.shape[0], len(self.grid.rows), len(zmap)))
            gc = [a for a, (o, _) in enumerate(self.grid.rows) if o == "GCph"]
            out[:, gc, :] = b[:, zmap][:, None, :]
            return out
        return b[:, :, zmap]

    def plan(self) -> _lib.PhotoPlan:
        if self._plan is None:
            g, cs = self.grid, self.cosmo_set
            self._plan = _lib.PhotoPlan(
                cs.context(), cs.bank(), cosmo_of_s=self.cosmo_of_s, ell=g.ell, z=g.z, dz=g.dz,
                chi=self.chi, hub=self.hub, wlpref=self.wlpref, kind=g.kind, ngal=self.ngal, bias=self.bias,
                ia=self.ia, ia_model=getattr(self, "ia_model", None), lmin=g.lmin, lmax=g.lmax, noise=g.noise,
                sqrtfsky=g.sqrtfsky, stw=self.stw,
                stden=self.stden, kmax=g.kmax, tab_p=g.tab_p, zmap=getattr(self, "zmap", None),
                win_of_s=getattr(self, "win_of_s", None), tab_pg=g.tab_pg)
        return self._plan

    def run(self):
        p = self.plan()
        p.run()
        return p

    def close(self):
        if self._plan is not None:
            self._plan.close()
            self._plan = None


def _cosmo_set_of(cfg) -> CosmoSet:
    cs = cfg.__dict__.get("_cosmo_set")
    if cs is None:
        nl = cfg.settings["nonlinear"]
        slots = [_lib.TAB_PNL if nl else _lib.TAB_PLIN]
        if cfg.settings["GCph_Tracer"] == "clustering":
            slots.append(_lib.TAB_PNL_CB if nl else _lib.TAB_PLIN_CB)
        cs = CosmoSet(cfg, slots=tuple(slots))
        cfg.__dict__["_cosmo_set"] = cs
    return cs


def _cls_dict(grid: _PhotoGrid, cl: np.ndarray) -> dict:
    """[Nl][Nb][Nb] -> the reference's dictionary {"ells": ..., "WL 1xGCph 3": (Nl,), ...}."""
    out = {"ells": grid.ell}
    for a, ca in enumerate(grid.cols):
        for b, cb in enumerate(grid.cols):
            out[ca + "x" + cb] = np.ascontiguousarray(cl[:, a, b])
    return out


class ComputeCls:
    """Angular power spectra of one parameter point (reference: photo_obs.py:163-928)."""

    def __init__(self, cosmopars, photopars, IApars, biaspars, print_info_specs=False, fiducial_cosmo=None,
                 configuration=None):
        configuration, cfg = _cfg_of(configuration)
        self.configuration, self._cfg = configuration, cfg
        self.feed_lvl = configuration.settings["feedback"]
        self.cosmopars, self.photopars, self.IApars, self.biaspars = cosmopars, photopars, IApars, biaspars
        self.cosmo_set = _cosmo_set_of(cfg)
        self.cosmo = self.cosmo_set.cosmos[self.cosmo_set.add(cosmopars)] if fiducial_cosmo is None else fiducial_cosmo
        g = self.grid = _photo_grid(configuration)
        self.observables, self.binrange = g.observables, g.binrange
        self.binrange_WL, self.binrange_GCph = g.binrange["WL"], g.binrange["GCph"]
        configuration.specs["ellmax"], configuration.specs["ellmin"] = g.ellmax, g.ellmin
        self.tracer = configuration.settings["GCph_Tracer"]
        self.zsamp, self.ellsamp, self.ell = g.zsamp, g.ellsamp, g.ell
        self.z_min, self.z_max, self.z, self.dz = g.z_min, g.z_max, g.z, g.dz
        self.window = galaxy_photo_dist(photopars, configuration)
        self.result = None
        if print_info_specs and self.feed_lvl >= 2:
            print(f"*** ell in [{g.ellmin}, {g.ellmax}] x{g.ellsamp}, z in [{g.z_min}, {g.z_max}] x{g.zsamp} ***")

    def compute_all(self):
        allpars = {**self.cosmopars, **self.IApars, **self.biaspars, **self.photopars}
        job = PhotoJob(self.configuration, [allpars], np.zeros((1, 1)), np.ones(1), self.cosmo_set)
        plan = job.run()
        _, cl, T = plan.fetch(cls=True, sqrtp=True)
        job.close()
        self.result = _cls_dict(self.grid, cl[0])
        sq = np.ascontiguousarray(T[job.cosmo_of_s[0]].T)  # (Nz, Nl) like the reference's sqrtPell tables
        self.sqrtPell = {"WL": sq, "WL_IA": sq, "GCph": sq, "GCph_IA": np.zeros_like(sq)}
        return None

    computecls = computecls_vectorized = lambda self: (self.compute_all(), self.result)[1]

    # --- radial kernels and single spectra (photo_obs.py:514-615, 878-928), from the device run ------------------
    def _windows(self):
        if getattr(self, "_win", None) is None:
            allpars = {**self.cosmopars, **self.IApars, **self.biaspars, **self.photopars}
            job = PhotoJob(self.configuration, [allpars], np.zeros((1, 1)), np.ones(1), self.cosmo_set)
            plan = job.run()
            U, V = plan.fetch_windows()
            sqrtH = np.sqrt(self.cosmo_set.cosmos[job.cosmo_of_s[0]].on_grid("Hubble", self.grid.z))
            bias = job.bias[0] if getattr(job, "zmap", None) is None else job.bias[0][:, job.zmap]
            job.close()
            self._win = (U[0] * sqrtH[None, :], V[0] * sqrtH[None, :], bias)
        return self._win

    def _row(self, obs, i):
        try:
            return self.grid.rows.index((obs, int(i)))
        except ValueError:
            raise ValueError(f"{obs} bin {i} is not part of this job") from None

    def _at(self, z, values):
        """Values on the job's z grid; other redshifts through a cubic interpolant of the grid values (the device
        integrates on the grid; the reference evaluates its closed forms wherever asked)."""
        z = np.asarray(z, dtype=float)
        if z.shape == self.z.shape and np.array_equal(z, self.z):
            return values.copy()
        from scipy.interpolate import interp1d

        return interp1d(self.z, values, kind="cubic")(z)

    def lensing_kernel(self, z, i):
        """(W_i^WL, W_i^IA)(z): lensing efficiency term and intrinsic-alignment term of bin i (photo_obs.py:558-615)."""
        U, V, _ = self._windows()
        a = self._row("WL", i)
        return self._at(z, U[a]), self._at(z, V[a])

    def galaxy_kernel(self, z, i):
        """n_i(z) H(z) of clustering bin i -- without the galaxy bias, as in the reference (photo_obs.py:514-556)."""
        U, _, bias = self._windows()
        a = self._row("GCph", i)
        return self._at(z, U[a] / bias[a])

    def clsintegral(self, obs1, bin1, obs2, bin2):
        """C_ell of one pair of tomographic rows (photo_obs.py:878-928) from the device's Limber contraction."""
        if self.result is None:
            self.compute_all()
        return self.result[f"{obs1} {bin1}x{obs2} {bin2}"]


class PhotoCov:
    """C_ell cache, noise, covariance and the Fisher of the photometric probes
    (reference: photo_cov.py:24-381 + cosmicfish.py:643-744, fused on the device)."""

    def __init__(self, cosmopars, photopars, IApars, biaspars, fiducial_Cls=None, enable_cache=True, cache_size=32,
                 configuration=None, stencil="3PT"):
        configuration, cfg = _cfg_of(configuration)
        self.configuration, self._cfg = configuration, cfg
        self.cosmopars, self.photopars, self.IApars, self.biaspars = cosmopars, photopars, IApars, biaspars
        self.allparsfid = {**cosmopars, **IApars, **biaspars, **photopars}
        self.fiducial_Cls = fiducial_Cls
        g = self.grid = _photo_grid(configuration)
        sp = configuration.specs
        self.observables, self.binrange = g.observables, g.binrange
        self.binrange_WL, self.binrange_GCph = sp["binrange_WL"], sp["binrange_GCph"]
        self.feed_lvl = configuration.settings["feedback"]
        self.fsky_WL, self.fsky_GCph = sp.get("fsky_WL"), sp.get("fsky_GCph")
        self.ngalbin_WL, self.ngalbin_GCph = np.array(sp["ngalbin_WL"]), np.array(sp["ngalbin_GCph"])
        self.numbins_WL, self.numbins_GCph = len(sp["z_bins_WL"]) - 1, len(sp["z_bins_GCph"]) - 1
        self.ellipt_error = sp["ellipt_error"]
        self.stencil = stencil
        self.timings = {}
        self._job = None
        self._cls_all = None
        self._cache = {}

    # --- single-point interface ------------------------------------------------------------------
    def getcls(self, allpars):
        """{"ells", "X ixY j"} at an arbitrary parameter point (photo_cov.py:116-162)."""
        key = tuple(sorted((k, v) for k, v in allpars.items() if not isinstance(v, (dict, list))))
        if key not in self._cache:
            job = PhotoJob(self.configuration, [dict(allpars)], np.zeros((1, 1)), np.ones(1))
            plan = job.run()
            _, cl, _ = plan.fetch(cls=True)
            job.close()
            self._cache[key] = _cls_dict(self.grid, cl[0])
        return self._cache[key]

    def getclsnoise(self, cls):
        noisy = dict(cls)
        for a, (o, i) in enumerate(self.grid.rows):
            k = f"{o} {i}x{o} {i}"
            noisy[k] = noisy[k] + self.grid.noise[a]
        return noisy

    def get_covmat(self, noisy_cls):
        """Per-multipole covariance (C + N)/sqrt(f_sky) as the reference returns it: a list of Nl pandas DataFrames
        indexed by the tomographic rows (photo_cov.py:198-245).  `np.array(covmat)` is the [Nl][Nb][Nb] array."""
        import pandas as pd

        cols = self.grid.cols
        self.cov_cols = cols
        cov = np.empty((len(noisy_cls["ells"]), len(cols), len(cols)))
        for a, ca in enumerate(cols):
            for b, cb in enumerate(cols):
                cov[:, a, b] = noisy_cls[ca + "x" + cb] / self.grid.sqrtfsky[a]
        return [pd.DataFrame(m, index=cols, columns=cols) for m in cov]

    def __getattr__(self, name):
        # the reference's FisherMatrix.compute() leaves cls / noisy_cls / covmat / derivs on photo_LSS; here they are
        # produced on first access (the Fisher matrix itself never needs them on the host)
        if name in ("cls", "noisy_cls", "covmat") and "_job" in self.__dict__:
            self.compute_covmat()
            return self.__dict__[name]
        if name == "derivs" and "_job" in self.__dict__:
            return self.compute_derivs()
        raise AttributeError(name)

    # --- the whole job ---------------------------------------------------------------------------
    def _freeparams(self, freeparams=None):
        fp = dict(self._cfg.freeparams if freeparams is None else freeparams)
        missing = [k for k in fp if k not in self.allparsfid]
        for k in missing:  # photo_cov.py:270-293
            if k in self._cfg.fiducialparams:
                self.allparsfid[k] = self._cfg.fiducialparams[k]
            elif k == "Omegab":
                self.allparsfid[k] = 0.05
            else:
                raise ValueError("Fiducial values missing for free parameters: " + k)
        return fp

    @property
    def points(self):
        """Flat {name: number} dictionary of every stencil point of the job (index 0: the fiducial)."""
        if self._points is None:
            pts = []
            for st in self._steps:
                ap = dict(self.allparsfid)
                if st is not None:
                    ap[st[0]] = ap[st[0]] + st[1]
                pts.append(ap)
            self._points = pts
        return self._points

    def build_job(self, freeparams=None) -> PhotoJob:
        fp = self._freeparams(freeparams)
        ext_eps = (self._cfg.external or {}).get("eps_values") if self.stencil == "STEM" else None
        # stencil points as (parameter, offset) steps off the fiducial; the dictionaries of `self.points` are formed on
        # demand (the job itself works on the matrix of parameter values)
        steps = [None]
        rows, stden = [], np.ones(len(fp))
        self._offsets = []
        for ip, par in enumerate(fp):
            offs, stden[ip] = stencils.offsets_and_weights(self.stencil, self.allparsfid[par], fp[par], ext_eps)
            row, used = {}, []
            for off, w in offs:
                if off == 0:
                    row[0] = row.get(0, 0.0) + w
                    used.append((0, off))
                    continue
                steps.append((par, off))
                row[len(steps) - 1] = w
                used.append((len(steps) - 1, off))
            rows.append(row)
            self._offsets.append(used)
        stw = np.zeros((len(fp), len(steps)))
        for ip, row in enumerate(rows):
            for s, w in row.items():
                stw[ip, s] = w
        self.parnames = list(fp)
        self._steps, self._points = steps, None
        cfg = self._cfg
        if self.stencil != "STEM" and not any(k in cfg.photoparams for k in fp):
            # stencil points as rows of a matrix: windows, IA and bias assembled with array operations (the photo-z
            # parameters, whose stencil points need their own n_i(z) windows, take the point-by-point constructor)
            names = [k for k, v in self.allparsfid.items() if isinstance(v, (int, float, np.floating, np.integer))
                     and not isinstance(v, bool)]
            col = {n: i for i, n in enumerate(names)}
            mat = np.tile(np.array([self.allparsfid[k] for k in names], dtype=float), (len(steps), 1))
            for s, st in enumerate(steps):
                if st is not None:
                    mat[s, col[st[0]]] = self.allparsfid[st[0]] + st[1]
            self._job = PhotoJob.from_matrix(self.configuration, names, mat)
            self._job.stw, self._job.stden = np.asarray(stw, float), np.asarray(stden, float)
        else:
            self._job = PhotoJob(self.configuration, self.points, stw, stden)
        self._cls_all = None
        return self._job

    def compute_fisher(self, freeparams=None, shard=(0, 1)) -> np.ndarray:
        """Fisher matrix of the job; `shard=(rank, world)` restricts this process to its slice of the
        multipole bins and all-reduces the partial sums (cosmicfishpie_b200.distributed)."""
        return self.finish_fisher(self.launch_fisher(freeparams, shard))

    def launch_fisher(self, freeparams=None, shard=(0, 1)):
        """Host preparation, H2D and the asynchronous kernel launches; `finish_fisher` collects the result."""
        from . import distributed as cdist

        t0 = time()
        job = self.build_job(freeparams)
        plan = job.plan()
        t1 = time()
        rank, world = shard
        nbins = plan.Nl - 1
        plan.set_slice(*cdist.split_range(nbins, rank, world))
        plan.run()
        return plan, world, nbins, t0, t1

    def finish_fisher(self, launched) -> np.ndarray:
        from . import distributed as cdist

        plan, world, nbins, t0, t1 = launched
        reduced = world > 1 and cdist.allreduce_plan_result(plan, (plan.npar, plan.npar))  # NCCL, in place on the device
        F, _, _ = plan.fetch()
        if world > 1:
            if not reduced:
                F = cdist.allreduce_numpy(F, device=getattr(getattr(plan, "ctx", None), "device", None))
            plan.set_slice(0, nbins)
            plan.run()  # every rank keeps the full C_ell set for compute_covmat()/compute_derivs()
        if self.stencil == "STEM":
            self._check_stem_residuals()
        t2 = time()
        self.timings = {"host_prep_s": t1 - t0, "device_s": t2 - t1}
        self.fisher = F
        return F

    def _check_stem_residuals(self):
        """derivatives.py:402-419 trims the outermost points while the straight-line fit of any C_ell leaves a sum of
        squared residuals above an ABSOLUTE 1e-3.  The device stencil uses all points: verify that the reference
        would have, too (spectra <= 1e-4 cannot fail this), and refuse loudly otherwise."""
        cl = self._all_cls()
        for par, used in zip(self.parnames, self._offsets):
            x = np.array([off for _, off in used])
            y = np.array([cl[s] for s, _ in used])
            w, den = stencils.slope_weights(x)
            slope = np.tensordot(w, y, axes=(0, 0)) / den
            fit = np.mean(y, axis=0) + (x - np.mean(x)).reshape((-1,) + (1,) * (y.ndim - 1)) * slope
            if np.max(np.sum((y - fit) ** 2, axis=0)) > stencils.STEM_THRESHOLD:
                raise NotImplementedError(f"SteM derivative of {par}: the fit residual exceeds {stencils.STEM_THRESHOLD}; "
                                          "the reference would trim the stencil per multipole, which is not supported")

    def _all_cls(self):
        if self._job is None:
            self.build_job()
            self._job.run()
        if self._cls_all is None:
            _, self._cls_all, _ = self._job.plan().fetch(cls=True)
        return self._cls_all

    def compute_covmat(self):
        cl = self._all_cls()[0]
        self.cls = _cls_dict(self.grid, cl)
        self.noisy_cls = self.getclsnoise(self.cls)
        self.covmat = self.get_covmat(self.noisy_cls)
        return self.noisy_cls, self.covmat

    def compute_derivs(self, print_info_specs=False):
        """{par: {"ells": ..., "X ixY j": dC/dpar (Nl,)}} from the device C_ell of every stencil point."""
        cl = self._all_cls()
        job = self._job
        derivs = {}
        for ip, par in enumerate(self.parnames):
            acc = np.zeros_like(cl[0])
            for s in np.nonzero(job.stw[ip])[0]:
                acc = acc + job.stw[ip, s] * cl[s]
            derivs[par] = _cls_dict(self.grid, acc / job.stden[ip])
        self.derivs = derivs
        return derivs
