"""Host-side cosmology input: cached Boltzmann tables -> spline facade -> device table bank.

This is the INPUT BOUNDARY of the hot path.  Boltzmann output (P(k,z), H(z),
growth ...) is computed once elsewhere, cached as tables, and consumed identically by the reference
and by this package.  Here the tables are turned into the same FITPACK splines the reference builds
(`external_input.calculate_interpol_results`, cosmicfishpie/cosmology/cosmology.py:1423-1540) so that
the knots/coefficients shipped to the GPU are *the* piecewise polynomials the reference evaluates.

Two sources are accepted for `extfiles`:
  * the reference's directory layout (cosmology.py:1249-1373), `extfiles["directory"]`
  * an in-memory bank `extfiles["tables"] = {folder_name: {z,k,H,s8,D,f,Pl,Pnl}}`
    (what `cosmicfishpie_b200.toy_tables.make_bank` returns) -- same numbers, no text parsing.

Only `code="external"` exists here: the Boltzmann solvers themselves are out of scope.
"""
from __future__ import annotations

import glob
import math
import os
from typing import Dict, Optional

import numpy as np
from scipy.integrate import trapezoid
from scipy.interpolate import InterpolatedUnivariateSpline, RectBivariateSpline
from scipy.signal import savgol_filter

from . import _lib

C_KMS = 299792.458  # scipy.constants.speed_of_light / 1000

DEFAULT_FILE_NAMES = {  # config.py:289-304
    "H_z": "background_Hz",
    "D_zk": "D_Growth-zk",
    "f_zk": "f_GrowthRate-zk",
    "Pl_zk": "Plin-zk",
    "Pnl_zk": "Pnonlin-zk",
    "s8_z": "sigma8-z",
    "z_arr": "z_values_list",
    "k_arr": "k_values_list",
    "fcb_zk": "fcb_GrowthRate-zk",
    "Plcb_zk": "Plincb-zk",
    "Pnlcb_zk": "Pnonlincb-zk",
    "s8cb_z": "sigma8cb-z",
}
_KEYMAP = {"z": "z_arr", "k": "k_arr", "H": "H_z", "s8": "s8_z", "D": "D_zk", "f": "f_zk", "Pl": "Pl_zk", "Pnl": "Pnl_zk"}
# cold dark matter + baryon tables, read when GCsp_Tracer / GCph_Tracer is "clustering" (cosmology.py:1240-1245)
_KEYMAP_CB = {"s8cb": "s8cb_z", "fcb": "fcb_zk", "Plcb": "Plcb_zk", "Pnlcb": "Pnlcb_zk"}


def round_decimals_up(number: float) -> float:
    """One-decimal mantissa ceiling (utilities/utils.py numerics.round_decimals_up)."""
    exponent = math.floor(math.log10(number))
    return math.ceil(number / (10**exponent) * 10) / 10 * (10**exponent)


def folder_for(cosmopars, fiducial, external) -> str:
    """Which table folder serves `cosmopars`: snap (theta/theta_fid - 1) to the nearest tabulated
    epsilon of the first parameter that differs from the fiducial (cosmology.py:1375-1421)."""
    names = external["paramnames"]
    fnames = external.get("folder_paramnames") or names
    eps = np.array(external["eps_values"], dtype=float)
    allowed = np.concatenate((eps, -eps, np.array([0.0])))
    for par, fpar in zip(names, fnames):
        if np.isclose(cosmopars[par], fiducial[par], rtol=1e-5):
            continue
        if fiducial[par] != 0:
            d = cosmopars[par] / fiducial[par] - 1.0
        else:
            d = cosmopars[par] - fiducial[par]
        tag = "pl" if d > 0 else "mn"
        d = allowed[np.abs(allowed - d).argmin()]
        tok = "{:.1E}".format(round_decimals_up(abs(d))).replace(".", "p")
        if external.get("E-00", True) is False:
            tok = tok.replace("E-0", "E-")
        return f"{fpar}_{tag}_eps_{tok}"
    return external.get("fiducial_folder", "fiducial_eps_0")


_TABLE_CACHE: Dict[tuple, dict] = {}


def load_tables(external: dict, folder: str, cb: bool = False) -> dict:
    """Raw arrays of one folder, cached per (source, folder); `cb` adds the cdm+baryon tables."""
    if "tables" in external and external["tables"] is not None:
        t = external["tables"][folder]
        if cb:
            missing = [k for k in _KEYMAP_CB if k not in t]
            if missing:
                raise FileNotFoundError(f"tracer 'clustering' needs the cb tables {missing} in folder {folder}")
        return t
    key = (os.path.abspath(external["directory"]), folder, bool(cb))
    if key in _TABLE_CACHE:
        return _TABLE_CACHE[key]
    names = dict(DEFAULT_FILE_NAMES)
    names.update({k: v for k, v in (external.get("file_names") or {}).items() if v is not None})
    out = {}
    for short, long in {**_KEYMAP, **(_KEYMAP_CB if cb else {})}.items():
        path = None
        for base in (folder, external.get("fiducial_folder", "fiducial_eps_0")):  # fiducial fallback for grids/H
            hits = glob.glob(os.path.join(external["directory"], base, names[long] + ".*"))
            if hits:
                path = hits[0]
                break
        if path is None:
            raise FileNotFoundError(f"{names[long]}.* not found under {external['directory']}/{folder}")
        out[short] = np.loadtxt(path)
    _TABLE_CACHE[key] = out
    return out


def _dcom_all(zgrid, hfunc):
    """Comoving distance at every table redshift by the reference's 100-point trapezoid of 1/H
    (cosmology.py:36-55, 1467-1469), all redshifts in one spline call (bit-identical to the per-z loop)."""
    zt = np.array([np.linspace(0.0, zi, 100) for zi in zgrid])
    return trapezoid(1 / hfunc(zt.ravel()).reshape(zt.shape), zt, axis=1)


class _LazySpline:
    """RectBivariateSpline built on first use (a GCsp job never touches P_nl, a photo job never f)."""

    def __init__(self, zg, kg, table):
        self._args = (zg, kg, table)
        self._spl = None

    def get(self):
        if self._spl is None:
            zg, kg, tab = self._args
            self._spl = RectBivariateSpline(zg, kg, tab, kx=3, ky=3)
            self._args = None
        return self._spl


class cosmo_functions:
    """Spline facade of one cosmology (reference: cosmology.py:1546-2010, external branch).

    Host accessors keep the reference names (`Hubble`, `angdist`, `comoving`, `matpow`, `f_growthrate`,
    `growth`, `sigma8_of_z`, `Omegam_of_z`, `nonwiggle_pow`, `E_hubble`).  `bank_row()` exports the
    bicubic tck of the (z,k) tables for the device bank.
    """

    c = C_KMS

    def __init__(self, cosmopars, input=None, config=None):
        if config is None:
            from . import config as config_module

            config = config_module.current()
        self.config = config
        self.settings = config.settings
        self.input = input if input is not None else config.input_type
        if self.input != "external":
            raise NotImplementedError(
                f"cosmicfishpie_b200 consumes cached Boltzmann tables only (code='external'); got {self.input!r}"
            )
        self.code = "external"
        self.external = config.external
        self.cosmopars = cosmopars
        self.fiducialcosmopars = config.fiducialparams
        self.cb_files_on = "clustering" in (self.settings.get("GCsp_Tracer"), self.settings.get("GCph_Tracer"))
        emulator = self.external.get("emulator")
        if emulator is not None and not emulator.is_node(cosmopars):
            # a cosmology between the tabulated ones: its tables from the bank's emulator (cosmicfishpie_b200/bank.py)
            # instead of the reference's snap to the nearest folder (cosmology.py:1375-1421)
            t = emulator.tables_at(cosmopars)
            self.folder = "<emulated>"
        else:
            self.folder = folder_for(cosmopars, self.fiducialcosmopars, self.external)
            t = load_tables(self.external, self.folder, cb=self.cb_files_on)
        pkf = rf = kf = 1.0
        if self.external.get("k-units") == "h/Mpc":  # cosmology.py:1438-1444
            pkf = (1 / cosmopars["h"]) ** 3
            kf = cosmopars["h"]
        ru = self.external.get("r-units")
        if ru == "Mpc/h":
            rf = 1 / cosmopars["h"]
        elif ru == "km/s/Mpc":
            rf = C_KMS
        zg = np.asarray(t["z"], float).flatten()
        kg = kf * np.asarray(t["k"], float).flatten()
        self.zgrid, self.kgrid = zg, kg
        self.h_of_z = InterpolatedUnivariateSpline(zg, (1 / rf) * np.asarray(t["H"], float).flatten())
        dcom = _dcom_all(zg, self.h_of_z)
        self.com_dist = InterpolatedUnivariateSpline(zg, dcom)
        self.ang_dist = InterpolatedUnivariateSpline(zg, dcom / (1 + zg))
        self._lazy = {
            "D_growth_zk": _LazySpline(zg, kg, t["D"]),
            "f_growthrate_zk": _LazySpline(zg, kg, t["f"]),
            "Pk_l": _LazySpline(zg, kg, pkf * np.asarray(t["Pl"], float)),
            "Pk_nl": _LazySpline(zg, kg, pkf * np.asarray(t["Pnl"], float)),
        }
        self.s8_of_z = InterpolatedUnivariateSpline(zg, np.asarray(t["s8"], float).flatten())
        if self.cb_files_on:  # cosmology.py:1508-1531
            self._lazy.update({
                "f_growthrate_cb_zk": _LazySpline(zg, kg, t["fcb"]),
                "Pk_cb_l": _LazySpline(zg, kg, pkf * np.asarray(t["Plcb"], float)),
                "Pk_cb_nl": _LazySpline(zg, kg, pkf * np.asarray(t["Pnlcb"], float)),
            })
            self.s8_cb_of_z = InterpolatedUnivariateSpline(zg, np.asarray(t["s8cb"], float).flatten())
        self.results = self  # reference code reaches splines through `.results`
        self._nw_cache = {}

    def __getattr__(self, name):  # the four (z,k) splines, fitted on first access
        lazy = self.__dict__.get("_lazy")
        if lazy is not None and name in lazy:
            return lazy[name].get()
        raise AttributeError(name)

    # --- host accessors --------------------------------------------------------------------
    def scalar(self, name, z):
        """`getattr(self, name)(z)` for a scalar redshift, remembered: jobs ask for the same few background
        values at the same bin centres again and again."""
        memo = self.__dict__.setdefault("_scalar_memo", {})
        key = (name, float(z))
        if key not in memo:
            memo[key] = getattr(self, name)(z)
        return memo[key]

    def on_grid(self, name, z):
        """`getattr(self, name)(z)` on a redshift grid, remembered per grid (read-only result)."""
        memo = self.__dict__.setdefault("_grid_memo", {})
        z = np.ascontiguousarray(z, dtype=float)
        key = (name, z.tobytes())
        if key not in memo:
            v = np.asarray(getattr(self, name)(z))
            v.setflags(write=False)
            memo[key] = v
        return memo[key]

    def Hubble(self, z, physical=False):
        return (self.c if physical else 1) * self.h_of_z(z)

    def E_hubble(self, z):
        return self.Hubble(z) / self.Hubble(0.0)

    def angdist(self, z):
        return self.ang_dist(z)

    def comoving(self, z):
        return self.com_dist(z)

    def matpow(self, z, k, nonlinear=False, tracer="matter"):
        if tracer == "clustering":  # cosmology.py:1715-1716, 1741-1758
            return (self.Pk_cb_nl if nonlinear else self.Pk_cb_l)(z, k, grid=False)
        return (self.Pk_nl if nonlinear else self.Pk_l)(z, k, grid=False)

    def f_growthrate(self, z, k=None, gamma=False, tracer="matter"):
        if tracer == "clustering":  # cosmology.py:1937-1939
            return self.f_growthrate_cb_zk(z, 0.0001 if k is None else k, grid=False)
        return self.f_growthrate_zk(z, 0.0001 if k is None else k, grid=False)

    def growth(self, z, k=None):
        return self.D_growth_zk(z, 0.0001 if k is None else k, grid=False)

    def sigma8_of_z(self, z, tracer="matter"):
        return self.s8_cb_of_z(z) if tracer == "clustering" else self.s8_of_z(z)  # cosmology.py:1851-1857

    def sigma8_cb_of_z(self, z):
        return self.s8_cb_of_z(z)

    def fsigma8_of_z(self, z, k=None, gamma=False, tracer="matter"):
        return self.f_growthrate(z, k, gamma, tracer=tracer) * self.sigma8_of_z(z, tracer=tracer)

    def Omegam_of_z(self, z):
        return (self.cosmopars["Omegam"] * (1 + z) ** 3) / self.E_hubble(z) ** 2

    def SigmaMG(self, z, k):
        return np.array(1)

    def nonwiggle_table(self, z, nonlinear=False, tracer="matter"):
        """Knots and coefficients of the degree-1 no-wiggle spline at redshift z
        (Savitzky-Golay smoothing in ln k, cosmology.py:1760-1813)."""
        key = (float(z), bool(nonlinear), tracer == "clustering")
        if key not in self._nw_cache:
            s = self.settings
            h = self.cosmopars["h"]
            lk = np.linspace(np.log(h * s["savgol_internalkmin"]), np.log(h * np.max(self.kgrid)),
                             s["savgol_internalsamples"])
            dlnk = np.mean(lk[1:] - lk[0:-1])
            n_sg = int(np.round(s["savgol_width"] / np.log(1 + dlnk)))
            lp = InterpolatedUnivariateSpline(
                lk, np.log(self.matpow(z, np.exp(lk), nonlinear=nonlinear, tracer=tracer).flatten()), k=1)
            sg = savgol_filter(lp(lk), n_sg, s["savgol_polyorder"])
            spl = InterpolatedUnivariateSpline(np.exp(lk), np.exp(sg), k=1)
            t, c, _ = spl._eval_args
            n = len(t) - 2
            self._nw_cache[key] = (np.ascontiguousarray(t[1:-1]), np.ascontiguousarray(c[:n]), spl)
        return self._nw_cache[key]

    def nonwiggle_pow(self, z, k, nonlinear=False, tracer="matter"):
        return self.nonwiggle_table(z, nonlinear, tracer)[2](k)

    # --- device export -----------------------------------------------------------------------
    def bank_row(self, slots=None):
        """Bicubic (tx, ty, c) per table slot of include/cfp_b200.h; `slots` limits which tables are fitted."""
        names = {_lib.TAB_PLIN: "Pk_l", _lib.TAB_PNL: "Pk_nl", _lib.TAB_F: "f_growthrate_zk", _lib.TAB_D: "D_growth_zk",
                 _lib.TAB_PLIN_CB: "Pk_cb_l", _lib.TAB_PNL_CB: "Pk_cb_nl", _lib.TAB_F_CB: "f_growthrate_cb_zk"}
        if not self.cb_files_on:
            names = {k: v for k, v in names.items() if k < _lib.TAB_PLIN_CB}
        row = [None] * _lib.NTAB
        for slot, name in names.items():
            if slots is None or slot in slots:
                row[slot] = getattr(self, name).tck
        return row

    def packed_row(self, ctx, slots=None) -> "_lib.BankRow":
        """`bank_row` packed once into page-locked memory (per device context and table selection)."""
        cache = self.__dict__.setdefault("_packed_rows", {})
        key = (id(ctx), None if slots is None else tuple(slots))
        hit = cache.get(key)
        if hit is None or hit[0] is not ctx:
            hit = (ctx, _lib.BankRow(ctx, self.bank_row(slots)))
            cache[key] = hit
        return hit[1]


_COSMO_CACHE: Dict[tuple, tuple] = {}
_COSMO_CACHE_MAX = 256  # fitted cosmologies kept (oldest evicted first)


def clear_caches():
    """Forget parsed tables and fitted cosmologies (bench.py uses this to time cold end-to-end steps)."""
    _TABLE_CACHE.clear()
    _COSMO_CACHE.clear()


def cached_cosmo(cosmopars, config) -> "cosmo_functions":
    """Boltzmann input is computed once and cached (north_star): one spline facade per (table source, folder,
    parameter values, units, no-wiggle settings), shared by every FisherMatrix / likelihood of the process."""
    ext, st = config.external, config.settings
    src = id(ext["tables"]) if ext.get("tables") is not None else os.path.abspath(ext["directory"])
    key = (src, tuple(sorted((k, v) for k, v in cosmopars.items())), tuple(sorted(config.fiducialparams.items())),
           tuple(ext["paramnames"]), tuple(ext["eps_values"]), ext.get("k-units"), ext.get("r-units"),
           st["savgol_internalkmin"], st["savgol_internalsamples"], st["savgol_width"], st["savgol_polyorder"],
           st.get("GCsp_Tracer"), st.get("GCph_Tracer"), id(ext.get("emulator")))
    hit = _COSMO_CACHE.get(key)
    if hit is not None and hit[0] is ext.get("tables"):  # identity check: id() values can be recycled
        return hit[1]
    c = cosmo_functions(dict(cosmopars), config.input_type, config)
    _COSMO_CACHE[key] = (ext.get("tables"), c)
    while len(_COSMO_CACHE) > _COSMO_CACHE_MAX:  # a sampler that moves in cosmology would grow this without bound
        _COSMO_CACHE.pop(next(iter(_COSMO_CACHE)))
    return c


class CosmoSet:
    """The distinct cosmologies of one job (fiducial + stencil points), their host facades and the
    device bank.  Index 0 is always the fiducial."""

    def __init__(self, config, ctx: Optional[_lib.Context] = None, slots=None):
        self.config = config
        self.ctx = ctx
        self.slots = slots  # table slots the device bank needs (None = all four)
        self._by_key = {}
        self.cosmos = []
        self.add(config.fiducialparams, config.fiducialcosmo)
        self._bank = None

    @staticmethod
    def _key(cosmopars):
        return tuple(sorted((k, v) for k, v in cosmopars.items()))

    def add(self, cosmopars, cosmo=None) -> int:
        key = self._key(cosmopars)
        if key not in self._by_key:
            if cosmo is None:
                cosmo = cached_cosmo(cosmopars, self.config)
            self._by_key[key] = len(self.cosmos)
            self.cosmos.append(cosmo)
            self._bank = None
        return self._by_key[key]

    def __len__(self):
        return len(self.cosmos)

    def begin_batch(self, limit: int = 64):
        """Called by the likelihoods before a batch indexes into the set: once more than `limit` cosmologies have
        accumulated (a sampler that varies cosmological parameters adds one per distinct point) the set is cut back
        to the fiducial and the device bank released, so host and device memory stay bounded over a long run."""
        if len(self.cosmos) > limit:
            if self._bank is not None:
                self._bank.close()
                self._bank = None
            self.cosmos = self.cosmos[:1]
            self._by_key = {k: v for k, v in self._by_key.items() if v == 0}

    def context(self) -> _lib.Context:
        """The device context every bank, plan and likelihood helper of this job uses: options["device"]
        (default: LOCAL_RANK or 0), created on first use."""
        if self.ctx is None:
            self.ctx = _lib.default_context((getattr(self.config, "settings", None) or {}).get("device"))
        return self.ctx

    def bank(self) -> _lib.Bank:
        if self._bank is None:
            ctx = self.context()
            self._bank = _lib.Bank.from_rows(ctx, [c.packed_row(ctx, self.slots) for c in self.cosmos])
        return self._bank
