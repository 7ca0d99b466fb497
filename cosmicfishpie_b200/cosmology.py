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
    """RectBivariateSpline built on first use (a GCsp job never touches P_nl, a photo job never f); the table is read
    from the cosmology's table mapping at that moment."""

    def __init__(self, zg, kg, tables, name, scale=1.0):
        self._args = (zg, kg, tables, name, scale)
        self._spl = None

    def get(self):
        if self._spl is None:
            zg, kg, tables, name, scale = self._args
            tab = np.asarray(tables[name], float)
            self._spl = RectBivariateSpline(zg, kg, tab if scale == 1.0 else scale * tab, kx=3, ky=3)
            self._args = None
        return self._spl


def device_prep(settings) -> bool:
    """Spline fits, comoving distances, no-wiggle spectra and velocity moments of a job's cosmologies on the device
    (csrc/cfp_prep.cu) -- the default; `options["device_prep"] = False` or CFP_HOST_PREP=1 keep them on the host
    (scipy, as the reference does them)."""
    if os.environ.get("CFP_HOST_PREP", "0") not in ("", "0"):
        return False
    return bool((settings or {}).get("device_prep", True))


def nak_knots(x):
    """FITPACK's knots of an interpolating cubic spline through the abscissae x (fpcurf.f with s = 0)."""
    x = np.asarray(x, float)
    return np.concatenate((np.full(4, x[0]), x[2:-2], np.full(4, x[-1])))


_SAVGOL_OPS: Dict[tuple, tuple] = {}


def savgol_operator(win: int, order: int):
    """scipy.signal.savgol_filter(x, win, order) (mode="interp") as a linear operator, for cfp_bank_nowiggle:
    (taps[win], off, EL[half][win], ER[half][win]) -- interior row s is taps . x[s + off : s + off + win]; the first and
    last half = win // 2 rows are the least-squares polynomial through the first / last `win` samples evaluated at
    those positions (scipy's _fit_edges_polyfit), here in a centred Legendre basis (agrees with scipy to 1e-14)."""
    key = (int(win), int(order))
    if key not in _SAVGOL_OPS:
        from scipy.signal import savgol_coeffs

        half = win // 2
        taps = np.ascontiguousarray(savgol_coeffs(win, order)[::-1])  # convolution -> correlation order
        off = -half if win % 2 else -half + 1  # scipy.ndimage.convolve1d's centre of an even-length kernel
        x = (np.arange(win, dtype=float) - 0.5 * (win - 1)) / (0.5 * (win - 1))
        V = np.polynomial.legendre.legvander(x, order)
        P = np.linalg.pinv(V)
        _SAVGOL_OPS[key] = (taps, off, np.ascontiguousarray(V[:half] @ P), np.ascontiguousarray(V[win - half:] @ P))
    return _SAVGOL_OPS[key]


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
        # raw tables in the bank's units: what the device fits (CosmoSet.prepare / bank) and the host fallbacks fit.
        # (z, k) surfaces are held by NAME and read from `t` on demand: the tables of an emulated cosmology are formed
        # lazily (bank.LazyTables), and never on the host when its bank row is combined on the device
        self._tables = t
        self._tab1d = {"h_of_z": (1 / rf) * np.asarray(t["H"], float).flatten(), "s8_of_z": np.asarray(t["s8"], float).flatten()}
        self._tab2d = {_lib.TAB_PLIN: ("Pl", pkf), _lib.TAB_PNL: ("Pnl", pkf), _lib.TAB_F: ("f", 1.0), _lib.TAB_D: ("D", 1.0)}
        self._lazy = {
            "D_growth_zk": _LazySpline(zg, kg, t, "D"),
            "f_growthrate_zk": _LazySpline(zg, kg, t, "f"),
            "Pk_l": _LazySpline(zg, kg, t, "Pl", pkf),
            "Pk_nl": _LazySpline(zg, kg, t, "Pnl", pkf),
        }
        if self.cb_files_on:  # cosmology.py:1508-1531
            self._lazy.update({
                "f_growthrate_cb_zk": _LazySpline(zg, kg, t, "fcb"),
                "Pk_cb_l": _LazySpline(zg, kg, t, "Plcb", pkf),
                "Pk_cb_nl": _LazySpline(zg, kg, t, "Pnlcb", pkf),
            })
            self._tab1d["s8_cb_of_z"] = np.asarray(t["s8cb"], float).flatten()
            self._tab2d.update({_lib.TAB_PLIN_CB: ("Plcb", pkf), _lib.TAB_PNL_CB: ("Pnlcb", pkf), _lib.TAB_F_CB: ("fcb", 1.0)})
        # D(z, k) at the first tabulated wavenumber: the growth of the intrinsic-alignment window (k = 1e-4 below it)
        self._growth_col0 = np.ascontiguousarray(t.column0("D") if hasattr(t, "column0") else np.asarray(t["D"], float)[:, 0])
        # an emulated cosmology whose bank row can be combined on the device from the tabulated rows: its weights
        self._emu = None
        if self.folder == "<emulated>" and hasattr(t, "weights") and pkf == 1.0 and kf == 1.0:
            self._emu = (emulator, t.weights)
        if not device_prep(self.settings):
            for name in ("h_of_z", "com_dist", "ang_dist", "s8_of_z"):  # the reference fits these eagerly
                getattr(self, name)
        self.results = self  # reference code reaches splines through `.results`
        # the no-wiggle settings this facade is cached under: kept by value (the settings dictionary may be edited in
        # place after the facade was made; what is computed lazily must still match the cache key)
        st = self.settings
        self._savgol = (st["savgol_internalkmin"], st["savgol_internalsamples"], st["savgol_width"], st["savgol_polyorder"])
        self._nw_cache = {}
        self._moment_cache = {}

    def __getattr__(self, name):
        """The splines, fitted on first access: the (z, k) surfaces (a GCsp job never touches P_nl, a photo job never f)
        and the functions of z.  `CosmoSet.prepare` fits the latter on the device for all cosmologies of a job at once
        and stores them under the same names, so this host fit (scipy, the reference's own calls,
        cosmology.py:1447-1506) only runs for a facade used on its own."""
        d = self.__dict__
        lazy = d.get("_lazy")
        if lazy is not None and name in lazy:
            return lazy[name].get()
        coefs = d.get("_coef1d")
        if coefs is not None and name in coefs:  # fitted on the device (CosmoSet.prepare): wrap the coefficients
            t, c = coefs[name]
            spl = d[name] = InterpolatedUnivariateSpline._from_tck((t, c, 3))
            return spl
        tab = d.get("_tab1d")
        if tab is not None:
            if name in tab:
                spl = InterpolatedUnivariateSpline(self.zgrid, tab[name])
                d[name] = spl
                return spl
            if name in ("com_dist", "ang_dist"):
                dcom = _dcom_all(self.zgrid, self.h_of_z)
                d["com_dist"] = InterpolatedUnivariateSpline(self.zgrid, dcom)
                d["ang_dist"] = InterpolatedUnivariateSpline(self.zgrid, dcom / (1 + self.zgrid))
                return d[name]
        raise AttributeError(name)

    def table2d(self, slot):
        """(raw (z, k) table, unit factor) of a bank slot."""
        name, scale = self._tab2d[slot]
        return self._tables[name], scale

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
        if k is None and ("_growth_lowk" in self.__dict__ or "_growth_lowk" in self.__dict__.get("_coef1d", ())):
            return self._growth_lowk(z)  # D(z, 1e-4) below the first tabulated k: the spline along z of that column
        return self.D_growth_zk(z, 0.0001 if k is None else k, grid=False)

    def sigma8_of_z(self, z, tracer="matter"):
        return self.s8_cb_of_z(z) if tracer == "clustering" else self.s8_of_z(z)  # cosmology.py:1851-1857

    def sigma8_cb_of_z(self, z):
        return self.s8_cb_of_z(z)

    def fsigma8_of_z(self, z, k=None, gamma=False, tracer="matter"):
        return self.f_growthrate(z, k, gamma, tracer=tracer) * self.sigma8_of_z(z, tracer=tracer)

    def Omegam_of_z(self, z):
        if np.ndim(z) == 0 and z == 0:  # E(0) = H(0) / H(0) = 1 exactly: no spline call
            return self.cosmopars["Omegam"] * (1 + z) ** 3 / 1.0
        return (self.cosmopars["Omegam"] * (1 + z) ** 3) / self.E_hubble(z) ** 2

    def SigmaMG(self, z, k):
        return np.array(1)

    def nonwiggle_samples(self):
        """(ln k samples, Savitzky-Golay window) of the no-wiggle smoothing (cosmology.py:1767-1779)."""
        hit = self.__dict__.get("_nw_samples")
        if hit is None:
            h = self.cosmopars["h"]
            kmin, nsamp, width, _ = self._savgol
            lk = np.linspace(np.log(h * kmin), np.log(h * np.max(self.kgrid)), nsamp)
            dlnk = np.mean(lk[1:] - lk[0:-1])
            hit = self.__dict__["_nw_samples"] = (lk, int(np.round(width / np.log(1 + dlnk))))
        return hit

    def nonwiggle_table(self, z, nonlinear=False, tracer="matter"):
        """Knots and coefficients of the degree-1 no-wiggle spline at redshift z
        (Savitzky-Golay smoothing in ln k, cosmology.py:1760-1813); the third entry is the host spline or None when the
        table came from the device (`CosmoSet.prepare_spectro`)."""
        key = (float(z), bool(nonlinear), tracer == "clustering")
        if key not in self._nw_cache:
            lk, n_sg = self.nonwiggle_samples()
            lp = InterpolatedUnivariateSpline(
                lk, np.log(self.matpow(z, np.exp(lk), nonlinear=nonlinear, tracer=tracer).flatten()), k=1)
            sg = savgol_filter(lp(lk), n_sg, self._savgol[3])
            spl = InterpolatedUnivariateSpline(np.exp(lk), np.exp(sg), k=1)
            t, c, _ = spl._eval_args
            n = len(t) - 2
            self._nw_cache[key] = (np.ascontiguousarray(t[1:-1]), np.ascontiguousarray(c[:n]), spl)
        return self._nw_cache[key]

    def nonwiggle_pow(self, z, k, nonlinear=False, tracer="matter"):
        kk, vv, spl = self.nonwiggle_table(z, nonlinear, tracer)
        if spl is None:
            spl = InterpolatedUnivariateSpline(kk, vv, k=1)
            self._nw_cache[(float(z), bool(nonlinear), tracer == "clustering")] = (kk, vv, spl)
        return spl(k)

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
    _BANK_CACHE.clear()
    from . import spectro

    spectro._PNW_MEMO.clear()


def _config_key_part(config):
    """Everything of a facade's cache key that belongs to the configuration, rebuilt only when one of the values it is
    made of changes (a forecast asks for thirty facades; sorting and tupling the same dictionaries thirty times was a
    tenth of its host time).  The check compares the raw values, so editing a setting in place is still seen."""
    ext, st = config.external, config.settings
    fid = config.fiducialparams
    tables, emulator, names = ext.get("tables"), ext.get("emulator"), ext["paramnames"]
    # (objects by identity -- kept alive beside the cached part, so an id cannot be recycled --, values by value)
    sig = (id(tables), id(emulator), id(names), len(names), ext.get("directory"), ext.get("k-units"), ext.get("r-units"),
           tuple(ext["eps_values"]), st["savgol_internalkmin"], st["savgol_internalsamples"], st["savgol_width"],
           st["savgol_polyorder"], st.get("GCsp_Tracer"), st.get("GCph_Tracer"), tuple(fid), tuple(fid.values()))
    hit = config.__dict__.get("_cosmo_key_part")
    if hit is not None and hit[0] == sig:
        return hit[1]
    src = id(ext["tables"]) if ext.get("tables") is not None else os.path.abspath(ext["directory"])
    part = (src, tuple(sorted(fid.items())), tuple(ext["paramnames"]), tuple(ext["eps_values"]), ext.get("k-units"),
            ext.get("r-units"), st["savgol_internalkmin"], st["savgol_internalsamples"], st["savgol_width"],
            st["savgol_polyorder"], st.get("GCsp_Tracer"), st.get("GCph_Tracer"), id(ext.get("emulator")))
    config.__dict__["_cosmo_key_part"] = (sig, part, (tables, emulator, names))
    return part


def cached_cosmo(cosmopars, config, cosmo_key=None) -> "cosmo_functions":
    """Boltzmann input is computed once and cached (north_star): one spline facade per (table source, folder,
    parameter values, units, no-wiggle settings), shared by every FisherMatrix / likelihood of the process.
    `cosmo_key`: tuple(sorted(cosmopars.items())) when the caller already has it."""
    ext = config.external
    if cosmo_key is None:
        cosmo_key = tuple(sorted(cosmopars.items()))
    key = (_config_key_part(config), cosmo_key)
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
                cosmo = cached_cosmo(cosmopars, self.config, cosmo_key=key)
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
            self.release_bank()
            self.cosmos = self.cosmos[:1]
            self._by_key = {k: v for k, v in self._by_key.items() if v == 0}

    def release_bank(self):
        """Drop this set's device bank.  A bank of host-fitted rows belongs to the set and is freed; a bank fitted on the
        device is shared through the resident cache and lives until it is evicted there (or `clear_caches`)."""
        if self._bank is not None:
            if not any(hit[2] is self._bank for hit in _BANK_CACHE.values()):
                self._bank.close()
            self._bank = None

    def context(self) -> _lib.Context:
        """The device context every bank, plan and likelihood helper of this job uses: options["device"]
        (default: LOCAL_RANK or 0), created on first use."""
        if self.ctx is None:
            self.ctx = _lib.default_context((getattr(self.config, "settings", None) or {}).get("device"))
        return self.ctx

    def bank(self) -> _lib.Bank:
        if self._bank is None:
            ctx = self.context()
            if device_prep(self.config.settings):
                self._bank = _fitted_bank(ctx, self.cosmos, self.slots)
            if self._bank is None:
                self._bank = _lib.Bank.from_rows(ctx, [c.packed_row(ctx, self.slots) for c in self.cosmos])
        return self._bank

    # --- batched preparation on the device ------------------------------------------
    _ACCESSOR_OF = {"h_of_z": "Hubble", "com_dist": "comoving", "ang_dist": "angdist", "s8_of_z": "sigma8_of_z",
                    "s8_cb_of_z": "sigma8_cb_of_z", "_growth_lowk": "growth"}

    def prepare(self, grids=(), scalars=()):
        """Functions of z of every cosmology of the set that has none yet, in ONE device call (cfp_background_fit):
        H(z), the comoving-distance integrals, both distances, sigma8(z) and the low-k growth as cubic splines whose
        coefficients come back to the facades (cosmology.py:36-55, 1447-1506; wrapped in scipy spline objects on first
        use).  `grids` (redshift arrays) and `scalars` (redshifts) a job is about to ask for are evaluated in the same
        call and land in the facades' memos, so a batch of thousands of cosmologies costs no per-cosmology spline
        call.  A facade that is never prepared fits the same splines on the host on first access."""
        if not device_prep(self.config.settings):
            return
        todo = [c for c in self.cosmos if "h_of_z" not in c.__dict__ and "com_dist" not in c.__dict__
                and "h_of_z" not in c.__dict__.get("_coef1d", ())]
        if not todo:
            return
        grids = [np.ascontiguousarray(g, dtype=float).ravel() for g in grids]
        scalars = [float(z) for z in scalars]
        zq = np.concatenate(grids + [np.array(scalars, dtype=float)]) if (grids or scalars) else None
        groups: Dict[tuple, list] = {}
        for c in todo:
            groups.setdefault((c.zgrid.tobytes(), tuple(sorted(c._tab1d)), bool(c.kgrid[0] >= 1e-4)), []).append(c)
        for (_, names, lowk), cs in groups.items():
            zg = cs[0].zgrid
            if zg.size < 5 or zg.size > 2048:
                continue
            extras = [n for n in names if n != "h_of_z"]
            ex = [[c._tab1d[n] for n in extras] + ([c._growth_col0] if lowk else []) for c in cs]
            out, coef = _lib.background_fit(self.context(), zg, np.stack([c._tab1d["h_of_z"] for c in cs]), np.array(ex),
                                            zq=zq, coef=True)
            t = nak_knots(zg)
            rows = ["h_of_z", "com_dist", "ang_dist"] + extras + (["_growth_lowk"] if lowk else [])
            for i, c in enumerate(cs):
                d = c.__dict__
                cf = d.setdefault("_coef1d", {})
                ci = coef[i].copy()  # (a facade keeps its own rows: views would pin the whole batch's arrays)
                for j, n in enumerate(rows):
                    cf[n] = (t, ci[j])
                if out is None:
                    continue
                oi = out[i].copy()
                gm, sm = d.setdefault("_grid_memo", {}), d.setdefault("_scalar_memo", {})
                for j, n in enumerate(rows):
                    acc = self._ACCESSOR_OF[n]
                    o = 0
                    for g in grids:
                        v = oi[j, o:o + g.size]
                        v.setflags(write=False)
                        gm.setdefault((acc, g.tobytes()), v)
                        o += g.size
                    for z in scalars:
                        sm.setdefault((acc, z), oi[j, o])
                        o += 1

    def prepare_spectro(self, zs, tracer="matter"):
        """Velocity moments and no-wiggle spectra of every cosmology of the set at the redshifts zs, two device calls
        on the fitted bank (cfp_bank_theta_moments, cfp_bank_nowiggle; spectro_obs.py:724-755, cosmology.py:1760-1813)
        whose results land in the facades' caches."""
        if not device_prep(self.config.settings):
            return
        from .spectro import _K_MOMENTS  # (the fixed 1024-point grid of spectro_obs.py:261-263)

        cb = tracer == "clustering"
        tab_p, tab_f = (_lib.TAB_PLIN_CB, _lib.TAB_F_CB) if cb else (_lib.TAB_PLIN, _lib.TAB_F)
        if self.slots is not None and not (tab_p in self.slots and tab_f in self.slots):
            return
        zs = [float(z) for z in zs]
        need_m = [z for z in zs if any(z not in c._moment_cache for c in self.cosmos)]
        need_w = [z for z in zs if any((z, False, cb) not in c._nw_cache for c in self.cosmos)]
        if not need_m and not need_w:
            return
        bank = self.bank()
        if need_m:
            # the velocity moments are those of the MATTER spectrum whatever the tracer (spectro_obs.py:740-747 calls
            # matpow / f_growthrate without one): with the cdm+baryon tracer they need a second small bank
            mbank = bank if not cb else _fitted_bank(self.context(), self.cosmos, (_lib.TAB_PLIN, _lib.TAB_F))
            if mbank is None:
                return
            mom = mbank.theta_moments(_lib.TAB_PLIN, _lib.TAB_F, need_m, _K_MOMENTS)
            for i, c in enumerate(self.cosmos):
                for j, z in enumerate(need_m):
                    c._moment_cache.setdefault(z, tuple(mom[i, j]))
        if need_w:
            samples = [c.nonwiggle_samples() for c in self.cosmos]
            nsamp = samples[0][0].size
            order = self.cosmos[0]._savgol[3]
            wins = sorted({n for _, n in samples})
            if any(lk.size != nsamp for lk, _ in samples) or any(not order < w <= nsamp for w in wins):
                return  # (scipy raises on such windows: leave the error to the host path)
            ks = np.exp(np.stack([lk for lk, _ in samples]))
            out = bank.nowiggle(tab_p, need_w, ks, [wins.index(n) for _, n in samples],
                                [savgol_operator(w, order) for w in wins])
            for i, c in enumerate(self.cosmos):
                for j, z in enumerate(need_w):
                    c._nw_cache.setdefault((z, False, cb), (ks[i], np.ascontiguousarray(out[i, j]), None))


_BANK_CACHE: Dict[tuple, tuple] = {}
_BANK_CACHE_MAX = 8
_BANK_CACHE_MAX_COSMOS = 64  # larger sets are one-off batches: not kept


def _combined_bank(ctx, cosmos, slots):
    """Bank of a set that contains EMULATED cosmologies, without their tables ever existing on the host: the rows of
    the tabulated cosmologies are fitted on the device once (`TaylorEmulator.base_bank`, kept), and every cosmology of
    the set -- tabulated ones through one-hot weights -- is a weighted sum of those coefficient rows (interpolating
    spline fits are linear in the table values): one `cfp_bank_combine_rows`.  None when the set has no emulated
    cosmology or does not qualify (several emulators, tables in h units, a cosmology from neither source)."""
    emus = [c._emu for c in cosmos if getattr(c, "_emu", None) is not None]
    if not emus or os.environ.get("CFP_NO_COMBINE", "0") not in ("", "0"):  # (the switch: tests compare the two routes)
        return None
    emulator = emus[0][0]
    if any(e[0] is not emulator for e in emus):
        return None
    order = emulator.folder_order
    col = {f: i for i, f in enumerate(order)}
    W = np.zeros((len(cosmos), len(order)))
    for m, c in enumerate(cosmos):
        if getattr(c, "_emu", None) is not None:
            for f, x in c._emu[1].items():
                W[m, col[f]] = x
        elif c.folder in col and c.external.get("tables") is emulator.tables and c.external.get("k-units") != "h/Mpc":
            W[m, col[c.folder]] = 1.0  # a tabulated cosmology of the same bank
        else:
            return None
    names = {s: cosmos[0]._tab2d[s][0] for s in slots if s in cosmos[0]._tab2d}
    if len(names) != len(slots):
        return None
    base = emulator.base_bank(ctx, slots, names)
    if base is None:
        return None
    return base.combined(W, keep_base=False)


def _fitted_bank(ctx, cosmos, slots):
    """Bank of the cosmologies' raw tables, fitted on the device (cfp_bank_fit) and kept resident: a later job on the same
    cosmologies (every forecast of a fiducial, every batch of a nuisance-only sampler) uploads nothing.  None when
    the tables do not share the z grid and the table shapes (the per-cosmology host fits handle that)."""
    slots = tuple(sorted(slots)) if slots is not None else tuple(sorted(cosmos[0]._tab2d))
    key = (id(ctx), slots, tuple(id(c) for c in cosmos))
    hit = _BANK_CACHE.get(key)
    if hit is not None and hit[0] is ctx and all(a is b for a, b in zip(hit[1], cosmos)) and hit[2].handle:
        _BANK_CACHE[key] = _BANK_CACHE.pop(key)  # most recently used last
        hit[2].h2d_last = 0  # resident: this job uploads nothing
        return hit[2]
    zg, nk = cosmos[0].zgrid, cosmos[0].kgrid.size
    if zg.size < 5 or nk < 5:
        return None
    bank = _combined_bank(ctx, cosmos, slots)
    if bank is not None:
        bank.h2d_last = bank.nbytes
        _remember_bank(key, ctx, cosmos, bank)
        return bank
    tables, scales = [], []
    for c in cosmos:
        if c.kgrid.size != nk or not np.array_equal(c.zgrid, zg) or any(s not in c._tab2d for s in slots):
            return None
        row, sc = [None] * _lib.NTAB, [1.0] * _lib.NTAB
        for s in slots:
            tab, scale = c.table2d(s)
            if np.shape(tab) != (zg.size, nk):
                return None
            row[s], sc[s] = tab, scale
        tables.append(row)
        scales.append(sc)
    bank = _lib.Bank.fit(ctx, zg, np.stack([c.kgrid for c in cosmos]), tables, scales)
    bank.h2d_last = bank.nbytes
    _remember_bank(key, ctx, cosmos, bank)
    return bank


def _remember_bank(key, ctx, cosmos, bank):
    """Keep a fitted bank resident for the next job on the same cosmologies -- unless it is the bank of a large one-off
    set (a sampler batch in which every point has its own cosmology: gigabytes that no later job will ask for again;
    such a bank belongs to its CosmoSet and is freed with it)."""
    if len(cosmos) > _BANK_CACHE_MAX_COSMOS:
        return
    _BANK_CACHE[key] = (ctx, list(cosmos), bank)
    while len(_BANK_CACHE) > _BANK_CACHE_MAX:
        _BANK_CACHE.pop(next(iter(_BANK_CACHE)))
