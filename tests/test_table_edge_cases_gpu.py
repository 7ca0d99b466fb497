"""GPU vs oracle on table sets that exercise the slow paths of the lattice evaluation:
  * tables that END inside the wavenumber range of the lattice -- FITPACK clamps the argument to the knot range
    (fpbisp.f) and the degree-1 no-wiggle spline extrapolates; the kernel's clamp branch (no SD_NOCLAMP);
  * f(z,k) fitted on its own wavenumber grid -- P_lin and f no longer share their k knots, the kernel searches both
    tables (no SD_SHARED_KNOTS).  Tables read from files always share one grid, so this is built by swapping the
    fitted spline on both sides."""
import os

import numpy as np
import pytest
import yaml
from scipy.interpolate import RectBivariateSpline

from conftest import REPO, base_options

pytestmark = pytest.mark.gpu
SPECDIR = os.path.join(REPO, "cosmicfishpie_b200", "data", "specs")
SMALL = {"spectro_Pk_k_samples": 257, "spectro_Pk_mu_samples": 9}
FREE = {"Omegam": 0.01, "h": 0.01}


def _oracle_fisher(bank, fm, hook=None, **units):
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.cosmo import Bank
    from oracle.spectro import spectro_fisher

    sp = yaml.safe_load(open(os.path.join(SPECDIR, "Euclid-Spectroscopic-ISTF-Pessimistic.yaml")))["specifications"]
    ob = Bank(bank, tt.FIDUCIAL, tt.COSMO_KEYS, **units)
    if hook is not None:
        make = ob.cosmo

        def cosmo(cosmopars, cache=True):
            c = make(cosmopars, cache=cache)
            hook(c)
            return c

        ob.cosmo = cosmo
    return spectro_fisher(ob, sp, SMALL, dict(tt.FIDUCIAL), dict(sp["sp_bias_parametrization"]["linear_log"]),
                          dict(sp["shot_noise_parametrization"]["constant"]), {}, dict(fm.config.freeparams), max_z_bins=2)


def _check(F, ref, tol=1e-9):
    assert list(F.get_param_names()) == ref["param_names"]
    scale = np.sqrt(np.outer(np.diag(ref["fisher"]), np.diag(ref["fisher"])))
    assert np.max(np.abs(F.fisher_matrix - ref["fisher"]) / scale) < tol


def test_lattice_beyond_the_table_range_is_clamped(tmp_path):
    from cosmicfishpie_b200 import cosmology as cc
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    full = tt.make_bank(eps_values=(0.01,), pars=tuple(FREE))
    keep = (full["fiducial_eps_0"]["k"] >= 2e-3) & (full["fiducial_eps_0"]["k"] <= 0.16)  # lattice: 6.7e-4 .. 0.1675
    bank = {name: {key: (v[..., keep] if v.shape[-1] == keep.size else v) for key, v in t.items()} for name, t in full.items()}
    ext = {"tables": bank, "directory": "<memory>", "paramnames": list(tt.COSMO_KEYS), "folder_paramnames": list(tt.COSMO_KEYS),
           "k-units": "1/Mpc", "r-units": "Mpc", "eps_values": [0.01]}
    cc.clear_caches()
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars=dict(FREE), options=base_options(tmp_path, **SMALL),
                      observables=["GCsp"], extfiles=ext)
    F = fm.compute(max_z_bins=2)
    assert fm.Pk_kgrid.max() > bank["fiducial_eps_0"]["k"].max() and fm.Pk_kgrid.min() < bank["fiducial_eps_0"]["k"].min()
    _check(F, _oracle_fisher(bank, fm))
    cc.clear_caches()


def test_growth_rate_on_its_own_wavenumber_grid(tmp_path, monkeypatch):
    from cosmicfishpie_b200 import cosmology as cc
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    bank = tt.make_bank(eps_values=(0.01,), pars=tuple(FREE))
    ext = {"tables": bank, "directory": "<memory>", "paramnames": list(tt.COSMO_KEYS), "folder_paramnames": list(tt.COSMO_KEYS),
           "k-units": "1/Mpc", "r-units": "Mpc", "eps_values": [0.01]}

    def coarse_f(zg, kg, table):  # every third wavenumber (first and last kept): different knots than P_lin
        sel = np.unique(np.r_[np.arange(0, kg.size, 3), kg.size - 1])
        return RectBivariateSpline(zg, kg[sel], np.asarray(table)[:, sel], kx=3, ky=3)

    # this package: the lazily fitted f spline of every cosmology
    orig_init = cc.cosmo_functions.__init__

    def patched_init(self, *a, **k):
        orig_init(self, *a, **k)
        lazy = self.__dict__["_lazy"]["f_growthrate_zk"]
        zg, kg, tables, name, _ = lazy._args
        lazy._spl, lazy._args = coarse_f(zg, kg, tables[name]), None

    # (tables whose surfaces live on different knots only arise from host fits: the device fit of a bank shares them)
    monkeypatch.setenv("CFP_HOST_PREP", "1")
    monkeypatch.setattr(cc.cosmo_functions, "__init__", patched_init)
    cc.clear_caches()
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars=dict(FREE), options=base_options(tmp_path, **SMALL),
                      observables=["GCsp"], extfiles=ext)
    F = fm.compute(max_z_bins=2)
    row = fm.fiducialcosmo.bank_row()
    from cosmicfishpie_b200 import _lib

    assert row[_lib.TAB_F][1].size != row[_lib.TAB_PLIN][1].size  # the k knots really differ

    def hook(c):  # the oracle's cosmology: the same coarse f spline
        if not getattr(c, "_coarse_f", False):
            c.f_growthrate_zk = coarse_f(c.zgrid, c.kgrid, c.f_growthrate_zk(c.zgrid, c.kgrid))
            c._coarse_f = True

    _check(F, _oracle_fisher(bank, fm, hook))
    cc.clear_caches()


@pytest.mark.parametrize("k_units,r_units", [("h/Mpc", "Mpc/h"), ("h/Mpc", "km/s/Mpc"), ("1/Mpc", "km/s/Mpc")])
def test_tables_in_other_units(tmp_path, k_units, r_units):
    """Tables tabulated in h/Mpc and (Mpc/h)^3, H in h/Mpc or km/s/Mpc (cosmology.py:1438-1444): the unit factors depend
    on the h of each cosmology, so the stencil cosmologies of h have their own wavenumber grids (the device fit takes
    one k grid per cosmology).  Same physics as the 1/Mpc bank, so the Fisher matrix must also stay close to that one."""
    from cosmicfishpie_b200 import cosmology as cc
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    full = tt.make_bank(eps_values=(0.01,), pars=tuple(FREE))
    bank = {}
    for name, t in full.items():
        h = tt.FIDUCIAL["h"] * (1.01 if name.startswith("h_pl") else 0.99 if name.startswith("h_mn") else 1.0)
        u = dict(t)
        if k_units == "h/Mpc":
            u["k"] = np.asarray(t["k"]) / h
            for key in ("Pl", "Pnl", "Plcb", "Pnlcb"):
                if key in t:
                    u[key] = np.asarray(t[key]) * h**3
        u["H"] = np.asarray(t["H"]) / h if r_units == "Mpc/h" else np.asarray(t["H"]) * cc.C_KMS
        bank[name] = u

    def run(tables, ku, ru):
        ext = {"tables": tables, "directory": "<memory>", "paramnames": list(tt.COSMO_KEYS),
               "folder_paramnames": list(tt.COSMO_KEYS), "k-units": ku, "r-units": ru, "eps_values": [0.01]}
        cc.clear_caches()
        fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars=dict(FREE), options=base_options(tmp_path, **SMALL),
                          observables=["GCsp"], extfiles=ext)
        return fm, fm.compute(max_z_bins=2)

    fm, F = run(bank, k_units, r_units)
    _check(F, _oracle_fisher(bank, fm, k_units=k_units, r_units=r_units))
    _, F0 = run(full, "1/Mpc", "Mpc")
    scale = np.sqrt(np.outer(np.diag(F0.fisher_matrix), np.diag(F0.fisher_matrix)))
    assert np.max(np.abs(F.fisher_matrix - F0.fisher_matrix) / scale) < 1e-6  # (knots moved by an ulp: not identical)
    cc.clear_caches()


def test_non_positive_spectrum_gives_nan_like_the_reference(tmp_path):
    """Tables that end well inside the lattice: the linearly extrapolated no-wiggle spectrum goes negative at high k,
    ln P_obs does not exist there.  The reference (numpy) propagates NaN into the Fisher matrix; so does the kernel."""
    from cosmicfishpie_b200 import cosmology as cc
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    full = tt.make_bank(eps_values=(0.01,), pars=tuple(FREE))
    keep = (full["fiducial_eps_0"]["k"] >= 2e-3) & (full["fiducial_eps_0"]["k"] <= 0.12)
    bank = {name: {key: (v[..., keep] if v.shape[-1] == keep.size else v) for key, v in t.items()} for name, t in full.items()}
    ext = {"tables": bank, "directory": "<memory>", "paramnames": list(tt.COSMO_KEYS), "folder_paramnames": list(tt.COSMO_KEYS),
           "k-units": "1/Mpc", "r-units": "Mpc", "eps_values": [0.01]}
    cc.clear_caches()
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars=dict(FREE),
                      options=base_options(tmp_path, outroot="", **SMALL), observables=["GCsp"], extfiles=ext)
    fm.compute(max_z_bins=2)
    ref = _oracle_fisher(bank, fm)
    assert np.array_equal(np.isnan(fm.fisher_array), np.isnan(ref["fisher"])) and np.isnan(ref["fisher"]).any()
    cc.clear_caches()
