#!/usr/bin/env python
"""Generate golden vectors by RUNNING THE UNMODIFIED REFERENCE in this container.

Usage (build container only; /root/reference does not exist on the GPU box)::

    python tools/make_goldens.py [case ...]      # default: all cases

The reference (pure Python, imported from /root/reference) is driven through its
public API in ``code="external"`` mode on the deterministic toy table bank of
``cosmicfishpie_b200.toy_tables``; outputs are frozen as small ``.npz`` fixtures
under ``tests/golden/``.  These fixtures are what pins the oracle (``oracle/``)
and, through it, the CUDA path.  Nothing in the product imports this file.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(REPO, "tools", "_stubs"))  # 6-line `nautilus` stub

from cosmicfishpie_b200 import toy_tables as tt  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
SCRATCH = os.environ.get("CFP_SCRATCH", "/tmp/cfp_gold")
PARS7 = list(tt.COSMO_KEYS)
KSTRIDE = 16  # lattices are stored every 16th k sample to keep fixtures small


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def bank_dir():
    d = os.path.join(SCRATCH, "bank")
    if not os.path.isfile(os.path.join(d, "fiducial_eps_0", "Plincb-zk.txt")):  # (cb tables were added later)
        tt.write_bank(d, eps_values=(0.01,))
    return d


def base_options(outroot, spectro=True, photo=False, **kw):
    o = {
        "accuracy": 1,
        "feedback": 0,
        "code": "external",
        "outroot": outroot,
        "results_dir": os.path.join(SCRATCH, "results") + "/",
        "survey_name": "Euclid",
        "survey_name_spectro": "Euclid-Spectroscopic-ISTF-Pessimistic" if spectro else False,
        "survey_name_photo": "Euclid-Photometric-ISTF-Pessimistic" if photo else False,
        "cosmo_model": "w0waCDM",
        "derivatives": "3PT",
    }
    o.update(kw)
    return o


def fisher_payload(F, fm):
    names = list(F.get_param_names())
    return {
        "fisher": np.array(F.fisher_matrix),
        "param_names": np.array(names),
        "sigma": np.array(F.get_confidence_bounds()),
        "freeparams_keys": np.array(list(fm.freeparams.keys())),
        "freeparams_eps": np.array([fm.freeparams[k] for k in fm.freeparams]),
    }


def run_gcsp(tag, freepars, options_extra=None, specifications=None, max_z_bins=None, spectro_yaml=None):
    import cosmicfishpie.fishermatrix.cosmicfish as cff

    ext = tt.extfiles_for(bank_dir())
    o = base_options(tag, spectro=True, photo=False, **(options_extra or {}))
    if spectro_yaml:
        o["survey_name_spectro"] = spectro_yaml
    t0 = time.time()
    with quiet():
        fm = cff.FisherMatrix(
            fiducialpars=dict(tt.FIDUCIAL),
            freepars=dict(freepars),
            options=o,
            specifications=specifications or {},
            observables=["GCsp"],
            extfiles=ext,
            cosmoModel="w0waCDM",
            surveyName="Euclid",
        )
        F = fm.compute(max_z_bins=max_z_bins)
    wall = time.time() - t0
    out = fisher_payload(F, fm)
    out["wall_s"] = wall
    out["zmids"] = np.array(fm.zmids)
    out["k_grid"] = fm.Pk_kgrid
    out["mu_grid"] = fm.Pk_mugrid
    out["veff_strided"] = np.array([v[:, ::KSTRIDE] for v in fm.veff_arr])
    # fiducial ln P_obs lattice (strided) straight from the fiducial observable
    lnp = np.array(
        [fm.pk_obs_fid.lnpobs_gg(z, fm.Pk_kmesh, fm.Pk_mumesh)[:, ::KSTRIDE] for z in fm.zmids]
    )
    out["lnpobs_fid_strided"] = lnp
    for par in fm.freeparams:
        out["deriv_strided__" + par] = np.array(
            [np.asarray(fm.derivs_dict[par][i])[:, ::KSTRIDE] * np.ones_like(lnp[0]) for i in range(len(fm.zmids))]
        )
    out["volume_survey"] = np.array([fm.pk_cov.volume_survey(i) for i in range(len(fm.zmids))])
    out["n_density"] = np.array([fm.pk_cov.n_density(z) for z in fm.zmids])
    pkF = np.array([fm.fisher_per_bin(i) for i in range(len(fm.zmids))])
    out["fisher_per_bin"] = pkF
    np.savez_compressed(os.path.join(GOLD, tag + ".npz"), **out)
    print(f"{tag}: npar={len(out['param_names'])} wall={wall:.2f}s F00={out['fisher'][0,0]:.10e}")
    return fm


def run_photo(tag, observables, freepars, options_extra=None, specifications=None, photo_yaml=None, save_derivs=()):
    import cosmicfishpie.fishermatrix.cosmicfish as cff

    ext = tt.extfiles_for(bank_dir())
    o = base_options(tag, spectro=False, photo=True, **(options_extra or {}))
    if photo_yaml:
        o["survey_name_photo"] = photo_yaml
    t0 = time.time()
    with quiet():
        fm = cff.FisherMatrix(
            fiducialpars=dict(tt.FIDUCIAL),
            freepars=dict(freepars),
            options=o,
            specifications=specifications or {},
            observables=list(observables),
            extfiles=ext,
            cosmoModel="w0waCDM",
            surveyName="Euclid",
        )
        F = fm.compute()
    wall = time.time() - t0
    out = fisher_payload(F, fm)
    out["wall_s"] = wall
    cls = fm.photo_LSS.cls
    keys = [k for k in cls if k != "ells"]
    out["ells"] = np.array(cls["ells"])
    out["cl_keys"] = np.array(keys)
    out["cls"] = np.array([cls[k] for k in keys])
    out["covmat"] = np.array([np.array(c, dtype=float) for c in fm.photo_LSS.covmat])
    out["cov_cols"] = np.array(list(fm.photo_LSS.covmat[0].columns))
    for par in save_derivs:
        d = fm.photo_derivs[par]
        out["dcl__" + par] = np.array([d[k] for k in keys])
    pc = fm.photo_obs_fid
    pc.compute_all()
    out["z_grid"] = pc.z
    out["sqrtP_WL"] = pc.sqrtPell["WL"][::8, ::4]
    np.savez_compressed(os.path.join(GOLD, tag + ".npz"), **out)
    print(f"{tag}: npar={len(out['param_names'])} wall={wall:.2f}s F00={out['fisher'][0,0]:.10e}")
    return fm


CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


@case
def gcsp_cfg1():
    """BASELINE cfg1 analogue: GCsp {Omegam,h} + auto nuisances, 3PT, 17x1025."""
    run_gcsp("gcsp_cfg1", {"Omegam": 0.01, "h": 0.01})


@case
def gcsp_w0wa():
    """cfg4 analogue on the default lattice: 7 cosmo + 4 lnbg + 4 Ps."""
    run_gcsp("gcsp_w0wa", {p: 0.01 for p in PARS7})


@case
def gcsp_small():
    """Small lattice with EVEN sample counts (scipy simpson's even-N correction) and 2 bins."""
    run_gcsp(
        "gcsp_small",
        {"Omegam": 0.01, "h": 0.01, "w0": 0.01},
        options_extra={"spectro_Pk_k_samples": 64, "spectro_Pk_mu_samples": 8},
        max_z_bins=2,
    )


@case
def gcsp_kat_bias():
    """Known-answer lattices of ComputeGalSpectro.observed_Pgg with the two bias overrides no Fisher path uses:
    b_i (a number replacing the whole bias term, spectro_obs.py:617-618) and use_bias_funcs=True (per-bin biases
    interpolated linearly in z, nuisance.py:327-344), at redshifts between the bin centres."""
    import cosmicfishpie.fishermatrix.cosmicfish as cff
    import cosmicfishpie.LSSsurvey.spectro_obs as spobs

    ext = tt.extfiles_for(bank_dir())
    with quiet():
        fm = cff.FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"h": 0.01}, options=base_options("gcsp_kat_bias"),
                              specifications={}, observables=["GCsp"], extfiles=ext, cosmoModel="w0waCDM", surveyName="Euclid")
        cp = dict(tt.FIDUCIAL)
        cp["h"] = tt.FIDUCIAL["h"] * 1.01
        k = np.linspace(0.002, 0.3, 48)
        mu = np.linspace(0.0, 1.0, 7)
        kk, mm = np.meshgrid(k, mu)
        plain = spobs.ComputeGalSpectro(cp, fiducial_cosmopars=dict(tt.FIDUCIAL))
        funcs = spobs.ComputeGalSpectro(cp, fiducial_cosmopars=dict(tt.FIDUCIAL), use_bias_funcs=True)
        zs = np.array([1.0, 1.13, 1.47, 1.65])
        out = {"k": k, "mu": mu, "z": zs, "b_i": np.array(1.7),
               "pgg_default": np.array([plain.observed_Pgg(z, kk, mm) for z in zs]),
               "pgg_b_i": np.array([plain.observed_Pgg(z, kk, mm, b_i=1.7) for z in zs]),
               "lnp_b_i": np.array([plain.lnpobs_gg(z, kk, mm, b_i=1.7) for z in zs]),
               "pgg_bias_funcs": np.array([funcs.observed_Pgg(z, kk, mm) for z in zs])}
    np.savez_compressed(os.path.join(GOLD, "gcsp_kat_bias.npz"), **out)
    print("gcsp_kat_bias:", out["pgg_default"][1, 3, 5], out["pgg_b_i"][1, 3, 5], out["pgg_bias_funcs"][1, 3, 5])


@case
def wl_cfg2():
    """BASELINE cfg2 analogue: WL only, 5 cosmo + 3 IA."""
    run_photo("wl_cfg2", ["WL"], {p: 0.01 for p in PARS7[:5]}, save_derivs=("Omegam", "AIA", "etaIA"))


@case
def gcph_small():
    """GCph only, coarse grids (ell_sampling=20, accuracy unchanged)."""
    run_photo("gcph_small", ["GCph"], {"Omegam": 0.01, "sigma8": 0.01}, options_extra={"ell_sampling": 20},
              save_derivs=("Omegam", "b3"))


@case
def photo_3x2pt():
    """3x2pt w0waCDM 7 cosmo + 10 b + 3 IA (ISTF pessimistic, 10+10 bins)."""
    run_photo("photo_3x2pt", ["WL", "GCph"], {p: 0.01 for p in PARS7},
              save_derivs=("Omegam", "w0", "b1", "b7", "AIA", "betaIA"))


@case
def photo_cfg3():
    """BASELINE cfg3: 3x2pt, 13+13 bins, lmax 5000, 7 cosmo + 13 b + 3 IA (custom spec YAML shipped in
    cosmicfishpie_b200/data/specs, loaded by the reference through options['specs_dir'])."""
    run_photo("photo_cfg3", ["WL", "GCph"], {p: 0.01 for p in PARS7},
              options_extra={"specs_dir": os.path.join(REPO, "cosmicfishpie_b200", "data", "specs")},
              photo_yaml="Euclid-Photometric-13bins-lmax5000", save_derivs=("Omegam", "b13", "etaIA"))


SMALL = {"spectro_Pk_k_samples": 65, "spectro_Pk_mu_samples": 9}
MYSPECS = os.path.join(REPO, "cosmicfishpie_b200", "data", "specs")
# option / specification variants of the spectroscopic path (each exercises one branch of spectro_obs.py)
GCSP_VARIANTS = {
    "gcsp_var_noap_nofog": dict(options_extra=dict(SMALL, AP_effect=False, FoG_switch=False)),
    "gcsp_var_linear": dict(options_extra=dict(SMALL, GCsp_linear=True)),
    "gcsp_var_bfs8": dict(options_extra=dict(SMALL, bfs8terms=True)),
    "gcsp_var_dzdep": dict(options_extra=dict(SMALL), specifications={"spec_sigma_dz_type": "z-dependent", "spec_sigma_dz": 0.001}),
    "gcsp_var_nlcosmo": dict(options_extra=dict(SMALL, fix_cosmo_nl_terms=False)),
    "gcsp_var_qbias": dict(options_extra=dict(SMALL, specs_dir=MYSPECS), spectro_yaml="Euclid-Spectroscopic-ISTF-Pessimistic-Qbias"),
    # per-bin rescaling of sigma_p and sigma_v as free parameters (reference's own sigma_pv specification): 19 parameters
    # with max_z_bins=2 -> three 8-row blocks in the Gram kernel
    # cdm+baryon tracer: P_cb, f_cb, sigma8_cb tables (the theta-theta moments stay on total matter, spectro_obs.py:748-752)
    "gcsp_var_clustering": dict(options_extra=dict(SMALL, GCsp_Tracer="clustering", bfs8terms=True)),
    "gcsp_var_sigmapv": dict(options_extra=dict(SMALL, specs_dir=MYSPECS), spectro_yaml="Euclid-Spectroscopic-ISTF-Pessimistic-sigma_pv",
                             specifications={"nonlinear_parametrization": dict(
                                 rescale_sigmap=True, rescale_sigmav=True, default=None,
                                 rescale_sigma_pv={f"sigma{c}_{i}": 1.0 for c in "pv" for i in (1, 2, 3, 4)})}),
}
PHOTO_VARIANTS = {
    "photo_var_linear_pk": dict(obs=["WL", "GCph"], free={"Omegam": 0.01, "h": 0.01}, options_extra={"ell_sampling": 16, "nonlinear": False}),
    "photo_var_sqrtbias": dict(obs=["GCph"], free={"sigma8": 0.01}, options_extra={"ell_sampling": 16}, specifications={"ph_bias_model": "sqrt"}),
    "photo_var_gcfirst": dict(obs=["GCph", "WL"], free={"Omegam": 0.01}, options_extra={"ell_sampling": 12, "pivot_z_IA": 0.3}),
    # GCph on the cdm+baryon spectrum, lensing on total matter
    "photo_var_clustering": dict(obs=["WL", "GCph"], free={"Omegam": 0.01, "h": 0.01},
                                 options_extra={"ell_sampling": 12, "GCph_Tracer": "clustering"}),
    # photo-z parameters as free parameters: each of their stencil points has its own n_i(z) windows (zb has fiducial 0:
    # absolute step)
    "photo_var_photoz": dict(obs=["WL", "GCph"], free={"Omegam": 0.01, "sigma_b": 0.02, "zb": 0.002, "fout": 0.02},
                             options_extra={"ell_sampling": 12}),
}


def _make_variant(name):
    def fn():
        if name in GCSP_VARIANTS:
            v = GCSP_VARIANTS[name]
            run_gcsp(name, {"Omegam": 0.01, "h": 0.01, "w0": 0.01}, max_z_bins=2, **v)
        else:
            v = dict(PHOTO_VARIANTS[name])
            run_photo(name, v.pop("obs"), v.pop("free"), **v)
    fn.__name__ = name
    return fn


for _n in list(GCSP_VARIANTS) + list(PHOTO_VARIANTS):
    CASES[_n] = _make_variant(_n)


STEM_EPS = (0.01, 0.02, 0.03)


@case
def photo_stem():
    """SteM derivatives (derivatives.py:348-441): no production caller of the reference selects them
    (PhotoCov.compute_derivs hard-wires 3PT), so the derivative engine is driven directly on the reference's own
    getcls, and its own einsum Fisher assembly turns the result into a matrix.  With external tables the stencil is
    +-eps for every eps in extfiles["eps_values"], for every free parameter (nuisance included)."""
    import cosmicfishpie.configs.config as cfg
    import cosmicfishpie.fishermatrix.cosmicfish as cff
    import cosmicfishpie.fishermatrix.derivatives as fishderiv
    import cosmicfishpie.LSSsurvey.photo_cov as photo_cov
    import cosmicfishpie.LSSsurvey.photo_obs as photo_obs

    d = os.path.join(SCRATCH, "bank_stem")
    pars = ("Omegam", "h")
    if not os.path.isdir(os.path.join(d, "h_mn_eps_3p0E-02")):
        tt.write_bank(d, eps_values=STEM_EPS, pars=pars)
    ext = tt.extfiles_for(d, eps_values=STEM_EPS, pars=pars)
    obs = ["WL", "GCph"]
    o = base_options("photo_stem", spectro=False, photo=True, ell_sampling=10)
    t0 = time.time()
    with quiet():
        fm = cff.FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01, "h": 0.01}, options=o,
                              specifications={}, observables=obs, extfiles=ext, cosmoModel="w0waCDM", surveyName="Euclid")
        fid = photo_obs.ComputeCls(fm.fiducialcosmopars, fm.photopars, fm.IApars, fm.photobiaspars)
        lss = photo_cov.PhotoCov(fm.fiducialcosmopars, fm.photopars, fm.IApars, fm.photobiaspars, fiducial_Cls=fid)
        noisy, covmat = lss.compute_covmat()
        eng = fishderiv.derivatives(lss.getcls, lss.allparsfid, derivatives_type="STEM", observables_type=obs,
                                    external_settings=cfg.external)
        derivs = eng.result
        F = fm.photo_LSS_fishermatrix_einsum(noisy_cls=noisy, covmat=covmat, derivs=derivs)
    wall = time.time() - t0
    F = np.array(F)
    out = {"fisher": F, "param_names": np.array(list(fm.freeparams.keys())),
           "sigma": np.sqrt(np.diag(np.linalg.inv(F))), "wall_s": wall, "eps_values": np.array(STEM_EPS)}
    keys = [k for k in lss.cls if k != "ells"]
    out["ells"] = np.array(lss.cls["ells"])
    out["cl_keys"] = np.array(keys)
    for par in ("Omegam", "b3", "AIA"):
        out["dcl__" + par] = np.array([derivs[par][k] for k in keys])
    np.savez_compressed(os.path.join(GOLD, "photo_stem.npz"), **out)
    print(f"photo_stem: npar={len(out['param_names'])} wall={wall:.2f}s F00={out['fisher'][0,0]:.10e}")


LIKE_POINTS_PHOTO = {
    "fid": {},
    "nuis": {"b1": 1.02, "b7": 0.985, "AIA": 1.04, "etaIA": 0.97, "betaIA": 1.01},     # multiplicative shifts
    "cosmo": {"Omegam": 1.01},
    "both": {"h": 0.99, "b3": 1.03, "AIA": 0.95},
}
LIKE_POINTS_SPECTRO = {
    "fid": {},
    "nuis": {"lnbg_1": 1.01, "lnbg_4": 0.99},
    "cosmo": {"h": 1.01},
    "shot": {"Ps_2": ("abs", 35.0), "lnbg_3": 1.005, "w0": 0.99},
}


def _shift(allpars, shifts):
    out = {}
    for k, v in shifts.items():
        out[k] = v[1] if isinstance(v, tuple) else allpars[k] * v
    return out


@case
def like_photo():
    """Photometric likelihood (3x2pt 10+10 bins): data cells and log-likelihood at a few parameter points."""
    from cosmicfishpie.likelihood.photo_like import PhotometricLikelihood

    fm = run_photo("_tmp_like_photo", ["WL", "GCph"], {p: 0.01 for p in PARS7})
    os.remove(os.path.join(GOLD, "_tmp_like_photo.npz"))
    with quiet():
        like = PhotometricLikelihood(cosmo_data=fm)
    out = {k: np.array(v) for k, v in like.data_obs.items()}
    names, vals = [], []
    allp = dict(fm.allparams)
    for name, sh in LIKE_POINTS_PHOTO.items():
        with quiet():
            vals.append(like.loglike(param_dict=_shift(allp, sh)))
        names.append(name)
    out["point_names"] = np.array(names)
    out["loglike"] = np.array(vals)
    np.savez_compressed(os.path.join(GOLD, "like_photo.npz"), **out)
    print("like_photo:", dict(zip(names, vals)))


@case
def like_spectro():
    """Spectroscopic likelihood functions (spectro_like.py:25-233): wedges and Legendre multipoles."""
    from cosmicfishpie.likelihood import spectro_like as sl

    fm = run_gcsp("_tmp_like_spectro", {p: 0.01 for p in PARS7})
    os.remove(os.path.join(GOLD, "_tmp_like_spectro.npz"))
    # attribute the reference forgets to define (SURVEY 8a-L2')
    fm.pk_cov.volume_survey_array = np.array([fm.pk_cov.volume_survey(i) for i in range(len(fm.zmids))])
    with quiet():
        data = sl.observable_Pgg(fm.pk_cov, fm)
        pl_data = sl.legendre_Pgg(data, fm)
        cov, icov = sl.compute_covariance_legendre(pl_data, fm)
    out = {"data_strided": data[:, :, ::KSTRIDE], "pl_data_strided": pl_data[::KSTRIDE], "cov_strided": cov[::KSTRIDE]}
    names, w, lg = [], [], []
    allp = dict(fm.allparams)
    for name, sh in LIKE_POINTS_SPECTRO.items():
        with quiet():
            th = sl.compute_theory_spectro(_shift(allp, sh), fm, "wedges")
            w.append(sl.compute_wedge_chi2(data, th, fm))
            lg.append(sl.compute_chi2_legendre(pl_data, sl.legendre_Pgg(th, fm), icov))
        if name == "shot":
            out["theory_shot_strided"] = th[:, :, ::KSTRIDE]
        names.append(name)
    out["point_names"] = np.array(names)
    out["chi2_wedges"] = np.array(w)
    out["chi2_legendre"] = np.array(lg)
    np.savez_compressed(os.path.join(GOLD, "like_spectro.npz"), **out)
    print("like_spectro wedges:", dict(zip(names, w)), "legendre:", dict(zip(names, lg)))


# ------------------------------------------------------------------ the workloads bench.py times (round 2)
FINE = {"spectro_Pk_k_samples": 8193, "spectro_Pk_mu_samples": 129}
FINE_KSTRIDE, FINE_MUSTRIDE = 64, 8


@case
def gcsp_cfg4_fine():
    """BASELINE cfg4 as bench.py runs it: GCsp, 4 bins, 7 cosmological + 4 lnbg + 4 Ps parameters, 129 x 8193
    lattice.  Lattices are stored every 64th wavenumber and every 8th mu row."""
    import cosmicfishpie.fishermatrix.cosmicfish as cff

    ext = tt.extfiles_for(bank_dir())
    o = base_options("gcsp_cfg4_fine", spectro=True, photo=False, **FINE)
    t0 = time.time()
    with quiet():
        fm = cff.FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in PARS7}, options=o,
                              specifications={}, observables=["GCsp"], extfiles=ext, cosmoModel="w0waCDM",
                              surveyName="Euclid")
        F = fm.compute()
    wall = time.time() - t0
    out = fisher_payload(F, fm)
    out["wall_s"] = wall
    out["zmids"] = np.array(fm.zmids)
    sl = (slice(None, None, FINE_MUSTRIDE), slice(None, None, FINE_KSTRIDE))
    out["veff_strided"] = np.array([v[sl] for v in fm.veff_arr])
    lnp = np.array([fm.pk_obs_fid.lnpobs_gg(z, fm.Pk_kmesh, fm.Pk_mumesh)[sl] for z in fm.zmids])
    out["lnpobs_fid_strided"] = lnp
    for par in fm.freeparams:
        out["deriv_strided__" + par] = np.array(
            [(np.asarray(fm.derivs_dict[par][i]) * np.ones((129, 8193)))[sl] for i in range(len(fm.zmids))])
    out["fisher_per_bin"] = np.array([fm.fisher_per_bin(i) for i in range(len(fm.zmids))])
    np.savez_compressed(os.path.join(GOLD, "gcsp_cfg4_fine.npz"), **out)
    print(f"gcsp_cfg4_fine: npar={len(out['param_names'])} wall={wall:.1f}s F00={out['fisher'][0,0]:.10e}")


def _bank_steps(rng, n, pars=PARS7):
    """n cosmologies of the bank: one parameter each at fiducial*(1 +- 0.01) (wa: +-0.01 absolute)."""
    rows = []
    for _ in range(n):
        p = pars[rng.integers(len(pars))]
        sgn = 1.0 if rng.integers(2) else -1.0
        rows.append((p, tt.FIDUCIAL[p] * (1 + 0.01 * sgn) if tt.FIDUCIAL[p] != 0 else 0.01 * sgn))
    return rows


@case
def like_photo_cfg3():
    """Photometric likelihood on the bench workload (3x2pt, 13+13 bins, lmax 5000): 40 points -- the fiducial,
    a group that varies only the galaxy biases, a group that varies biases and IA, and 24 points on cosmologies of the
    bank with their own nuisance draws.  The parameter matrix is stored with the fixture."""
    from cosmicfishpie.likelihood.photo_like import PhotometricLikelihood

    fm = run_photo("_tmp_like_photo_cfg3", ["WL", "GCph"], {p: 0.01 for p in PARS7},
                   options_extra={"specs_dir": MYSPECS}, photo_yaml="Euclid-Photometric-13bins-lmax5000")
    os.remove(os.path.join(GOLD, "_tmp_like_photo_cfg3.npz"))
    with quiet():
        like = PhotometricLikelihood(cosmo_data=fm)
    allp = dict(fm.allparams)
    bkeys = [f"b{i}" for i in range(1, 14)]
    ikeys = ["AIA", "betaIA", "etaIA"]
    keys = PARS7 + bkeys + ikeys
    fid = np.array([allp[k] for k in keys])
    rng = np.random.default_rng(2024)
    mat = np.tile(fid, (40, 1))
    nb, ni = len(bkeys), len(ikeys)
    mat[1:9, 7:7 + nb] *= 1 + 0.02 * rng.standard_normal((8, nb))
    mat[9:16, 7:7 + nb] *= 1 + 0.02 * rng.standard_normal((7, nb))
    mat[9:16, 7 + nb:] *= 1 + 0.03 * rng.standard_normal((7, ni))
    for r, (p, v) in zip(range(16, 40), _bank_steps(rng, 24)):
        mat[r, keys.index(p)] = v
        mat[r, 7:7 + nb] *= 1 + 0.02 * rng.standard_normal(nb)
        if r % 2:
            mat[r, 7 + nb:] *= 1 + 0.03 * rng.standard_normal(ni)
    vals = []
    t0 = time.time()
    for row in mat:
        with quiet():
            vals.append(like.loglike(param_dict=dict(zip(keys, row))))
    out = {k: np.array(v) for k, v in like.data_obs.items()}
    out.update(keys=np.array(keys), points=mat, loglike=np.array(vals), wall_per_point_s=(time.time() - t0) / len(mat))
    np.savez_compressed(os.path.join(GOLD, "like_photo_cfg3.npz"), **out)
    print("like_photo_cfg3:", vals[:3], "...", out["wall_per_point_s"], "s/point")


@case
def like_spectro_fine():
    """Spectroscopic likelihood on the bench lattice (129 x 8193): wedges and Legendre multipoles at 22 points --
    groups of 4-7 points that share a cosmology and differ in lnbg_i / Ps_i (the moment kernels' case), plus single
    points.  The parameter matrix is stored with the fixture."""
    import cosmicfishpie.fishermatrix.cosmicfish as cff
    from cosmicfishpie.likelihood import spectro_like as sl

    ext = tt.extfiles_for(bank_dir())
    o = base_options("_tmp_like_spectro_fine", spectro=True, photo=False, **FINE)
    with quiet():
        # a forecast object without the derivative pass: the likelihood needs the fiducial observable and covariance
        fm = cff.FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in PARS7}, options=o,
                              specifications={}, observables=["GCsp"], extfiles=ext, cosmoModel="w0waCDM",
                              surveyName="Euclid")
        # the prelude of FisherMatrix.compute (cosmicfish.py:262-290) without its derivative pass
        import cosmicfishpie.LSSsurvey.spectro_cov as spec_cov
        import cosmicfishpie.LSSsurvey.spectro_obs as spec_obs
        fm.set_pk_settings()
        fm.obs_spectrum = ["g", "g"]
        fm.pk_obs_fid = spec_obs.ComputeGalSpectro(cosmopars=fm.fiducialcosmopars, fiducial_cosmopars=fm.fiducialcosmopars,
                                                   spectrobiaspars=fm.Spectrobiaspars, spectrononlinearpars=fm.Spectrononlinpars,
                                                   IMbiaspars=fm.IMbiaspars, PShotpars=fm.PShotpars, configuration=fm)
        fm.pk_cov = spec_cov.SpectroCov(fm.fiducialcosmopars, fiducial_specobs=fm.pk_obs_fid, bias_samples=fm.obs_spectrum,
                                        configuration=fm)
        fm.zmids = fm.pk_cov.inter_z_bin_mids
    fm.pk_cov.volume_survey_array = np.array([fm.pk_cov.volume_survey(i) for i in range(len(fm.zmids))])
    with quiet():
        data = sl.observable_Pgg(fm.pk_cov, fm)
        pl_data = sl.legendre_Pgg(data, fm)
        cov, icov = sl.compute_covariance_legendre(pl_data, fm)
    allp = dict(fm.allparams)
    keys = PARS7 + [f"lnbg_{i}" for i in range(1, 5)] + [f"Ps_{i}" for i in range(1, 5)]
    fid = np.array([allp[k] for k in keys])
    rng = np.random.default_rng(77)
    mat = np.tile(fid, (22, 1))

    def nuis(rows, shot=True):
        mat[rows, 7:11] += 0.01 * rng.standard_normal((len(rows), 4))
        if shot:
            mat[rows, 11:15] = rng.uniform(0.0, 50.0, size=(len(rows), 4))

    nuis(list(range(1, 8)))                       # 7 points on the fiducial cosmology (+ the fiducial itself)
    mat[8:13, keys.index("h")] = tt.FIDUCIAL["h"] * 1.01
    nuis(list(range(8, 13)))                      # group of 5
    mat[13:17, keys.index("Omegam")] = tt.FIDUCIAL["Omegam"] * 0.99
    nuis(list(range(13, 17)), shot=False)         # group of 4, biases only
    for r, (p, v) in zip(range(17, 22), [("w0", -1.01), ("wa", 0.01), ("sigma8", tt.FIDUCIAL["sigma8"] * 1.01),
                                         ("ns", tt.FIDUCIAL["ns"] * 0.99), ("Omegab", tt.FIDUCIAL["Omegab"] * 1.01)]):
        mat[r, keys.index(p)] = v                 # single points
    nuis([17, 19, 21])
    w, lg = [], []
    t0 = time.time()
    sl2 = (slice(None), slice(None, None, FINE_MUSTRIDE), slice(None, None, FINE_KSTRIDE))
    out = {"data_strided": data[sl2], "pl_data_strided": pl_data[::FINE_KSTRIDE], "cov_strided": cov[::FINE_KSTRIDE]}
    for i, row in enumerate(mat):
        with quiet():
            th = sl.compute_theory_spectro(dict(zip(keys, row)), fm, "wedges")
            w.append(sl.compute_wedge_chi2(data, th, fm))
            lg.append(sl.compute_chi2_legendre(pl_data, sl.legendre_Pgg(th, fm), icov))
        if i == 9:
            out["theory_9_strided"] = th[sl2]
    out.update(keys=np.array(keys), points=mat, chi2_wedges=np.array(w), chi2_legendre=np.array(lg),
               wall_per_point_s=(time.time() - t0) / len(mat))
    np.savez_compressed(os.path.join(GOLD, "like_spectro_fine.npz"), **out)
    print("like_spectro_fine wedges:", w[:4], "legendre:", lg[:4], out["wall_per_point_s"], "s/point")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    os.makedirs(SCRATCH, exist_ok=True)
    os.chdir(SCRATCH)  # the reference creates ./cache and ./memory_cache at import
    which = sys.argv[1:] or list(CASES)
    for name in which:
        CASES[name]()
