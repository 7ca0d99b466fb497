"""CPU: the oracle (oracle/) reproduces the goldens of the unmodified reference BIT FOR BIT.

The goldens (tests/golden/*.npz) were produced by tools/make_goldens.py running the reference from
/root/reference in the build container on the toy table bank.  This is what pins the oracle."""
import os

import numpy as np
import pytest
import yaml

from conftest import REPO, golden

SPECDIR = os.path.join(REPO, "cosmicfishpie_b200", "data", "specs")
P7 = ["Omegam", "Omegab", "h", "ns", "sigma8", "w0", "wa"]


@pytest.fixture(scope="module")
def obank(toy_bank):
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.cosmo import Bank

    return Bank(toy_bank, tt.FIDUCIAL, tt.COSMO_KEYS)


def _spectro_specs():
    return yaml.safe_load(open(os.path.join(SPECDIR, "Euclid-Spectroscopic-ISTF-Pessimistic.yaml")))["specifications"]


@pytest.mark.parametrize("tag,free,settings,mzb", [
    ("gcsp_cfg1", {"Omegam": 0.01, "h": 0.01}, None, None),
    ("gcsp_small", {"Omegam": 0.01, "h": 0.01, "w0": 0.01}, {"spectro_Pk_k_samples": 64, "spectro_Pk_mu_samples": 8}, 2),
    ("gcsp_w0wa", {p: 0.01 for p in P7}, None, None),
])
def test_spectro_oracle_bit_exact(obank, tag, free, settings, mzb):
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.spectro import spectro_fisher

    sp = _spectro_specs()
    bias = dict(sp["sp_bias_parametrization"]["linear_log"])
    ps = dict(sp["shot_noise_parametrization"]["constant"])
    fp = dict(free)
    for k in list(bias) + list(ps):
        fp.setdefault(k, 1e-4)
    r = spectro_fisher(obank, sp, settings, dict(tt.FIDUCIAL), bias, ps, {}, fp, max_z_bins=mzb, return_all=True)
    g = golden(tag)
    assert r["param_names"] == list(g["param_names"])
    assert np.array_equal(r["fisher"], g["fisher"])
    assert np.array_equal(r["fisher_per_bin"], g["fisher_per_bin"])
    assert np.array_equal(np.array([x[:, ::16] for x in r["lnp_fid"]]), g["lnpobs_fid_strided"])
    assert np.array_equal(np.array([x[:, ::16] for x in r["veff"]]), g["veff_strided"])
    for par in r["param_names"]:
        d = np.array([np.asarray(r["derivs"][par][i])[:, ::16] for i in range(len(r["zmids"]))])
        assert np.array_equal(d, g["deriv_strided__" + par]), par
    s = np.sqrt(np.diag(np.linalg.inv(r["fisher"])))
    from cosmicfishpie_b200.fisher_matrix import confidence_coefficient
    assert np.allclose(confidence_coefficient(0.6827) * s, g["sigma"], rtol=1e-12)


@pytest.mark.parametrize("tag,obs,free,settings,spec", [
    ("gcph_small", ["GCph"], {"Omegam": 0.01, "sigma8": 0.01}, {"ell_sampling": 20}, "Euclid-Photometric-ISTF-Pessimistic"),
    ("wl_cfg2", ["WL"], {p: 0.01 for p in P7[:5]}, None, "Euclid-Photometric-ISTF-Pessimistic"),
    ("photo_3x2pt", ["WL", "GCph"], {p: 0.01 for p in P7}, None, "Euclid-Photometric-ISTF-Pessimistic"),
    ("photo_cfg3", ["WL", "GCph"], {p: 0.01 for p in P7}, None, "Euclid-Photometric-13bins-lmax5000"),
])
def test_photo_oracle_bit_exact(obank, tag, obs, free, settings, spec):
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.photo import default_photo_bias, finish_photo_specs, photo_fisher

    sp = finish_photo_specs(yaml.safe_load(open(os.path.join(SPECDIR, spec + ".yaml")))["specifications"])
    lum = np.load(os.path.join(REPO, "cosmicfishpie_b200", "data", "lumratio.npy"))
    bias = default_photo_bias(sp) if "GCph" in obs else {}
    ia = dict(sp["IA_params"])
    fp = dict(free)
    if "GCph" in obs:
        for k in bias:
            if k != "bias_model":
                fp.setdefault(k, sp["vary_ph_bias"])
    if "WL" in obs:
        for k in ia:
            if k != "IA_model":
                fp.setdefault(k, sp["vary_IA_pars"])
    r = photo_fisher(obank, sp, settings, obs, dict(tt.FIDUCIAL), dict(sp["photo_z_params"]), ia, bias, fp, lum,
                     return_all=True)
    g = golden(tag)
    assert r["param_names"] == list(g["param_names"])
    assert np.array_equal(r["fisher"], g["fisher"])
    assert np.array_equal(np.array([r["cls"][k] for k in g["cl_keys"]]), g["cls"])
    assert np.array_equal(r["cov"], g["covmat"])
    assert r["cols"] == list(g["cov_cols"])
    for f in g.files:
        if f.startswith("dcl__"):
            assert np.array_equal(np.array([r["derivs"][f[5:]][k] for k in g["cl_keys"]]), g[f]), f


SMALL = {"spectro_Pk_k_samples": 65, "spectro_Pk_mu_samples": 9}
F3 = {"Omegam": 0.01, "h": 0.01, "w0": 0.01}


@pytest.mark.parametrize("tag,settings,spec_over,yaml_name", [
    ("gcsp_var_noap_nofog", dict(SMALL, AP_effect=False, FoG_switch=False), {}, None),
    ("gcsp_var_linear", dict(SMALL, GCsp_linear=True), {}, None),
    ("gcsp_var_bfs8", dict(SMALL, bfs8terms=True), {}, None),
    ("gcsp_var_dzdep", dict(SMALL), {"spec_sigma_dz_type": "z-dependent", "spec_sigma_dz": 0.001}, None),
    ("gcsp_var_nlcosmo", dict(SMALL, fix_cosmo_nl_terms=False), {}, None),
    ("gcsp_var_qbias", dict(SMALL), {}, "Euclid-Spectroscopic-ISTF-Pessimistic-Qbias"),
    ("gcsp_var_clustering", dict(SMALL, GCsp_Tracer="clustering", bfs8terms=True), {}, None),
    ("gcsp_var_sigmapv", dict(SMALL), {"nonlinear_parametrization": dict(
        rescale_sigmap=True, rescale_sigmav=True, default=None,
        rescale_sigma_pv={f"sigma{c}_{i}": 1.0 for c in "pv" for i in (1, 2, 3, 4)})},
     "Euclid-Spectroscopic-ISTF-Pessimistic-sigma_pv"),
])
def test_spectro_oracle_variants_bit_exact(obank, tag, settings, spec_over, yaml_name):
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.spectro import spectro_fisher

    sp = _spectro_specs()  # defaults first, then the named survey (config.py:481-567)
    if yaml_name:
        sp.update(yaml.safe_load(open(os.path.join(SPECDIR, yaml_name + ".yaml")))["specifications"])
    sp.update(spec_over)
    bias = dict(sp["sp_bias_parametrization"][sp["sp_bias_model"]])
    ps = dict(sp["shot_noise_parametrization"][sp["shot_noise_model"]])
    fp = dict(F3)
    for k in list(bias) + list(ps):
        fp.setdefault(k, 1e-4)
    nonlin = {}
    if sp.get("nonlinear_model", "default") == "rescale_sigma_pv":  # config.py:722-732
        nonlin = dict(sp["nonlinear_parametrization"]["rescale_sigma_pv"])
        for k in nonlin:
            fp.setdefault(k, 0.01)
    r = spectro_fisher(obank, sp, settings, dict(tt.FIDUCIAL), bias, ps, nonlin, fp, max_z_bins=2, return_all=True)
    g = golden(tag)
    assert r["param_names"] == list(g["param_names"])
    assert np.array_equal(r["fisher"], g["fisher"])
    assert np.array_equal(np.array([x[:, ::16] for x in r["lnp_fid"]]), g["lnpobs_fid_strided"])


@pytest.mark.parametrize("tag,obs,free,settings,spec_over", [
    ("photo_var_linear_pk", ["WL", "GCph"], {"Omegam": 0.01, "h": 0.01}, {"ell_sampling": 16, "nonlinear": False}, {}),
    ("photo_var_sqrtbias", ["GCph"], {"sigma8": 0.01}, {"ell_sampling": 16}, {"ph_bias_model": "sqrt"}),
    ("photo_var_gcfirst", ["GCph", "WL"], {"Omegam": 0.01}, {"ell_sampling": 12, "pivot_z_IA": 0.3}, {}),
    ("photo_var_clustering", ["WL", "GCph"], {"Omegam": 0.01, "h": 0.01}, {"ell_sampling": 12, "GCph_Tracer": "clustering"}, {}),
    ("photo_var_photoz", ["WL", "GCph"], {"Omegam": 0.01, "sigma_b": 0.02, "zb": 0.002, "fout": 0.02},
     {"ell_sampling": 12}, {}),
])
def test_photo_oracle_variants_bit_exact(obank, tag, obs, free, settings, spec_over):
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.photo import default_photo_bias, finish_photo_specs, photo_fisher

    raw = yaml.safe_load(open(os.path.join(SPECDIR, "Euclid-Photometric-ISTF-Pessimistic.yaml")))["specifications"]
    sp = finish_photo_specs(raw)
    sp.update(spec_over)
    lum = np.load(os.path.join(REPO, "cosmicfishpie_b200", "data", "lumratio.npy"))
    bias = default_photo_bias(sp) if "GCph" in obs else {}
    ia = dict(sp["IA_params"])
    fp = dict(free)
    if "GCph" in obs:
        for k in bias:
            if k != "bias_model":
                fp.setdefault(k, sp["vary_ph_bias"])
    if "WL" in obs:
        for k in ia:
            if k != "IA_model":
                fp.setdefault(k, sp["vary_IA_pars"])
    r = photo_fisher(obank, sp, settings, obs, dict(tt.FIDUCIAL), dict(sp["photo_z_params"]), ia, bias, fp, lum,
                     return_all=True)
    g = golden(tag)
    assert r["param_names"] == list(g["param_names"])
    assert r["cols"] == list(g["cov_cols"])
    assert np.array_equal(np.array([r["cls"][k] for k in g["cl_keys"]]), g["cls"])
    assert np.array_equal(r["fisher"], g["fisher"])


def test_photo_oracle_stem_bit_exact():
    """SteM derivatives: the reference's derivative engine run on its own getcls with +-1, 2, 3 % tables, its
    einsum assembly for the matrix (tools/make_goldens.py::photo_stem)."""
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.cosmo import Bank
    from oracle.photo import default_photo_bias, finish_photo_specs, photo_fisher

    g = golden("photo_stem")
    eps = tuple(g["eps_values"])
    ob = Bank(tt.make_bank(eps_values=eps, pars=("Omegam", "h")), tt.FIDUCIAL, tt.COSMO_KEYS, eps_values=eps)
    sp = finish_photo_specs(yaml.safe_load(open(os.path.join(SPECDIR, "Euclid-Photometric-ISTF-Pessimistic.yaml")))["specifications"])
    lum = np.load(os.path.join(REPO, "cosmicfishpie_b200", "data", "lumratio.npy"))
    bias, ia = default_photo_bias(sp), dict(sp["IA_params"])
    fp = {"Omegam": 0.01, "h": 0.01}
    fp.update({k: sp["vary_ph_bias"] for k in bias if k != "bias_model"})
    fp.update({k: sp["vary_IA_pars"] for k in ia if k != "IA_model"})
    r = photo_fisher(ob, sp, {"ell_sampling": 10}, ["WL", "GCph"], dict(tt.FIDUCIAL), dict(sp["photo_z_params"]), ia, bias,
                     fp, lum, return_all=True, stencil="STEM")
    assert r["param_names"] == list(g["param_names"])
    for par in ("Omegam", "b3", "AIA"):
        assert np.array_equal(np.array([r["derivs"][par][k] for k in g["cl_keys"]]), g["dcl__" + par])
    assert np.array_equal(r["fisher"], g["fisher"])


def test_spectro_oracle_bias_overrides_bit_exact(obank):
    """observed_Pgg with b_i and with use_bias_funcs=True, between the bin centres (golden gcsp_kat_bias)."""
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.spectro import DEFAULT_SETTINGS, GalSpectro

    g = golden("gcsp_kat_bias")
    sp = _spectro_specs()
    bias = dict(sp["sp_bias_parametrization"]["linear_log"])
    cp = dict(tt.FIDUCIAL)
    cp["h"] = tt.FIDUCIAL["h"] * 1.01
    fid = obank.cosmo(dict(tt.FIDUCIAL))
    kk, mm = np.meshgrid(g["k"], g["mu"])
    plain = GalSpectro(obank.cosmo(cp), fid, sp, dict(DEFAULT_SETTINGS), bias, {})
    funcs = GalSpectro(obank.cosmo(cp), fid, sp, dict(DEFAULT_SETTINGS), bias, {}, use_bias_funcs=True)
    for i, z in enumerate(g["z"]):
        assert np.array_equal(plain.observed_Pgg(z, kk, mm), g["pgg_default"][i])
        assert np.array_equal(plain.observed_Pgg(z, kk, mm, b_i=float(g["b_i"])), g["pgg_b_i"][i])
        assert np.array_equal(funcs.observed_Pgg(z, kk, mm), g["pgg_bias_funcs"][i])


LIKE_PHOTO = {"fid": {}, "nuis": {"b1": 1.02, "b7": 0.985, "AIA": 1.04, "etaIA": 0.97, "betaIA": 1.01},
              "cosmo": {"Omegam": 1.01}, "both": {"h": 0.99, "b3": 1.03, "AIA": 0.95}}
LIKE_SPECTRO = {"fid": {}, "nuis": {"lnbg_1": 1.01, "lnbg_4": 0.99}, "cosmo": {"h": 1.01},
                "shot": {"Ps_2": ("abs", 35.0), "lnbg_3": 1.005, "w0": 0.99}}


def test_photo_likelihood_oracle_bit_exact(obank):
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.likelihood import PhotoLikelihood
    from oracle.photo import default_photo_bias, finish_photo_specs

    sp = finish_photo_specs(yaml.safe_load(open(os.path.join(SPECDIR, "Euclid-Photometric-ISTF-Pessimistic.yaml")))["specifications"])
    lum = np.load(os.path.join(REPO, "cosmicfishpie_b200", "data", "lumratio.npy"))
    bias, ia = default_photo_bias(sp), dict(sp["IA_params"])
    L = PhotoLikelihood(obank, sp, None, ["WL", "GCph"], dict(tt.FIDUCIAL), dict(sp["photo_z_params"]), ia, bias, lum)
    g = golden("like_photo")
    for k in ("Cell_LL", "Cell_GG", "Cell_GL"):
        assert np.array_equal(L.data[k], g[k])
    allp = {**tt.FIDUCIAL, **bias, **ia}
    for name, ref in zip(g["point_names"], g["loglike"]):
        par = {k: allp[k] * v for k, v in LIKE_PHOTO[str(name)].items()}
        assert L.loglike(par) == ref, name


def test_spectro_likelihood_oracle_bit_exact(obank):
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.likelihood import SpectroLikelihoodOracle

    sp = _spectro_specs()
    bias = dict(sp["sp_bias_parametrization"]["linear_log"])
    ps = dict(sp["shot_noise_parametrization"]["constant"])
    L = SpectroLikelihoodOracle(obank, sp, None, dict(tt.FIDUCIAL), bias, ps, {})
    g = golden("like_spectro")
    assert np.array_equal(L.data[:, :, ::16], g["data_strided"])
    pld = L.legendre(L.data)
    cov, icov = L.legendre_cov(pld)
    assert np.array_equal(pld[::16], g["pl_data_strided"]) and np.array_equal(cov[::16], g["cov_strided"])
    allp = {**tt.FIDUCIAL, **bias, **ps}
    for i, name in enumerate(g["point_names"]):
        par = {k: (v[1] if isinstance(v, tuple) else allp[k] * v) for k, v in LIKE_SPECTRO[str(name)].items()}
        th = L.theory(par)
        assert L.wedge_chi2(th) == g["chi2_wedges"][i]
        assert L.legendre_chi2(pld, L.legendre(th), icov) == g["chi2_legendre"][i]
    # the backend-free known answer of the reference's own test (tests/spectro_like_test.py:60-73): 5/(4 pi^2)


# ---------------------------------------------------------------- the workloads bench.py times (round-2 goldens)
FINE = {"spectro_Pk_k_samples": 8193, "spectro_Pk_mu_samples": 129}


def test_spectro_oracle_fine_lattice_bit_exact(obank):
    """GCsp on the bench lattice (129 x 8193, 15 parameters): golden gcsp_cfg4_fine of the unmodified reference."""
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.spectro import spectro_fisher

    sp = _spectro_specs()
    bias = dict(sp["sp_bias_parametrization"]["linear_log"])
    ps = dict(sp["shot_noise_parametrization"]["constant"])
    fp = {p: 0.01 for p in P7}
    fp.update({k: 1e-4 for k in list(bias) + list(ps)})
    r = spectro_fisher(obank, sp, dict(FINE), dict(tt.FIDUCIAL), bias, ps, {}, fp, return_all=True)
    g = golden("gcsp_cfg4_fine")
    assert r["param_names"] == list(g["param_names"])
    assert np.array_equal(r["fisher"], g["fisher"])
    assert np.array_equal(r["fisher_per_bin"], g["fisher_per_bin"])
    assert np.array_equal(np.array([x[::8, ::64] for x in r["lnp_fid"]]), g["lnpobs_fid_strided"])
    assert np.array_equal(np.array([x[::8, ::64] for x in r["veff"]]), g["veff_strided"])
    for par in ("Omegam", "wa", "lnbg_2", "Ps_4"):
        d = np.array([(np.asarray(r["derivs"][par][i]) * np.ones((129, 8193)))[::8, ::64] for i in range(4)])
        assert np.array_equal(d, g["deriv_strided__" + par]), par


def test_photo_likelihood_oracle_cfg3_bit_exact(obank):
    """13+13-bin photometric likelihood (golden like_photo_cfg3): data cells and a third of the stored points."""
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.likelihood import PhotoLikelihood
    from oracle.photo import default_photo_bias, finish_photo_specs

    sp = finish_photo_specs(yaml.safe_load(open(os.path.join(SPECDIR, "Euclid-Photometric-13bins-lmax5000.yaml")))["specifications"])
    lum = np.load(os.path.join(REPO, "cosmicfishpie_b200", "data", "lumratio.npy"))
    bias, ia = default_photo_bias(sp), dict(sp["IA_params"])
    L = PhotoLikelihood(obank, sp, None, ["WL", "GCph"], dict(tt.FIDUCIAL), dict(sp["photo_z_params"]), ia, bias, lum)
    g = golden("like_photo_cfg3")
    for k in ("Cell_LL", "Cell_GG", "Cell_GL"):
        assert np.array_equal(L.data[k], g[k])
    keys = [str(k) for k in g["keys"]]
    for i in list(range(0, 40, 3)):
        assert L.loglike(dict(zip(keys, g["points"][i]))) == g["loglike"][i], i


def test_spectro_likelihood_oracle_fine_bit_exact(obank):
    """Spectroscopic likelihood on the 129 x 8193 lattice (golden like_spectro_fine): data, multipoles and six points."""
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.likelihood import SpectroLikelihoodOracle

    sp = _spectro_specs()
    bias = dict(sp["sp_bias_parametrization"]["linear_log"])
    ps = dict(sp["shot_noise_parametrization"]["constant"])
    L = SpectroLikelihoodOracle(obank, sp, dict(FINE), dict(tt.FIDUCIAL), bias, ps, {})
    g = golden("like_spectro_fine")
    assert np.array_equal(L.data[:, ::8, ::64], g["data_strided"])
    pld = L.legendre(L.data)
    cov, icov = L.legendre_cov(pld)
    assert np.array_equal(pld[::64], g["pl_data_strided"]) and np.array_equal(cov[::64], g["cov_strided"])
    keys = [str(k) for k in g["keys"]]
    for i in (0, 3, 9, 14, 18, 21):
        th = L.theory(dict(zip(keys, g["points"][i])))
        if i == 9:
            assert np.array_equal(th[:, ::8, ::64], g["theory_9_strided"])
        assert L.wedge_chi2(th) == g["chi2_wedges"][i], i
        assert L.legendre_chi2(pld, L.legendre(th), icov) == g["chi2_legendre"][i], i
