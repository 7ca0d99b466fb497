"""GPU parity: CUDA photometric path (through the C ABI) vs goldens of the unmodified reference."""
import numpy as np
import pytest

from conftest import base_options, elem_rel_err, golden, rel_err

pytestmark = pytest.mark.gpu

P7 = ["Omegam", "Omegab", "h", "ns", "sigma8", "w0", "wa"]
CASES = {
    "gcph_small": dict(obs=["GCph"], free={"Omegam": 0.01, "sigma8": 0.01}, opts={"ell_sampling": 20}),
    "wl_cfg2": dict(obs=["WL"], free={p: 0.01 for p in P7[:5]}, opts={}),
    "photo_3x2pt": dict(obs=["WL", "GCph"], free={p: 0.01 for p in P7}, opts={}),
    # BASELINE.json configs[2]: 13+13 bins, lmax 5000, 23 parameters (custom spec shipped in data/specs)
    "photo_cfg3": dict(obs=["WL", "GCph"], free={p: 0.01 for p in P7}, opts={},
                       spec="Euclid-Photometric-13bins-lmax5000"),
}


CASES.update({  # goldens: tools/make_goldens.py PHOTO_VARIANTS
    "photo_var_linear_pk": dict(obs=["WL", "GCph"], free={"Omegam": 0.01, "h": 0.01},
                                opts={"ell_sampling": 16, "nonlinear": False}),
    "photo_var_sqrtbias": dict(obs=["GCph"], free={"sigma8": 0.01}, opts={"ell_sampling": 16},
                               specs={"ph_bias_model": "sqrt"}),
    "photo_var_gcfirst": dict(obs=["GCph", "WL"], free={"Omegam": 0.01}, opts={"ell_sampling": 12, "pivot_z_IA": 0.3}),
    # GCph on the cdm+baryon ("clustering") spectrum, lensing on total matter
    "photo_var_clustering": dict(obs=["WL", "GCph"], free={"Omegam": 0.01, "h": 0.01},
                                 opts={"ell_sampling": 12, "GCph_Tracer": "clustering"}),
    # photo-z parameters free: their stencil points carry their own n_i(z) windows and lensing efficiencies
    "photo_var_photoz": dict(obs=["WL", "GCph"], free={"Omegam": 0.01, "sigma_b": 0.02, "zb": 0.002, "fout": 0.02},
                             opts={"ell_sampling": 12}),
})


def run_case(name, extfiles, tmp_path):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    c = CASES[name]
    opts = base_options(tmp_path, survey_name_photo=c.get("spec", "Euclid-Photometric-ISTF-Pessimistic"),
                        survey_name_spectro=False, **c["opts"])
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars=dict(c["free"]), options=opts, observables=list(c["obs"]),
                      specifications=dict(c.get("specs", {})), extfiles=extfiles, cosmoModel="w0waCDM", surveyName="Euclid")
    return fm, fm.compute()


@pytest.mark.parametrize("name", list(CASES))
def test_fisher_matches_reference(name, extfiles, tmp_path):
    g = golden(name)
    fm, F = run_case(name, extfiles, tmp_path)
    assert list(F.get_param_names()) == list(g["param_names"])
    # north_star tolerance: 1e-6 on Fisher elements, 1e-4 on marginalised 1-sigma errors
    err = elem_rel_err(F.fisher_matrix, g["fisher"])
    assert err < 1e-7, err
    serr = np.max(np.abs(F.get_confidence_bounds() - g["sigma"]) / g["sigma"])
    assert serr < 1e-5, serr


@pytest.mark.parametrize("name", list(CASES))
def test_cls_cov_derivs_match_reference(name, extfiles, tmp_path):
    g = golden(name)
    fm, _ = run_case(name, extfiles, tmp_path)
    noisy, cov = fm.photo_LSS.compute_covmat()
    assert np.array_equal(noisy["ells"], g["ells"])
    keys = list(g["cl_keys"])
    cl = np.array([fm.photo_LSS.cls[k] for k in keys])
    # per-spectrum relative error (cross spectra with strong cancellation are compared on their own scale)
    for k, a, b in zip(keys, cl, g["cls"]):
        assert rel_err(a, b) < 1e-12, k
    assert list(fm.photo_LSS.cov_cols) == list(g["cov_cols"])
    assert rel_err(cov, g["covmat"]) < 1e-13
    derivs = fm.photo_derivs
    for f in g.files:
        if f.startswith("dcl__"):
            par = f[5:]
            d = np.array([derivs[par][k] for k in keys])
            assert rel_err(d, g[f]) < 1e-9, par


def test_sqrtp_table_matches_reference(extfiles, tmp_path):
    g = golden("wl_cfg2")
    fm, _ = run_case("wl_cfg2", extfiles, tmp_path)
    fm.photo_obs_fid.compute_all()
    sq = fm.photo_obs_fid.sqrtPell["WL"][::8, ::4]
    assert np.array_equal(sq == 0, g["sqrtP_WL"] == 0)  # identical kappa < kmax mask
    assert rel_err(sq, g["sqrtP_WL"]) < 1e-14
