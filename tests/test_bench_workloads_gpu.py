"""GPU parity on the workloads bench.py times, against goldens of the UNMODIFIED reference
(tools/make_goldens.py cases gcsp_cfg4_fine, like_photo_cfg3, like_spectro_fine; round 2):

* GCsp Fisher, 4 bins, 15 parameters, 129 x 8193 lattice (reference: fishermatrix/cosmicfish.py:366-542);
* photometric 3x2pt likelihood, 13+13 bins, lmax 5000, 40 points (likelihood/photo_like.py:180-256);
* spectroscopic wedge and multipole likelihoods on the 129 x 8193 lattice, 22 points in groups that share a
  cosmology -- the moment kernels' case -- against the reference itself (likelihood/spectro_like.py:143-233).

Tolerances (north_star asks 1e-6 on Fisher elements, 1e-4 on sigma): 1e-8 on elements, 1e-9 on log-likelihoods
(1e-7 for the multipole chi2 whose 3x3 covariances have cond ~ 1e6)."""
import numpy as np
import pytest

from conftest import base_options, elem_rel_err, golden, rel_err

pytestmark = pytest.mark.gpu

P7 = ["Omegam", "Omegab", "h", "ns", "sigma8", "w0", "wa"]
FINE = {"spectro_Pk_k_samples": 8193, "spectro_Pk_mu_samples": 129}


@pytest.fixture(scope="module")
def fine_fm(extfiles, tmp_path_factory):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7},
                      options=base_options(tmp_path_factory.mktemp("fine"), **FINE), observables=["GCsp"],
                      extfiles=extfiles)
    fm.result = fm.compute()
    return fm


@pytest.fixture(scope="module")
def cfg3_fm(extfiles, tmp_path_factory):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    o = base_options(tmp_path_factory.mktemp("cfg3"), survey_name_photo="Euclid-Photometric-13bins-lmax5000",
                     survey_name_spectro=False)
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7}, options=o,
                      observables=["WL", "GCph"], extfiles=extfiles)
    fm.compute()
    return fm


def test_gcsp_fine_lattice_fisher_matches_reference(fine_fm):
    g = golden("gcsp_cfg4_fine")
    F = fine_fm.result
    assert list(F.get_param_names()) == list(g["param_names"])
    err = elem_rel_err(F.fisher_matrix, g["fisher"])
    assert err < 1e-8, err
    serr = np.max(np.abs(F.get_confidence_bounds() - g["sigma"]) / g["sigma"])
    assert serr < 1e-6, serr
    perbin = elem_rel_err(fine_fm.pk_LSS_Fisher(len(fine_fm.zmids)), g["fisher_per_bin"])
    assert perbin < 1e-8, perbin


def test_gcsp_fine_lattice_lattices_match_reference(fine_fm):
    g = golden("gcsp_cfg4_fine")
    fm = fine_fm
    sl = (slice(None, None, 8), slice(None, None, 64))
    lnp = np.array([fm.pk_obs_fid.lnpobs_gg(z, fm.Pk_kmesh, fm.Pk_mumesh)[sl] for z in fm.zmids])
    assert np.max(np.abs(lnp - g["lnpobs_fid_strided"])) < 1e-12
    assert rel_err(np.array([v[sl] for v in fm.veff_arr]), g["veff_strided"]) < 1e-12
    for par in fm.freeparams:
        d = np.array([np.asarray(fm.derivs_dict[par][i])[sl] for i in range(len(fm.zmids))])
        ref = g["deriv_strided__" + par]
        assert np.max(np.abs(d - ref)) < 1e-8 * max(1.0, np.max(np.abs(ref))), par


def test_photo_likelihood_cfg3_matches_reference(cfg3_fm):
    from cosmicfishpie_b200.likelihood import PhotometricLikelihood

    g = golden("like_photo_cfg3")
    like = PhotometricLikelihood(cosmo_data=cfg3_fm)
    for k in ("Cell_LL", "Cell_GG", "Cell_GL"):
        assert rel_err(like.data_obs[k], g[k]) < 1e-12, k
    keys = [str(k) for k in g["keys"]]
    ref = g["loglike"]
    batch = like.loglike_batch(g["points"], keys=keys)
    assert abs(batch[0]) < 1e-6  # the fiducial: rounding noise in the reference too (4e-9)
    assert np.max(np.abs(batch[1:] - ref[1:]) / np.abs(ref[1:])) < 1e-9
    # the one-object-per-point route on a few of them
    for i in (3, 12, 20, 33):
        one = like.loglike(param_dict=dict(zip(keys, g["points"][i])))
        assert abs(one - ref[i]) / abs(ref[i]) < 1e-9, i


def test_spectro_likelihood_fine_lattice_matches_reference(fine_fm):
    """The groups of 4-8 points go through the moment kernels: compared with the reference, not with the direct kernels."""
    from cosmicfishpie_b200 import likelihood as lk

    g = golden("like_spectro_fine")
    keys = [str(k) for k in g["keys"]]
    sl = (slice(None), slice(None, None, 8), slice(None, None, 64))
    like = lk.SpectroLikelihood(cosmoFM_data=fine_fm)
    assert rel_err(like.data_obs[sl], g["data_strided"]) < 1e-13
    ref = g["chi2_wedges"]
    chi2 = like.chi2_matrix(keys, g["points"])
    assert chi2[0] < 1e-18 and ref[0] == 0.0
    assert np.max(np.abs(chi2[1:] - ref[1:]) / ref[1:]) < 1e-9
    ll = like.loglike_batch(g["points"], keys=keys)
    assert np.max(np.abs(ll[1:] + 0.5 * ref[1:]) / ref[1:]) < 1e-9
    th = like.compute_theory(dict(zip(keys, g["points"][9])))
    assert rel_err(th[sl], g["theory_9_strided"]) < 1e-13
    # multipoles
    leg = lk.SpectroLikelihood(cosmoFM_data=fine_fm, leg_flag="legendre")
    assert rel_err(leg.data_obs[::64], g["pl_data_strided"]) < 1e-13
    refl = g["chi2_legendre"]
    chi2l = leg.chi2_matrix(keys, g["points"])
    assert chi2l[0] < 1e-15
    assert np.max(np.abs(chi2l[1:] - refl[1:]) / refl[1:]) < 1e-7
