"""GPU parity: batched photometric and spectroscopic likelihoods vs goldens of the unmodified reference
(tools/make_goldens.py cases like_photo / like_spectro)."""
import numpy as np
import pytest

from conftest import base_options, golden, rel_err

pytestmark = pytest.mark.gpu

P7 = ["Omegam", "Omegab", "h", "ns", "sigma8", "w0", "wa"]
POINTS_PHOTO = {
    "fid": {},
    "nuis": {"b1": 1.02, "b7": 0.985, "AIA": 1.04, "etaIA": 0.97, "betaIA": 1.01},
    "cosmo": {"Omegam": 1.01},
    "both": {"h": 0.99, "b3": 1.03, "AIA": 0.95},
}
POINTS_SPECTRO = {
    "fid": {},
    "nuis": {"lnbg_1": 1.01, "lnbg_4": 0.99},
    "cosmo": {"h": 1.01},
    "shot": {"Ps_2": ("abs", 35.0), "lnbg_3": 1.005, "w0": 0.99},
}


def shift(allp, sh):
    return {k: (v[1] if isinstance(v, tuple) else allp[k] * v) for k, v in sh.items()}


@pytest.fixture(scope="module")
def photo_fm(extfiles, tmp_path_factory):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    o = base_options(tmp_path_factory.mktemp("lp"), survey_name_photo="Euclid-Photometric-ISTF-Pessimistic",
                     survey_name_spectro=False)
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7}, options=o,
                      observables=["WL", "GCph"], extfiles=extfiles)
    fm.compute()
    return fm


@pytest.fixture(scope="module")
def spectro_fm(extfiles, tmp_path_factory):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7},
                      options=base_options(tmp_path_factory.mktemp("ls")), observables=["GCsp"], extfiles=extfiles)
    fm.compute()
    return fm


def test_photo_likelihood_matches_reference(photo_fm):
    from cosmicfishpie_b200.likelihood import PhotometricLikelihood

    g = golden("like_photo")
    like = PhotometricLikelihood(cosmo_data=photo_fm)
    for k in ("Cell_LL", "Cell_GG", "Cell_GL"):
        assert rel_err(like.data_obs[k], g[k]) < 1e-12, k
    assert np.array_equal(like.data_obs["ells"], g["ells"])
    allp = dict(photo_fm.allparams)
    dicts = [shift(allp, POINTS_PHOTO[str(n)]) for n in g["point_names"]]
    single = np.array([like.loglike(param_dict=d) for d in dicts])
    keys = sorted({k for d in dicts for k in d})
    mat = np.array([[d.get(k, allp[k]) for k in keys] for d in dicts])
    batch = like.loglike_batch(mat, keys=keys)
    ref = g["loglike"]
    assert abs(single[0]) < 1e-7 and abs(batch[0]) < 1e-7  # fiducial: reference gives 4.4e-11 (rounding noise)
    assert np.max(np.abs(single[1:] - ref[1:]) / np.abs(ref[1:])) < 1e-9
    assert np.max(np.abs(batch[1:] - ref[1:]) / np.abs(ref[1:])) < 1e-9
    # operator seam: chi2 of an explicitly computed theory vector
    th = like.compute_theory(dicts[1])
    assert abs(-0.5 * like.compute_chi2(th) - ref[1]) / abs(ref[1]) < 1e-9


def test_photo_likelihood_large_batch_is_consistent(photo_fm):
    from cosmicfishpie_b200.likelihood import PhotometricLikelihood

    like = PhotometricLikelihood(cosmo_data=photo_fm)
    rng = np.random.default_rng(0)
    keys = [f"b{i}" for i in range(1, 11)] + ["AIA", "betaIA", "etaIA"]
    fid = np.array([photo_fm.allparams[k] for k in keys])
    mat = fid * (1 + 0.03 * rng.standard_normal((300, len(keys))))
    mat[0] = fid
    ll = like.loglike_batch(mat, keys=keys)
    assert ll.shape == (300,) and np.all(np.isfinite(ll)) and abs(ll[0]) < 1e-7 and np.all(ll[1:] < 0)
    again = like.loglike_batch(mat[[5, 17, 123]], keys=keys)
    assert np.array_equal(again, ll[[5, 17, 123]])  # deterministic, independent of batch composition


def test_spectro_likelihood_matches_reference(spectro_fm):
    from cosmicfishpie_b200 import likelihood as lk

    g = golden("like_spectro")
    like = lk.SpectroLikelihood(cosmoFM_data=spectro_fm)
    assert rel_err(like.data_obs[:, :, ::16], g["data_strided"]) < 1e-13
    allp = dict(spectro_fm.allparams)
    dicts = [shift(allp, POINTS_SPECTRO[str(n)]) for n in g["point_names"]]
    ref = g["chi2_wedges"]
    chi2_batch = like.chi2_batch(dicts)
    assert chi2_batch[0] < 1e-18 and ref[0] == 0.0
    assert np.max(np.abs(chi2_batch[1:] - ref[1:]) / ref[1:]) < 1e-10
    th = like.compute_theory(dicts[3])
    assert rel_err(th[:, :, ::16], g["theory_shot_strided"]) < 1e-13
    assert abs(like.compute_chi2(th) - ref[3]) / ref[3] < 1e-10
    assert abs(like.loglike(param_dict=dicts[2]) + 0.5 * ref[2]) / ref[2] < 1e-10
    assert abs(lk.loglike(param_vec=dicts[1], cosmoFM_data=spectro_fm) + 0.5 * ref[1]) / ref[1] < 1e-10


def test_matrix_fast_paths_agree_with_point_by_point(photo_fm, spectro_fm):
    """loglike_batch assembles the per-point inputs with array operations; it must agree with the
    one-object-per-point path (which is the one checked against the reference above)."""
    from cosmicfishpie_b200 import likelihood as lk

    rng = np.random.default_rng(7)
    Lp = lk.PhotometricLikelihood(cosmo_data=photo_fm)
    keys = ["Omegam", "b2", "b9", "AIA", "etaIA", "betaIA", "not_a_parameter"]
    fid = np.array([photo_fm.allparams.get(k, 1.0) for k in keys])
    mat = fid * (1 + 0.02 * rng.standard_normal((12, len(keys))))
    mat[:, 0] = fid[0] * rng.choice([0.99, 1.0, 1.01], size=12)  # tabulated cosmologies
    a = Lp.loglike_batch(mat, keys=keys)
    b = -0.5 * Lp.chi2_batch([dict(zip(keys, r)) for r in mat])
    assert np.max(np.abs(a - b) / np.abs(b)) < 1e-10
    Ls = lk.SpectroLikelihood(cosmoFM_data=spectro_fm)
    keys = ["h", "lnbg_1", "lnbg_3", "Ps_2", "Ps_4"]
    fid = np.array([spectro_fm.allparams[k] for k in keys])
    mat = np.tile(fid, (10, 1))
    mat[:, 1:3] *= 1 + 0.01 * rng.standard_normal((10, 2))
    mat[:, 3:] = rng.uniform(0, 40, size=(10, 2))
    mat[:, 0] = fid[0] * rng.choice([0.99, 1.0, 1.01], size=10)
    a = Ls.loglike_batch(mat, keys=keys)
    b = -0.5 * Ls.chi2_batch([dict(zip(keys, r)) for r in mat])
    assert np.max(np.abs(a - b) / np.abs(b)) < 1e-12


def test_spectro_legendre_matches_reference(spectro_fm):
    from cosmicfishpie_b200 import likelihood as lk

    g = golden("like_spectro")
    like = lk.SpectroLikelihood(cosmoFM_data=spectro_fm, leg_flag="legendre")
    assert rel_err(like.data_obs[::16], g["pl_data_strided"]) < 1e-13
    cov, icov = lk.compute_covariance_legendre(like.data_obs, spectro_fm)
    for a in range(3):
        for b in range(3):
            assert rel_err(cov[::16, :, a, b], g["cov_strided"][:, :, a, b]) < 1e-12, (a, b)
    allp = dict(spectro_fm.allparams)
    dicts = [shift(allp, POINTS_SPECTRO[str(n)]) for n in g["point_names"]]
    ref = g["chi2_legendre"]
    chi2 = like.chi2_batch(dicts)
    assert chi2[0] < 1e-15
    # the 3x3 multipole covariances are ill-conditioned (cond ~ 1e6), so inverse-covariance products agree to ~1e-9
    assert np.max(np.abs(chi2[1:] - ref[1:]) / ref[1:]) < 1e-7
    th = like.compute_theory(dicts[2])
    assert abs(like.compute_chi2(th) - ref[2]) / ref[2] < 1e-7
    # the array fast path (fused multipole kernel, points grouped by shared blocks) against the point-by-point route
    keys = ["lnbg_1", "lnbg_2", "lnbg_3", "lnbg_4", "Ps_2"]
    fid = np.array([allp[k] for k in keys])
    mat = fid + np.array([0.01, 0.01, 0.01, 0.01, 5.0]) * np.random.default_rng(2).standard_normal((21, len(keys)))
    fast = like.chi2_matrix(keys, mat)
    slow = np.array([like.compute_chi2(like.compute_theory(dict(zip(keys, row)))) for row in mat[:5]])
    assert np.max(np.abs(fast[:5] - slow) / slow) < 1e-9


def test_photo_batch_pipeline_equals_single_chunks(photo_fm):
    """Batches longer than one chunk run as a two-stage pipeline (worker thread + second stream): same numbers as
    the chunks evaluated one by one."""
    from cosmicfishpie_b200.likelihood import PhotometricLikelihood

    like = PhotometricLikelihood(cosmo_data=photo_fm)
    keys = ["b1", "b5", "AIA", "etaIA"]
    fid = np.array([photo_fm.allparams[k] for k in keys])
    mat = fid * (1 + 0.02 * np.random.default_rng(11).standard_normal((4096 + 700, len(keys))))
    whole = like.loglike_batch(mat, keys=keys)
    parts = np.concatenate([like.loglike_batch(mat[:4096], keys=keys), like.loglike_batch(mat[4096:], keys=keys)])
    assert np.array_equal(whole, parts)
    assert np.all(np.isfinite(whole)) and whole.max() <= 1e-9


@pytest.mark.parametrize("flag", ["wedges", "legendre"])
def test_spectro_moment_path_equals_direct_kernels(spectro_fm, monkeypatch, flag):
    """Groups of >= 4 points that differ only in bias and shot noise are evaluated through the moments of the group
    (one lattice pass per bin); CFP_NO_MOMENTS=1 sends the same batch through the direct kernels.  The two agree to
    rounding, a member equal to the group's reference exactly, and the fiducial's chi2 stays (numerically) zero."""
    from cosmicfishpie_b200 import likelihood as lk

    like = lk.SpectroLikelihood(cosmoFM_data=spectro_fm, leg_flag=flag)
    keys = ["h", "lnbg_1", "lnbg_2", "lnbg_3", "lnbg_4", "Ps_1", "Ps_3"]
    fid = np.array([spectro_fm.allparams[k] for k in keys])
    rng = np.random.default_rng(31)
    mat = np.tile(fid, (150, 1))
    mat[1:, 1:5] += 0.02 * rng.standard_normal((149, 4))
    mat[1:, 5:] = rng.uniform(0.0, 60.0, size=(149, 2))
    mat[100:, 0] = fid[0] * 1.01       # a second cosmology of the bank: its own groups
    mat[148:, 0] = fid[0] * 0.99       # ... and a group of 2 (direct kernels either way)
    mom = like.chi2_matrix(keys, mat)
    monkeypatch.setenv("CFP_NO_MOMENTS", "1")
    direct = like.chi2_matrix(keys, mat)
    monkeypatch.delenv("CFP_NO_MOMENTS")
    assert mom[0] < 1e-15 and direct[0] < 1e-15
    tol = 1e-11 if flag == "wedges" else 1e-8  # (the 3x3 multipole covariances are ill-conditioned, cond ~ 1e6)
    # the group of two takes the direct kernels either way (the split of the lattice over CTAs may differ)
    assert np.max(np.abs(mom[148:] - direct[148:]) / direct[148:]) < 1e-12
    assert np.max(np.abs(mom[1:] - direct[1:]) / direct[1:]) < tol
    # a large, bias-only batch (what bench.py times) stays finite and ordered like the direct evaluation
    big = np.tile(fid, (1500, 1))
    big[:, 1:5] *= 1 + 0.01 * rng.standard_normal((1500, 4))
    a = like.chi2_matrix(keys, big)
    monkeypatch.setenv("CFP_NO_MOMENTS", "1")
    b = like.chi2_matrix(keys, big[:64])
    monkeypatch.delenv("CFP_NO_MOMENTS")
    assert np.all(np.isfinite(a)) and np.max(np.abs(a[:64] - b) / b) < tol


def test_batch_edge_sizes(extfiles, tmp_path):
    """Empty batches, single points and a batch one point longer than a chunk: shapes, finiteness, and the last point
    equal to its point-by-point evaluation."""
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix
    from cosmicfishpie_b200.likelihood import PhotometricLikelihood, SpectroLikelihood

    fp = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01},
                      options=base_options(tmp_path / "p", survey_name_photo="Euclid-Photometric-ISTF-Pessimistic",
                                           survey_name_spectro=False, ell_sampling=12),
                      observables=["WL", "GCph"], extfiles=extfiles)
    fp.compute()
    fs = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01},
                      options=base_options(tmp_path / "s", spectro_Pk_k_samples=65, spectro_Pk_mu_samples=9),
                      observables=["GCsp"], extfiles=extfiles)
    fs.compute()
    rng = np.random.default_rng(0)
    for like, fm, big in ((PhotometricLikelihood(cosmo_data=fp), fp, 4097), (SpectroLikelihood(cosmoFM_data=fs), fs, 8193),
                          (SpectroLikelihood(cosmoFM_data=fs, leg_flag="legendre"), fs, 8193)):
        keys = [k for k in fm.freeparams if k != "Omegam"][:3]
        fid = np.array([fm.allparams[k] for k in keys])
        for B in (0, 1, 2, big):
            mat = fid * (1 + 0.01 * rng.standard_normal((B, len(keys))))
            ll = like.loglike_batch(mat, keys=keys)
            assert ll.shape == (B,) and np.all(np.isfinite(ll))
            if B:
                one = like.loglike(param_dict=dict(zip(keys, mat[-1])))
                assert abs(ll[-1] - one) <= 1e-9 * abs(one)
