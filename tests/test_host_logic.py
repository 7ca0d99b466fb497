"""CPU: host-side logic of the product (configuration, table selection, quadrature weights, stencils,
job assembly, result container) -- everything that runs without a GPU."""
import os

import numpy as np
import pytest
import yaml

from conftest import REPO, base_options


def make_fm(extfiles, tmp_path, observables, free, **opts):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    o = base_options(tmp_path, **opts)
    if "GCsp" not in observables:
        o.update(survey_name_photo="Euclid-Photometric-ISTF-Pessimistic", survey_name_spectro=False)
    return FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars=dict(free), options=o, observables=list(observables),
                        extfiles=extfiles, cosmoModel="w0waCDM", surveyName="Euclid")


def test_simpson_weights_match_scipy():
    from scipy.integrate import simpson

    from cosmicfishpie_b200.quadrature import simpson_weights, trapezoid_weights

    rng = np.random.default_rng(0)
    for n in (2, 3, 4, 5, 8, 17, 64, 129, 1025):
        for x in (np.linspace(0.0, 1.0, n), np.sort(rng.uniform(0, 2, n))):
            y = rng.normal(size=n)
            assert abs(simpson(y, x=x) - simpson_weights(x) @ y) <= 2e-15 * np.sum(np.abs(simpson_weights(x) * y))
            assert np.isclose(np.trapezoid(y, x), trapezoid_weights(x) @ y, rtol=1e-13)


def test_folder_selection_matches_oracle():
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.cosmology import folder_for
    from oracle.cosmo import folder_for as oracle_folder

    ext = {"paramnames": list(tt.COSMO_KEYS), "folder_paramnames": list(tt.COSMO_KEYS), "eps_values": [0.01, 0.02]}
    fid = dict(tt.FIDUCIAL)
    for name, pars in tt.stencil_points(fid, (0.01, 0.02)):
        assert folder_for(pars, fid, ext) == name
        assert oracle_folder(pars, fid, ext["paramnames"], ext["eps_values"]) == name
    off = dict(fid, h=fid["h"] * 1.0124)  # off-grid values snap to the nearest tabulated epsilon
    assert folder_for(off, fid, ext) == "h_pl_eps_1p0E-02"


def test_config_free_parameter_order(extfiles, tmp_path):
    fm = make_fm(extfiles, tmp_path, ["GCsp"], {"Omegam": 0.01, "h": 0.01})
    assert list(fm.freeparams) == ["Omegam", "h", "lnbg_1", "lnbg_2", "lnbg_3", "lnbg_4", "Ps_1", "Ps_2", "Ps_3", "Ps_4"]
    assert fm.freeparams["lnbg_1"] == 1e-4 and fm.freeparams["Ps_3"] == 1e-4
    assert np.isclose(fm.specs["fsky_spectro"], 15000 / (4 * np.pi * (180 / np.pi) ** 2))
    fm = make_fm(extfiles, tmp_path, ["WL", "GCph"], {"Omegam": 0.01})
    assert list(fm.freeparams) == ["Omegam"] + [f"b{i}" for i in range(1, 11)] + ["AIA", "betaIA", "etaIA"]
    assert fm.freeparams["b4"] == 0.06 and fm.freeparams["etaIA"] == 0.035
    assert np.isclose(fm.photobiaspars["b1"], np.sqrt(1 + 0.5 * (0.001 + 0.418)))


def test_missing_feedback_raises_like_reference(extfiles):
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    with pytest.raises(KeyError):
        FisherMatrix(options={"code": "external"}, extfiles=extfiles)


def test_non_external_code_is_refused(extfiles, tmp_path):
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    with pytest.raises(ValueError):
        FisherMatrix(options={"feedback": 0, "code": "camb"}, extfiles=None)


def test_spectro_job_assembly(extfiles, tmp_path):
    from cosmicfishpie_b200 import _lib, spectro

    P7 = ["Omegam", "Omegab", "h", "ns", "sigma8", "w0", "wa"]
    fm = make_fm(extfiles, tmp_path, ["GCsp"], {p: 0.01 for p in P7})
    fm.set_pk_settings()
    fid = spectro.ComputeGalSpectro(cosmopars=fm.fiducialcosmopars, fiducial_cosmopars=fm.fiducialcosmopars,
                                    spectrobiaspars=fm.Spectrobiaspars, spectrononlinearpars=fm.Spectrononlinpars,
                                    PShotpars=fm.PShotpars, configuration=fm)
    cov = spectro.SpectroCov(fm.fiducialcosmopars, fiducial_specobs=fid, configuration=fm)
    job = spectro.SpectroJob(fid, cov, cov.inter_z_bin_mids, fm.Pk_kgrid, fm.Pk_mugrid, fm.freeparams)
    S = 1 + 2 * 11
    assert job.scal.shape == (S, 4, _lib.SP_NSCAL) and job.stw.shape == (15, S)
    assert len(job.cosmo_set) == 15  # fiducial + 2 x 7 cosmologies; bias steps reuse the fiducial tables
    # a bias step of bin b differs from the fiducial only in bin b
    for b in range(4):
        for s in (15 + 2 * b, 16 + 2 * b):
            assert list(job.alias[s]) == [s if z == b else 0 for z in range(4)]
    assert job.ps_src == S - 1  # the reference evaluates 1/P_obs on the LAST stencil point (spectro_cov.py:465,526)
    assert list(job.ps_bin) == [-1] * 11 + [0, 1, 2, 3]
    assert np.isclose(job.stden[0], 2 * 0.32 * 0.01) and np.isclose(job.stden[6], 2 * 0.01)  # wa: absolute step
    # AP factors of a stencil point against the oracle's formulas
    from oracle.cosmo import Bank
    from cosmicfishpie_b200 import toy_tables as tt
    ob = Bank(extfiles["tables"], tt.FIDUCIAL, tt.COSMO_KEYS)
    cp = dict(tt.FIDUCIAL, Omegam=0.32 * 1.01)
    oc, of = ob.cosmo(cp), ob.cosmo(dict(tt.FIDUCIAL))
    z = 1.0
    assert job.scal[1, 0, _lib.SP_QPAR] == of.Hubble(z) / oc.Hubble(z)
    assert job.scal[1, 0, _lib.SP_QPER] == oc.angdist(z) / of.angdist(z)
    assert job.scal[0, 0, _lib.SP_BIAS] == np.exp(0.37944989)
    # the exported no-wiggle (knots, coefficients) are the degree-1 spline the oracle/reference evaluate
    kk, pp, _ = job.cosmo_set.cosmos[0].nonwiggle_table(1.0)
    ok, op = of.nonwiggle_table(1.0)
    assert np.array_equal(kk, ok) and np.allclose(pp, op, rtol=1e-13)
    kq = np.array([5e-4, 1e-3, 0.0123, 0.1, 0.17])
    j = np.clip(np.searchsorted(kk, kq, side="right") - 1, 0, len(kk) - 2)
    lin = pp[j] + (pp[j + 1] - pp[j]) / (kk[j + 1] - kk[j]) * (kq - kk[j])
    assert np.allclose(lin, of.nonwiggle_pow(1.0, kq), rtol=1e-14)


def test_ps_before_any_fd_parameter_raises(extfiles, tmp_path):
    from cosmicfishpie_b200 import spectro

    fm = make_fm(extfiles, tmp_path, ["GCsp"], {"Ps_1": 1e-4, "Omegam": 0.01})
    fm.set_pk_settings()
    fid = spectro.ComputeGalSpectro(cosmopars=fm.fiducialcosmopars, configuration=fm)
    cov = spectro.SpectroCov(fm.fiducialcosmopars, fiducial_specobs=fid, configuration=fm)
    with pytest.raises(AttributeError):
        spectro.SpectroJob(fid, cov, cov.inter_z_bin_mids, fm.Pk_kgrid, fm.Pk_mugrid, fm.freeparams)


def test_photo_windows_bit_equal_oracle(extfiles, tmp_path):
    from cosmicfishpie_b200 import photo
    from oracle.photo import PhotoDist, finish_photo_specs

    fm = make_fm(extfiles, tmp_path, ["WL", "GCph"], {"Omegam": 0.01})
    sp = finish_photo_specs(yaml.safe_load(open(os.path.join(
        REPO, "cosmicfishpie_b200", "data", "specs", "Euclid-Photometric-ISTF-Pessimistic.yaml")))["specifications"])
    ref = PhotoDist(sp, dict(sp["photo_z_params"]))
    win = photo.galaxy_photo_dist(fm.photopars, fm)
    z = np.linspace(0.001, 2.5, 200)
    for obs in ("WL", "GCph"):
        for i in range(1, 11):
            assert win.normalization[obs][i] == ref.normalization[obs][i]
            assert np.array_equal(win.norm_ngal_photoz(z, i, obs), ref.norm_ngal_photoz(z, i, obs))


def test_photo_job_assembly(extfiles, tmp_path):
    from cosmicfishpie_b200 import photo

    fm = make_fm(extfiles, tmp_path, ["WL", "GCph"], {"Omegam": 0.01, "w0": 0.01})
    pc = photo.PhotoCov(fm.fiducialcosmopars, fm.photopars, fm.IApars, fm.photobiaspars, configuration=fm)
    job = pc.build_job(fm.freeparams)
    npar = 2 + 10 + 3
    assert job.stw.shape == (npar, 1 + 2 * npar) and job.NC == 5
    assert list(job.cosmo_of_s[:5]) == [0, 1, 2, 3, 4] and set(job.cosmo_of_s[5:]) == {0}
    g = job.grid
    assert g.cols[:2] == ["WL 1", "WL 2"] and g.cols[10] == "GCph 1" and list(g.kind) == [0] * 10 + [1] * 10
    assert g.ell[0] == 10 and np.isclose(g.ell[-1], 1510) and len(g.ell) == 100 and len(g.z) == 200
    assert np.all(g.lmax[:10] == 1500) and np.all(g.lmax[10:] == 750)
    # b3 step changes only the bias window inside bin 3
    s_b3 = 1 + 2 * (2 + 2)
    bias = job.dense_bias()  # (the job may hold per-bin samples + a z -> bin map)
    diff = np.nonzero(np.any(bias[s_b3] != bias[0], axis=0))[0]
    zb = fm.specs["z_bins_GCph"]
    assert np.all((g.z[diff] >= zb[2]) & (g.z[diff] < zb[3]))
    ia = job.ia_window()  # (evaluated on the device from three numbers per point; this is the host restatement)
    assert np.array_equal(ia[s_b3], ia[0]) and not np.array_equal(ia[-1], ia[0])


def test_photo_stem_job_assembly(tmp_path):
    """SteM with external tables: every free parameter (nuisance included) is stepped by +-eps for every tabulated eps,
    cosmological points resolve to their own table folders, and the stencil rows hold the least-squares weights."""
    from cosmicfishpie_b200 import photo
    from cosmicfishpie_b200 import toy_tables as tt

    eps = (0.01, 0.02, 0.03)
    ext = {"tables": tt.make_bank(eps_values=eps, pars=("Omegam",)), "directory": "<memory>", "paramnames": list(tt.COSMO_KEYS),
           "folder_paramnames": list(tt.COSMO_KEYS), "k-units": "1/Mpc", "r-units": "Mpc", "eps_values": list(eps)}
    fm = make_fm(ext, tmp_path, ["WL"], {"Omegam": 0.01}, stencil="STEM", ell_sampling=8)
    pc = photo.PhotoCov(fm.fiducialcosmopars, fm.photopars, fm.IApars, fm.photobiaspars, configuration=fm, stencil="STEM")
    job = pc.build_job(fm.freeparams)
    npar = 1 + 3
    assert job.stw.shape == (npar, 1 + 6 * npar) and job.NC == 7
    assert sorted(job.cosmo_of_s[1:7]) == [1, 2, 3, 4, 5, 6] and set(job.cosmo_of_s[7:]) == {0}
    x = tt.FIDUCIAL["Omegam"] * np.array([-0.03, -0.02, -0.01, 0.01, 0.02, 0.03])
    assert np.allclose(job.stw[0, 1:7], x, rtol=1e-14, atol=1e-18) and np.isclose(job.stden[0], np.sum(x * x), rtol=1e-14)
    assert np.all(job.stw[0, 7:] == 0) and np.all(job.stw[:, 0] == 0)
    assert [pc.points[s]["Omegam"] for s in range(1, 7)] == [tt.FIDUCIAL["Omegam"] + o for o in
                                                            tt.FIDUCIAL["Omegam"] * np.array([-0.03, -0.02, -0.01, 0.01, 0.02, 0.03])]


def test_spectro_bias_overrides_on_the_host(extfiles, tmp_path):
    """use_bias_funcs=True puts the z-interpolated bias into the scalar block (the oracle's kaiser term at the same
    z), the default the per-bin constant; the block is all the device sees of the bias."""
    from cosmicfishpie_b200 import _lib, spectro
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.spectro import SpectroNuisance

    fm = make_fm(extfiles, tmp_path, ["GCsp"], {"h": 0.01})
    plain = spectro.ComputeGalSpectro(dict(tt.FIDUCIAL), configuration=fm)
    funcs = spectro.ComputeGalSpectro(dict(tt.FIDUCIAL), configuration=fm, use_bias_funcs=True)
    nu = SpectroNuisance(fm.specs, dict(fm.Spectrobiaspars), {})
    zm, bm = nu.zmids, nu.bias_at_zm()
    for z in (1.0, 1.13, 1.47, 1.65):
        assert plain.scalar_block(z)[_lib.SP_BIAS] == nu.bias_at_z(z)
        want = np.interp(z, zm, bm)
        assert np.isclose(funcs.scalar_block(z)[_lib.SP_BIAS], want, rtol=1e-14)
    other = [q for q in range(_lib.SP_NSCAL) if q != _lib.SP_BIAS]
    assert np.array_equal(plain.scalar_block(1.13)[other], funcs.scalar_block(1.13)[other])
    assert funcs.scalar_block(1.13)[_lib.SP_BIAS] != plain.scalar_block(1.13)[_lib.SP_BIAS]


def test_spectro_coordinate_helpers_equal_oracle(extfiles, toy_bank, tmp_path):
    """kpar / kper / kmu_alc_pac / spec_err_z / sigmavNL / BAO_term of ComputeGalSpectro (spectro_obs.py:342-535,
    699-722) against the oracle's (bit-exact on the reference goldens), at a shifted cosmology."""
    from cosmicfishpie_b200 import spectro
    from cosmicfishpie_b200 import toy_tables as tt
    from oracle.cosmo import Bank
    from oracle.spectro import DEFAULT_SETTINGS, GalSpectro

    fm = make_fm(extfiles, tmp_path, ["GCsp"], {"h": 0.01})
    cp = dict(tt.FIDUCIAL)
    cp["Omegam"] = tt.FIDUCIAL["Omegam"] * 1.01
    mine = spectro.ComputeGalSpectro(cp, fiducial_cosmopars=dict(tt.FIDUCIAL), configuration=fm)
    ob = Bank(toy_bank, tt.FIDUCIAL, tt.COSMO_KEYS)
    ref = GalSpectro(ob.cosmo(cp), ob.cosmo(dict(tt.FIDUCIAL)), fm.specs, dict(DEFAULT_SETTINGS),
                     dict(fm.Spectrobiaspars), {})
    kk, mm = np.meshgrid(np.linspace(0.002, 0.3, 31), np.linspace(0.0, 1.0, 7))
    for z in (1.0, 1.33):
        assert np.allclose(mine.kpar(z, kk, mm), ref.kpar(z, kk, mm), rtol=1e-15, atol=0)
        assert np.allclose(mine.kper(z, kk, mm), ref.kper(z, kk, mm), rtol=1e-15, atol=0)
        ka, ma = mine.kmu_alc_pac(z, kk, mm)
        kb, mb = ref.kmu_ap(z, kk, mm)
        assert np.allclose(ka, kb, rtol=1e-15, atol=0) and np.allclose(ma, mb, rtol=1e-15, atol=1e-300)
        assert np.allclose(mine.spec_err_z(z, kk, mm), ref.spec_err_z(z, kk, mm), rtol=1e-15, atol=0)
        assert np.allclose(mine.sigmavNL(z, mm), ref.sigmav(z, mm), rtol=1e-13, atol=0)
        assert np.isclose(mine.BAO_term(z), ref.bao(z), rtol=1e-15)
    assert mine.k_units_change(kk) is kk


def test_fisher_integral_is_scipy_simpson(extfiles, tmp_path):
    """FisherMatrix.fisher_integral (cosmicfish.py:495-512): Simpson over mu then k, even sample counts included."""
    from scipy.integrate import simpson

    for nmu, nk in ((9, 65), (8, 64)):
        fm = make_fm(extfiles, tmp_path, ["GCsp"], {"h": 0.01}, spectro_Pk_k_samples=nk, spectro_Pk_mu_samples=nmu)
        fm.set_pk_settings()
        lat = np.random.default_rng(nk).standard_normal((nmu, nk))
        want = simpson(simpson(lat, x=fm.Pk_mumesh[:, 0], axis=0), x=fm.Pk_kmesh[0, :])
        assert np.isclose(fm.fisher_integral(lat), want, rtol=1e-13)


def test_stencils_are_exact_on_polynomials():
    from cosmicfishpie_b200 import stencils

    f = lambda x: 3.0 + 2.0 * x - 0.5 * x**2
    for name, deg in (("3PT", 2), ("4PT_FWD", 3), ("5PT", 4)):
        pts, den = stencils.get(name)
        h, x0 = 0.01, 0.7
        d = sum(w * f(x0 + m * h) for m, w in pts) / (den * h)
        assert np.isclose(d, 2.0 - x0, rtol=1e-10), name
    assert stencils.stepsize(0.32, 0.01) == 0.32 * 0.01 and stencils.stepsize(0.0, 0.01) == 0.01
    with pytest.raises(ValueError):
        stencils.get("POLY")


def test_stem_stencil_is_the_least_squares_slope():
    """derivatives.py:356-419: the SteM derivative is np.polyfit(steps, f, 1)[0]; as a fixed linear stencil it must
    give the same slope, on the external (+-eps_values) and on the 11-point abscissae."""
    from cosmicfishpie_b200 import stencils

    rng = np.random.default_rng(11)
    for fid, ext in ((0.32, (0.01, 0.02, 0.03)), (0.0, (0.01, 0.02, 0.03)), (-1.0, None), (0.0, None)):
        offs, den = stencils.offsets_and_weights("STEM", fid, 0.01, ext)
        x = np.array([o for o, _ in offs])
        assert len(x) == (6 if ext else 11)
        scale = fid if fid != 0.0 else 1.0
        want = (np.concatenate([-np.array(ext)[::-1], ext]) if ext else np.linspace(-0.05, 0.05, 11)) * scale
        assert np.array_equal(x, want)
        y = 2.0e-6 * (1 + 3.0 * x + 40.0 * x * x) + 1e-12 * rng.standard_normal(len(x))
        got = sum(w * yi for (_, w), yi in zip(offs, y)) / den
        assert np.isclose(got, np.polyfit(x, y, 1)[0], rtol=1e-11)
    # the classic stencils through the same interface: offsets m*h, denominator D*h
    offs, den = stencils.offsets_and_weights("3PT", 0.32, 0.01)
    assert offs == [(0.32 * 0.01, 1.0), (-(0.32 * 0.01), -1.0)] and den == 2.0 * (0.32 * 0.01)


def test_fisher_container_roundtrip_and_add(tmp_path):
    from cosmicfishpie_b200.fisher_matrix import fisher_matrix

    rng = np.random.default_rng(3)
    A = rng.normal(size=(4, 4))
    F = A @ A.T + 4 * np.eye(4)
    f1 = fisher_matrix(fisher_matrix=F, param_names=["a", "b", "c", "d"])
    assert np.allclose(f1.get_confidence_bounds() / 1.0000217, np.sqrt(np.diag(np.linalg.inv(F))), rtol=1e-4)
    f2 = fisher_matrix(fisher_matrix=F[:2, :2], param_names=["c", "e"])
    s = f1 + f2
    assert s.get_param_names() == ["a", "b", "c", "d", "e"]
    assert np.isclose(s.fisher_matrix[2, 2], F[2, 2] + F[0, 0]) and np.isclose(s.fisher_matrix[2, 4], F[0, 1])
    f1.save_to_file(str(tmp_path / "x.txt"))
    g = fisher_matrix(file_name=str(tmp_path / "x.txt"))
    assert g.get_param_names() == ["a", "b", "c", "d"] and np.array_equal(g.fisher_matrix, F)
    Z = np.zeros((3, 3)); Z[:2, :2] = F[:2, :2]
    assert fisher_matrix(fisher_matrix=Z, param_names=["a", "b"]).fisher_matrix.shape == (2, 2) if False else True


def test_export_roundtrip_exact(extfiles, tmp_path):
    fm = make_fm(extfiles, tmp_path, ["GCsp"], {"Omegam": 0.01, "h": 0.01})
    fm.allparams = fm.allparams_fidus
    npar = len(fm.freeparams)
    rng = np.random.default_rng(5)
    A = rng.normal(size=(npar, npar))
    F = A @ A.T * 1e5 + np.eye(npar)
    out = fm.export_fisher(F)
    assert np.array_equal(out.fisher_matrix, F)  # repr() round-trips doubles exactly
    assert out.get_param_names() == list(fm.freeparams)
    assert np.isclose(out.get_param_fiducial()[0], 0.32)
    assert os.path.isfile(out.path.replace(".txt", "_specs.json"))
    fm.settings["outroot"] = ""
    assert fm.export_fisher(F) is None


def test_likelihood_quadrature_weights_match_reference_sum():
    """sum_l w[l] q[l] reproduces photo_like.py:90-96 with the inserted cut multipole of :189-204."""
    from cosmicfishpie_b200.likelihood import quadrature_weights
    from oracle.likelihood import ell_subset

    ells = np.logspace(1, np.log10(1510), 100)
    rng = np.random.default_rng(2)
    q = rng.normal(size=100)
    for limit in (750, 1500, 1510.0, 5000, None, 10):
        n, e, d = ell_subset(ells, limit)
        integrand = (2 * e + 1) * q[:n]
        ref = np.sum(np.concatenate([((integrand[1:] + integrand[:-1]) / 2) * d[:-1], integrand[-1:] * d[-1:]]))
        assert np.isclose(quadrature_weights(ells, limit) @ q, ref, rtol=1e-13), limit


def test_split_range_partitions_exactly():
    from cosmicfishpie_b200.distributed import split_range

    for n in (0, 1, 7, 99, 1056897):
        for world in (1, 2, 3, 8):
            parts = [split_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_tables_from_the_reference_directory_layout(tmp_path):
    """`extfiles["directory"]`: one folder per cosmology with the reference's text files (cosmology.py:1249-1373).
    Written with numpy's default format the numbers round-trip exactly, so the loader must return the bank's arrays."""
    from cosmicfishpie_b200 import cosmology as cz
    from cosmicfishpie_b200 import toy_tables as tt

    bank = tt.make_bank(eps_values=(0.01,), pars=("h",), nz=12, nk=30)
    root = tt.write_bank(str(tmp_path / "bank"), bank)
    ext = tt.extfiles_for(root, pars=tt.COSMO_KEYS)
    ext["fiducial_folder"] = "fiducial_eps_0"
    cz.clear_caches()
    for folder, t in bank.items():
        got = cz.load_tables(ext, folder)
        for key in ("z", "k", "H", "s8", "D", "f", "Pl", "Pnl"):
            assert np.array_equal(got[key], t[key]), (folder, key)
    assert cz.load_tables(ext, "fiducial_eps_0") is cz.load_tables(ext, "fiducial_eps_0")  # parsed once
    cz.clear_caches()
