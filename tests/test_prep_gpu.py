"""Device-side preparation of the cosmology input (csrc/cfp_prep.cu) against scipy's FITPACK -- the calls
the reference makes per cosmology (cosmology.py:1447-1532 spline fits, :36-55 comoving distance, :1760-1813 no-wiggle
spectrum; spectro_obs.py:724-755 velocity moments) -- and end to end: forecasts with the device preparation (the
default) equal forecasts with the host fits, and both sit on the reference goldens."""
import numpy as np
import pytest
from scipy.interpolate import InterpolatedUnivariateSpline, RectBivariateSpline

from conftest import base_options, elem_rel_err, golden, rel_err

pytestmark = pytest.mark.gpu

P7 = ["Omegam", "Omegab", "h", "ns", "sigma8", "w0", "wa"]


@pytest.fixture(scope="module")
def ctx():
    from cosmicfishpie_b200 import _lib

    return _lib.default_context()


def _grids(rng, nz, nk):
    z = np.sort(np.concatenate(([0.0], rng.uniform(0.01, 3.0, nz - 2), [3.0])))
    k = np.logspace(-4, 1.5, nk) * (1 + 0.02 * rng.uniform(-1, 1, nk))
    return z, np.sort(k)


@pytest.mark.parametrize("nz,nk", [(5, 5), (6, 7), (7, 6), (8, 33), (100, 500), (37, 1000)])
def test_bank_fit_matches_fitpack(ctx, nz, nk):
    """cfp_bank_fit: knots identical to FITPACK's, coefficients and evaluated values to rounding."""
    from cosmicfishpie_b200 import _lib

    rng = np.random.default_rng(nz * 1000 + nk)
    z, k0 = _grids(rng, nz, nk)
    ks = np.stack([k0, 0.97 * k0, 1.01 * k0])  # (the k-units factor differs per cosmology)
    tabs, scales, ref = [], [], []
    for c in range(3):
        pl = 1e4 * ks[c][None, :] / (1 + (ks[c][None, :] / 0.02) ** 2.7) / (1 + z[:, None]) ** 2 * (1 + 0.1 * c)
        pl = pl * (1 + 0.03 * np.sin(40 * ks[c][None, :] + z[:, None]))
        f = 0.5 + 0.4 * np.tanh(z[:, None] - 1) + 0.01 * np.log(ks[c][None, :]) * (1 + 0.2 * c)
        row, sc = [None] * _lib.NTAB, [1.0] * _lib.NTAB
        row[_lib.TAB_PLIN], row[_lib.TAB_F] = pl, f
        sc[_lib.TAB_PLIN] = 0.7 + 0.1 * c
        tabs.append(row)
        scales.append(sc)
        ref.append((RectBivariateSpline(z, ks[c], sc[_lib.TAB_PLIN] * pl, kx=3, ky=3), RectBivariateSpline(z, ks[c], f, kx=3, ky=3)))
    bank = _lib.Bank.fit(ctx, z, ks, tabs, scales)
    zq = rng.uniform(-0.1, 3.2, 4000)  # beyond the grid: FITPACK clamps the argument
    for c in range(3):
        kq = np.exp(rng.uniform(np.log(ks[c][0] * 0.8), np.log(ks[c][-1] * 1.1), zq.size))
        for slot, spl in zip((_lib.TAB_PLIN, _lib.TAB_F), ref[c]):
            got = bank.eval(c, slot, zq, kq)
            want = spl(zq, kq, grid=False)
            assert np.max(np.abs(got - want) / (np.abs(want) + 1e-3 * np.max(np.abs(want)))) < 2e-12, (c, slot)
    many = bank.eval_many(_lib.TAB_F, zq[:100], np.full(100, ks[0][3]))
    for c in range(3):
        assert np.allclose(many[c], ref[c][1](zq[:100], ks[0][3], grid=False), rtol=1e-12, atol=1e-14)
    bank.close()


def test_bank_fit_rejects_bad_input(ctx):
    from cosmicfishpie_b200 import _lib

    z, k = np.linspace(0, 1, 6), np.logspace(-3, 0, 8)
    row = [None] * _lib.NTAB
    row[0] = np.ones((6, 8))
    with pytest.raises(_lib.CfpError):
        _lib.Bank.fit(ctx, z[:4], k[None, :], [[np.ones((4, 8))] + [None] * (_lib.NTAB - 1)], [[1.0] * _lib.NTAB])  # < 5 samples
    zz = z.copy()
    zz[2] = zz[1]
    with pytest.raises(_lib.CfpError):
        _lib.Bank.fit(ctx, zz, k[None, :], [row], [[1.0] * _lib.NTAB])  # not strictly increasing
    row2 = [None] * _lib.NTAB
    row2[1] = np.ones((6, 8))
    with pytest.raises(_lib.CfpError):
        _lib.Bank.fit(ctx, z, np.stack([k, k]), [row, row2], [[1.0] * _lib.NTAB] * 2)  # different slots per cosmology


@pytest.mark.parametrize("nz", [5, 6, 9, 100, 513])
def test_background_fit_matches_scipy(ctx, nz):
    from cosmicfishpie_b200 import _lib
    from cosmicfishpie_b200.cosmology import _dcom_all, nak_knots

    rng = np.random.default_rng(nz)
    z = np.sort(np.concatenate(([0.0], rng.uniform(0.01, 3.0, nz - 2), [3.0])))
    NC = 4
    hub = np.stack([(67 + 2 * c) / 299792.458 * np.sqrt(0.3 * (1 + z) ** 3 + 0.7) for c in range(NC)])
    extra = np.stack([np.stack([0.8 / (1 + z) ** (0.9 + 0.02 * c), 1 / (1 + z) + 0.01 * c]) for c in range(NC)])
    zq = np.concatenate((rng.uniform(0, 3, 300), z, [0.0, 3.0]))
    out, coef = _lib.background_fit(ctx, z, hub, extra, zq, coef=True)
    t = nak_knots(z)
    for c in range(NC):
        h = InterpolatedUnivariateSpline(z, hub[c])
        dcom = _dcom_all(z, h)
        refs = [h, InterpolatedUnivariateSpline(z, dcom), InterpolatedUnivariateSpline(z, dcom / (1 + z)),
                InterpolatedUnivariateSpline(z, extra[c, 0]), InterpolatedUnivariateSpline(z, extra[c, 1])]
        for j, r in enumerate(refs):
            tt_, cc_, _ = r._eval_args
            assert np.array_equal(tt_, t)
            scale = np.max(np.abs(cc_[:nz]))
            assert np.max(np.abs(coef[c, j] - cc_[:nz])) < 1e-12 * scale, (c, j)
            assert np.max(np.abs(out[c, j] - r(zq))) < 1e-12 * scale, (c, j)
            # the injected spline object (what the facade holds) evaluates like scipy's own
            inj = InterpolatedUnivariateSpline._from_tck((t, coef[c, j], 3))
            assert np.max(np.abs(inj(zq) - r(zq))) < 1e-12 * scale


@pytest.fixture(scope="module")
def spectro_cfg(extfiles, tmp_path_factory):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    def make(**kw):
        o = base_options(tmp_path_factory.mktemp("prep"), **kw)
        return FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7}, options=o,
                            observables=["GCsp"], extfiles=extfiles)

    return make


def test_prepared_facades_match_host_facades(spectro_cfg, extfiles, monkeypatch):
    """CosmoSet.prepare / prepare_spectro fill the facades with what the host (scipy) path computes: H, distances,
    sigma8, low-k growth, velocity moments, no-wiggle tables -- for every cosmology of the toy bank, h-steps included
    (their Savitzky-Golay window and k samples differ)."""
    from cosmicfishpie_b200 import cosmology as cz
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.spectro import _cosmo_set_of, _theta_moments

    cz.clear_caches()
    fm = spectro_cfg()
    cfg = fm.config
    cs = _cosmo_set_of(cfg)
    fid = dict(tt.FIDUCIAL)
    pars = [fid]
    for p in P7:
        for s in (1, -1):
            q = dict(fid)
            q[p] = fid[p] * (1 + s * 0.01) if fid[p] != 0 else s * 0.01
            pars.append(q)
    for q in pars:
        cs.add(q)
    zs = [1.0, 1.2, 1.4, 1.65]
    cs.prepare()
    cs.prepare_spectro(zs)
    zq = np.linspace(0, 2.5, 301)
    worst = {}
    assert len(cs.cosmos) == len(pars)
    for q, dev in zip(pars, cs.cosmos):
        assert "_growth_lowk" in dev._coef1d and "com_dist" in dev._coef1d  # really the prepared ones
        monkeypatch.setenv("CFP_HOST_PREP", "1")
        host = cz.cosmo_functions(dict(dev.cosmopars), "external", cfg)
        monkeypatch.delenv("CFP_HOST_PREP")
        for name in ("Hubble", "comoving", "angdist", "sigma8_of_z", "growth"):
            a, b = np.asarray(getattr(dev, name)(zq)).ravel(), np.asarray(getattr(host, name)(zq)).ravel()
            worst[name] = max(worst.get(name, 0.0), float(np.max(np.abs(a - b) / np.max(np.abs(b)))))
        for z in zs:
            a, b = np.array(dev._moment_cache[z]), np.array(_theta_moments(host, z))
            worst["moments"] = max(worst.get("moments", 0.0), float(np.max(np.abs(a / b - 1))))
            ka, va, _ = dev._nw_cache[(z, False, False)]
            kb, vb, _ = host.nonwiggle_table(z)
            assert np.array_equal(ka, kb)
            worst["nowiggle"] = max(worst.get("nowiggle", 0.0), float(np.max(np.abs(va / vb - 1))))
    print("device preparation vs host (max relative difference):", worst)
    assert worst["Hubble"] < 1e-13 and worst["comoving"] < 1e-13 and worst["angdist"] < 1e-13
    assert worst["sigma8_of_z"] < 1e-13 and worst["growth"] < 1e-12
    assert worst["moments"] < 1e-12 and worst["nowiggle"] < 1e-11


def test_forecasts_agree_between_device_and_host_preparation(extfiles, tmp_path, monkeypatch):
    """GCsp w0waCDM forecast and a 3x2pt forecast: device preparation (default) against host fits, and against the
    reference golden."""
    from cosmicfishpie_b200 import cosmology as cz
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    def run(obs, **kw):
        cz.clear_caches()
        o = base_options(tmp_path, **kw)
        fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in P7}, options=o,
                          observables=obs, extfiles=extfiles)
        return np.array(fm.compute().fisher_matrix)

    g = golden("gcsp_w0wa")
    Fd = run(["GCsp"])
    monkeypatch.setenv("CFP_HOST_PREP", "1")
    Fh = run(["GCsp"])
    monkeypatch.delenv("CFP_HOST_PREP")
    print("GCsp device vs host prep:", elem_rel_err(Fd, Fh), "device vs reference:", elem_rel_err(Fd, g["fisher"]))
    assert elem_rel_err(Fd, Fh) < 1e-9
    assert elem_rel_err(Fd, g["fisher"]) < 1e-8
    kw = dict(survey_name_photo="Euclid-Photometric-ISTF-Pessimistic", survey_name_spectro=False, ell_sampling=24)
    Pd = run(["WL", "GCph"], **kw)
    monkeypatch.setenv("CFP_HOST_PREP", "1")
    Ph = run(["WL", "GCph"], **kw)
    monkeypatch.delenv("CFP_HOST_PREP")
    print("3x2pt device vs host prep:", elem_rel_err(Pd, Ph))
    assert elem_rel_err(Pd, Ph) < 1e-7  # (the photometric off-diagonals cancel to 1e-3: DESIGN section 2)
    assert rel_err(np.sqrt(np.diag(np.linalg.inv(Pd))), np.sqrt(np.diag(np.linalg.inv(Ph)))) < 1e-9
    cz.clear_caches()


@pytest.mark.parametrize("nz", [4, 5, 7])
def test_short_redshift_grids(tmp_path, monkeypatch, nz):
    """Tables on very few redshifts: with fewer than five the device preparation declines (a not-a-knot cubic needs
    five samples per axis) and the host fits take over silently; with five or seven it runs -- either way the forecast
    equals the one prepared on the host."""
    from cosmicfishpie_b200 import cosmology as cz
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    bank = tt.make_bank(eps_values=(0.01,), pars=("Omegam", "h"), nz=nz, nk=120)
    ext = {"tables": bank, "directory": "<memory>", "paramnames": list(tt.COSMO_KEYS), "folder_paramnames": list(tt.COSMO_KEYS),
           "k-units": "1/Mpc", "r-units": "Mpc", "eps_values": [0.01]}

    def run():
        cz.clear_caches()
        fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01, "h": 0.01},
                          options=base_options(tmp_path, spectro_Pk_k_samples=65, spectro_Pk_mu_samples=9),
                          observables=["GCsp"], extfiles=ext)
        F = np.array(fm.compute(max_z_bins=2).fisher_matrix)
        fitted_on_device = any(not hit[2].rows for hit in cz._BANK_CACHE.values())
        return F, fitted_on_device

    Fd, on_device = run()
    assert on_device == (nz >= 5)
    monkeypatch.setenv("CFP_HOST_PREP", "1")
    Fh, _ = run()
    monkeypatch.delenv("CFP_HOST_PREP")
    cz.clear_caches()
    assert np.all(np.isfinite(Fd)) and elem_rel_err(Fd, Fh) < 1e-8


def test_forecast_from_the_reference_directory_layout(tmp_path, toy_bank, extfiles):
    """The reference's own input format -- a directory tree of text files -- gives the forecast of the in-memory bank."""
    from cosmicfishpie_b200 import cosmology as cz
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    sub = {name: t for name, t in toy_bank.items() if name.startswith(("fiducial", "Omegam", "h_"))}
    root = tt.write_bank(str(tmp_path / "bank"), sub)

    def run(ext):
        cz.clear_caches()
        fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01, "h": 0.01},
                          options=base_options(tmp_path / "out", spectro_Pk_k_samples=129, spectro_Pk_mu_samples=9),
                          observables=["GCsp"], extfiles=ext)
        return np.array(fm.compute(max_z_bins=2).fisher_matrix)

    F_dir = run(tt.extfiles_for(root))
    F_mem = run(extfiles)
    cz.clear_caches()
    assert elem_rel_err(F_dir, F_mem) < 1e-13


def test_forecast_from_a_memory_mapped_bank_file(tmp_path, toy_bank, extfiles):
    """`extfiles["tables"] = load_bank(path)`: read-only arrays of a memory map go to the device fit as they are."""
    from cosmicfishpie_b200 import cosmology as cz
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.bank import load_bank, save_bank
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    tables, _ = load_bank(save_bank(str(tmp_path / "toy.cfpbank"), toy_bank))
    assert not tables["fiducial_eps_0"]["Pl"].flags.writeable

    def run(ext, kind):
        cz.clear_caches()
        opts = base_options(tmp_path / "out")
        if kind == "photo":
            opts.update(survey_name_photo="Euclid-Photometric-ISTF-Pessimistic", survey_name_spectro=False, ell_sampling=12)
        fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01, "h": 0.01}, options=opts,
                          observables=["WL", "GCph"] if kind == "photo" else ["GCsp"], extfiles=ext)
        return np.array(fm.compute().fisher_matrix)

    for kind in ("spectro", "photo"):
        assert np.array_equal(run(dict(extfiles, tables=tables), kind), run(extfiles, kind)), kind
    cz.clear_caches()


@pytest.mark.parametrize("win,order", [(101, 3), (58, 2), (7, 3)])
def test_nowiggle_and_moments_kernels_against_scipy(ctx, win, order):
    """cfp_bank_nowiggle (odd and even Savitzky-Golay windows, several polynomial orders) and cfp_bank_theta_moments on
    a synthetic two-cosmology bank, against scipy.signal.savgol_filter / scipy.integrate.trapezoid of FITPACK values."""
    from scipy.integrate import trapezoid
    from scipy.signal import savgol_filter

    from cosmicfishpie_b200 import _lib
    from cosmicfishpie_b200.cosmology import savgol_operator

    rng = np.random.default_rng(win)
    z, k = _grids(rng, 24, 160)
    ks = np.stack([k, 1.03 * k])
    tabs, scales, spl = [], [], []
    for c in range(2):
        kk = ks[c][None, :]
        pl = 2e4 * kk / (1 + (kk / 0.02) ** 2.6) / (1 + z[:, None]) ** 1.7 * (1 + 0.04 * np.sin(55 * kk + 0.3 * c))
        f = 0.55 + 0.35 * np.tanh(z[:, None] - 0.8) + 0.005 * np.log(kk)
        row = [None] * _lib.NTAB
        row[_lib.TAB_PLIN], row[_lib.TAB_F] = pl, f
        tabs.append(row)
        scales.append([1.0] * _lib.NTAB)
        spl.append((RectBivariateSpline(z, ks[c], pl, kx=3, ky=3), RectBivariateSpline(z, ks[c], f, kx=3, ky=3)))
    bank = _lib.Bank.fit(ctx, z, ks, tabs, scales)
    zq = np.array([0.3, 1.1, 2.4])
    nsamp = 300
    samp = np.stack([np.exp(np.linspace(np.log(ks[c][0] * 1.5), np.log(ks[c][-1] * 0.7), nsamp)) for c in range(2)])
    got = bank.nowiggle(_lib.TAB_PLIN, zq, samp, [0, 0], [savgol_operator(win, order)])
    for c in range(2):
        for j, zz in enumerate(zq):
            ref = np.exp(savgol_filter(np.log(spl[c][0](zz, samp[c], grid=False)), win, order))
            assert np.max(np.abs(got[c, j] / ref - 1)) < 1e-11, (c, zz)
    kq = np.logspace(np.log10(k[0] * 2), np.log10(k[-1] * 0.5), 777)
    mom = bank.theta_moments(_lib.TAB_PLIN, _lib.TAB_F, zq, kq)
    for c in range(2):
        for j, zz in enumerate(zq):
            p, f = spl[c][0](zz, kq, grid=False), spl[c][1](zz, kq, grid=False)
            ref = [trapezoid(p * f**m, x=kq) / (6 * np.pi**2) for m in (0, 1, 2)]
            assert np.allclose(mom[c, j], ref, rtol=1e-12, atol=0), (c, zz)
    bank.close()
