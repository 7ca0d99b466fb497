"""Host-side pieces around the device preparation: the Savitzky-Golay filter as the linear operator cfp_bank_nowiggle
applies, FITPACK's interpolation knots, the lazily formed tables of emulated cosmologies, the facade cache key, the
helper thread that writes the result files of a batch of forecasts."""
import numpy as np
from scipy.interpolate import InterpolatedUnivariateSpline


def test_savgol_operator_is_scipys_filter():
    from scipy.signal import savgol_filter

    from cosmicfishpie_b200.cosmology import savgol_operator

    rng = np.random.default_rng(3)
    x = np.cumsum(rng.normal(size=800)) * 0.05 + np.linspace(5, -3, 800)
    for win, order in ((101, 3), (100, 3), (57, 3), (58, 2), (31, 2), (5, 3)):
        taps, off, EL, ER = savgol_operator(win, order)
        ne, wc = EL.shape
        y = np.empty_like(x)
        for s in range(x.size):
            if s < ne:
                y[s] = EL[s] @ x[:wc]
            elif s >= x.size - ne:
                y[s] = ER[s - (x.size - ne)] @ x[x.size - wc:]
            else:
                y[s] = taps @ x[s + off:s + off + win]
        assert np.max(np.abs(y - savgol_filter(x, win, order))) < 1e-13, (win, order)


def test_nak_knots_are_fitpacks():
    from cosmicfishpie_b200.cosmology import nak_knots

    rng = np.random.default_rng(0)
    for n in (5, 6, 17, 100):
        x = np.sort(rng.uniform(0, 3, n))
        t = InterpolatedUnivariateSpline(x, np.sin(x))._eval_args[0]
        assert np.array_equal(nak_knots(x), t)


def test_lazy_emulated_tables(toy_bank):
    """bank.LazyTables forms a table on first access only; functions of z and single columns go through the stacked
    copies and equal the weighted sums of the full tables."""
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.bank import LazyTables, TaylorEmulator

    emu = TaylorEmulator(toy_bank, tt.FIDUCIAL, tt.COSMO_KEYS, 0.01)
    theta = dict(tt.FIDUCIAL, Omegam=0.3223, h=0.6725, ns=0.9633, wa=0.004)
    t = emu.tables_at(theta)
    assert isinstance(t, LazyTables) and not t._done
    w = emu.weights(theta)
    col0 = t.column0("D")
    assert "D" not in t._done  # the column alone, the table was not formed
    full = sum(x * np.asarray(toy_bank[f]["D"]) for f, x in w.items())
    assert np.allclose(col0, full[:, 0], rtol=1e-14, atol=0)
    assert np.allclose(t["H"], sum(x * np.asarray(toy_bank[f]["H"]) for f, x in w.items()), rtol=1e-14, atol=0)
    assert np.allclose(t["D"], full, rtol=1e-14, atol=0) and "D" in t._done
    assert np.array_equal(t["k"], toy_bank["fiducial_eps_0"]["k"]) and "Pl" in t and "nope" not in t
    assert abs(emu.weight_vector(w).sum() - 1.0) < 1e-14 and emu.weight_vector(w).shape == (len(toy_bank),)


def test_facade_cache_sees_settings_edited_in_place(toy_bank, extfiles, tmp_path):
    """The configuration part of the facade cache key is remembered per Config object but re-derived when one of the
    values it is built from changes -- editing a setting in place must not hand back a facade fitted for the old one."""
    import os

    from conftest import base_options
    from cosmicfishpie_b200 import config as config_module
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.cosmology import cached_cosmo

    os.environ["CFP_HOST_PREP"] = os.environ.get("CFP_HOST_PREP", "1")
    cfg = config_module.Config(options=base_options(tmp_path), observables=["GCsp"], extfiles=extfiles,
                               fiducialpars=dict(tt.FIDUCIAL), freepars={"h": 0.01})
    a = cached_cosmo(dict(tt.FIDUCIAL), cfg)
    assert cached_cosmo(dict(tt.FIDUCIAL), cfg) is a
    cfg.settings["savgol_width"] = cfg.settings["savgol_width"] * 1.25
    b = cached_cosmo(dict(tt.FIDUCIAL), cfg)
    assert b is not a
    cfg.external["k-units"] = "h/Mpc"
    c = cached_cosmo(dict(tt.FIDUCIAL), cfg)
    assert c is not b and c.kgrid[0] != b.kgrid[0]


def test_deferred_result_files(tmp_path):
    """The helper thread of FisherMatrix.compute_many: files are on disk after flush(), the first error is re-raised
    there, close() drains the queue."""
    import pytest

    from cosmicfishpie_b200 import fishermatrix as fmod

    w = fmod._DeferredWrites()
    for i in range(20):
        w.put(str(tmp_path / f"f{i}.txt"), "x" * i)
    w.flush()
    assert all((tmp_path / f"f{i}.txt").read_text() == "x" * i for i in range(20))
    w.put(str(tmp_path / "no_such_dir" / "f.txt"), "y")
    with pytest.raises(OSError):
        w.flush()
    w.put(str(tmp_path / "after.txt"), "z")
    w.close()
    assert (tmp_path / "after.txt").read_text() == "z"
