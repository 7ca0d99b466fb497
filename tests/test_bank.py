"""CPU: the binary table bank and the Taylor emulator over it (cosmicfishpie_b200/bank.py)."""
import numpy as np
import pytest

from conftest import base_options


def test_bank_file_round_trip(toy_bank, tmp_path):
    from cosmicfishpie_b200.bank import load_bank, save_bank

    path = save_bank(str(tmp_path / "toy.cfpbank"), toy_bank, meta={"eps_values": [0.01], "note": "toy"})
    tables, meta = load_bank(path)
    assert meta["eps_values"] == [0.01] and list(tables) == list(toy_bank)
    for folder, t in toy_bank.items():
        assert list(tables[folder]) == list(t)
        for name, a in t.items():
            assert np.array_equal(tables[folder][name], a), (folder, name)
    assert not tables["fiducial_eps_0"]["Pl"].flags.writeable  # a read-only view of the memory map
    with open(path, "r+b") as fh:
        fh.write(b"XXXXXXXX")
    with pytest.raises(ValueError):
        load_bank(path)


def test_emulator_weights_and_nodes(toy_bank):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.bank import TaylorEmulator

    emu = TaylorEmulator(toy_bank, tt.FIDUCIAL, tt.COSMO_KEYS, 0.01)
    # tabulated cosmologies are reproduced exactly: one-hot weights
    for folder, pars in tt.stencil_points(dict(tt.FIDUCIAL), (0.01,)):
        assert emu.is_node(pars)
        w = emu.weights(pars)
        assert w == {folder: 1.0} or (folder == "fiducial_eps_0" and w == {folder: 1.0})
        t = emu.tables_at(pars)
        for name in ("H", "Pl", "Pnl", "f", "D", "s8"):
            assert np.array_equal(t[name], toy_bank[folder][name]), (folder, name)
    rng = np.random.default_rng(3)
    theta = {k: (v * (1 + 0.01 * rng.uniform(-1, 1)) if v != 0 else 0.01 * rng.uniform(-1, 1)) for k, v in tt.FIDUCIAL.items()}
    assert not emu.is_node(theta)
    assert abs(sum(emu.weights(theta).values()) - 1.0) < 1e-14
    W = emu.weight_matrix([theta, dict(tt.FIDUCIAL)], list(toy_bank))
    assert W.shape == (2, 15) and np.allclose(W.sum(axis=1), 1.0) and W[1, 0] == 1.0 and np.count_nonzero(W[1]) == 1


def test_emulator_accuracy_off_node(toy_bank):
    """Stated interpolation error: all seven parameters moved by up to +-1 % at once, against the toy model's exact
    tables -- the diagonal second-order expansion leaves the cross terms (not tabulated in a +-eps bank): measured
    worst case 2e-5 in H, 5e-5 in sigma8, 2.6e-3 in P_lin (where the BAO wiggles move with h, Omegam and Omegab together)."""
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.bank import TaylorEmulator

    emu = TaylorEmulator(toy_bank, tt.FIDUCIAL, tt.COSMO_KEYS, 0.01)
    zg, kg = toy_bank["fiducial_eps_0"]["z"], toy_bank["fiducial_eps_0"]["k"]
    rng = np.random.default_rng(11)
    worst = {}
    for _ in range(8):
        theta = {k: (v * (1 + 0.01 * rng.uniform(-1, 1)) if v != 0 else 0.01 * rng.uniform(-1, 1))
                 for k, v in tt.FIDUCIAL.items()}
        truth = tt.make_tables(theta, zg, kg)
        got = emu.tables_at(theta)
        for name in ("H", "s8", "f", "Pl", "Pnl"):
            err = float(np.max(np.abs(got[name] - truth[name]) / np.abs(truth[name])))
            worst[name] = max(worst.get(name, 0.0), err)
    assert worst["H"] < 5e-5 and worst["s8"] < 1e-4 and worst["f"] < 1e-4 and worst["Pl"] < 5e-3 and worst["Pnl"] < 5e-3, worst
    # single-parameter moves (no cross terms): third-order accurate
    theta = dict(tt.FIDUCIAL, h=tt.FIDUCIAL["h"] * 1.005)
    err = np.max(np.abs(emu.tables_at(theta)["Pl"] / tt.make_tables(theta, zg, kg)["Pl"] - 1))
    assert err < 1e-4, err  # (measured 4e-5: the third-order term of the BAO phase in h)


def test_facade_uses_the_emulator_off_node_only(toy_bank, extfiles, tmp_path):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.bank import TaylorEmulator
    from cosmicfishpie_b200.cosmology import cached_cosmo
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    emu = TaylorEmulator(toy_bank, tt.FIDUCIAL, tt.COSMO_KEYS, 0.01)
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"h": 0.01}, options=base_options(tmp_path),
                      observables=["GCsp"], extfiles=dict(extfiles, emulator=emu))
    plain = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"h": 0.01}, options=base_options(tmp_path),
                         observables=["GCsp"], extfiles=extfiles)
    node = dict(tt.FIDUCIAL, h=tt.FIDUCIAL["h"] * 1.01)
    assert cached_cosmo(node, fm.config).folder == "h_pl_eps_1p0E-02"
    off = dict(tt.FIDUCIAL, h=tt.FIDUCIAL["h"] * 1.008, Omegam=tt.FIDUCIAL["Omegam"] * 1.007)
    c = cached_cosmo(off, fm.config)
    assert c.folder == "<emulated>"
    zg, kg = toy_bank["fiducial_eps_0"]["z"], toy_bank["fiducial_eps_0"]["k"]
    truth = tt.make_tables(off, zg, kg)
    assert abs(c.Hubble(1.0) / np.interp(1.0, zg, truth["H"]) - 1) < 1e-4
    # without the emulator the same point snaps to a tabulated folder, as the reference does
    assert cached_cosmo(off, plain.config).folder in ("Omegam_pl_eps_1p0E-02", "h_pl_eps_1p0E-02")
