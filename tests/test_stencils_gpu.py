"""GPU: 5PT and 4PT_FWD derivative stencils vs the oracle run with the same stencil.

The reference has no 5PT stencil and its production callers ignore `settings["derivatives"]`, so these
modes are pinned by the oracle only (whose 3PT mode is bit-exact against the reference goldens)."""
import os

import numpy as np
import pytest
import yaml

from conftest import REPO, base_options

pytestmark = pytest.mark.gpu
SPECDIR = os.path.join(REPO, "cosmicfishpie_b200", "data", "specs")
EPS = (0.01, 0.02, 0.03)


@pytest.fixture(scope="module")
def bank3():
    from cosmicfishpie_b200 import toy_tables as tt

    return tt.make_bank(eps_values=EPS, pars=("Omegam", "h", "w0"))


def ext_of(bank):
    from cosmicfishpie_b200 import toy_tables as tt

    return {"tables": bank, "directory": "<memory>", "paramnames": list(tt.COSMO_KEYS),
            "folder_paramnames": list(tt.COSMO_KEYS), "k-units": "1/Mpc", "r-units": "Mpc", "eps_values": list(EPS)}


@pytest.mark.parametrize("stencil", ["5PT", "4PT_FWD"])
def test_spectro_stencil_vs_oracle(bank3, tmp_path, stencil):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix
    from oracle.cosmo import Bank
    from oracle.spectro import spectro_fisher

    small = {"spectro_Pk_k_samples": 65, "spectro_Pk_mu_samples": 9}
    free = {"Omegam": 0.01, "h": 0.01, "w0": 0.01}
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars=dict(free), options=base_options(tmp_path, stencil=stencil, **small),
                      observables=["GCsp"], extfiles=ext_of(bank3))
    F = fm.compute(max_z_bins=2)
    sp = yaml.safe_load(open(os.path.join(SPECDIR, "Euclid-Spectroscopic-ISTF-Pessimistic.yaml")))["specifications"]
    ob = Bank(bank3, tt.FIDUCIAL, tt.COSMO_KEYS, eps_values=EPS)
    ref = spectro_fisher(ob, sp, small, dict(tt.FIDUCIAL), dict(sp["sp_bias_parametrization"]["linear_log"]),
                         dict(sp["shot_noise_parametrization"]["constant"]), {}, dict(fm.config.freeparams), max_z_bins=2,
                         stencil=stencil)
    assert list(F.get_param_names()) == ref["param_names"]
    scale = np.sqrt(np.outer(np.diag(ref["fisher"]), np.diag(ref["fisher"])))
    assert np.max(np.abs(F.fisher_matrix - ref["fisher"]) / scale) < 1e-9


def test_photo_5pt_vs_oracle(bank3, tmp_path):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix
    from oracle.cosmo import Bank
    from oracle.photo import default_photo_bias, finish_photo_specs, photo_fisher

    o = base_options(tmp_path, stencil="5PT", ell_sampling=14, survey_name_photo="Euclid-Photometric-ISTF-Pessimistic",
                     survey_name_spectro=False)
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01, "w0": 0.01}, options=o,
                      observables=["WL"], extfiles=ext_of(bank3))
    F = fm.compute()
    sp = finish_photo_specs(yaml.safe_load(open(os.path.join(SPECDIR, "Euclid-Photometric-ISTF-Pessimistic.yaml")))["specifications"])
    lum = np.load(os.path.join(REPO, "cosmicfishpie_b200", "data", "lumratio.npy"))
    ob = Bank(bank3, tt.FIDUCIAL, tt.COSMO_KEYS, eps_values=EPS)
    ref = photo_fisher(ob, sp, {"ell_sampling": 14}, ["WL"], dict(tt.FIDUCIAL), dict(sp["photo_z_params"]),
                       dict(sp["IA_params"]), {}, dict(fm.config.freeparams), lum, stencil="5PT")
    assert list(F.get_param_names()) == ref["param_names"]
    scale = np.sqrt(np.outer(np.diag(ref["fisher"]), np.diag(ref["fisher"])))
    assert np.max(np.abs(F.fisher_matrix - ref["fisher"]) / scale) < 1e-8


def test_photo_stem_vs_reference_golden(tmp_path):
    """SteM derivatives against the reference's own derivative engine and einsum assembly (golden photo_stem):
    6-point straight-line fits through the +-1, 2, 3 % tables for all 15 parameters of a 3x2pt forecast."""
    from conftest import golden
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    g = golden("photo_stem")
    eps = tuple(float(e) for e in g["eps_values"])
    ext = ext_of(tt.make_bank(eps_values=eps, pars=("Omegam", "h")))
    ext["eps_values"] = list(eps)
    o = base_options(tmp_path, stencil="STEM", ell_sampling=10, survey_name_photo="Euclid-Photometric-ISTF-Pessimistic",
                     survey_name_spectro=False)
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01, "h": 0.01}, options=o,
                      observables=["WL", "GCph"], extfiles=ext)
    F = fm.compute()
    assert list(F.get_param_names()) == list(g["param_names"])
    assert fm.photo_LSS._job.stw.shape == (15, 1 + 15 * 6)
    scale = np.sqrt(np.outer(np.diag(g["fisher"]), np.diag(g["fisher"])))
    assert np.max(np.abs(F.fisher_matrix - g["fisher"]) / scale) < 1e-6      # north_star tolerance on Fisher elements
    assert np.max(np.abs(np.array(F.get_confidence_bounds()) / g["sigma"] - 1)) < 1e-4
    d = fm.photo_LSS.compute_derivs()
    for par in ("Omegam", "b3", "AIA"):
        got = np.array([d[par][k] for k in g["cl_keys"]])
        assert np.max(np.abs(got - g["dcl__" + par])) <= 1e-9 * np.max(np.abs(g["dcl__" + par]))


def test_spectro_refuses_stem(bank3, tmp_path):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"h": 0.01}, options=base_options(tmp_path, stencil="STEM"),
                      observables=["GCsp"], extfiles=ext_of(bank3))
    with pytest.raises(ValueError, match="STEM derivative not availabe"):  # derivatives.py:427
        fm.compute(max_z_bins=1)


def test_photo_accuracy_2_vs_oracle(bank3, tmp_path):
    """accuracy = 2 doubles the redshift and multipole sampling (400 z, 200 multipoles): other tile shapes in the
    window, C_ell and Fisher kernels than any golden case."""
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix
    from oracle.cosmo import Bank
    from oracle.photo import default_photo_bias, finish_photo_specs, photo_fisher

    o = base_options(tmp_path, accuracy=2, survey_name_photo="Euclid-Photometric-ISTF-Pessimistic", survey_name_spectro=False)
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01, "h": 0.01}, options=o,
                      observables=["GCph"], extfiles=ext_of(bank3))
    F = fm.compute()
    assert fm.photo_LSS.grid.zsamp == 400 and fm.photo_LSS.grid.ellsamp == 200
    sp = finish_photo_specs(yaml.safe_load(open(os.path.join(SPECDIR, "Euclid-Photometric-ISTF-Pessimistic.yaml")))["specifications"])
    lum = np.load(os.path.join(REPO, "cosmicfishpie_b200", "data", "lumratio.npy"))
    ob = Bank(bank3, tt.FIDUCIAL, tt.COSMO_KEYS, eps_values=EPS)
    ref = photo_fisher(ob, sp, {"accuracy": 2}, ["GCph"], dict(tt.FIDUCIAL), dict(sp["photo_z_params"]), {},
                       default_photo_bias(sp), dict(fm.config.freeparams), lum)
    assert list(F.get_param_names()) == ref["param_names"]
    scale = np.sqrt(np.outer(np.diag(ref["fisher"]), np.diag(ref["fisher"])))
    assert np.max(np.abs(F.fisher_matrix - ref["fisher"]) / scale) < 1e-9
