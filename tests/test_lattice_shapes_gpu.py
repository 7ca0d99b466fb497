"""GPU vs oracle on awkward lattice shapes: fewer wavenumbers than one 256-wide tile, exactly one tile, one more than
a tile, two mu rows, even sample counts (scipy's Simpson correction), and slices of those lattices.  The oracle's 3PT
mode is bit-exact against the reference goldens, so it stands in for the reference here."""
import os

import numpy as np
import pytest
import yaml

from conftest import REPO, base_options

pytestmark = pytest.mark.gpu
SPECDIR = os.path.join(REPO, "cosmicfishpie_b200", "data", "specs")
SHAPES = [(2, 3), (3, 256), (4, 257), (5, 64), (9, 513), (2, 1000)]


@pytest.mark.parametrize("nmu,nk", SHAPES)
def test_spectro_fisher_on_edge_shapes(toy_bank, extfiles, tmp_path, nmu, nk):
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix
    from oracle.cosmo import Bank
    from oracle.spectro import spectro_fisher

    small = {"spectro_Pk_k_samples": nk, "spectro_Pk_mu_samples": nmu}
    free = {"Omegam": 0.01, "h": 0.01, "ns": 0.01}
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars=dict(free), options=base_options(tmp_path, **small),
                      observables=["GCsp"], extfiles=extfiles)
    F = fm.compute(max_z_bins=2)
    sp = yaml.safe_load(open(os.path.join(SPECDIR, "Euclid-Spectroscopic-ISTF-Pessimistic.yaml")))["specifications"]
    ob = Bank(toy_bank, tt.FIDUCIAL, tt.COSMO_KEYS)
    ref = spectro_fisher(ob, sp, small, dict(tt.FIDUCIAL), dict(sp["sp_bias_parametrization"]["linear_log"]),
                         dict(sp["shot_noise_parametrization"]["constant"]), {}, dict(fm.config.freeparams), max_z_bins=2)
    assert list(F.get_param_names()) == ref["param_names"]
    scale = np.sqrt(np.outer(np.diag(ref["fisher"]), np.diag(ref["fisher"])))
    assert np.max(np.abs(F.fisher_matrix - ref["fisher"]) / scale) < 1e-10
    # the same lattice in three ragged slices
    plan = fm._spectro_job.plan()
    ncell = plan.Nmu * plan.Nk
    plan.run()
    full = plan.fetch()[0]
    acc = np.zeros_like(full)
    cuts = sorted({0, ncell // 3 + 1, (2 * ncell) // 3, ncell})
    for a, b in zip(cuts[:-1], cuts[1:]):
        plan.set_slice(a, b)
        plan.run()
        acc += plan.fetch()[0]
    plan.set_slice(0, ncell)
    assert np.max(np.abs(acc - full)) <= 1e-11 * np.max(np.abs(full))


def test_spectro_fisher_three_gram_blocks(toy_bank, extfiles, tmp_path):
    """19 free parameters (7 cosmological + 3 bins x (bias, shot noise, sigma_p and sigma_v rescaling)): the Gram
    tile has three 8-row blocks and the kernel runs 3 CTAs per SM."""
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix
    from oracle.cosmo import Bank
    from oracle.spectro import spectro_fisher

    small = {"spectro_Pk_k_samples": 300, "spectro_Pk_mu_samples": 7}
    nl = dict(rescale_sigmap=True, rescale_sigmav=True, default=None,
              rescale_sigma_pv={f"sigma{c}_{i}": 1.0 for c in "pv" for i in (1, 2, 3, 4)})
    free = {p: 0.01 for p in tt.COSMO_KEYS}
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars=dict(free),
                      options=base_options(tmp_path, survey_name_spectro="Euclid-Spectroscopic-ISTF-Pessimistic-sigma_pv",
                                           **small),
                      specifications={"nonlinear_parametrization": nl}, observables=["GCsp"], extfiles=extfiles)
    F = fm.compute(max_z_bins=3)
    assert F.fisher_matrix.shape == (19, 19)
    sp = yaml.safe_load(open(os.path.join(SPECDIR, "Euclid-Spectroscopic-ISTF-Pessimistic.yaml")))["specifications"]
    sp.update(yaml.safe_load(open(os.path.join(SPECDIR, "Euclid-Spectroscopic-ISTF-Pessimistic-sigma_pv.yaml")))["specifications"])
    sp["nonlinear_parametrization"] = nl
    ob = Bank(toy_bank, tt.FIDUCIAL, tt.COSMO_KEYS)
    ref = spectro_fisher(ob, sp, small, dict(tt.FIDUCIAL), dict(sp["sp_bias_parametrization"][sp["sp_bias_model"]]),
                         dict(sp["shot_noise_parametrization"][sp["shot_noise_model"]]), dict(nl["rescale_sigma_pv"]),
                         dict(fm.config.freeparams), max_z_bins=3)
    assert list(F.get_param_names()) == ref["param_names"]
    scale = np.sqrt(np.outer(np.diag(ref["fisher"]), np.diag(ref["fisher"])))
    assert np.max(np.abs(F.fisher_matrix - ref["fisher"]) / scale) < 1e-10
