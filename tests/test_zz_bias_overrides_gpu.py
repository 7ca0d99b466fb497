"""GPU: ComputeGalSpectro.observed_Pgg with the two bias overrides outside the Fisher path -- b_i (a number replacing
the bias term) and use_bias_funcs=True (per-bin biases interpolated in z) -- against lattices of the unmodified
reference at redshifts between the bin centres (golden gcsp_kat_bias, tools/make_goldens.py)."""
import numpy as np
import pytest

from conftest import base_options, golden, rel_err

pytestmark = pytest.mark.gpu


def test_observed_pgg_bias_overrides_match_reference(extfiles, tmp_path):
    from cosmicfishpie_b200 import spectro
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    g = golden("gcsp_kat_bias")
    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"h": 0.01}, options=base_options(tmp_path),
                      observables=["GCsp"], extfiles=extfiles)
    cp = dict(tt.FIDUCIAL)
    cp["h"] = tt.FIDUCIAL["h"] * 1.01
    kk, mm = np.meshgrid(g["k"], g["mu"])
    plain = spectro.ComputeGalSpectro(cp, fiducial_cosmopars=dict(tt.FIDUCIAL), configuration=fm)
    funcs = spectro.ComputeGalSpectro(cp, fiducial_cosmopars=dict(tt.FIDUCIAL), configuration=fm, use_bias_funcs=True)
    b_i = float(g["b_i"])
    for i, z in enumerate(g["z"]):
        assert rel_err(plain.observed_Pgg(z, kk, mm), g["pgg_default"][i]) < 1e-10
        assert rel_err(plain.observed_Pgg(z, kk, mm, b_i=b_i), g["pgg_b_i"][i]) < 1e-10
        assert np.max(np.abs(plain.lnpobs_gg(z, kk, mm, b_i=b_i) - g["lnp_b_i"][i])) < 1e-10
        assert rel_err(funcs.observed_Pgg(z, kk, mm), g["pgg_bias_funcs"][i]) < 1e-10
    with pytest.raises(NotImplementedError):
        plain.observed_Pgg(1.0, kk, mm, b_i=np.ones_like(kk))


def test_single_element_helpers_agree_with_the_fused_kernel(extfiles, tmp_path):
    """fisher_calculation / fish_integrand / fisher_integral and compute_binned_derivs (cosmicfish.py:383-542) rebuild
    single Fisher elements from the materialised lattices: the same numbers the fused kernel reduced on the device."""
    from cosmicfishpie_b200 import toy_tables as tt
    from cosmicfishpie_b200.fishermatrix import FisherMatrix

    fm = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={"Omegam": 0.01, "h": 0.01},
                      options=base_options(tmp_path, spectro_Pk_k_samples=65, spectro_Pk_mu_samples=9),
                      observables=["GCsp"], extfiles=extfiles)
    fm.compute(max_z_bins=2)
    names = list(fm.freeparams)
    for zi in (0, 1):
        Fz = fm.fisher_per_bin(zi)
        for a, b in ((0, 0), (0, 1), (1, names.index("lnbg_%d" % (zi + 1)))):
            got = fm.fisher_calculation(zi, names[a], names[b])
            assert abs(got - Fz[a, b]) <= 1e-10 * np.sqrt(Fz[a, a] * Fz[b, b])
    binned = fm.compute_binned_derivs(np.array(fm.zmids))
    assert set(binned) == set(names) and binned["h"][0].shape == (9, 65)
    assert np.array_equal(binned["h"][1], fm.derivs_dict["h"][1])
