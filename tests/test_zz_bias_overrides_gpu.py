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
