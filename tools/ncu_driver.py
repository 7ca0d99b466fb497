"""Minimal workload for `ncu`: GCsp Fisher (spectro_fisher_kernel), 3x2pt Fisher (photo kernels), a spectroscopic
wedge-likelihood batch (spectro_chi2_kernel) and a photometric likelihood batch (photo_chi2_kernel), each twice,
on the bench configuration.
Usage: python tools/ncu_driver.py [n_like_points]"""
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import bench  # noqa: E402
from cosmicfishpie_b200 import toy_tables as tt  # noqa: E402
from cosmicfishpie_b200.likelihood import PhotometricLikelihood, SpectroLikelihood  # noqa: E402


def main():
    n_like = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    tmp = tempfile.mkdtemp(prefix="cfp_ncu_")
    bank = tt.make_bank(eps_values=(0.01,))
    for _ in range(2):
        a = bench.make_photo_fm(bank, tmp)
        a.compute()
        b = bench.make_spectro_fm(bank, tmp)
        b.compute()
    like = SpectroLikelihood(cosmoFM_data=b)
    keys = [k for k in b.freeparams if k.startswith("lnbg")]
    fid = np.array([b.allparams[k] for k in keys])
    mat = fid * (1.0 + 0.02 * np.random.default_rng(0).standard_normal((n_like, len(keys))))
    for _ in range(2):
        ll = like.loglike_batch(mat, keys=keys)
    plike = PhotometricLikelihood(cosmo_data=a)
    pkeys = [k for k in a.freeparams if k not in bench.P7]
    pfid = np.array([a.allparams[k] for k in pkeys])
    pmat = pfid * (1.0 + 0.02 * np.random.default_rng(1).standard_normal((4 * n_like, len(pkeys))))
    for _ in range(2):
        pl = plike.loglike_batch(pmat, keys=pkeys)
    print("ok", float(np.sum(ll)), float(np.sum(pl)))


if __name__ == "__main__":
    main()
