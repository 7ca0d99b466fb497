"""cProfile of the batched likelihoods (`loglike_batch`) on the bench configuration: where the host side goes.
Usage: python tools/profile_like.py [out.txt]"""
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import bench  # noqa: E402
from cosmicfishpie_b200 import toy_tables as tt  # noqa: E402
from cosmicfishpie_b200.likelihood import PhotometricLikelihood, SpectroLikelihood  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    tmp = tempfile.mkdtemp(prefix="cfp_prof_")
    bank = tt.make_bank(eps_values=(0.01,))
    a = bench.make_photo_fm(bank, tmp)
    a.compute()
    b = bench.make_spectro_fm(bank, tmp)
    b.compute()
    buf = io.StringIO()
    for name, like, keys, allp, B in (
            ("photo", PhotometricLikelihood(cosmo_data=a), [k for k in a.freeparams if k not in bench.P7], a.allparams, 16384),
            ("spectro", SpectroLikelihood(cosmoFM_data=b), [k for k in b.freeparams if k.startswith("lnbg")], b.allparams, 2048),
            ("spectro 16k", SpectroLikelihood(cosmoFM_data=b), [k for k in b.freeparams if k.startswith("lnbg")], b.allparams, 16384),
            ("legendre 16k", SpectroLikelihood(cosmoFM_data=b, leg_flag="legendre"),
             [k for k in b.freeparams if k.startswith("lnbg")], b.allparams, 16384)):
        fid = np.array([allp[k] for k in keys])
        mat = fid * (1.0 + 0.02 * np.random.default_rng(0).standard_normal((B, len(keys))))
        like.loglike_batch(mat, keys=keys)
        t0 = time.perf_counter()
        like.loglike_batch(mat, keys=keys)
        dt = time.perf_counter() - t0
        pr = cProfile.Profile()
        pr.enable()
        like.loglike_batch(mat, keys=keys)
        pr.disable()
        buf.write(f"==== {name}: {B / dt:.0f} evals/s, {dt * 1e3:.1f} ms per batch of {B}\n")
        pstats.Stats(pr, stream=buf).strip_dirs().sort_stats("cumulative").print_stats(28)
    txt = buf.getvalue()
    if out:
        open(out, "w").write(txt)
    else:
        print(txt)


if __name__ == "__main__":
    main()
