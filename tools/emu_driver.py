"""Throughput of likelihood batches in which EVERY point has its own emulated cosmology (bank.TaylorEmulator): the
realistic sampler case.  Usage: python tools/emu_driver.py [B]"""
import os
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import bench  # noqa: E402
from cosmicfishpie_b200 import toy_tables as tt  # noqa: E402
from cosmicfishpie_b200.bank import TaylorEmulator  # noqa: E402
from cosmicfishpie_b200.fishermatrix import FisherMatrix  # noqa: E402
from cosmicfishpie_b200.likelihood import PhotometricLikelihood, SpectroLikelihood  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    bank = tt.make_bank()
    emu = TaylorEmulator(bank, tt.FIDUCIAL, tt.COSMO_KEYS, 0.01)
    tmp = tempfile.mkdtemp()
    ext = dict(bench.ext_dict(bank), emulator=emu)
    rng = np.random.default_rng(0)
    fp = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in bench.P7},
                      options=bench.base_options(tmp, survey_name_photo=bench.PHOTO_SPEC, survey_name_spectro=False),
                      observables=["WL", "GCph"], extfiles=ext)
    fp.compute()
    fs = FisherMatrix(fiducialpars=dict(tt.FIDUCIAL), freepars={p: 0.01 for p in bench.P7},
                      options=bench.base_options(tmp, survey_name_spectro=bench.SPECTRO_SPEC, survey_name_photo=False,
                                                 spectro_Pk_k_samples=1025, spectro_Pk_mu_samples=17),
                      observables=["GCsp"], extfiles=ext)
    fs.compute()
    for name, like, fm in (("photo 3x2pt 13+13", PhotometricLikelihood(cosmo_data=fp), fp),
                           ("spectro wedges 17x1025", SpectroLikelihood(cosmoFM_data=fs), fs)):
        keys = list(bench.P7) + [k for k in fm.freeparams if k not in bench.P7][:4]
        fid = np.array([fm.allparams[k] for k in keys])
        mat = np.tile(fid, (B, 1))
        mat[:, :7] = np.where(fid[:7] != 0, fid[:7] * (1 + 0.008 * rng.uniform(-1, 1, (B, 7))), 0.008 * rng.uniform(-1, 1, (B, 7)))
        mat[:, 7:] *= 1 + 0.01 * rng.standard_normal((B, len(keys) - 7))
        for label, env in (("device combine", "0"), ("host tables", "1")):
            if env == "1" and B > 256:
                sub = mat[:256]
            else:
                sub = mat
            os.environ["CFP_NO_COMBINE"] = env
            like.loglike_batch(sub[:8], keys=keys)
            t0 = time.perf_counter()
            ll = like.loglike_batch(sub, keys=keys)
            dt = time.perf_counter() - t0
            print(f"{name}, {len(sub)} points, every point its own emulated cosmology, {label}: "
                  f"{len(sub) / dt:.0f} evals/s ({dt * 1e3:.1f} ms), finite {bool(np.all(np.isfinite(ll)))}")
        os.environ["CFP_NO_COMBINE"] = "0"
        if "--profile" in sys.argv:
            import cProfile
            import pstats

            from cosmicfishpie_b200 import cosmology as cz

            cz.clear_caches()
            pr = cProfile.Profile()
            pr.enable()
            like.loglike_batch(mat, keys=keys)
            pr.disable()
            pstats.Stats(pr).sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
