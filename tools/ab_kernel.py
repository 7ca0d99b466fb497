"""Time the lattice kernel of the bench forecast (spectro_fisher_kernel, CUDA events inside the library, L2 flushed
between launches) for the libcfp_b200.so currently in the package.  Used to A/B library variants on the SAME box:
    for v in build_ab/*.so; do cp $v cosmicfishpie_b200/libcfp_b200.so; python tools/ab_kernel.py $v; done
Usage: python tools/ab_kernel.py [label] [n_runs]"""
import os
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

import bench  # noqa: E402
from cosmicfishpie_b200 import toy_tables as tt  # noqa: E402


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "lib"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    tmp = tempfile.mkdtemp(prefix="cfp_ab_")
    fm = bench.make_spectro_fm(tt.make_bank(eps_values=(0.01,)), tmp)
    fm.compute()
    plan = fm._spectro_job.plan()
    stream = torch.cuda.Stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    ms = []
    for i in range(n + 5):
        with torch.cuda.stream(stream):
            flush.zero_()
        plan.run(materialize=0, stream=stream.cuda_stream)
        if i >= 5:
            ms.append(plan.kernel_ms())
    F = plan.fetch()[0]
    print(f"{label}: kernel ms mean {np.mean(ms):.4f} min {np.min(ms):.4f} max {np.max(ms):.4f}  trace {np.trace(F.sum(axis=0)):.10e}")
    if len(sys.argv) > 3:  # likelihood kernels too: device time of a 2048-point nuisance batch and of 64 lone points
        from cosmicfishpie_b200.likelihood import SpectroLikelihood

        ctx = plan.ctx
        rng = np.random.default_rng(7)
        for flag in ("wedges", "legendre"):
            like = SpectroLikelihood(cosmoFM_data=fm, leg_flag=flag)
            for keys, B, sig in (([k for k in fm.freeparams if k.startswith("lnbg")], 2048, 0.01), (["h", "lnbg_1"], 64, 0.0)):
                fid = np.array([fm.allparams[k] for k in keys])
                mat = fid * (1.0 + sig * rng.standard_normal((B, len(keys))))
                ll = like.loglike_batch(mat, keys=keys)
                t = []
                for _ in range(5):
                    ll = like.loglike_batch(mat, keys=keys)
                    t.append(ctx.last_run_ms)
                print(f"{label}: {flag} {keys[0]}.. B={B}: device ms {np.mean(t):.3f}  sum loglike {np.sum(ll):.12e}")


def photo():
    """device time of photometric likelihood chunks (13+13 bins): python tools/ab_kernel.py label photo [B]"""
    from cosmicfishpie_b200.likelihood import PhotometricLikelihood
    import time

    label = sys.argv[1]
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
    tmp = tempfile.mkdtemp(prefix="cfp_ab_")
    fm = bench.make_photo_fm(tt.make_bank(eps_values=(0.01,)), tmp)
    fm.compute()
    like = PhotometricLikelihood(cosmo_data=fm)
    keys = [k for k in fm.freeparams if k not in bench.P7]
    fid = np.array([fm.allparams[k] for k in keys])
    mat = fid * (1.0 + 0.02 * np.random.default_rng(1).standard_normal((B, len(keys))))
    ll = like.loglike_batch(mat, keys=keys)
    ctx = fm.ctx
    t0 = time.perf_counter()
    ll = like.loglike_batch(mat, keys=keys)
    dt = time.perf_counter() - t0
    print(f"{label}: photo likelihood B={B}: {B / dt:.0f} evals/s wall, device ms of the last chunk {ctx.last_run_ms:.3f}, sum {np.sum(ll):.12e}")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "photo":
        photo()
    else:
        main()
